"""Summarise an .ncu-rep (one kernel, --set full --import-source on) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ssd_fwd_ncu.txt [units_per_launch]
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|" \
       r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__warps_active\.avg\.pct_of_peak_sustained_active|" \
       r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)|sm__cycles_elapsed\.max|" \
       r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|l1tex__data_pipe_lsu_wavefronts\.sum\.pct_of_peak_sustained_elapsed|" \
       r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|lts__t_sector_hit_rate\.pct|" \
       r"sm__inst_executed_pipe_(alu|fma|lsu|xu|tmem|uniform)\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$"


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none --import-source on; numbers under the profiler are",
             "# for SHARES and traffic, never bench values)"]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    traffic = None
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        lines.append(f"\n== kernel: {name}")
        vals = {}
        for i, h in enumerate(hdr):
            if re.match(KEYS, h):
                lines.append(f"{h:85s} {row[i]:>18s} {units[i]}")
                vals[h] = (row[i], units[i])
        try:
            mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            r, w = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            traffic = float(r[0]) * mul[r[1]] + float(w[0]) * mul[w[1]]
            lines.append(f"dram traffic per launch (read + write): {traffic / 1e6:.1f} MB")
        except Exception:
            pass
    src = ncu_csv(rep, "source")
    h2 = src[1]
    idx = {h: i for i, h in enumerate(h2)}
    # (a report with several launches repeats the two header rows per kernel: keep the instruction rows only)
    data = [r for r in src[2:] if len(r) == len(h2) and r[idx["# Samples"]].isdigit()]
    st = collections.Counter()
    for r in data:
        for h in h2:
            if h.startswith("stall_") and "Not" not in h:
                st[h] += int(r[idx[h]])
    tot = sum(st.values()) or 1
    lines.append(f"\n== warp-state samples over {len(data)} SASS instructions (all warps)")
    for k, v in st.most_common(10):
        lines.append(f"{k:28s} {v:8d}  {100.0 * v / tot:5.1f} %")
    lines.append("\n== top SASS lines by samples: samples, executed, instruction, main stalls")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:25]:
        stalls = {h[6:]: int(r[idx[h]]) for h in h2 if h.startswith("stall_") and "Not" not in h and int(r[idx[h]]) > 0}
        top = ", ".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:2])
        lines.append(f"{r[idx['# Samples']]:>7s} {r[idx['Instructions Executed']]:>10s}  {r[idx['Source']][:64]:64s} {top}")
    open(dst, "w").write("\n".join(lines) + "\n")
    if traffic is not None and len(sys.argv) > 3:
        json.dump({"dram_bytes_per_launch": traffic, "source": f"{dst} (ncu --set full, one launch)"},
                  open(sys.argv[3], "w"))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
