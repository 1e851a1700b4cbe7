"""Join the SASS page of an .ncu-rep (one kernel, --import-source on) with nvdisasm's line info of the built library and
print, per source line of a .cu file, the stall samples (and their main reasons) and instruction counts.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep ssd_tc 'ssd_tc_fwd_kernelILi0ELb0' [first_line last_line]

(read-only analysis helper; needs cuobjdump / nvdisasm / ncu on PATH, no GPU)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(tu, mangled_part):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "omnimamba_b200/lib/libomnissm.so")], cwd=tmp, capture_output=True)
    out = subprocess.run(["nvdisasm", "-gi", "-c", f"{tu}.sm_100a.cubin"], cwd=tmp, capture_output=True, text=True).stdout
    res, active, cur = {}, False, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            active = mangled_part in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
            # attribute inlined helpers to the outermost call site (the last "inlined at" of the chain)
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if chain:
                cur = (os.path.basename(chain[-1][0]), int(chain[-1][1]), m.group(3))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            res[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return res


def main():
    rep, tu, part = sys.argv[1], sys.argv[2], sys.argv[3]
    lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10 ** 9)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(rows[2][ia], 16)
    lines = sass_lines(tu, part)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
    total = 0
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        info = lines.get(off)
        key = info[0][:2] if info and info[0] else ("?", 0)
        a = agg[key]
        s = int(r[isamp] or 0)
        a[0] += s
        a[1] += int(r[iexec] or 0)
        a[3] += 1
        total += s
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[2][h[6:]] += v
    print(f"# {rep}: {total} samples; per source line: samples, share, warp instructions executed, SASS instructions, top stalls")
    for key in sorted(agg, key=lambda k: (k[0], k[1])):
        f, l = key
        a = agg[key]
        if a[0] == 0 and a[1] == 0:
            continue
        if f.endswith(".cu") and not (lo <= l <= hi):
            continue
        top = ", ".join(f"{k} {v}" for k, v in a[2].most_common(3))
        print(f"{f}:{l:5d}  {a[0]:7d} {100.0 * a[0] / max(total, 1):5.1f}%  exec {a[1]:9d}  sass {a[3]:4d}  {top}")


if __name__ == "__main__":
    main()
