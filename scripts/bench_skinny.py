"""Decode-shaped projections (M = batch 64) with the weights coming from HBM: eight weight sets (> L2) rotate, as the 48 layers
of a decode step do.  Ours (csrc/gemm_skinny.cu) vs the tile kernel of gemm_tc.cu (debug mode 3) vs cuBLAS (torch.mm).

    python scripts/bench_skinny.py > gpurun_out/bench_skinny.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200 import _cabi  # noqa: E402


def timed(fns, reps=10, warmup=3):
    """The calls of `fns` (one per weight set) captured back to back in ONE CUDA graph, as the decode step replays them: per-call
    host work (ctypes, tensor-map encodes: ~20 us, more than the kernels take) stays outside the measurement."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    for _ in range(warmup):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns))


def main():
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6548.5
    dev = "cuda"
    lib = _cabi.lib()
    rows = []
    for M in (64, 128, 16):
        for name, N, K in (("in_proj", 8512, 2048), ("out_proj", 2048, 4096)):
            ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(12)]
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ours = timed([lambda w=w: _cabi.gemm(x, w, torch.bfloat16, out=out) for w in ws])
            forced = {}
            if M == 64:
                for ks in (1, 2, 4, 8):
                    lib.omni_debug_set_gemm_mode(10 + ks)
                    forced[ks] = timed([lambda w=w: _cabi.gemm(x, w, torch.bfloat16, out=out) for w in ws]) * 1e3
                lib.omni_debug_set_gemm_mode(10)
            lib.omni_debug_set_gemm_mode(3)
            tile = timed([lambda w=w: _cabi.gemm(x, w, torch.bfloat16, out=out) for w in ws])
            lib.omni_debug_set_gemm_mode(0)
            cub = timed([lambda w=w: torch.mm(x, w.t(), out=out) for w in ws])
            by = N * K * 2 + M * K * 2 + M * N * 2
            rows.append({"case": name, "M": M, "N": N, "K": K, "ours_us": ours * 1e3, "tile_kernel_us": tile * 1e3, "cublas_us": cub * 1e3,
                         "forced_ksplit_us": forced, "bytes": by, "ours_gbs": by / ours / 1e6, "frac_of_hbm": by / ours / 1e6 / hbm})
            print(rows[-1], file=sys.stderr, flush=True)
    # fp32 operands (the model as inference_t2i.py runs it): 3xTF32 kernel vs cuBLAS SGEMM
    for M in (64,):
        for name, N, K in (("in_proj fp32", 8512, 2048), ("out_proj fp32", 2048, 4096)):
            ws = [torch.randn(N, K, device=dev) * 0.02 for _ in range(8)]
            x = torch.randn(M, K, device=dev)
            out = torch.empty(M, N, device=dev)
            ours = timed([lambda w=w: _cabi.gemm_f32_decode(x, w, out=out) for w in ws])
            cub = timed([lambda w=w: torch.mm(x, w.t(), out=out) for w in ws])
            by = N * K * 4 + M * K * 4 + M * N * 4
            rows.append({"case": name, "M": M, "N": N, "K": K, "ours_us": ours * 1e3, "cublas_us": cub * 1e3, "bytes": by,
                         "ours_gbs": by / ours / 1e6, "frac_of_hbm": by / ours / 1e6 / hbm})
            print(rows[-1], file=sys.stderr, flush=True)
    print(json.dumps({"hbm_gbs": hbm, "rows": rows}))


if __name__ == "__main__":
    main()
