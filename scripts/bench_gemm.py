"""The tcgen05 GEMM of libomnissm (csrc/gemm_tc.cu) beside cuBLAS (torch.mm) at the projection shapes of the d_model=2048
Mamba-2 block: forward, dgrad and wgrad of in_proj (2048 -> 8512, + LoRA r=8) and out_proj (4096 -> 2048) for the bench
token counts (65 536 = 16 x 4096, 29 610 = 90 x 329, 8 192 = 8 x 1024) and the decode batch (64).  CUDA events, 5 warm-ups +
20 steps per case; the roofline denominator is MEASURED_PEAKS.json bf16_tflops (cuBLAS burst on this pool's B200s).

    python scripts/bench_gemm.py > gpurun_out/bench_gemm.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200 import _cabi  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["bf16_tflops"]) if os.path.exists(peaks) else 1670.0
    dev = "cuda"
    rn = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
    rows = []
    for M in (65536, 29610 - 29610 % 8 + 8, 8192, 64):
        x, w_in = rn(M, 2048), rn(8512, 2048) * 0.02
        y, w_out = rn(M, 4096), rn(2048, 4096) * 0.02
        dz, do = rn(M, 8512), rn(M, 2048)
        t, lb = rn(M, 8), rn(8512, 8)
        cases = [
            ("in_proj fwd", x, w_in, None, None),
            ("in_proj fwd + LoRA pair", x, w_in, t, lb),
            ("in_proj dgrad", dz, w_in.t(), None, None),
            ("in_proj wgrad", dz.t(), x.t(), None, None),
            ("out_proj fwd", y, w_out, None, None),
            ("out_proj dgrad", do, w_out.t(), None, None),
            ("out_proj wgrad", do.t(), y.t(), None, None),
        ]
        for name, a, b, a2, b2 in cases:
            if M == 64 and "wgrad" in name:
                continue
            flops = 2.0 * a.shape[0] * b.shape[0] * (a.shape[1] + (a2.shape[1] if a2 is not None else 0))
            out = torch.empty(a.shape[0], b.shape[0], device=dev, dtype=torch.bfloat16)
            ours = timed(lambda: _cabi.gemm(a, b, torch.bfloat16, a2, b2, out=out))
            if a2 is None:
                ref = timed(lambda: torch.mm(a, b.t(), out=out))
            else:
                ref = timed(lambda: torch.addmm(torch.mm(a2, b2.t()), a, b.t(), out=out))
            rows.append({"case": name, "M": a.shape[0], "N": b.shape[0], "K": a.shape[1], "ours_ms": ours, "cublas_ms": ref,
                         "ours_tflops": flops / ours / 1e9, "cublas_tflops": flops / ref / 1e9, "frac_of_measured_peak": flops / ours / 1e9 / peak,
                         "ours_over_cublas": ref / ours})
            print(rows[-1], file=sys.stderr, flush=True)
    print(json.dumps({"peak_bf16_tflops": peak, "rows": rows}))


if __name__ == "__main__":
    main()
