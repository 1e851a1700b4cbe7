"""Experiment harness for schedule variants of the tensor-core SSD forward.

Apply `ssd_tc_variants.patch` to omnimamba_b200/csrc/ssd_tc.cu (it adds a `VAR` template parameter whose bits switch schedule
changes on and off, and a launcher that picks the variant per launch from the OMNI_TC_VARIANT environment variable), build with
    OMNI_NVCC_EXTRA=-DOMNI_TC_VARIANTS python -m omnimamba_b200.build
and run on the GPU box:
    python scripts/experiments/gpu_variants.py 1 33 49 ... > gpurun_out/variants.txt
Every variant is timed like bench.py times the scan (5 warm-up + 20 steps after an idle gap: repeatable to ~0.1 %; seconds of
back-to-back launches settle ~10 % slower and scatter by +-3 %) and its y is compared bit for bit with variant 0.
Results of the round-1 run: profiles/r1_ssd_fwd_variants_v13.txt."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw  # noqa: E402


def main():
    variants = [0] + [int(v) for v in sys.argv[1:]]
    B, L, H, P, N = 16, 4096, 64, 64, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=dev, generator=g) * 4 - 6
    D = torch.ones(H, device=dev)
    run = lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, algo="chunked_tc")

    def timed(n=20):
        time.sleep(1.0)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ref = None
    res = {}
    for rep in range(3):
        for v in variants:
            os.environ["OMNI_TC_VARIANT"] = str(v)
            y = run()[0]
            torch.cuda.synchronize()
            if ref is None:
                ref = y.clone()
            same = bool(torch.equal(y, ref))
            ms = timed()
            res.setdefault(v, []).append(ms)
            print(f"rep {rep} variant {v:3d}  {ms:.4f} ms  bit-equal-to-v0 {same}", flush=True)
    med = {v: sorted(r)[len(r) // 2] for v, r in res.items()}
    for v in variants:
        print(f"variant {v:3d}  min {min(res[v]):.4f}  median {med[v]:.4f}")
    best = min(med, key=med.get)
    print("best (median)", best, med[best], "baseline", med[0])


if __name__ == "__main__":
    main()
