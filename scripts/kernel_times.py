"""Device time per kernel of the bench-size scan forward (CUPTI via torch.profiler): python scripts/kernel_times.py [L]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B, H, P, N = 65536 // L, 64, 64, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=dev, generator=g) * 4 - 6
    D = torch.ones(H, device=dev)
    run = lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, algo="chunked_tc")
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            run()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    for e in rows[:8]:
        print(f"{e.device_time_total / e.count:9.1f} us x {e.count:3d}  {e.key[:90]}")


if __name__ == "__main__":
    main()
