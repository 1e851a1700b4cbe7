"""Device time of the backward kernels at the bench shape (CUPTI): python scripts/bwd_kernel_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, L, H, P, N = 16, 4096, 64, 64, 128
rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
x, dt, Bm, Cm, dy = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N), rn(B, L, H, P)
A = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
dt_bias = torch.rand(H, device=dev, generator=g) * 4 - 6
D = torch.ones(H, device=dev)
run = lambda: ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, algo="chunked_tc")
for _ in range(3):
    run()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(8):
        run()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:4]:
    print(f"{e.device_time_total / e.count:9.1f} us x {e.count:3d}  {e.key[:70]}")
