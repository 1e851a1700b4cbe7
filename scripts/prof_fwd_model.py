"""Per-kernel device time of one config-2 forward (48 layers, (8, 1024), bf16 autocast) via torch.profiler."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.backbone import MixerStack  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
stack = MixerStack(2048, 48, device=dev).eval()
x = torch.randn(8, 1024, 2048, device=dev, dtype=torch.bfloat16)


def step():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return stack(x)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"wall per forward: {e0.elapsed_time(e1) / 5:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
    if t:
        rows.append((t, ev.count, ev.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"forward: {tot / 1e3:.2f} ms of kernel time, {sum(r[1] for r in rows)} launches")
for t, n, k in rows[:16]:
    print(f"  {t / 1e3:7.2f} ms {100 * t / tot:5.1f} %  {n:4d} x {t / n:8.1f} us   {k}")
