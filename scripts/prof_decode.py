"""Per-kernel device time of one captured decode step (config 4), via torch.profiler on graph replays.
    python scripts/prof_decode.py [bf16|fp32] > gpurun_out/prof_decode.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.backbone import InferenceParams, MixerStack  # noqa: E402

dtype = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else torch.bfloat16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
stack = MixerStack(2048, 48, device=dev, dtype=dtype, lora=False).eval()
ip = InferenceParams(max_seqlen=512, max_batch_size=64)
with torch.no_grad():
    stack(torch.randn(64, 72, 2048, device=dev, dtype=dtype), ip)
    ip.seqlen_offset = 72
    tok = torch.randn(64, 1, 2048, device=dev, dtype=dtype)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            stack(tok, ip)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = stack(tok, ip)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    steps = 10
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            graph.replay()
        torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
    if t:
        rows.append((t / steps, ev.count / steps, ev.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"decode step ({dtype}): {tot / 1e3:.3f} ms of kernel time per step")
for t, n, k in rows[:14]:
    print(f"  {t:9.1f} us/step  {n:6.1f} launches  {t / n:7.2f} us each   {k}")
