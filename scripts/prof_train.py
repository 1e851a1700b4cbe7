"""Per-kernel device time of one config-3 train step (bench_workloads.train) via torch.profiler.
    python scripts/prof_train.py [layers] > gpurun_out/prof_train.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.backbone import T2IModel  # noqa: E402
from omnimamba_b200.dist import BucketedGradReducer  # noqa: E402

n_layer = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = T2IModel(2048, n_layer, device=dev).freeze_backbones("align")
model.train()
for layer in model.backbone.layers:
    for nme, p in layer.mixer.in_proj.named_parameters():
        if "mmu_lora" in nme:
            p.requires_grad_(False)
params = [p for p in model.parameters() if p.requires_grad]
red = BucketedGradReducer(params)
opt = torch.optim.AdamW(params, lr=8e-4, betas=(0.9, 0.95), weight_decay=0.0, fused=True)
image_ids = torch.randint(0, 16384, (90, 256), device=dev)
caption_ids = torch.randint(0, 50277, (90, 73), device=dev)


def step():
    red.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(image_ids, caption_ids)
    loss.backward()
    red.finish()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
    if t:
        rows.append((t, ev.count, ev.key[:120]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"train step ({n_layer} layers): {tot / 1e3:.1f} ms of kernel time")
for t, n, k in rows[:40]:
    print(f"  {t / 1e3:8.2f} ms  {100 * t / tot:5.1f} %  {n:5d} launches  {t / n:9.1f} us each   {k}")
