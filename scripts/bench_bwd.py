"""Time the SSD backward (tensor-core vs recurrent) at the bench size on the GPU box."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw, ssd_fwd_raw
import bench
B, L = 16, 4096
host = bench.make_inputs(B, L)
dev = {k: v.cuda() for k, v in host.items()}
dy = torch.randn(B, L, 64, 64, device="cuda").bfloat16()
dx = torch.empty_like(dev["x"]); ddt = torch.empty_like(dev["dt"])
def run(algo):
    return ssd_bwd_raw(dy, dev["x"], dev["dt"], dev["A"], dev["B"], dev["C"], 256, D=dev["D"], dt_bias=dev["dt_bias"], dt_softplus=True, dx=dx, ddt=ddt, algo=algo)
for algo in (["chunked_tc"] + (["recurrent"] if "--rec" in sys.argv else [])):
    for _ in range(2): run(algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): r = run(algo)
    e1.record(); torch.cuda.synchronize()
    print(f"{algo:12s}: {e0.elapsed_time(e1)/n:8.3f} ms/step  ({B*L/(e0.elapsed_time(e1)/n)*1e3/1e6:.1f} M tokens/s)")
if "--rec" in sys.argv:
    a = run("chunked_tc"); b = run("recurrent")
    names = ["dx", "ddt", "dA", "dB", "dC", "dD", "dz", "ddt_bias"]
    for nme, u, v in zip(names, a, b):
        if u is not None:
            print(nme, ((u.double() - v.double()).norm() / v.double().norm()).item())
