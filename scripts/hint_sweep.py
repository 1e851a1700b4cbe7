"""Sweep the mbarrier suspend-time hint of the tensor-core SSD forward (run on the GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200 import _cabi
from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
import bench
lib = _cabi.lib()
host = bench.make_inputs(16, 4096)
dev = {k: v.cuda() for k, v in host.items()}
out = torch.empty(16, 4096, 64, 64, device="cuda", dtype=torch.bfloat16)
run = lambda: ssd_fwd_raw(dev["x"], dev["dt"], dev["A"], dev["B"], dev["C"], 256, D=dev["D"], dt_bias=dev["dt_bias"], dt_softplus=True, out=out)
for ns in [int(a) for a in sys.argv[1:]] or [20000, 5000, 1000, 300, 100, 30, 0]:
    lib.omni_debug_set_mbar_hint(ns)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"hint {ns:6d} ns: {e0.elapsed_time(e1)/10*1e3:8.1f} us/step")
