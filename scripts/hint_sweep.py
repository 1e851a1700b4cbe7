"""Sweep the mbarrier suspend-time hint of the tensor-core SSD forward (run on the GPU box).
Timed as bench.py times the scan: 5 warm-up + 20 steps after an idle gap (back-to-back runs of seconds settle ~10 % slower
and are noisier), three rounds, median reported."""
import os, sys, time
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
hints = [int(a) for a in sys.argv[1:]] or [20000, 100000, 5000, 1000, 300, 100, 0]
res = {h: [] for h in hints}
for rep in range(3):
    for ns in hints:
        lib.omni_debug_set_mbar_hint(ns)
        time.sleep(1.0)
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        res[ns].append(e0.elapsed_time(e1) / 20 * 1e3)
for ns in hints:
    print(f"hint {ns:6d} ns: median {sorted(res[ns])[1]:8.1f} us/step   {['%.1f' % v for v in res[ns]]}")
