import sys, torch, ctypes
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
import bench
from omnimamba_b200 import _cabi
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw, ssd_fwd_raw
lib = _cabi.lib()
lib.omni_debug_set_mbar_hint.argtypes = [ctypes.c_uint]
host = bench.make_inputs(16, 4096)
d = {k: v.cuda() for k, v in host.items()}
dy = torch.randn(16, 4096, 64, 64, device="cuda").bfloat16()
def bwd():
    ssd_bwd_raw(dy, d["x"], d["dt"], d["A"], d["B"], d["C"], 256, D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
def fwd():
    ssd_fwd_raw(d["x"], d["dt"], d["A"], d["B"], d["C"], 256, D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for hint in (20000, 2000, 500, 100, 0):
    lib.omni_debug_set_mbar_hint(hint)
    torch.cuda.synchronize()
    print(f"hint {hint:6d} ns: bwd {timed(bwd):.3f} ms  fwd {timed(fwd, 20):.4f} ms", flush=True)
