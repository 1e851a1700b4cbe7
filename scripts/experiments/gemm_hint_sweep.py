import sys, torch, ctypes
sys.path.insert(0, ".")
from omnimamba_b200 import _cabi
lib = _cabi.lib()
lib.omni_debug_set_mbar_hint.argtypes = [ctypes.c_uint]
x = torch.randn(64, 2048, device="cuda", dtype=torch.bfloat16)
ws = [torch.randn(8512, 2048, device="cuda", dtype=torch.bfloat16) for _ in range(8)]   # 280 MB of weights: HBM-cold each call
out = torch.empty(64, 8512, device="cuda", dtype=torch.bfloat16)
xb = torch.randn(65536, 2048, device="cuda", dtype=torch.bfloat16)
outb = torch.empty(65536, 8512, device="cuda", dtype=torch.bfloat16)
def timed(fn, n):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for hint in (20000, 1000, 100, 0):
    lib.omni_debug_set_mbar_hint(hint)
    torch.cuda.synchronize()
    a = timed(lambda i: _cabi.gemm(x, ws[i % 8], torch.bfloat16, out=out), 40)
    b = timed(lambda i: torch.mm(x, ws[i % 8].t(), out=out), 40)
    c = timed(lambda i: _cabi.gemm(xb, ws[i % 8], torch.bfloat16, out=outb), 10)
    print(f"hint {hint:6d} ns: skinny ours {a:.2f} us (cuBLAS {b:.2f} us), in_proj 65536 ours {c:.1f} us", flush=True)
