"""Per-phase timeline of the tensor-core SSD backward gradient kernel (CTA 0), run on the GPU box."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200 import _cabi
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw
import bench
NAMES = ["item start", "tables done", "tiles arrived", "x/dy fp16", "G1,G3 done", "PT built", "G2 done", "dx epilogue done", "G4(h0) done",
         "M/MT(h0) built", "G5,G7(h0)+G4(h1) done", "M/MT(h1) built", "G5,G7(h1) done", "scaled + zc", "G6,G8,G10 done", "roff done",
         "da/ddt done", "dC/dB reduced"]
B, L = 16, 4096
host = bench.make_inputs(B, L)
dev = {k: v.cuda() for k, v in host.items()}
dy = torch.randn(B, L, 64, 64, device="cuda").bfloat16()
run = lambda: ssd_bwd_raw(dy, dev["x"], dev["dt"], dev["A"], dev["B"], dev["C"], 256, D=dev["D"], dt_bias=dev["dt_bias"], dt_softplus=True, algo="chunked_tc")
run(); torch.cuda.synchronize()
n = 12
buf = torch.zeros(n * 32, dtype=torch.int64, device="cuda")
lib = _cabi.lib()
lib.omni_debug_set_bwd_trace(ctypes.c_void_p(buf.data_ptr()), n)
run(); torch.cuda.synchronize()
lib.omni_debug_set_bwd_trace(None, 0)
t = buf.cpu().view(n, 32)
for it in range(4, 8):
    print(f"--- item {it}: total {int(t[it + 1, 0] - t[it, 0])} cycles")
    for e in range(1, len(NAMES)):
        print(f"   {int(t[it, e] - t[it, e - 1]):7d}  -> {NAMES[e]}")
    print(f"   {int(t[it + 1, 0] - t[it, len(NAMES) - 1]):7d}  -> next item")
