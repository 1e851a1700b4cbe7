"""Per-role timeline of the state sweeps (modes 1 / 2 of the tensor-core SSD kernel, CTA 0), run on the GPU box:
    python scripts/trace_sweep.py [mode] > gpurun_out/trace_sweep.txt"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200 import _cabi
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw
import bench
NAMES = {0: "tma:x_issue(g+1)", 1: "tma:B_issue(g+2)", 5: "mma:U", 8: "tab:start", 9: "tab:free", 10: "tab:ready", 15: "S:u_done(g-1)",
         16: "S:s_ready", 17: "X:full_x", 18: "X:x16_ready", 26: "tab:dt done", 27: "tab:scan done", 28: "tab:exp done"}
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, L = 16, 4096
host = bench.make_inputs(B, L)
dev = {k: v.cuda() for k, v in host.items()}
dy = torch.randn(B, L, 64, 64, device="cuda").bfloat16()
run = lambda: ssd_bwd_raw(dy, dev["x"], dev["dt"], dev["A"], dev["B"], dev["C"], 256, D=dev["D"], dt_bias=dev["dt_bias"], dt_softplus=True, algo="chunked_tc")
run(); torch.cuda.synchronize()
n = 24
buf = torch.zeros(n * 32, dtype=torch.int64, device="cuda")
lib = _cabi.lib()
lib.omni_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), 1000 * mode + n)
run(); torch.cuda.synchronize()
lib.omni_debug_set_trace(None, 0)
t = buf.cpu().view(n, 32)
t0 = int(t[0][t[0] > 0].min())
for c in range(6, 12):
    print(f"--- chunk {c}")
    for ts, name in sorted((int(t[c, e]) - t0, NAMES[e]) for e in NAMES if t[c, e] > 0):
        print(f"  {ts:9d}  {name}")
print("U-issue period (cycles):", [int(t[c + 1, 5]) - int(t[c, 5]) for c in range(4, n - 1)])
