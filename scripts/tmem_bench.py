"""TMEM <-> register throughput per SM (run on the GPU box)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200 import _cabi
lib = _cabi.lib()
out = torch.zeros(128, dtype=torch.int64, device="cuda")
iters = 256
for mode, name, bytes_per_it in [(0, "ld 32x32b.x32 (wait each)", 4096), (2, "ld 32x32b.x32 x2 per wait", 4096), (1, "st 32x32b.x16 x2", 4096)]:
    for nw in (1, 4, 8, 12, 16):
        out.zero_()
        lib.omni_debug_tmem_bench(ctypes.c_void_p(out.data_ptr()), mode, nw, iters, None)
        torch.cuda.synchronize()
        cyc = out[:nw].max().item()
        tot = bytes_per_it * iters * nw
        print(f"{name:28s} warps={nw:2d}: {cyc:8d} cycles  -> {tot / cyc:7.1f} B/clk/SM, {cyc / iters:6.1f} cyc/iter/warp")
