"""Per-role timeline of the tensor-core SSD kernel (CTA 0): prints, per chunk, cycles between pipeline events.
Run on the GPU box:  python scripts/trace_tc.py [L] > gpurun_out/trace.txt"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200 import _cabi  # noqa: E402
from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw  # noqa: E402

NAMES = {0: "tma:x_issue(g+1)", 1: "tma:B_issue(g+2)", 2: "tma:C_issue(g+1)", 3: "mma:CB(g+1)", 4: "mma:Yoff", 5: "mma:U", 6: "mma:Yd",
         8: "tab:start", 9: "tab:free", 10: "tab:ready", 12: "P:cb_done", 13: "P:done", 14: "X:dx_ready", 15: "S:u_done(g-1)",
         16: "S:s_ready", 17: "X:full_x", 18: "X:x16_ready", 19: "E:acc_done", 22: "E:stored", 23: "E:s4=0", 24: "E:s4=1", 25: "E:s4=2"}


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B, H, P, N = 16, 64, 64, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=dev, generator=g) * 4 - 6
    D = torch.ones(H, device=dev)
    nch = 24
    buf = torch.zeros(nch * 32, dtype=torch.int64, device=dev)
    lib = _cabi.lib()
    run = lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, algo="chunked_tc")
    run()
    torch.cuda.synchronize()
    lib.omni_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), nch)
    run()
    torch.cuda.synchronize()
    lib.omni_debug_set_trace(None, 0)
    t = buf.cpu().view(nch, 32)
    t0 = int(t[0][t[0] > 0].min())
    for c in range(nch):
        ev = sorted((int(t[c, e]) - t0, NAMES[e]) for e in NAMES if t[c, e] > 0)
        print(f"--- chunk {c}")
        prev = None
        for ts, name in ev:
            print(f"  {ts:9d}  {name}")
    # per-chunk period
    per = [(int(t[c + 1, 3]) - int(t[c, 3])) for c in range(4, nch - 1)]
    print("CB-issue period (cycles):", per)


if __name__ == "__main__":
    main()
