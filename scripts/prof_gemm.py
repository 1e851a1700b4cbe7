"""One call of the tcgen05 GEMM at a given shape, for ncu:  python scripts/prof_gemm.py M N K [iters]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnimamba_b200 import _cabi
M, N, K = (int(v) for v in sys.argv[1:4])
it = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(6)]  # different weights per call: HBM, not L2
for i in range(it):
    _cabi.gemm(a, ws[i % 6], torch.bfloat16)
torch.cuda.synchronize()
