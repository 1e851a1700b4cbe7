#!/bin/bash
# compute-sanitizer over the tensor-core kernels (run on the GPU box): memcheck + racecheck + synccheck of one small forward /
# backward / GEMM call each.  Logs -> gpurun_out/r2_sanitizer_*.log (copied to profiles/ when clean).
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
from cases import scan_inputs
from omnimamba_b200 import _cabi
from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw, ssd_fwd_raw
which = sys.argv[1]
c = lambda t: t.cuda()
if which in ("fwd", "bwd"):
    # 5 x 300 x 64 heads: 160 items on 148 SMs -> the half-item hand-off schedule runs as well
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(5, 300, 64, 64, 1, 128, 0, torch.bfloat16)
    if which == "fwd":
        ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), dt_softplus=True, return_final_states=True, algo="chunked_tc")
    else:
        dy = torch.randn_like(x)
        ssd_bwd_raw(c(dy), c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), dt_softplus=True, algo="chunked_tc")
else:
    a = torch.randn(704, 520, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(1000, 520, device="cuda", dtype=torch.bfloat16)
    _cabi.gemm(a, b, torch.bfloat16)
    _cabi.gemm(a.t().contiguous().t(), b.t().contiguous().t(), torch.float32)
torch.cuda.synchronize()
print("case", which, "done")
PY
for tool in memcheck racecheck synccheck; do
  for case in fwd bwd gemm; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py $case > gpurun_out/r2_sanitizer_${tool}_${case}.log 2>&1
    echo "$tool $case rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/r2_sanitizer_${tool}_${case}.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_${case}.log | tail -1)"
  done
done
