#!/bin/bash
# compute-sanitizer over the decode-chain kernels added in the second half of round 2 (run on the GPU box): the
# weight-streaming GEMM (every cluster size, LoRA pair), the layer core at 4 CTAs per SM, and a two-layer decode step launched
# with programmatic dependent launch.  Logs -> gpurun_out/r2b_sanitizer_*.log
set -u
mkdir -p gpurun_out
cat > /tmp/san_dec.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from omnimamba_b200 import _cabi
which = sys.argv[1]
dev = "cuda"
if which == "skinny":
    lib = _cabi.lib()
    for (M, N, K) in ((64, 8512, 2048), (3, 2048, 4096), (128, 1000, 520)):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
        t, bl = torch.randn(M, 8, device=dev, dtype=torch.bfloat16), torch.randn(N, 8, device=dev, dtype=torch.bfloat16)
        for ks in (0, 1, 2, 4):
            lib.omni_debug_set_gemm_mode(10 + ks)
            _cabi.gemm(x, w, torch.bfloat16)
            _cabi.gemm(x, w, torch.float32, t, bl)
        lib.omni_debug_set_gemm_mode(10)
else:
    from omnimamba_b200.backbone import InferenceParams, MixerStack
    torch.manual_seed(0)
    stack = MixerStack(2048, 2, device=dev, dtype=torch.bfloat16, lora=False).eval()
    ip = InferenceParams(max_seqlen=32, max_batch_size=8)
    with torch.no_grad():
        stack(torch.randn(8, 5, 2048, device=dev, dtype=torch.bfloat16), ip)
        ip.seqlen_offset = 5
        for _ in range(2):
            stack(torch.randn(8, 1, 2048, device=dev, dtype=torch.bfloat16), ip)
torch.cuda.synchronize()
print("case", which, "done")
PY
for tool in memcheck racecheck synccheck; do
  for case in skinny step; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_dec.py $case > gpurun_out/r2b_sanitizer_${tool}_${case}.log 2>&1
    echo "[$tool $case] rc=$? $(grep 'case .* done' gpurun_out/r2b_sanitizer_${tool}_${case}.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2b_sanitizer_${tool}_${case}.log | tail -1)"
  done
done
