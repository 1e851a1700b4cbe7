#!/bin/bash
# compute-sanitizer over the kernels added / changed in round 2c (run on the GPU box): the forward that keeps its chunk states
# (kernel MODE 3: TMA stores of the S16 tile beside the MMA's reads; half-item hand-off schedule included) followed by the backward
# on the kept states, the packed-pair conv1d fast kernels (ragged L, both directions) and the add + RMSNorm backward fast kernel.
set -u
mkdir -p gpurun_out
cat > /tmp/san_r2c.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests/golden")
from cases import scan_inputs
from omnimamba_b200.interface.ssd_combined import _alloc_chunk_states, ssd_bwd_raw, ssd_fwd_raw
which = sys.argv[1]
c = lambda t: t.cuda()
if which == "fwdkeep":
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(5, 300, 64, 64, 1, 128, 0, torch.bfloat16)
    cs = _alloc_chunk_states(5, 300, 64, 64, 128, "cuda", torch.bfloat16)
    kw = dict(D=c(D), dt_bias=c(dt_bias), dt_softplus=True, algo="chunked_tc")
    out, fin, kept = ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, return_final_states=True, chunk_states=cs, **kw)
    assert kept is cs
    ssd_bwd_raw(c(torch.randn_like(x)), c(x), c(dt), c(A), c(Bm), c(Cm), 256, chunk_states=kept, **kw)
elif which == "conv":
    from omnimamba_b200.interface import causal_conv1d_fn
    for (B, L) in ((2, 329), (1, 67)):
        zx = torch.randn(B, L, 8512, device="cuda").bfloat16().requires_grad_()
        w, b = torch.randn(4352, 4, device="cuda").requires_grad_(), torch.randn(4352, device="cuda").requires_grad_()
        y = causal_conv1d_fn(zx[..., 4096:4096 + 4352].transpose(1, 2), w, b, activation="silu")
        y.backward(torch.randn_like(y))
else:
    from omnimamba_b200.interface.layer_norm import layer_norm_fn
    x = torch.randn(300, 2048, device="cuda").bfloat16().requires_grad_()
    r = torch.randn(300, 2048, device="cuda").requires_grad_()
    w = torch.ones(2048, device="cuda").requires_grad_()
    y, ro = layer_norm_fn(x, w, None, residual=r, prenorm=True, residual_in_fp32=True, eps=1e-5, is_rms_norm=True)
    (y.float().sum() + ro.sum()).backward()
torch.cuda.synchronize()
print("case", which, "done")
PY
run() { tool=$1; case=$2
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_r2c.py $case > gpurun_out/r2c_sanitizer_${tool}_${case}.log 2>&1
  echo "[$tool $case] rc=$? $(grep 'case .* done' gpurun_out/r2c_sanitizer_${tool}_${case}.log | tail -1) | $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2c_sanitizer_${tool}_${case}.log | tail -1)"
}
run memcheck fwdkeep; run racecheck fwdkeep; run synccheck fwdkeep
run memcheck conv; run memcheck addnorm; run racecheck addnorm
