"""HBM roofline of the other hot-path rows of SURVEY.md 8(a) at d_model=2048 (run on the GPU box):

    python scripts/bench_ops.py > gpurun_out/bench_ops.json

Each op is called through the public operator surface on device-resident tensors larger than L2 (or with an L2 flush
between iterations for the small decode tensors), timed with CUDA events on the launch stream; `achieved` = ALGORITHMIC
bytes / time against MEASURED_PEAKS.json hbm_gbs.  One JSON object per line.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface import causal_conv1d_fn, causal_conv1d_update, selective_state_update  # noqa: E402
from omnimamba_b200.interface.layer_norm import layer_norm_fn  # noqa: E402
from omnimamba_b200.interface.layernorm_gated import rmsnorm_fn  # noqa: E402
from omnimamba_b200.interface.ssd_combined import mamba_split_conv1d_scan_combined  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, steps=20, warmup=5, flush=None):
    """`steps` launches back to back between two events (host-side wrapper time overlaps the previous launch); with `flush`
    an L2-sized memset runs before every launch and its own time, measured the same way, is subtracted."""
    def loop(body):
        for _ in range(warmup):
            body()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            body()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps / 1e3
    if flush is None:
        return loop(fn)
    both = loop(lambda: (flush.zero_(), fn()))
    return max(both - loop(lambda: flush.zero_()), 1e-9)


def main():
    dev = "cuda"
    hbm = peak()
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, dtype=torch.bfloat16: torch.randn(*s, device=dev, generator=g).to(dtype)
    B, L, d_inner, conv_dim, H, P, N = 16, 4096, 4096, 4352, 64, 64, 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out = []

    # a3 causal_conv1d_fn: channel-last view of zxbcdt's xBC slice (row pitch 8512), SiLU
    zx = rn(B, L, 8512)
    xBC = zx[..., d_inner:d_inner + conv_dim].transpose(1, 2)
    w, b = rn(conv_dim, 4, dtype=torch.float32), rn(conv_dim, dtype=torch.float32)
    t = timeit(lambda: causal_conv1d_fn(xBC, w, b, activation="silu"))
    by = 2 * B * L * conv_dim * 2
    out.append(dict(op="causal_conv1d_fn fwd (B=16, L=4096, D=4352, channel-last slice, silu)", ms=t * 1e3, bytes=by))

    # a5 gated RMSNorm (norm_before_gate=False), rows of 4096
    y, z, nw = rn(B * L, d_inner), rn(B * L, d_inner), torch.ones(d_inner, device=dev)
    t = timeit(lambda: rmsnorm_fn(y, nw, None, z=z, eps=1e-5, group_size=d_inner, norm_before_gate=False))
    by = 3 * B * L * d_inner * 2
    out.append(dict(op="rmsnorm_fn gated fwd (65536 rows x 4096)", ms=t * 1e3, bytes=by))

    # a3 / a5 backward (training path): conv1d bwd reads x, dout and writes dx; gated norm bwd reads x, z, dy and writes dx, dz
    xg = xBC.detach().requires_grad_()
    wg, bg = w.clone().requires_grad_(), b.clone().requires_grad_()
    yc = causal_conv1d_fn(xg, wg, bg, activation="silu")
    dyc = torch.randn_like(yc)
    t = timeit(lambda: torch.autograd.grad(yc, (xg, wg, bg), dyc, retain_graph=True))
    out.append(dict(op="causal_conv1d_fn bwd (same shape; incl. torch partial-sum reductions)", ms=t * 1e3, bytes=3 * B * L * conv_dim * 2))
    del yc, dyc, xg
    yg, zg, nwg = y.detach().requires_grad_(), z.detach().requires_grad_(), nw.clone().requires_grad_()
    yn = rmsnorm_fn(yg, nwg, None, z=zg, eps=1e-5, group_size=d_inner, norm_before_gate=False)
    dyn = torch.randn_like(yn)
    t = timeit(lambda: torch.autograd.grad(yn, (yg, zg, nwg), dyn, retain_graph=True))
    out.append(dict(op="rmsnorm_fn gated bwd (65536 rows x 4096; incl. torch partial-sum reduction)", ms=t * 1e3, bytes=5 * B * L * d_inner * 2))
    del yn, dyn, yg, zg

    # a2 the fused training op without out_proj (conv1d + SiLU -> SSD -> gated RMSNorm): SURVEY 8(d) secondary figure,
    # 25 216 B/token forward (read zxbcdt 17 024, write y 8 192); fwd+bwd adds dy 8 192 + re-read 17 024 + dzxbcdt 17 024
    import math
    dt0 = torch.exp(torch.rand(H, device=dev, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3)).clamp(min=1e-4)
    dtb_f = dt0 + torch.log(-torch.expm1(-dt0))
    A_f = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
    D_f = torch.ones(H, device=dev)
    fused = lambda zz: mamba_split_conv1d_scan_combined(zz, w, b, dtb_f, A_f, D_f, 256, activation="silu", rmsnorm_weight=nw,
                                                        rmsnorm_eps=1e-5, outproj_weight=None, headdim=P, ngroups=1,
                                                        norm_before_gate=False)
    with torch.no_grad():
        t = timeit(lambda: fused(zx), steps=10)
    out.append(dict(op="mamba_split_conv1d_scan_combined fwd, no out_proj (B=16, L=4096, d_model=2048)", ms=t * 1e3,
                    bytes=B * L * 25216, tokens_per_s=B * L / t))
    zg2 = zx.detach().requires_grad_()
    wg2, bg2, nwg2 = w.clone().requires_grad_(), b.clone().requires_grad_(), nw.clone().requires_grad_()
    dtbg, Ag, Dg = dtb_f.clone().requires_grad_(), A_f.clone().requires_grad_(), D_f.clone().requires_grad_()
    dyf = rn(B, L, d_inner)

    def fused_fb():
        yy = mamba_split_conv1d_scan_combined(zg2, wg2, bg2, dtbg, Ag, Dg, 256, activation="silu", rmsnorm_weight=nwg2,
                                              rmsnorm_eps=1e-5, outproj_weight=None, headdim=P, ngroups=1, norm_before_gate=False)
        torch.autograd.grad(yy, (zg2, wg2, bg2, dtbg, Ag, Dg, nwg2), dyf)
    t = timeit(fused_fb, steps=5, warmup=2)
    out.append(dict(op="mamba_split_conv1d_scan_combined fwd+bwd, no out_proj (same shape)", ms=t * 1e3,
                    bytes=B * L * (25216 + 8192 + 17024 + 17024), tokens_per_s=B * L / t))
    del zg2, dyf

    # f1 fused residual add + RMSNorm (block.py:86-95): x bf16 + residual fp32 -> y bf16 + residual fp32
    xh, res, w2 = rn(B * L, 2048), rn(B * L, 2048, dtype=torch.float32), torch.ones(2048, device=dev)
    t = timeit(lambda: layer_norm_fn(xh, w2, None, residual=res, prenorm=True, residual_in_fp32=True, eps=1e-5, is_rms_norm=True))
    by = (B * L) * 2048 * (2 + 4 + 2 + 4)
    out.append(dict(op="layer_norm_fn add+rmsnorm prenorm fwd (65536 rows x 2048)", ms=t * 1e3, bytes=by))
    # ... and its backward (the training loop): reads the fp32 residual stream, dy (bf16), the incoming residual gradient
    # (fp32); writes dx (bf16) and the outgoing residual gradient (fp32)
    xg2, rg2, wg2 = xh.detach().requires_grad_(), res.detach().requires_grad_(), w2.clone().requires_grad_()
    yn2, ro2 = layer_norm_fn(xg2, wg2, None, residual=rg2, prenorm=True, residual_in_fp32=True, eps=1e-5, is_rms_norm=True)
    dyn2, dro2 = torch.randn_like(yn2), torch.randn_like(ro2)
    t = timeit(lambda: torch.autograd.grad((yn2, ro2), (xg2, rg2, wg2), (dyn2, dro2), retain_graph=True))
    out.append(dict(op="layer_norm_fn add+rmsnorm prenorm bwd (65536 rows x 2048; incl. torch partial-sum reduction)", ms=t * 1e3,
                    bytes=(B * L) * 2048 * (4 + 2 + 4 + 2 + 4)))
    del yn2, ro2, dyn2, dro2, xg2, rg2

    # a8 selective_state_update, batch 64, fp32 state, stride-0 broadcast operands exactly as Mamba2.step passes them
    Bd = 64
    state = rn(Bd, H, P, N, dtype=torch.float32)
    xs, dts = rn(Bd, H, P), rn(Bd, H)[:, :, None].expand(Bd, H, P)
    A = (-torch.rand(H, device=dev, generator=g) * 15 - 1)[:, None, None].expand(H, P, N)
    Bs, Cs = rn(Bd, 1, N), rn(Bd, 1, N)
    D = torch.ones(H, device=dev)[:, None].expand(H, P)
    dtb = torch.zeros(H, device=dev)[:, None].expand(H, P)
    t = timeit(lambda: selective_state_update(state, xs, dts, A, Bs, Cs, D, z=None, dt_bias=dtb, dt_softplus=True), flush=flush)
    by = 2 * Bd * H * P * N * 4 + 2 * (2 * Bd * H * P + Bd * H + 2 * Bd * N)
    out.append(dict(op="selective_state_update (B=64, fp32 state, L2 flushed)", ms=t * 1e3, bytes=by))
    state16 = state.bfloat16()
    t = timeit(lambda: selective_state_update(state16, xs, dts, A, Bs, Cs, D, z=None, dt_bias=dtb, dt_softplus=True), flush=flush)
    by = 2 * Bd * H * P * N * 2 + 2 * (2 * Bd * H * P + Bd * H + 2 * Bd * N)
    out.append(dict(op="selective_state_update (B=64, bf16 state, L2 flushed)", ms=t * 1e3, bytes=by))

    # a7 causal_conv1d_update, batch 64
    cs, xu = rn(Bd, conv_dim, 4), rn(Bd, conv_dim)
    t = timeit(lambda: causal_conv1d_update(xu, cs, w, b, "silu"), flush=flush)
    by = Bd * conv_dim * (2 * 4 * 2 + 2 * 2)
    out.append(dict(op="causal_conv1d_update (B=64, D=4352, L2 flushed)", ms=t * 1e3, bytes=by))

    for o in out:
        o["achieved_gbs"] = o["bytes"] / (o["ms"] / 1e3) / 1e9
        o["peak_gbs"] = hbm
        o["frac"] = o["achieved_gbs"] / hbm
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
