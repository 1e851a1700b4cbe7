"""GPU parity: every C-ABI entry point of libomnissm.so (called through the reference-facing Python surface)
against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 <= 1e-5 relative, bf16 <= 1e-3 relative, both as the norm-wise
error ||y - y_oracle||_2 / ||y_oracle||_2 with the oracle evaluated in fp32 (fp64 where noted) on the SAME
(already rounded) inputs.  Gradients: fp32 <= 1e-4 (sums over B*L terms), bf16 <= 1e-2.
"""
import math

import numpy as np
import pytest
import torch

import oracle
from cases import block_input, block_params, scan_inputs, BLOCK_CASES

pytestmark = pytest.mark.gpu

DEV = "cuda"
TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-3, torch.float16: 1e-3}
GTOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2, torch.float16: 1e-2}


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check(a, b, tol, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    e = rel_l2(a, b)
    assert e <= tol, f"{what}: rel_l2 {e:.3e} > {tol:.1e}"


@pytest.fixture(scope="module")
def ops():
    import omnimamba_b200.interface as I
    from omnimamba_b200 import _cabi
    _cabi.lib()
    return I


# ------------------------------------------------------------------------------------------------------------
# causal conv1d
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channel_last", [True, False])
@pytest.mark.parametrize("B,D,L,W", [(2, 768, 128, 4), (1, 4352, 329, 4), (3, 24, 37, 3), (2, 64, 1, 2), (2, 40, 5, 4)])
@pytest.mark.parametrize("act", [None, "silu"])
def test_conv1d_fwd(ops, dtype, channel_last, B, D, L, W, act):
    g = torch.Generator().manual_seed(B * 1000 + D + L)
    x = torch.randn(B, L, D, generator=g).to(dtype)
    x = x.transpose(1, 2) if channel_last else x.transpose(1, 2).contiguous()
    w = torch.randn(D, W, generator=g) / 2
    b = torch.randn(D, generator=g)
    ref = oracle.causal_conv1d_ref(x, w, b, activation=act, compute_dtype=torch.float32)
    out = ops.causal_conv1d_fn(x.to(DEV), w.to(DEV), b.to(DEV), activation=act)
    check(out, ref, TOL[dtype], "conv out")
    assert out.stride() == x.stride() or not channel_last


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("L", [1, 7, 64, 65, 329, 1024])
def test_conv1d_fwd_fast_path_zxbcdt_slice(ops, dtype, L):
    """The call Mamba2.forward makes at d_model=2048: xBC = zxbcdt[..., 4096:8448].transpose(1, 2) (row pitch 8512), width 4,
    SiLU - the 4-channel-per-thread fast kernel, segment boundaries (64 tokens per thread) and ragged tails included."""
    g = torch.Generator().manual_seed(L)
    zx = torch.randn(2, L, 8512, generator=g).to(dtype)
    x = zx[..., 4096:4096 + 4352].transpose(1, 2)
    w, b = torch.randn(4352, 4, generator=g) / 2, torch.randn(4352, generator=g)
    ref = oracle.causal_conv1d_ref(x, w, b, activation="silu", compute_dtype=torch.float32)
    xd = zx.to(DEV)[..., 4096:4096 + 4352].transpose(1, 2)
    out = ops.causal_conv1d_fn(xd, w.to(DEV), b.to(DEV), activation="silu")
    check(out, ref, TOL[dtype], "conv fast out")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv1d_states(ops, dtype):
    g = torch.Generator().manual_seed(5)
    B, D, L, W = 2, 256, 50, 4
    x = torch.randn(B, L, D, generator=g).to(dtype).transpose(1, 2)
    init = torch.randn(B, W - 1, D, generator=g).to(dtype).transpose(1, 2)
    w = torch.randn(D, W, generator=g) / 2
    ref, fin = oracle.causal_conv1d_ref(x, w, None, initial_states=init, return_final_states=True, activation="silu",
                                        compute_dtype=torch.float32)
    out, fo = ops.causal_conv1d_fn(x.to(DEV), w.to(DEV), None, initial_states=init.to(DEV), return_final_states=True,
                                   activation="silu")
    check(out, ref, TOL[dtype], "conv out (initial_states)")
    assert torch.equal(fo.cpu(), fin)
    # L < W-1: the final state still holds zeros / initial states on the left
    xs = x[:, :, :2]
    ref, fin = oracle.causal_conv1d_ref(xs, w, None, return_final_states=True, compute_dtype=torch.float32)
    out, fo = ops.causal_conv1d_fn(xs.to(DEV), w.to(DEV), None, return_final_states=True)
    check(out, ref, TOL[dtype], "conv out short")
    assert torch.equal(fo.cpu(), fin)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channel_last", [True, False])
@pytest.mark.parametrize("act", [None, "silu"])
def test_conv1d_bwd(ops, dtype, channel_last, act):
    g = torch.Generator().manual_seed(11)
    B, D, L, W = 2, 192, 150, 4
    x = torch.randn(B, L, D, generator=g).to(dtype)
    x = x.transpose(1, 2) if channel_last else x.transpose(1, 2).contiguous()
    w = torch.randn(D, W, generator=g) / 2
    b = torch.randn(D, generator=g)
    dy = torch.randn(B, L, D, generator=g).to(dtype).transpose(1, 2)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    oracle.causal_conv1d_ref(xr, wr, br, activation=act, compute_dtype=torch.float32).backward(dy)
    xg, wg, bg = (t.to(DEV).detach().requires_grad_() for t in (x, w, b))
    ops.causal_conv1d_fn(xg, wg, bg, activation=act).backward(dy.to(DEV))
    check(xg.grad, xr.grad, TOL[dtype] * 4, "conv dx")
    check(wg.grad, wr.grad, GTOL[dtype], "conv dweight")
    check(bg.grad, br.grad, GTOL[dtype], "conv dbias")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,L", [(2, 1), (2, 5), (2, 64), (2, 67), (1, 329), (2, 1030), (3, 9300)])
@pytest.mark.parametrize("act", [None, "silu"])
def test_conv1d_bwd_fast_path_zxbcdt_slice(ops, dtype, B, L, act):
    """Backward of the call Mamba2.forward makes at d_model=2048 (channel-last slices with row pitch 8512, width 4): the
    4-channel-per-thread fast kernel, with 64 tokens per thread and (B=3, L=9300) 256 tokens per thread, ragged tails."""
    if L > 2000 and (act is None or dtype == torch.float16):
        pytest.skip("large case once")
    g = torch.Generator().manual_seed(L + B)
    D = 4352
    zx = torch.randn(B, L, 8512, generator=g).to(dtype)
    dzx = torch.randn(B, L, 8512, generator=g).to(dtype)
    x, dy = zx[..., 4096:4096 + D].transpose(1, 2), dzx[..., 4096:4096 + D].transpose(1, 2)
    w, b = torch.randn(D, 4, generator=g) / 2, torch.randn(D, generator=g)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    oracle.causal_conv1d_ref(xr, wr, br, activation=act, compute_dtype=torch.float32).backward(dy)
    zxd = zx.to(DEV).requires_grad_()
    wg, bg = w.to(DEV).requires_grad_(), b.to(DEV).requires_grad_()
    out = ops.causal_conv1d_fn(zxd[..., 4096:4096 + D].transpose(1, 2), wg, bg, activation=act)
    out.backward(dzx.to(DEV)[..., 4096:4096 + D].transpose(1, 2))
    check(zxd.grad[..., 4096:4096 + D].transpose(1, 2), xr.grad, TOL[dtype] * 4, "conv fast dx")
    check(wg.grad, wr.grad, GTOL[dtype], "conv fast dweight")
    check(bg.grad, br.grad, GTOL[dtype], "conv fast dbias")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T", [None, 3])
def test_conv1d_update(ops, dtype, T):
    g = torch.Generator().manual_seed(3)
    B, D, W, S = 4, 4352, 4, 4
    state = torch.randn(B, S, D, generator=g).to(dtype).transpose(1, 2)  # Mamba2's (B, D, S) channel-contiguous cache
    x = torch.randn(B, D, generator=g).to(dtype) if T is None else torch.randn(B, D, T, generator=g).to(dtype)
    w = torch.randn(D, W, generator=g) / 2
    b = torch.randn(D, generator=g)
    st_ref = state.clone()
    ref = oracle.causal_conv1d_update_ref(x, st_ref, w, b, "silu")
    st = state.to(DEV)
    out = ops.causal_conv1d_update(x.to(DEV), st, w.to(DEV), b.to(DEV), "silu")
    check(out, ref, TOL[dtype], "conv update out")
    assert torch.equal(st.cpu(), st_ref)


def test_conv1d_update_equals_prefill(ops):
    """decode steps after a prefill reproduce the full-sequence conv (the contract generation.py relies on)."""
    g = torch.Generator().manual_seed(9)
    B, D, L, W = 2, 128, 20, 4
    x = torch.randn(B, L, D, generator=g).transpose(1, 2).to(DEV)
    w, b = (torch.randn(D, W, generator=g) / 2).to(DEV), torch.randn(D, generator=g).to(DEV)
    full = ops.causal_conv1d_fn(x, w, b, activation="silu")
    state = torch.zeros(B, W, D, device=DEV).transpose(1, 2)
    outs = [ops.causal_conv1d_update(x[:, :, t].contiguous(), state, w, b, "silu") for t in range(L)]
    check(torch.stack(outs, -1), full, 1e-6, "stepwise conv")


# ------------------------------------------------------------------------------------------------------------
# SSD scan (mamba_chunk_scan_combined)
# ------------------------------------------------------------------------------------------------------------
def _ssd_case(batch, L, H, P, G, N, seed, dtype):
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(batch, L, H, P, G, N, seed, dtype)
    return x, dt, A, Bm, Cm, D, dt_bias


SSD_SHAPES = [(2, 128, 8, 64, 1, 128), (1, 329, 4, 64, 1, 128), (2, 45, 3, 8, 1, 16), (1, 70, 4, 32, 2, 64), (2, 1, 2, 64, 1, 128)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", SSD_SHAPES)
@pytest.mark.parametrize("variant", ["plain", "z", "init", "nodt", "limit"])
@pytest.mark.parametrize("algo", ["recurrent", "auto"])
def test_ssd_fwd(ops, dtype, shape, variant, algo):
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    batch, L, H, P, G, N = shape
    x, dt, A, Bm, Cm, D, dt_bias = _ssd_case(batch, L, H, P, G, N, 7, dtype)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(batch, L, H, P, generator=g).to(dtype) if variant == "z" else None
    init = torch.randn(batch, H, P, N, generator=g) if variant == "init" else None
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True)
    if variant == "nodt":
        kw = dict(D=None, dt_bias=None, dt_softplus=False)
        dt = (dt.float().abs() * 0.1).to(dtype)
    if variant == "limit":
        kw["dt_limit"] = (0.01, 0.05)
    ref, fin_ref = oracle.mamba_chunk_scan_combined_ref(x, dt, A, Bm, Cm, 256, z=z, initial_states=init,
                                                        return_final_states=True, **kw)
    c = lambda t: None if t is None else t.to(DEV)
    out, fin = ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(kw.get("D")), z=c(z), dt_bias=c(kw.get("dt_bias")),
                           initial_states=c(init), dt_softplus=kw["dt_softplus"],
                           dt_limit=kw.get("dt_limit", (0.0, float("inf"))), return_final_states=True, algo=algo)
    if dtype == torch.bfloat16 and algo == "auto":
        # may run on the tcgen05 kernel (fp16 tensor-core operands): excess-over-output-rounding metric against the
        # UNROUNDED fp32 oracle, tests/parity_metric.py; its final states carry the fp16 rounding of X' (<= 3e-3 fp32)
        from parity_metric import excess_over_rounding
        f = lambda t: None if t is None else t.float()
        ref32 = oracle.mamba_chunk_scan_combined_ref(f(x), dt, A, f(Bm), f(Cm), 256, z=f(z), initial_states=init, **kw)
        assert ref32.dtype == torch.float32
        e = excess_over_rounding(out, ref32)
        assert e <= TOL[dtype], f"ssd out {variant}: excess over bf16 rounding {e:.3e}"
        check(out, ref, 2e-3, f"ssd out {variant} (vs the bf16-rounded oracle: amplified metric, reported bound)")
        check(fin, fin_ref, 3e-3, f"ssd final_states {variant}")
    else:
        check(out, ref, TOL[dtype], f"ssd out {variant}")
        check(fin, fin_ref, TOL[dtype], f"ssd final_states {variant}")


def test_ssd_fwd_golden(ops):
    """The committed dense-formula golden vectors (tests/golden/ops_lib.npz)."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ops_lib.npz"))
    t = lambda k: torch.from_numpy(d[k]).float().to(DEV)
    out, fin = ops.mamba_chunk_scan_combined(t("ssd_x"), t("ssd_dt"), t("ssd_A"), t("ssd_B"), t("ssd_C"), 32, D=t("ssd_D"),
                                             dt_bias=t("ssd_dt_bias"), dt_softplus=True, return_final_states=True)
    check(out, torch.from_numpy(d["ssd_y"]), 1e-5, "golden ssd y")
    check(fin, torch.from_numpy(d["ssd_final"]), 1e-5, "golden ssd final")


def test_ssd_seq_idx(ops):
    batch, L, H, P, G, N = 2, 90, 4, 64, 1, 128
    x, dt, A, Bm, Cm, D, dt_bias = _ssd_case(batch, L, H, P, G, N, 3, torch.float32)
    seq_idx = torch.zeros(batch, L, dtype=torch.int32)
    seq_idx[0, 30:] = 1
    seq_idx[0, 71:] = 2
    seq_idx[1, 45:] = 1
    ref = oracle.mamba_chunk_scan_combined_ref(x, dt, A, Bm, Cm, 64, D=D, dt_bias=dt_bias, seq_idx=seq_idx, dt_softplus=True)
    out = ops.mamba_chunk_scan_combined(*(t.to(DEV) for t in (x, dt, A, Bm, Cm)), 64, D=D.to(DEV), dt_bias=dt_bias.to(DEV),
                                        seq_idx=seq_idx.to(DEV), dt_softplus=True)
    check(out, ref, 1e-5, "ssd seq_idx")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant", ["plain", "z", "init_final"])
def test_ssd_bwd(ops, dtype, variant):
    batch, L, H, P, G, N = 2, 100, 4, 64, 1, 128
    x, dt, A, Bm, Cm, D, dt_bias = _ssd_case(batch, L, H, P, G, N, 21, dtype)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(batch, L, H, P, generator=g).to(dtype) if variant == "z" else None
    init = torch.randn(batch, H, P, N, generator=g) if variant == "init_final" else None
    dy = torch.randn(batch, L, H, P, generator=g).to(dtype)
    dfin = torch.randn(batch, H, P, N, generator=g) if variant == "init_final" else None
    names = ["x", "dt", "A", "B", "C", "D", "dt_bias"] + (["z"] if z is not None else []) + (["init"] if init is not None else [])

    def run(fn, dev, cd):
        ts = dict(x=x, dt=dt, A=A, B=Bm, C=Cm, D=D, dt_bias=dt_bias, z=z, init=init)
        ts = {k: (None if v is None else v.to(dev).detach().clone().requires_grad_()) for k, v in ts.items()}
        kw = dict(D=ts["D"], z=ts["z"], dt_bias=ts["dt_bias"], initial_states=ts["init"], dt_softplus=True,
                  return_final_states=dfin is not None)
        if cd is not None:
            kw["compute_dtype"] = cd
        r = fn(ts["x"], ts["dt"], ts["A"], ts["B"], ts["C"], 64, **kw)
        if dfin is not None:
            (r[0].float() * dy.to(dev).float()).sum().add((r[1] * dfin.to(dev)).sum()).backward()
        else:
            r.backward(dy.to(dev))
        return {k: ts[k].grad for k in names}

    gref = run(oracle.mamba_chunk_scan_combined_ref, "cpu", torch.float64 if dtype == torch.float32 else torch.float32)
    gout = run(ops.mamba_chunk_scan_combined, DEV, None)
    for k in names:
        check(gout[k], gref[k], GTOL[dtype], f"ssd d{k} ({variant})")


# ------------------------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,Dm,gs", [(256, 512, 512), (329, 4096, 4096), (64, 512, 128), (7, 96, 32)])
@pytest.mark.parametrize("nbg", [False, True])
@pytest.mark.parametrize("gated", [True, False])
def test_norm_gated(ops, dtype, M, Dm, gs, nbg, gated):
    g = torch.Generator().manual_seed(M + Dm)
    x = torch.randn(M, Dm, generator=g).to(dtype)
    z = torch.randn(M, Dm, generator=g).to(dtype) if gated else None
    w = torch.rand(Dm, generator=g) + 0.5
    dy = torch.randn(M, Dm, generator=g).to(dtype)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    zr = z.clone().requires_grad_() if gated else None
    ref = oracle.rmsnorm_gated_ref(xr, wr, None, z=zr, eps=1e-5, group_size=gs, norm_before_gate=nbg)
    ref.backward(dy)
    xg, wg = x.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    zg = z.to(DEV).requires_grad_() if gated else None
    out = ops.rmsnorm_fn(xg, wg, None, z=zg, eps=1e-5, group_size=gs, norm_before_gate=nbg)
    out.backward(dy.to(DEV))
    check(out, ref, TOL[dtype], "gated norm out")
    check(xg.grad, xr.grad, TOL[dtype] * 10, "gated norm dx")
    check(wg.grad, wr.grad, GTOL[dtype], "gated norm dw")
    if gated:
        check(zg.grad, zr.grad, TOL[dtype] * 10, "gated norm dz")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,Dm", [(1, 2048), (333, 4096), (1500, 4096), (40, 8192)])
def test_norm_gated_fast_path(ops, dtype, M, Dm):
    """rmsnorm(x * silu(z)) * w with one group of 2048 / 4096 / 8192 columns in a 16-bit type: the persistent fast kernel
    (more rows than resident CTAs at M=1500, fewer at M=1), strided x rows (y is a (B*L, H*P) view of the scan output)."""
    g = torch.Generator().manual_seed(M + Dm)
    xfull = torch.randn(M, Dm + 64, generator=g).to(dtype)
    x, z = xfull[:, :Dm], torch.randn(M, Dm, generator=g).to(dtype)
    w = torch.rand(Dm, generator=g) + 0.5
    dy = torch.randn(M, Dm, generator=g).to(dtype)
    xr, zr, wr = x.clone().requires_grad_(), z.clone().requires_grad_(), w.clone().requires_grad_()
    ref = oracle.rmsnorm_gated_ref(xr, wr, None, z=zr, eps=1e-5, group_size=Dm, norm_before_gate=False)
    ref.backward(dy)
    xf, zg, wg = xfull.to(DEV).requires_grad_(), z.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    out = ops.rmsnorm_fn(xf[:, :Dm], wg, None, z=zg, eps=1e-5, group_size=Dm, norm_before_gate=False)
    out.backward(dy.to(DEV))
    check(out, ref, TOL[dtype], "gated norm fast out")
    check(xf.grad[:, :Dm], xr.grad, TOL[dtype] * 10, "gated norm fast dx")
    check(zg.grad, zr.grad, TOL[dtype] * 10, "gated norm fast dz")
    check(wg.grad, wr.grad, GTOL[dtype], "gated norm fast dw")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("has_res", [True, False])
@pytest.mark.parametrize("rms", [True, False])
def test_add_norm(ops, dtype, has_res, rms):
    g = torch.Generator().manual_seed(4)
    M, Dm = 300, 2048
    x = torch.randn(2, M // 2, Dm, generator=g).to(dtype)
    res = torch.randn(2, M // 2, Dm, generator=g) if has_res else None  # fp32 residual stream
    w = torch.rand(Dm, generator=g) + 0.5
    b = None if rms else torch.randn(Dm, generator=g)
    dy = torch.randn(2, M // 2, Dm, generator=g).to(dtype)
    dres = torch.randn(2, M // 2, Dm, generator=g)
    leaves = lambda dev: [t.to(dev).detach().clone().requires_grad_() if t is not None else None for t in (x, res, w, b)]
    xr, rr, wr, br = leaves("cpu")
    y, ro = oracle.layer_norm_ref(xr, wr, br, residual=rr, eps=1e-5, prenorm=True, residual_in_fp32=True, is_rms_norm=rms)
    (y.float() * dy.float()).sum().add((ro * dres).sum()).backward()
    xg, rg, wg, bg = leaves(DEV)
    y2, ro2 = ops.layer_norm_fn(xg, wg, bg, residual=rg, eps=1e-5, prenorm=True, residual_in_fp32=True, is_rms_norm=rms)
    assert ro2.dtype == torch.float32
    (y2.float() * dy.to(DEV).float()).sum().add((ro2 * dres.to(DEV)).sum()).backward()
    check(y2, y, TOL[dtype], "add_norm y")
    check(ro2, ro, 1e-6, "add_norm residual")
    check(xg.grad, xr.grad, TOL[dtype] * 10, "add_norm dx")
    check(wg.grad, wr.grad, GTOL[dtype], "add_norm dw")
    if has_res:
        check(rg.grad, rr.grad, 1e-5, "add_norm dresidual")
    if b is not None:
        check(bg.grad, br.grad, GTOL[dtype], "add_norm db")


# ------------------------------------------------------------------------------------------------------------
# single-token state update
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("xdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("tied", [True, False])
@pytest.mark.parametrize("B,H,P,N,G", [(64, 64, 64, 128, 1), (3, 8, 64, 128, 1), (2, 4, 32, 64, 2), (2, 5, 6, 16, 1)])
def test_selective_state_update(ops, sdtype, xdtype, tied, B, H, P, N, G):
    g = torch.Generator().manual_seed(B + H)
    state = torch.randn(B, H, P, N, generator=g).to(sdtype)
    x = torch.randn(B, H, P, generator=g).to(xdtype)
    z = torch.randn(B, H, P, generator=g).to(xdtype)
    Bm, Cm = torch.randn(B, G, N, generator=g).to(xdtype), torch.randn(B, G, N, generator=g).to(xdtype)
    if tied:
        dt = torch.randn(B, H, generator=g).to(xdtype).unsqueeze(-1).expand(B, H, P)
        A = (-torch.rand(H, generator=g) * 15 - 1).view(H, 1, 1).expand(H, P, N)
        D = torch.rand(H, generator=g).view(H, 1).expand(H, P)
        dt_bias = (torch.rand(H, generator=g) - 4).view(H, 1).expand(H, P)
    else:
        dt = torch.randn(B, H, P, generator=g).to(xdtype)
        A = -torch.rand(H, P, N, generator=g) * 15 - 1
        D = torch.rand(H, P, generator=g)
        dt_bias = torch.rand(H, P, generator=g) - 4
    st_ref = state.clone()
    ref = oracle.selective_state_update_ref(st_ref, x, dt, A, Bm, Cm, D=D, z=z, dt_bias=dt_bias, dt_softplus=True)
    st = state.to(DEV)
    c = lambda t: t.to(DEV)  # .to keeps stride-0 expands
    out = ops.selective_state_update(st, c(x), c(dt), c(A), c(Bm), c(Cm), D=c(D), z=c(z), dt_bias=c(dt_bias), dt_softplus=True)
    check(out, ref, TOL[xdtype], "ssu out")
    check(st, st_ref, TOL[sdtype], "ssu state")


def test_selective_state_update_no_heads(ops):
    g = torch.Generator().manual_seed(8)
    B, Dm, N = 3, 96, 16
    state = torch.randn(B, Dm, N, generator=g)
    x, dt = torch.randn(B, Dm, generator=g), torch.randn(B, Dm, generator=g)
    A = -torch.rand(Dm, N, generator=g) - 0.5
    Bm, Cm = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    st_ref = state.clone()
    ref = oracle.selective_state_update_ref(st_ref, x, dt, A, Bm, Cm, dt_softplus=True)
    st = state.to(DEV)
    out = ops.selective_state_update(st, *(t.to(DEV) for t in (x, dt, A, Bm, Cm)), dt_softplus=True)
    check(out, ref, 1e-5, "ssu (dim form) out")
    check(st, st_ref, 1e-5, "ssu (dim form) state")


def test_state_update_under_cuda_graph(ops):
    """Path C must be capturable exactly as generation.py:383-424 captures it (no sync, no allocation surprises)."""
    g = torch.Generator().manual_seed(12)
    B, H, P, N = 4, 8, 64, 128
    state0 = torch.randn(B, H, P, N, generator=g)
    xs = torch.randn(5, B, H, P, generator=g)
    dt = torch.randn(B, H, generator=g).unsqueeze(-1).expand(B, H, P)
    A = (-torch.rand(H, generator=g) - 0.5).view(H, 1, 1).expand(H, P, N)
    Bm, Cm = torch.randn(B, 1, N, generator=g), torch.randn(B, 1, N, generator=g)
    st_ref = state0.clone()
    refs = [oracle.selective_state_update_ref(st_ref, xs[i], dt, A, Bm, Cm, dt_softplus=True) for i in range(5)]
    st = state0.to(DEV)
    x_static = torch.empty(B, H, P, device=DEV)
    dtd, Ad, Bd, Cd = dt.to(DEV), A.to(DEV), Bm.to(DEV), Cm.to(DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        x_static.copy_(xs[0])
        warm = st.clone()
        ops.selective_state_update(warm, x_static, dtd, Ad, Bd, Cd, dt_softplus=True)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_static = ops.selective_state_update(st, x_static, dtd, Ad, Bd, Cd, dt_softplus=True)
    st.copy_(state0)  # capture does not execute; start from the initial state
    for i in range(5):
        x_static.copy_(xs[i])
        graph.replay()
        check(y_static, refs[i], 1e-5, f"graph step {i}")
    check(st, st_ref, 1e-5, "graph state")


# ------------------------------------------------------------------------------------------------------------
# Mamba-1 selective scan
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Dm,L,N,G", [(2, 128, 100, 16, 1), (1, 96, 33, 16, 2), (2, 64, 257, 8, 1)])
def test_selective_scan_fwd(ops, dtype, B, Dm, L, N, G):
    g = torch.Generator().manual_seed(L)
    u, delta, z = (torch.randn(B, Dm, L, generator=g).to(dtype) for _ in range(3))
    A = -torch.rand(Dm, N, generator=g) * 4 - 0.5
    Bm, Cm = torch.randn(B, G, N, L, generator=g).to(dtype), torch.randn(B, G, N, L, generator=g).to(dtype)
    D, db = torch.rand(Dm, generator=g), torch.rand(Dm, generator=g) - 3
    ref, last_ref = oracle.selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=db, delta_softplus=True,
                                              return_last_state=True)
    out, last = ops.selective_scan_fn(*(t.to(DEV) for t in (u, delta, A, Bm, Cm, D)), z=z.to(DEV), delta_bias=db.to(DEV),
                                      delta_softplus=True, return_last_state=True)
    check(out, ref, TOL[dtype], "selective_scan out")
    check(last, last_ref, TOL[dtype], "selective_scan last_state")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Dm,L,N,G,with_z", [(2, 128, 100, 16, 1, True), (1, 96, 33, 16, 2, True), (2, 64, 257, 8, 1, False),
                                               (1, 80, 70, 32, 1, True), (1, 64, 21, 64, 1, True)])
def test_selective_scan_bwd(ops, dtype, B, Dm, L, N, G, with_z):
    """omni_selective_scan_bwd (two-pass checkpointed reverse recurrence) against the oracle's autograd; every tile
    length (16 / 8 / 4 tokens for d_state <= 16 / 32 / 64), ragged tails, grouped B/C, channel counts that are not a multiple of 64."""
    g = torch.Generator().manual_seed(L + N)
    u, delta = (torch.randn(B, Dm, L, generator=g).to(dtype) for _ in range(2))
    z = torch.randn(B, Dm, L, generator=g).to(dtype) if with_z else None
    A = -torch.rand(Dm, N, generator=g) * 4 - 0.5
    Bm, Cm = torch.randn(B, G, N, L, generator=g).to(dtype), torch.randn(B, G, N, L, generator=g).to(dtype)
    D, db = torch.rand(Dm, generator=g), torch.rand(Dm, generator=g) - 3
    dout = torch.randn(B, Dm, L, generator=g).to(dtype)
    leaves = [t.clone().float().requires_grad_() if t is not None else None for t in (u, delta, A, Bm, Cm, D, z, db)]
    ref = oracle.selective_scan_ref(*leaves[:6], z=leaves[6], delta_bias=leaves[7], delta_softplus=True)
    ref.backward(dout.float())
    dl = [t.to(DEV).requires_grad_() if t is not None else None for t in (u, delta, A, Bm, Cm, D, z, db)]
    out = ops.selective_scan_fn(*dl[:6], z=dl[6], delta_bias=dl[7], delta_softplus=True)
    out.backward(dout.to(DEV))
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias"]
    for n, a, b in zip(names, dl, leaves):
        if a is None:
            continue
        assert a.grad is not None and a.grad.dtype == a.dtype, n
        check(a.grad, b.grad, GTOL[dtype], "selective_scan " + n)


def test_mamba1_block_trains(ops):
    """Mamba (v1) through mamba_inner_fn (use_fast_path, the default create_block gives it when ssm_cfg has no "layer":
    /root/reference/models/stage2/mixer_seq_simple.py:197): forward and every parameter gradient against the same module
    evaluated with the oracle's selective_scan_ref / conv reference on the CPU."""
    import torch.nn.functional as F
    from omnimamba_b200.modules import Mamba
    torch.manual_seed(0)
    m = Mamba(64, d_state=16, layer_idx=0)
    u = torch.randn(2, 50, 64)

    def cpu_forward(m, u):
        batch, L, _ = u.shape
        xz = m.in_proj(u).transpose(1, 2)
        x, z = xz.chunk(2, dim=1)
        x = oracle.causal_conv1d_ref(x, m.conv1d.weight.squeeze(1), m.conv1d.bias, activation="silu")
        x_dbl = m.x_proj(x.transpose(1, 2).reshape(batch * L, m.d_inner))
        dt, Bm, Cm = torch.split(x_dbl, [m.dt_rank, m.d_state, m.d_state], dim=-1)
        dt = F.linear(dt, m.dt_proj.weight).view(batch, L, m.d_inner).transpose(1, 2)
        Bm = Bm.reshape(batch, L, m.d_state).transpose(1, 2)
        Cm = Cm.reshape(batch, L, m.d_state).transpose(1, 2)
        y = oracle.selective_scan_ref(x, dt, -torch.exp(m.A_log.float()), Bm, Cm, m.D.float(), z=z,
                                      delta_bias=m.dt_proj.bias.float(), delta_softplus=True)
        return m.out_proj(y.transpose(1, 2))

    yr = cpu_forward(m, u)
    yr.square().sum().backward()
    gref = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    md = Mamba(64, d_state=16, layer_idx=0).to(DEV)
    md.load_state_dict(m.state_dict())
    y = md(u.to(DEV))
    y.square().sum().backward()
    check(y, yr, 2e-5, "mamba1 out")
    for k, p in md.named_parameters():
        check(p.grad, gref[k], 2e-4, "mamba1 d" + k)


# ------------------------------------------------------------------------------------------------------------
# the block: fused path A, and paths B / C through the Mamba2 module (drop-in surface)
# ------------------------------------------------------------------------------------------------------------
def _load_block(name, dtype=torch.float32):
    from omnimamba_b200.modules import Mamba2
    d_model, batch, seqlen, chunk, seed = BLOCK_CASES[name]
    m = Mamba2(d_model, chunk_size=chunk, layer_idx=0, device=DEV, dtype=torch.float32)
    m.load_state_dict({k: v.to(DEV) for k, v in block_params(d_model, seed).items()})
    u = block_input(d_model, batch, seqlen, seed).to(DEV)
    return m, u


@pytest.mark.parametrize("name", list(BLOCK_CASES))
def test_block_matches_hf_golden(name):
    """Path A (fused op) against the committed outputs of transformers' Mamba2Mixer.torch_forward (BASELINE config 1)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "mamba2_block_hf.npz"))
    m, u = _load_block(name)
    out = m(u)
    check(out, torch.from_numpy(gold[name]), 2e-5, f"block {name} vs HF golden")


def test_block_paths_agree_and_match_oracle():
    from omnimamba_b200.modules import Mamba2

    class IP:
        def __init__(self):
            self.seqlen_offset, self.key_value_memory_dict = 0, {}

    d_model, B, L, extra = 256, 2, 72, 6
    m = Mamba2(d_model, layer_idx=0, device=DEV)
    p = oracle.mamba2_init_params(d_model, seed=5)
    sd = {"in_proj.weight": p.in_proj_weight, "conv1d.weight": p.conv1d_weight, "conv1d.bias": p.conv1d_bias,
          "dt_bias": p.dt_bias, "A_log": p.A_log, "D": p.D, "norm.weight": p.norm_weight, "out_proj.weight": p.out_proj_weight}
    m.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    u = torch.randn(B, L + extra, d_model, generator=torch.Generator().manual_seed(0))
    full_ref = oracle.mamba2_forward_ref(p, u)
    full = m(u.to(DEV))                                             # path A
    check(full, full_ref, 2e-5, "path A vs oracle")
    ip = IP()
    pre = m(u[:, :L].to(DEV), inference_params=ip)                  # path B (prefill fills the caches)
    check(pre, full_ref[:, :L], 2e-5, "path B vs oracle")
    conv_ref = torch.zeros(B, p.conv_dim, p.d_conv)
    ssm_ref = torch.zeros(B, p.nheads, p.headdim, p.d_state)
    oracle.mamba2_forward_ref(p, u[:, :L], conv_state=conv_ref, ssm_state=ssm_ref)
    conv_state, ssm_state = ip.key_value_memory_dict[0]
    check(conv_state, conv_ref, 1e-6, "prefill conv_state")
    check(ssm_state, ssm_ref, 1e-5, "prefill ssm_state")
    ip.seqlen_offset = L
    for t in range(L, L + extra):                                   # path C
        y = m(u[:, t:t + 1].to(DEV), inference_params=ip)
        check(y, full_ref[:, t:t + 1], 5e-5, f"path C token {t}")
        ip.seqlen_offset += 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_path_a_backward(dtype):
    """mamba_split_conv1d_scan_combined fwd+bwd against autograd through the oracle (L = 329 is the stage-1 length)."""
    import omnimamba_b200.interface as I
    d_model, B, L = 128, 2, 329 if dtype == torch.float32 else 140
    p = oracle.mamba2_init_params(d_model, seed=3)
    g = torch.Generator().manual_seed(1)
    zxbcdt = torch.randn(B, L, p.d_in_proj, generator=g).to(dtype)
    dout = torch.randn(B, L, d_model, generator=g).to(dtype)
    A = -torch.exp(p.A_log)
    params = dict(conv_w=p.conv1d_weight.squeeze(1), conv_b=p.conv1d_bias, dt_bias=p.dt_bias, A=A, D=p.D,
                  norm_w=p.norm_weight + 0.1 * torch.randn(p.d_inner, generator=g), out_w=p.out_proj_weight.to(dtype))

    def run(fn, dev, **kw):
        ts = {k: v.to(dev).detach().clone().requires_grad_() for k, v in dict(zxbcdt=zxbcdt, **params).items()}
        out = fn(ts["zxbcdt"], ts["conv_w"], ts["conv_b"], ts["dt_bias"], ts["A"], ts["D"], 64, activation="silu",
                 rmsnorm_weight=ts["norm_w"], rmsnorm_eps=1e-5, outproj_weight=ts["out_w"], headdim=p.headdim,
                 ngroups=p.ngroups, norm_before_gate=False, **kw)
        out.backward(dout.to(dev))
        return out, {k: v.grad for k, v in ts.items()}

    ref, gref = run(oracle.mamba_split_conv1d_scan_combined_ref, "cpu")
    out, gout = run(I.mamba_split_conv1d_scan_combined, DEV)
    check(out, ref, TOL[dtype] * (3 if dtype == torch.float32 else 8), "fused out")
    for k in gref:
        check(gout[k], gref[k], GTOL[dtype] * (1 if dtype == torch.float32 else 3), f"fused d{k}")
