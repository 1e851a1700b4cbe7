"""tcgen05 building blocks (omni_selftest) and the tensor-core chunked SSD kernel against the oracle."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def run_selftest(which):
    from omnimamba_b200 import _cabi
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(0)
    dt16 = torch.float16 if which & 256 else torch.bfloat16
    Cm = torch.randn(128, 128, generator=g).to(dt16)
    Bm = torch.randn(128, 128, generator=g).to(dt16)
    X = torch.randn(128, 64, generator=g).to(dt16)
    P = torch.randn(128, 128, generator=g)
    Xs = torch.randn(128, 128, generator=g)
    S = torch.randn(128, 128, generator=g)
    dev = [t.to(DEV) for t in (Cm, Bm, X, P, Xs, S)]
    D1, D3, D4 = (torch.zeros(128, 128, device=DEV) for _ in range(3))
    D2 = torch.zeros(128, 64, device=DEV)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.omni_selftest(*(ptr(t) for t in dev), ptr(D1), ptr(D2), ptr(D3), ptr(D4), which,
                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.omni_last_error()
    torch.cuda.synchronize()
    Cf, Bf, Xf = Cm.double(), Bm.double(), X.double()
    r16 = lambda t: t.to(dt16).double()
    errs = {}
    if which & 1:
        errs["D1 = C B^T (SS, K-major x K-major)"] = rel_l2(D1, Cf @ Bf.t())
    if which & 2:
        errs["D2 = r(P) X (TS, A in TMEM x MN-major)"] = rel_l2(D2, r16(P) @ Xf)
    if which & 8:
        errs["D3 = r(Xs) B (TS, MN-major with LBO)"] = rel_l2(D3, r16(Xs) @ Bf)
    if which & 64:
        errs["D3 = r(Xs) B (SS, A MN-major smem x B MN-major)"] = rel_l2(D3, r16(Xs) @ Bf)
    if which & 128:
        errs["D3 = r(Xs) B (SS, A K-major smem x B MN-major)"] = rel_l2(D3, r16(Xs) @ Bf)
    if which & 4:
        errs["D4 = C r(S)^T (SS, thread-written swizzled B)"] = rel_l2(D4, Cf @ r16(S).t())
    return errs


@pytest.mark.parametrize("which", [1, 2, 4, 8, 64, 128, 15, 257, 258, 260, 264, 320, 384, 271])
def test_umma_selftest(which):
    # each form in its own process: an illegal-instruction fault poisons the CUDA context
    code = f"import sys; sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r});" \
           f"import test_gpu_tc as t; e = t.run_selftest({which}); print(e); assert all(v < 1e-5 for v in e.values()), e"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-1500:])
    assert r.returncode == 0


# ------------------------------------------------------------------------------------------------------------
# tensor-core chunked SSD forward (algo="chunked_tc") vs the oracle; bf16 I/O, tolerance 1e-3 (north_star)
# ------------------------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _tc_case(batch, L, H, G, seed, variant):
    import oracle
    from cases import scan_inputs
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(batch, L, H, 64, G, 128, seed, torch.bfloat16)
    g = torch.Generator().manual_seed(seed + 1)
    init = torch.randn(batch, H, 64, 128, generator=g) if variant == "init" else None
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True)
    if variant == "nodt":
        kw = dict(D=None, dt_bias=None, dt_softplus=False)
        dt = (dt.float().abs() * 0.1).to(torch.bfloat16)
    if variant == "limit":
        kw["dt_limit"] = (0.01, 0.05)
    ref, fin_ref = oracle.mamba_chunk_scan_combined_ref(x.float(), dt, A, Bm.float(), Cm.float(), 256, initial_states=init,
                                                        return_final_states=True, **kw)
    assert ref.dtype == torch.float32  # unrounded oracle: see tests/parity_metric.py
    c = lambda t: None if t is None else t.to(DEV)
    out, fin = ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(kw.get("D")), dt_bias=c(kw.get("dt_bias")),
                           initial_states=c(init), dt_softplus=kw["dt_softplus"], dt_limit=kw.get("dt_limit", (0.0, float("inf"))),
                           return_final_states=True, algo="chunked_tc")
    torch.cuda.synchronize()
    from parity_metric import excess_over_rounding
    return excess_over_rounding(out, ref), rel_l2(fin, fin_ref)


@pytest.mark.parametrize("batch,L,H,G", [(1, 128, 2, 1), (2, 329, 4, 1), (1, 1024, 8, 1), (3, 72, 4, 2), (1, 1, 2, 1), (2, 257, 6, 1)])
@pytest.mark.parametrize("variant", ["plain", "init", "nodt", "limit"])
def test_ssd_tc_fwd(batch, L, H, G, variant):
    e_out, e_fin = _tc_case(batch, L, H, G, 11, variant)
    print(f"tc fwd B={batch} L={L} H={H} G={G} {variant}: out (excess over bf16 rounding) {e_out:.2e} final {e_fin:.2e}")
    assert e_out < 1e-3, e_out
    # final_states are returned in fp32 (no output rounding to hide behind): their error is the bf16 rounding of the
    # decay-scaled x operand of the state GEMM (2^-9 worst case, as in upstream's _chunk_state_fwd); y itself stays < 1e-3
    assert e_fin < 3e-3, e_fin


@pytest.mark.parametrize("batch,L,H,G,variant", [(2, 329, 4, 1, "plain"), (1, 1024, 8, 1, "init"), (3, 72, 4, 2, "limit")])
def test_ssd_tc_fwd_fp32_out(batch, L, H, G, variant):
    """The kernel's own accuracy, without the bf16 output rounding on either side: fp32 `out` (same code path, only
    the final convert differs) against the fp32 oracle evaluated on the same bf16-valued inputs.  North-star
    tolerance 1e-3; the fp16 rounding of the computed tensor-core operands (P, S16, X') puts it near 2e-4."""
    import oracle
    from cases import scan_inputs
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(batch, L, H, 64, G, 128, 5, torch.bfloat16)
    g = torch.Generator().manual_seed(6)
    init = torch.randn(batch, H, 64, 128, generator=g) if variant == "init" else None
    lim = (0.01, 0.05) if variant == "limit" else (0.0, float("inf"))
    ref = oracle.mamba_chunk_scan_combined_ref(x.float(), dt, A, Bm.float(), Cm.float(), 256, D=D, dt_bias=dt_bias,
                                               initial_states=init, dt_softplus=True, dt_limit=lim)
    assert ref.dtype == torch.float32
    c = lambda t: None if t is None else t.to(DEV)
    out32 = torch.empty(batch, L, H, 64, device=DEV, dtype=torch.float32)
    ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), initial_states=c(init), dt_softplus=True,
                dt_limit=lim, out=out32, algo="chunked_tc")
    out16, _ = ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), initial_states=c(init),
                           dt_softplus=True, dt_limit=lim, algo="chunked_tc")
    torch.cuda.synchronize()
    e = rel_l2(out32, ref)
    print(f"tc fwd fp32-out B={batch} L={L} H={H} G={G} {variant}: {e:.2e}")
    assert e < 5e-4, e
    assert torch.equal(out16.cpu(), out32.cpu().to(torch.bfloat16)), "bf16 output must be the rounded fp32 result"


def test_ssd_tc_matches_recurrent_at_bench_size():
    """BASELINE size (16 x 4096, d_model=2048 geometry): the CPU oracle is too slow here, so the exact fp32 SIMT
    recurrence (itself oracle-checked in test_gpu_parity.py) is the on-device reference; also many more items than SMs."""
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    g = torch.Generator(device=DEV).manual_seed(0)
    B, L, H, P, N = 16, 4096, 64, 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    o1, f1 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True, algo="recurrent")
    o2, f2 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True, algo="chunked_tc")
    torch.cuda.synchronize()
    e, ef = rel_l2(o2, o1), rel_l2(f2, f1)
    print(f"tc vs recurrent at 16x4096: out {e:.2e} final {ef:.2e}")
    assert e < 1e-3 and ef < 3e-3


@pytest.mark.parametrize("G", [1, 2])
def test_ssd_tc_prepass_layouts(G):
    """The fp16 pre-pass of B and C has a division-free path for rows a constant pitch apart (contiguous tensors, channel
    slices of a wider row such as the conv output) and a generic one for every other stride pattern (here: a batch stride
    that is not L rows, and B / C interleaved in one buffer).  All layouts must give the bit-identical y and final state."""
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    g = torch.Generator(device=DEV).manual_seed(3)
    B, L, H, P, N = 3, 391, 8, 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, G, N), rn(B, L, G, N)
    A = -(torch.rand(H, device=DEV, generator=g) * 4 + 0.5)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    run = lambda b, c: ssd_fwd_raw(x, dt, A, b, c, 256, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True,
                                   algo="chunked_tc")
    y0, f0 = run(Bm, Cm)
    # (a) channel slices of one wide row (B, L, 2 G N + 64): constant row pitch when G == 1 -> fast path
    wide = torch.zeros(B, L, 2 * G * N + 64, device=DEV, dtype=torch.bfloat16)
    wide[..., :G * N] = Bm.reshape(B, L, G * N)
    wide[..., G * N:2 * G * N] = Cm.reshape(B, L, G * N)
    y1, f1 = run(wide[..., :G * N].view(B, L, G, N), wide[..., G * N:2 * G * N].view(B, L, G, N))
    # (b) padded sequence: batch stride (L + 7) rows -> generic path
    padB = torch.zeros(B, L + 7, G, N, device=DEV, dtype=torch.bfloat16)
    padC = torch.zeros(B, L + 7, G, N, device=DEV, dtype=torch.bfloat16)
    padB[:, :L] = Bm
    padC[:, :L] = Cm
    y2, f2 = run(padB[:, :L], padC[:, :L])
    torch.cuda.synchronize()
    for y, f in ((y1, f1), (y2, f2)):
        assert torch.equal(y, y0) and torch.equal(f, f0)


@pytest.mark.parametrize("B,L,H", [(5, 633, 64), (3, 300, 128), (10, 129, 32)])
def test_ssd_tc_half_item_schedule(B, L, H):
    """More items than SMs with a remainder of at most half the grid: the left-over items are cut into two half sequences
    handed over through the fp32 state slots (odd chunk counts, initial and final states included).  Reference: the exact
    fp32 SIMT recurrence on the device (oracle-checked in test_gpu_parity.py)."""
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    g = torch.Generator(device=DEV).manual_seed(B * L)
    P, N = 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    init = torch.randn(B, H, P, N, device=DEV, generator=g)
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, return_final_states=True)
    o1, f1 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="recurrent", **kw)
    o2, f2 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)
    o3, f3 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)  # (flags are reset per launch)
    torch.cuda.synchronize()
    e, ef = rel_l2(o2, o1), rel_l2(f2, f1)
    print(f"half-item schedule B={B} L={L} H={H}: out {e:.2e} final {ef:.2e}")
    assert e < 1e-3 and ef < 3e-3
    assert torch.equal(o2, o3) and torch.equal(f2, f3)


def test_ssd_tc_half_item_handoff_waits_for_a_late_producer():
    """The consumer of a half-item hand-off must WAIT for the producer's flag: the producer is held back by 2 ms
    (debug knob) and the result must still be bit-identical to the undelayed run; with the schedule switched off the
    plain schedule gives the same y as well."""
    from omnimamba_b200 import _cabi
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    lib = _cabi.lib()
    g = torch.Generator(device=DEV).manual_seed(77)
    B, L, H, P, N = 5, 633, 64, 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    run = lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dt_bias, dt_softplus=True, return_final_states=True, algo="chunked_tc")
    y0, f0 = run()
    try:
        lib.omni_debug_set_handoff(2000, 3)
        y1, f1 = run()
        lib.omni_debug_set_handoff(0, 0)
        y2, f2 = run()
        torch.cuda.synchronize()
    finally:
        lib.omni_debug_set_handoff(0, 3)
    assert torch.equal(y0, y1) and torch.equal(f0, f1)
    assert torch.equal(y0, y2) and rel_l2(f2, f0) < 1e-6


@pytest.mark.parametrize("B,L,H,with_init", [(1, 8192, 8, True), (2, 4096, 4, False), (1, 16384, 64, True), (3, 3072, 2, True)])
def test_ssd_tc_piece_schedule(B, L, H, with_init):
    """Few (batch, head pair) items: every sequence is cut into k independent pieces (state sweep from a zero start, total
    decay per piece, chain over the pieces, forward per piece from its entering state).  Must agree with the plain schedule
    of the same kernel (fp16 state copies differ per chunk start, so not bit-for-bit) and with the exact fp32 SIMT recurrence;
    final states included."""
    from omnimamba_b200 import _cabi
    from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw
    lib = _cabi.lib()
    g = torch.Generator(device=DEV).manual_seed(B * L + H)
    P, N = 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    A[0] = -0.01   # a head that barely decays: the entering state matters over the whole piece
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    init = torch.randn(B, H, P, N, device=DEV, generator=g) if with_init else None
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, return_final_states=True)
    y_p, f_p = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)
    o32 = torch.empty(B, L, H, P, device=DEV, dtype=torch.float32)
    ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="chunked_tc", out=o32, **kw)
    try:
        lib.omni_debug_set_handoff(0, 1)           # piece schedule off
        y_s, f_s = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)
    finally:
        lib.omni_debug_set_handoff(0, 3)
    y_r = torch.empty(B, L, H, P, device=DEV, dtype=torch.float32)
    _, f_r = ssd_fwd_raw(x.float(), dt.float(), A, Bm.float(), Cm.float(), 256, algo="recurrent", out=y_r, **kw)
    torch.cuda.synchronize()
    e_plain, e_rec, e_fin = rel_l2(y_p, y_s), rel_l2(o32, y_r), rel_l2(f_p, f_r)
    print(f"piece schedule B={B} L={L} H={H}: vs plain schedule {e_plain:.2e}; fp32-out vs fp32 recurrence {e_rec:.2e}; final {e_fin:.2e}")
    assert e_plain < 2.5e-3      # two bf16 outputs of the same fp32 result up to the kernel's own 2e-4
    assert e_rec < 5e-4 and e_fin < 3e-3
    assert rel_l2(f_p, f_s) < 1e-3


# ------------------------------------------------------------------------------------------------------------
# tensor-core chunked SSD backward (state sweeps + per-chunk gradient kernel) vs the oracle's autograd
# ------------------------------------------------------------------------------------------------------------
def _tc_bwd_case(batch, L, H, G, seed, variant):
    import oracle
    from cases import scan_inputs
    from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw, ssd_fwd_raw
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(batch, L, H, 64, G, 128, seed, torch.bfloat16)
    g = torch.Generator().manual_seed(seed + 1)
    dy = torch.randn(batch, L, H, 64, generator=g).to(torch.bfloat16)
    init = torch.randn(batch, H, 64, 128, generator=g) if variant == "init_final" else None
    dfin = torch.randn(batch, H, 64, 128, generator=g) if variant == "init_final" else None
    lim = (0.002, 0.08) if variant == "limit" else (0.0, float("inf"))
    # oracle: fp32 autograd on the same (bf16-valued) inputs
    leaf = lambda t: None if t is None else t.float().detach().clone().requires_grad_()
    xr, dtr, Ar, Br, Cr, Dr, dbr, ir = (leaf(t) for t in (x, dt, A, Bm, Cm, D, dt_bias, init))
    res = oracle.mamba_chunk_scan_combined_ref(xr, dtr, Ar, Br, Cr, 256, D=Dr, dt_bias=dbr, initial_states=ir, dt_softplus=True,
                                               dt_limit=lim, return_final_states=dfin is not None)
    if dfin is not None:
        yr, fr = res
        (yr * dy.float()).sum().add((fr * dfin).sum()).backward()
    else:
        (res * dy.float()).sum().backward()
    c = lambda t: None if t is None else t.to(DEV)
    out, _ = ssd_fwd_raw(c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), initial_states=c(init),
                         dt_softplus=True, dt_limit=lim, algo="chunked_tc")
    dx, ddt, dA, dB, dC, dD, _, ddtb, dinit = ssd_bwd_raw(
        c(dy), c(x), c(dt), c(A), c(Bm), c(Cm), 256, D=c(D), dt_bias=c(dt_bias), initial_states=c(init), dt_softplus=True,
        dt_limit=lim, dfinal_states=c(dfin), want_dinitial=init is not None, algo="chunked_tc", out=out)
    torch.cuda.synchronize()

    def rel_sum(a, ref, terms):
        """Error of a per-head SUM over (batch, time) relative to the L2 norm of its summands: a sum of terms that each
        carry a relative error eps is off by ~eps * ||terms||_2 however much the terms cancel."""
        return ((a.double().cpu() - ref.double()).norm() / terms.double().norm().clamp_min(1e-30)).item()

    ddt_ref = dtr.grad  # (B, L, H): the summands of ddt_bias; dA_h = sum dt da has summands of the size of ddt / |A_h|
    errs = {"dx": rel_l2(dx, xr.grad), "ddt": rel_l2(ddt, ddt_ref), "dB": rel_l2(dB, Br.grad), "dC": rel_l2(dC, Cr.grad),
            "dD": rel_l2(dD, Dr.grad), "ddt_bias": rel_sum(ddtb, dbr.grad, ddt_ref),
            "dA": rel_sum(dA, Ar.grad, ddt_ref / Ar.detach().abs())}
    if init is not None:
        errs["dinitial_states"] = rel_l2(dinit, ir.grad)
    return errs


@pytest.mark.parametrize("batch,L,H,G", [(1, 128, 2, 1), (2, 329, 4, 1), (1, 1024, 8, 1), (3, 72, 4, 2), (1, 1, 2, 1), (2, 257, 6, 1)])
@pytest.mark.parametrize("variant", ["plain", "init_final", "limit"])
def test_ssd_tc_bwd(batch, L, H, G, variant):
    """Gradient tolerance for bf16 I/O: 1e-2 relative L2 (tests/test_gpu_parity.py GTOL) against fp32 autograd of the oracle;
    the per-head sums dA and ddt_bias are measured against the norm of their summands (see rel_sum)."""
    errs = _tc_bwd_case(batch, L, H, G, 23, variant)
    print(f"tc bwd B={batch} L={L} H={H} G={G} {variant}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        # dA / ddt_bias: the reverse cumulative sums inside a chunk make the fp16 operand rounding of neighbouring tokens
        # correlated, so these sums over (batch, time) get 3e-2 (upstream's own bf16 tolerance for them is rtol 3e-2 too)
        assert v < (3e-2 if k in ("dA", "ddt_bias") else 1e-2), (k, v)


@pytest.mark.parametrize("B,L,H", [(5, 633, 64), (3, 300, 128)])
def test_ssd_tc_bwd_half_item_schedule(B, L, H):
    """The state sweeps of the backward with the half-item schedule (more items than SMs, odd chunk counts, initial states,
    a gradient flowing into the final states): tensor-core backward against the exact fp32 SIMT backward on the device."""
    from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw
    g = torch.Generator(device=DEV).manual_seed(B + L)
    P, N = 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm, dy = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N), rn(B, L, H, P)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    init = torch.randn(B, H, P, N, device=DEV, generator=g)
    dfin = torch.randn(B, H, P, N, device=DEV, generator=g) * 0.1
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, dfinal_states=dfin, want_dinitial=True)
    ref = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, algo="recurrent", **kw)
    got = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)
    torch.cuda.synchronize()
    names = ["dx", "ddt", "dA", "dB", "dC", "dD", "dz", "ddt_bias", "dinit"]
    for nme, r, t in zip(names, ref, got):
        if r is None:
            continue
        e = rel_l2(t, r)
        print(f"bwd half-item schedule B={B} L={L} H={H}: {nme} {e:.2e}")
        # (the per-head sums dA / ddt_bias cancel heavily: 3e-2 as in test_ssd_tc_bwd, upstream's own bf16 tolerance)
        assert e < (3e-2 if nme in ("dA", "ddt_bias") else 1e-2), (nme, e)


@pytest.mark.parametrize("B,L,H", [(1, 8192, 8), (2, 4096, 4), (1, 16384, 64)])
def test_ssd_tc_bwd_piece_schedule(B, L, H):
    """Few long sequences: both state sweeps of the backward run on k independent pieces per sequence (store-free sweep ->
    piece decays -> chain, last piece first in the reverse sweep -> real sweep from the entering states).  Tensor-core backward
    against the exact fp32 SIMT backward on the device, with initial states and a gradient flowing into the final states."""
    from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw
    g = torch.Generator(device=DEV).manual_seed(B * L + H)
    P, N = 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm, dy = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N), rn(B, L, H, P)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    A[0] = -0.01   # a head that barely decays: the chain over the pieces carries real weight
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    init = torch.randn(B, H, P, N, device=DEV, generator=g)
    dfin = torch.randn(B, H, P, N, device=DEV, generator=g) * 0.1
    kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, dfinal_states=dfin, want_dinitial=True)
    ref = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, algo="recurrent", **kw)
    got = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, algo="chunked_tc", **kw)
    torch.cuda.synchronize()
    names = ["dx", "ddt", "dA", "dB", "dC", "dD", "dz", "ddt_bias", "dinit"]
    errs = {n: rel_l2(t, r) for n, r, t in zip(names, ref, got) if r is not None}
    print(f"bwd piece schedule B={B} L={L} H={H}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < (3e-2 if k in ("dA", "ddt_bias") else 1e-2), (k, v)


@pytest.mark.parametrize("B,L,H,with_init", [(2, 329, 4, False), (1, 1024, 8, True), (5, 633, 64, True), (3, 300, 128, False),
                                             (1, 8192, 8, True)])
def test_ssd_tc_forward_keeps_chunk_states_for_the_backward(B, L, H, with_init):
    """omnissm.h `chunk_states`: a forward that will be followed by a backward also stores the fp16 state entering every chunk
    (kernel MODE 3) and the backward given that tensor skips its forward state sweep.  The output must be bit-equal to the
    plain forward's and every gradient bit-equal to the backward that recomputes the states (same recurrence, same operands);
    shapes with the half-item schedule ((5, 633, 64), (3, 300, 128)) and one the piece schedule takes ((1, 8192, 8): the
    library then declines to fill the tensor and the caller falls back to the sweep)."""
    from omnimamba_b200 import _cabi as abi
    from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw, ssd_fwd_raw
    g = torch.Generator(device=DEV).manual_seed(B * L + H)
    P, N = 64, 128
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g).bfloat16()
    x, dt, Bm, Cm, dy = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N), rn(B, L, H, P)
    A = -(torch.rand(H, device=DEV, generator=g) * 15 + 1)
    dt_bias = torch.rand(H, device=DEV, generator=g) * 4 - 6
    D = torch.ones(H, device=DEV)
    init = torch.randn(B, H, P, N, device=DEV, generator=g) if with_init else None
    nbytes = abi.ssd_chunk_states_bytes(B, L, H, P, N)
    assert nbytes == B * ((L + 127) // 128) * H * P * N * 2
    cs = torch.full((nbytes // 2,), float("nan"), device=DEV, dtype=torch.float16)
    fkw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, return_final_states=True, algo="chunked_tc")
    out0, fin0 = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, **fkw)
    out1, fin1, kept = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, chunk_states=cs, **fkw)
    torch.cuda.synchronize()
    assert torch.equal(out0, out1) and torch.equal(fin0, fin1)
    if (B, L, H) == (1, 8192, 8):
        assert kept is None   # piece schedule: not filled, the backward runs its own sweep
        return
    assert kept is cs and not torch.isnan(cs.float()).any()
    if init is not None:   # the state entering chunk 0 is the initial state
        s0 = cs.view(B, (L + 127) // 128, H * P, N)[:, 0].float()
        assert torch.equal(s0, init.view(B, H * P, N).half().float())
    # the state entering chunk c is the final state of the first 128 c tokens (fp32, from the plain forward on the prefix)
    nch = (L + 127) // 128
    csv = cs.view(B, nch, H * P, N).float()
    for c_ in sorted({1, nch // 2, nch - 1} - {0}):
        t = 128 * c_
        _, fin_c = ssd_fwd_raw(x[:, :t].contiguous(), dt[:, :t].contiguous(), A, Bm[:, :t].contiguous(), Cm[:, :t].contiguous(), 256,
                               D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, return_final_states=True, algo="chunked_tc")
        ref_c = fin_c.view(B, H * P, N)
        dmax = (csv[:, c_] - ref_c).abs().max().item()
        assert rel_l2(csv[:, c_], ref_c) < 1e-3 and dmax < 2e-3 * ref_c.abs().max().item(), (c_, rel_l2(csv[:, c_], ref_c), dmax)
    bkw = dict(D=D, dt_bias=dt_bias, dt_softplus=True, initial_states=init, want_dinitial=init is not None, algo="chunked_tc")
    ref = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, **bkw)
    got = ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, chunk_states=cs, **bkw)
    torch.cuda.synchronize()
    names = ["dx", "ddt", "dA", "dB", "dC", "dD", "dz", "ddt_bias", "dinit"]
    for nme, r, t in zip(names, ref, got):
        if r is None:
            continue
        # Same recurrence on the same operands in both kernels.  dx / dinitial_states come out bit-equal; ddt, dA, ddt_bias, dD,
        # dB, dC pass through shared-memory / global atomics in the gradient kernel, whose order changes from run to run (two
        # runs of the SAME backward differ in a few ddt elements at 3e-4 of the maximum and in the last bits of the sums -
        # measured), so they get a tolerance: 1e-3 relative L2, 5e-3 for the heavily cancelling per-head sums
        e = rel_l2(t, r)
        print(f"kept chunk states B={B} L={L} H={H}: {nme} bit-equal {torch.equal(t, r)} rel {e:.1e}")
        if nme in ("dx", "dinit"):
            assert torch.equal(t, r), (nme, e)
        assert e < (5e-3 if nme in ("dA", "ddt_bias") else 1e-3), (nme, e)
