"""mamba_chunk_scan_combined / mamba_split_conv1d_scan_combined with the mamba_ssm==2.2.2 signatures
(mamba_ssm/ops/triton/ssd_combined.py upstream; Mamba2.forward paths A and B).

Both are autograd Functions over libomnissm.so entry points.  Everything that touches (B, L, ...)
activations is one of our CUDA kernels, the out_proj GEMM and its backward included (interface/gemm.py: the tcgen05 GEMM
of libomnissm for bf16 operands)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .. import _cabi as abi
from .causal_conv1d import conv1d_bwd_raw, conv1d_fwd_raw
from .gemm import cast_param, mm_nt
from .layernorm_gated import norm_gated_bwd_raw, norm_gated_fwd_raw

import os

_ALGO = {"auto": abi.SSD_AUTO, "recurrent": abi.SSD_RECURRENT, "chunked_tc": abi.SSD_CHUNKED_TC}
# Process-wide opt-out of the fp16-operand tensor-core scan ("Numerical range"): OMNI_SSD_ALGO=recurrent makes
# every "auto" call - mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined, Mamba2 - take the exact fp32 recurrence.
_DEFAULT_ALGO = os.environ.get("OMNI_SSD_ALGO", "auto")
assert _DEFAULT_ALGO in _ALGO, f"OMNI_SSD_ALGO must be one of {sorted(_ALGO)}"


# A forward that will be followed by a backward also keeps the fp16 state entering every chunk (omnissm.h: chunk_states; +2 H P N
# bytes per 128 tokens and sequence): the backward then skips its forward state sweep.  OMNI_SSD_SAVE_STATES=0 restores upstream's
# recompute-everything behaviour (less memory, one more sweep).
_SAVE_STATES = os.environ.get("OMNI_SSD_SAVE_STATES", "1") != "0"
# The fused path-A node keeps its conv output for the backward as well instead of recomputing it there as upstream does
# (2 (d_inner + 2 G N) bytes per token and layer: 12 GB over the 48 layers of the stage-1 step, on a 180 GB part);
# OMNI_KEEP_CONV_OUT=0 restores the recompute.
_KEEP_CONV = os.environ.get("OMNI_KEEP_CONV_OUT", "1") != "0"


def _alloc_chunk_states(batch, seqlen, nheads, headdim, dstate, device, dtype):
    if not _SAVE_STATES or dtype != torch.bfloat16 or _DEFAULT_ALGO == "recurrent":
        return None
    n = abi.ssd_chunk_states_bytes(batch, seqlen, nheads, headdim, dstate)
    return torch.empty(n // 2, device=device, dtype=torch.float16) if n > 0 else None


def _last_contig(t):
    return t if t is None or t.stride(-1) == 1 else t.contiguous()


def ssd_fwd_raw(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None, seq_idx=None,
                dt_softplus=False, dt_limit=(0.0, float("inf")), return_final_states=False, out=None, algo="auto",
                chunk_states=None):
    """x: (B, L, H, P); dt: (B, L, H); A: (H); B, C: (B, L, G, N).  Returns (out, final_states | None); with `chunk_states`
    (a tensor of _alloc_chunk_states) a third value: that tensor if the forward filled it, else None."""
    batch, seqlen, nheads, headdim = x.shape
    dstate = B.shape[-1]
    if out is None:
        out = torch.empty(batch, seqlen, nheads, headdim, device=x.device, dtype=x.dtype)
    fin = torch.empty(batch, nheads, headdim, dstate, device=x.device, dtype=torch.float32) if return_final_states else None
    p = abi.SsdFwd()
    p.x, p.dt, p.A, p.B, p.C = (abi.tdesc(t) for t in (x, dt, A, B, C))
    p.D, p.z, p.dt_bias = abi.tdesc(D), abi.tdesc(z), abi.tdesc(dt_bias)
    p.initial_states, p.seq_idx = abi.tdesc(initial_states), abi.tdesc(seq_idx)
    p.out, p.final_states = abi.tdesc(out), abi.tdesc(fin)
    nws = abi.ssd_fwd_workspace_bytes(batch, seqlen, nheads, headdim, B.shape[-2], dstate) if algo != "recurrent" else 0
    # scratch for the tensor-core path (fp16 copies of B and C); the caching allocator makes this free in steady state
    ws = torch.empty(nws, device=x.device, dtype=torch.uint8) if nws > 0 and x.dtype == torch.bfloat16 else None
    p.workspace = abi.tdesc(ws)
    p.chunk_size, p.dt_softplus = int(chunk_size), int(bool(dt_softplus))
    p.dt_min, p.dt_max = float(dt_limit[0]), float(min(dt_limit[1], 3.0e38))
    p.algo = _ALGO[_DEFAULT_ALGO if algo == "auto" else algo]
    if chunk_states is not None:
        p.chunk_states = abi.tdesc(chunk_states)
        if not abi.lib().omni_ssd_fwd_saves_chunk_states(abi.C.byref(p)):
            chunk_states = None
            p.chunk_states = abi.tdesc(None)
        abi.call("omni_ssd_chunk_scan_fwd", p, x.device)
        return out, fin, chunk_states
    abi.call("omni_ssd_chunk_scan_fwd", p, x.device)
    return out, fin


def ssd_bwd_raw(dout, x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None, seq_idx=None,
                dt_softplus=False, dt_limit=(0.0, float("inf")), dfinal_states=None, dx=None, ddt=None, dz=None,
                want_dinitial=False, algo="auto", out=None, chunk_states=None):
    """Returns dx, ddt (raw), dA (H), dB, dC (fp32, (B,L,G,N)), dD, dz, ddt_bias, dinitial_states.
    bf16 x/B/C/dout with the OmniMamba geometry run on the tensor-core kernels (algo "auto"); `out` is accepted and ignored."""
    batch, seqlen, nheads, headdim = x.shape
    ngroups, dstate = B.shape[-2], B.shape[-1]
    dev = x.device
    if dx is None:
        dx = torch.empty(batch, seqlen, nheads, headdim, device=dev, dtype=x.dtype)
    if ddt is None:
        ddt = torch.empty(batch, seqlen, nheads, device=dev, dtype=dt.dtype)
    if z is not None and dz is None:
        dz = torch.empty(batch, seqlen, nheads, headdim, device=dev, dtype=z.dtype)
    dB = torch.zeros(batch, seqlen, ngroups, dstate, device=dev, dtype=torch.float32)
    dC = torch.zeros(batch, seqlen, ngroups, dstate, device=dev, dtype=torch.float32)
    dA_part = torch.empty(batch, nheads, device=dev, dtype=torch.float32)
    ddtb_part = torch.empty(batch, nheads, device=dev, dtype=torch.float32)
    dD_part = torch.empty(batch, nheads, headdim, device=dev, dtype=torch.float32)
    dinit = torch.empty(batch, nheads, headdim, dstate, device=dev, dtype=torch.float32) if want_dinitial else None
    if algo == "auto":
        algo = _DEFAULT_ALGO
    p = abi.SsdBwd()
    p.x, p.dt, p.A, p.B, p.C = (abi.tdesc(t) for t in (x, dt, A, B, C))
    p.D, p.z, p.dt_bias = abi.tdesc(D), abi.tdesc(z), abi.tdesc(dt_bias)
    p.initial_states, p.seq_idx = abi.tdesc(initial_states), abi.tdesc(seq_idx)
    p.out = abi.tdesc(None)  # (reserved: no algorithm needs the forward output; r_i is re-formed in fp32)
    p.dout, p.dfinal_states = abi.tdesc(dout), abi.tdesc(dfinal_states)
    p.dx, p.ddt, p.dB, p.dC, p.dz = (abi.tdesc(t) for t in (dx, ddt, dB, dC, dz))
    p.dinitial_states = abi.tdesc(dinit)
    p.chunk_states = abi.tdesc(chunk_states)   # (the forward's chunk states, if it kept them: no forward state sweep)
    p.dA_part, p.ddt_bias_part = abi.tdesc(dA_part), abi.tdesc(ddtb_part)
    p.chunk_size, p.dt_softplus = int(chunk_size), int(bool(dt_softplus))
    p.dt_min, p.dt_max = float(dt_limit[0]), float(min(dt_limit[1], 3.0e38))
    # Tensor-core path?  The library decides (omni_ssd_bwd_tc_supported: ONE eligibility test for dtypes, strides, alignment,
    # geometry, driver) on the params it would be called with: the tensor-core workspace and dD layout are offered first.
    use_tc = False
    if algo != "recurrent" and x.dtype == torch.bfloat16:
        tc_bytes = abi.ssd_bwd_tc_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate)
        if tc_bytes > 0:
            ws = torch.empty((tc_bytes + 3) // 4, device=dev, dtype=torch.float32)   # fp16 B / C copies, chunk states (not zero-filled)
            dD_tc = torch.empty(batch, nheads, device=dev, dtype=torch.float32)
            p.dD_part, p.workspace = abi.tdesc(dD_tc), abi.tdesc(ws)
            use_tc = bool(abi.lib().omni_ssd_bwd_tc_supported(abi.C.byref(p)))
            if use_tc:
                dD_part = dD_tc
    if algo == "chunked_tc" and not use_tc:
        raise RuntimeError("ssd bwd: algo='chunked_tc' needs bf16 x/B/C/dout (16-byte aligned rows), headdim 64, "
                           "d_state 128, an even number of heads per group, D of shape (H), no z / seq_idx")
    if not use_tc:
        ws = torch.zeros(abi.ssd_bwd_workspace_elems(batch, seqlen, nheads, headdim, dstate), device=dev, dtype=torch.float32)
        p.dD_part, p.workspace = abi.tdesc(dD_part), abi.tdesc(ws)
    p.algo = _ALGO["chunked_tc" if use_tc else "recurrent"]
    abi.call("omni_ssd_chunk_scan_bwd", p, dev)
    dA = dA_part.sum(0)
    ddt_bias = ddtb_part.sum(0) if dt_bias is not None else None
    dD = None
    if D is not None:
        dD = dD_part.sum(0) if (D.dim() == 2 or dD_part.dim() == 2) else dD_part.sum((0, 2))
    return dx, ddt, dA, dB, dC, dD, dz, ddt_bias, dinit


class MambaChunkScanCombinedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None, seq_idx=None,
                cu_seqlens=None, dt_softplus=False, dt_limit=(0.0, float("inf")), return_final_states=False,
                return_varlen_states=False):
        if cu_seqlens is not None or return_varlen_states:
            raise NotImplementedError("cu_seqlens / return_varlen_states are not on the OmniMamba path")
        batch, seqlen, nheads, headdim = x.shape
        ngroups, dstate = B.shape[-2], B.shape[-1]
        assert nheads % ngroups == 0, "nheads must be divisible by ngroups"
        assert B.shape == (batch, seqlen, ngroups, dstate)
        assert C.shape == B.shape
        assert dt.shape == (batch, seqlen, nheads)
        assert A.shape == (nheads,)
        if z is not None:
            assert z.shape == x.shape
        if D is not None:
            assert D.shape == (nheads, headdim) or D.shape == (nheads,)
        if dt_bias is not None:
            assert dt_bias.shape == (nheads,)
        if initial_states is not None:
            assert initial_states.shape == (batch, nheads, headdim, dstate)
        if seq_idx is not None:
            assert seq_idx.shape == (batch, seqlen)
            seq_idx = seq_idx.to(torch.int32)
        x, z, B, C = _last_contig(x), _last_contig(z), _last_contig(B), _last_contig(C)
        A = A.float().contiguous()
        dt_bias = dt_bias.contiguous() if dt_bias is not None else None
        initial_states = _last_contig(initial_states)
        cs = None
        if any(ctx.needs_input_grad) and z is None and seq_idx is None:
            cs = _alloc_chunk_states(batch, seqlen, nheads, headdim, dstate, x.device, x.dtype)
        if cs is not None:
            out, fin, cs = ssd_fwd_raw(x, dt, A, B, C, chunk_size, D, z, dt_bias, initial_states, seq_idx, dt_softplus,
                                       dt_limit, return_final_states, chunk_states=cs)
        else:
            out, fin = ssd_fwd_raw(x, dt, A, B, C, chunk_size, D, z, dt_bias, initial_states, seq_idx, dt_softplus,
                                   dt_limit, return_final_states)
        ctx.save_for_backward(x, dt, A, B, C, D, z, dt_bias, initial_states, seq_idx, cs)
        ctx.chunk_size, ctx.dt_softplus, ctx.dt_limit = chunk_size, dt_softplus, dt_limit
        ctx.return_final_states = return_final_states
        return out if not return_final_states else (out, fin)

    @staticmethod
    def backward(ctx, dout, *args):
        x, dt, A, B, C, D, z, dt_bias, initial_states, seq_idx, cs = ctx.saved_tensors
        dfin = args[0] if ctx.return_final_states else None
        if dfin is not None:
            dfin = dfin.float().contiguous()
        dout = _last_contig(dout)
        dx, ddt, dA, dB, dC, dD, dz, ddt_bias, dinit = ssd_bwd_raw(
            dout, x, dt, A, B, C, ctx.chunk_size, D, z, dt_bias, initial_states, seq_idx, ctx.dt_softplus,
            ctx.dt_limit, dfinal_states=dfin, want_dinitial=initial_states is not None, chunk_states=cs)
        return (dx, ddt, dA, dB.to(B.dtype), dC.to(C.dtype), None,
                dD.to(D.dtype) if dD is not None else None, dz,
                ddt_bias.to(dt_bias.dtype) if ddt_bias is not None else None,
                dinit.to(initial_states.dtype) if dinit is not None else None, None, None, None, None, None, None)


def mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None,
                              seq_idx=None, cu_seqlens=None, dt_softplus=False, dt_limit=(0.0, float("inf")),
                              return_final_states=False, return_varlen_states=False):
    """x: (batch, seqlen, nheads, headdim); dt: (batch, seqlen, nheads); A: (nheads); B, C: (batch, seqlen,
    ngroups, dstate); D: (nheads, headdim) | (nheads,); z like x; dt_bias: (nheads,); initial_states: (batch,
    nheads, headdim, dstate); seq_idx: (batch, seqlen).  Returns out like x [, final_states fp32]."""
    return MambaChunkScanCombinedFn.apply(x, dt, A, B, C, chunk_size, D, z, dt_bias, initial_states, seq_idx,
                                          cu_seqlens, dt_softplus, dt_limit, return_final_states,
                                          return_varlen_states)


def _autocast_dtype(device_type="cuda"):
    return torch.get_autocast_dtype(device_type) if torch.is_autocast_enabled(device_type) else None


class MambaSplitConv1dScanCombinedFn(torch.autograd.Function):
    """The training hot op (path A): split zxbcdt -> causal conv1d + SiLU -> SSD scan -> gated RMSNorm -> out_proj
    in one autograd node.  Saves zxbcdt, the pre-norm scan output, rstd and - memory for time on a 180 GB part, both with an
    environment opt-out - the conv output and the fp16 chunk states of the scan; upstream saves only the first three and
    recomputes the conv output and the chunk states in its backward."""

    @staticmethod
    def forward(ctx, zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size, initial_states=None, seq_idx=None,
                dt_limit=(0.0, float("inf")), return_final_states=False, activation="silu", rmsnorm_weight=None,
                rmsnorm_eps=1e-6, outproj_weight=None, outproj_bias=None, headdim=None, ngroups=1,
                norm_before_gate=True):
        assert activation in (None, "silu", "swish")
        if D.dim() == 1:
            assert headdim is not None
            nheads, = D.shape
        else:
            nheads, headdim = D.shape
        batch, seqlen, _ = zxbcdt.shape
        dim = nheads * headdim
        assert nheads % ngroups == 0
        dstate = (conv1d_weight.shape[0] - dim) // ngroups // 2
        d_nonssm = (zxbcdt.shape[-1] - 2 * dim - 2 * ngroups * dstate - nheads) // 2
        assert d_nonssm >= 0
        if d_nonssm > 0:
            raise NotImplementedError("d_mlp > 0 (gated-MLP lanes in zxbcdt) is not on the OmniMamba path")
        assert zxbcdt.shape == (batch, seqlen, 2 * dim + 2 * ngroups * dstate + nheads)
        assert dt_bias.shape == (nheads,)
        assert A.shape == (nheads,)
        if rmsnorm_weight is not None:
            assert rmsnorm_weight.shape == (dim,)
        zxbcdt = zxbcdt.contiguous()
        if seq_idx is not None:
            seq_idx = seq_idx.to(torch.int32)
        conv_dim = dim + 2 * ngroups * dstate
        A = A.float().contiguous()
        z = zxbcdt[..., :dim]
        xBC = zxbcdt[..., dim:dim + conv_dim]
        dt = zxbcdt[..., dim + conv_dim:]
        act = abi.ACT_NONE if activation is None else abi.ACT_SILU
        dev, dt_ = zxbcdt.device, zxbcdt.dtype
        xBC_conv = torch.empty(batch, seqlen, conv_dim, device=dev, dtype=dt_)
        scan_out = torch.empty(batch, seqlen, dim, device=dev, dtype=dt_)
        fin = torch.empty(batch, nheads, headdim, dstate, device=dev, dtype=torch.float32) if return_final_states else None
        rstd = y = None
        if rmsnorm_weight is not None:
            rmsnorm_weight = rmsnorm_weight.contiguous()
            rstd = torch.empty(batch * seqlen * ngroups, device=dev, dtype=torch.float32)
            y = torch.empty(batch, seqlen, dim, device=dev, dtype=dt_)
        # out_proj inside the same call when its operands are bf16 (autocast / bf16 model): the tcgen05 GEMM
        ac = _autocast_dtype()
        w, b_ = outproj_weight, outproj_bias
        if w is not None and ac is not None:
            w = cast_param(w, ac)
            b_ = b_.to(ac) if b_ is not None else None
        fuse_gemm = (w is not None and w.dtype == torch.bfloat16 and dt_ == torch.bfloat16 and w.stride(-1) == 1
                     and w.stride(0) % 8 == 0 and w.data_ptr() % 16 == 0 and w.shape[0] % 8 == 0 and abi.gemm_supported())
        out = torch.empty(batch, seqlen, w.shape[0], device=dev, dtype=torch.bfloat16) if fuse_gemm else None
        nws = abi.ssd_fwd_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate) if dt_ == torch.bfloat16 else 0
        ws = torch.empty(nws, device=dev, dtype=torch.uint8) if nws > 0 else None
        p = abi.SplitConv1dScanFwd()
        p.zxbcdt, p.conv1d_weight, p.conv1d_bias = abi.tdesc(zxbcdt), abi.tdesc(conv1d_weight), abi.tdesc(conv1d_bias)
        p.dt_bias, p.A, p.D = abi.tdesc(dt_bias.contiguous()), abi.tdesc(A), abi.tdesc(D)
        p.initial_states, p.seq_idx = abi.tdesc(_last_contig(initial_states)), abi.tdesc(seq_idx)
        p.rmsnorm_weight, p.outproj_weight = abi.tdesc(rmsnorm_weight), abi.tdesc(w if fuse_gemm else None)
        p.xbc_conv, p.scan_out, p.rstd, p.y, p.out = (abi.tdesc(t) for t in (xBC_conv, scan_out, rstd, y, out))
        p.final_states, p.workspace = abi.tdesc(fin), abi.tdesc(ws)
        cs = None
        if any(ctx.needs_input_grad) and rmsnorm_weight is not None and seq_idx is None:
            cs = _alloc_chunk_states(batch, seqlen, nheads, headdim, dstate, dev, dt_)
        if cs is not None:   # ask the library whether the scan inside the call will fill it (same views as the C side builds)
            q = abi.SsdFwd()
            q.x = abi.tdesc(xBC_conv[..., :dim].view(batch, seqlen, nheads, headdim))
            q.B = abi.tdesc(xBC_conv[..., dim:dim + ngroups * dstate].view(batch, seqlen, ngroups, dstate))
            q.C = abi.tdesc(xBC_conv[..., dim + ngroups * dstate:].view(batch, seqlen, ngroups, dstate))
            q.dt, q.A, q.D, q.dt_bias = abi.tdesc(dt), abi.tdesc(A), abi.tdesc(D), abi.tdesc(dt_bias.contiguous())
            q.initial_states = abi.tdesc(_last_contig(initial_states))
            q.out, q.final_states = abi.tdesc(scan_out.view(batch, seqlen, nheads, headdim)), abi.tdesc(fin)
            q.workspace, q.chunk_states, q.algo = abi.tdesc(ws), abi.tdesc(cs), _ALGO[_DEFAULT_ALGO]
            if not abi.lib().omni_ssd_fwd_saves_chunk_states(abi.C.byref(q)):
                cs = None
        p.chunk_states = abi.tdesc(cs)
        p.nheads, p.headdim, p.ngroups, p.dstate, p.chunk_size = nheads, headdim, ngroups, dstate, int(chunk_size)
        p.activation, p.norm_before_gate, p.algo = act, int(bool(norm_before_gate)), _ALGO[_DEFAULT_ALGO]
        p.dt_min, p.dt_max, p.rmsnorm_eps = float(dt_limit[0]), float(min(dt_limit[1], 3.0e38)), float(rmsnorm_eps)
        abi.call("omni_split_conv1d_scan_fwd", p, dev)     # conv1d + SiLU -> scan -> gated norm (-> out_proj): one C-ABI call
        scan_out = scan_out.view(batch, seqlen, nheads, headdim)
        if y is None:
            y = scan_out.view(batch, seqlen, dim)
        ctx.outproj_weight_dtype = outproj_weight.dtype if outproj_weight is not None else None
        if fuse_gemm:
            if b_ is not None:
                out = out + b_
        elif w is not None:
            y2 = (y.to(ac) if ac is not None else y.to(w.dtype)).reshape(batch * seqlen, dim)
            out = mm_nt(y2, w).view(batch, seqlen, w.shape[0])
            if b_ is not None:
                out = out + b_
        else:
            out = y
        ctx.save_for_backward(zxbcdt, conv1d_weight, conv1d_bias, scan_out, A, D, dt_bias, initial_states, seq_idx,
                              rmsnorm_weight, rstd, outproj_weight, outproj_bias, cs,
                              xBC_conv if (_KEEP_CONV and any(ctx.needs_input_grad)) else None)
        ctx.dt_limit, ctx.return_final_states, ctx.act = dt_limit, return_final_states, act
        ctx.rmsnorm_eps, ctx.norm_before_gate, ctx.chunk_size = rmsnorm_eps, norm_before_gate, chunk_size
        ctx.headdim, ctx.ngroups = headdim, ngroups
        return out if not return_final_states else (out, fin)

    @staticmethod
    def backward(ctx, dout, *args):
        (zxbcdt, conv1d_weight, conv1d_bias, scan_out, A, D, dt_bias, initial_states, seq_idx, rmsnorm_weight, rstd,
         outproj_weight, outproj_bias, cs, xBC_conv) = ctx.saved_tensors
        dfin = args[0] if ctx.return_final_states else None
        if dfin is not None:
            dfin = dfin.float().contiguous()
        headdim, ngroups = ctx.headdim, ctx.ngroups
        nheads = D.shape[0]
        dim = nheads * headdim
        batch, seqlen, _ = zxbcdt.shape
        conv_dim = conv1d_weight.shape[0]
        dstate = (conv_dim - dim) // ngroups // 2
        dev = zxbcdt.device
        M = batch * seqlen
        z = zxbcdt[..., :dim]
        xBC = zxbcdt[..., dim:dim + conv_dim]
        dt = zxbcdt[..., dim + conv_dim:]
        if xBC_conv is None:   # recompute the conv output (OMNI_KEEP_CONV_OUT=0: upstream's behaviour)
            xBC_conv = torch.empty(batch, seqlen, conv_dim, device=dev, dtype=zxbcdt.dtype)
            conv1d_fwd_raw(xBC.transpose(1, 2), conv1d_weight, conv1d_bias, seq_idx, None, xBC_conv.transpose(1, 2), None,
                           ctx.act)
        x = xBC_conv[..., :dim].view(batch, seqlen, nheads, headdim)
        Bm = xBC_conv[..., dim:dim + ngroups * dstate].view(batch, seqlen, ngroups, dstate)
        Cm = xBC_conv[..., dim + ngroups * dstate:].view(batch, seqlen, ngroups, dstate)
        dzxbcdt = torch.empty_like(zxbcdt)
        dz = dzxbcdt[..., :dim]
        dxBC = dzxbcdt[..., dim:dim + conv_dim]
        ddt = dzxbcdt[..., dim + conv_dim:]
        dxBC_conv = torch.empty(batch, seqlen, conv_dim, device=dev, dtype=zxbcdt.dtype)
        dx = dxBC_conv[..., :dim].view(batch, seqlen, nheads, headdim)
        doutproj_weight = doutproj_bias = drmsnorm_weight = None
        if outproj_weight is not None:
            dout2 = dout.reshape(M, dout.shape[-1])
            if dout2.stride(-1) != 1 and dout2.stride(0) != 1:
                dout2 = dout2.contiguous()
            dy = mm_nt(dout2, cast_param(outproj_weight, dout2.dtype).t())
            doutproj_bias = dout2.sum(0).to(outproj_bias.dtype) if outproj_bias is not None else None
        else:
            dy = dout.reshape(M, dim)
        if dy.stride(-1) != 1:
            dy = dy.contiguous()
        if rmsnorm_weight is not None:
            z2 = z.reshape(M, dim)      # views: zxbcdt / dzxbcdt are contiguous, only the row pitch differs
            dz2 = dz.reshape(M, dim)
            assert dz2.data_ptr() == dz.data_ptr()
            dscan = torch.empty(M, dim, device=dev, dtype=scan_out.dtype)
            _, drmsnorm_weight, _, _, y_rec = norm_gated_bwd_raw(
                dy.to(scan_out.dtype), scan_out.view(M, dim), rmsnorm_weight, None, z2, None, rstd, ctx.rmsnorm_eps,
                dim // ngroups, ctx.norm_before_gate, True, dx=dscan, dz=dz2, recompute_output=outproj_weight is not None and ctx.needs_input_grad[14])
            if outproj_weight is not None and ctx.needs_input_grad[14]:   # (frozen out_proj - stage "align" - skips its wgrad)
                doutproj_weight = mm_nt(dout2.t(), y_rec.to(dout2.dtype).t()).to(outproj_weight.dtype)
            dscan = dscan.view(batch, seqlen, nheads, headdim)
            zscan, dzscan = None, None
        else:
            if outproj_weight is not None and ctx.needs_input_grad[14]:
                # y = scan_out (gated inside the scan): the saved tensor is the GEMM input
                doutproj_weight = mm_nt(dout2.t(), scan_out.view(M, dim).to(dout2.dtype).t()).to(outproj_weight.dtype)
            dscan = dy.to(scan_out.dtype).view(batch, seqlen, nheads, headdim)
            zscan = z.view(batch, seqlen, nheads, headdim)
            dzscan = dz.view(batch, seqlen, nheads, headdim)
        _, _, dA, dB, dC, dD, _, ddt_bias, dinit = ssd_bwd_raw(
            dscan, x, dt, A, Bm, Cm, ctx.chunk_size, D, zscan, dt_bias, initial_states, seq_idx, True, ctx.dt_limit,
            dfinal_states=dfin, dx=dx, ddt=ddt, dz=dzscan, want_dinitial=initial_states is not None, chunk_states=cs)
        dxBC_conv[..., dim:dim + ngroups * dstate].copy_(dB.view(batch, seqlen, ngroups * dstate))
        dxBC_conv[..., dim + ngroups * dstate:].copy_(dC.view(batch, seqlen, ngroups * dstate))
        dweight = torch.zeros(conv1d_weight.shape, device=dev, dtype=torch.float32)
        dbias = torch.zeros(conv_dim, device=dev, dtype=torch.float32) if conv1d_bias is not None else None
        conv1d_bwd_raw(xBC.transpose(1, 2), conv1d_weight, conv1d_bias, dxBC_conv.transpose(1, 2), seq_idx, None,
                       dxBC.transpose(1, 2), dweight, dbias, None, ctx.act)
        return (dzxbcdt, dweight.to(conv1d_weight.dtype), dbias.to(conv1d_bias.dtype) if dbias is not None else None,
                ddt_bias.to(dt_bias.dtype), dA, dD.to(D.dtype), None,
                dinit.to(initial_states.dtype) if dinit is not None else None, None, None, None, None,
                drmsnorm_weight, None, doutproj_weight, doutproj_bias, None, None, None)


def mamba_split_conv1d_scan_combined(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size, initial_states=None,
                                     seq_idx=None, dt_limit=(0.0, float("inf")), return_final_states=False,
                                     activation="silu", rmsnorm_weight=None, rmsnorm_eps=1e-6, outproj_weight=None,
                                     outproj_bias=None, headdim=None, ngroups=1, norm_before_gate=True):
    """zxbcdt: (batch, seqlen, 2*dim + 2*ngroups*dstate + nheads); conv1d_weight: (dim + 2*ngroups*dstate, width);
    dt_bias, A: (nheads,); D: (nheads, headdim) | (nheads,); rmsnorm_weight: (dim,); outproj_weight: (out_dim, dim).
    Returns out (batch, seqlen, out_dim | dim) [, final_states]."""
    return MambaSplitConv1dScanCombinedFn.apply(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size,
                                                initial_states, seq_idx, dt_limit, return_final_states, activation,
                                                rmsnorm_weight, rmsnorm_eps, outproj_weight, outproj_bias, headdim,
                                                ngroups, norm_before_gate)
