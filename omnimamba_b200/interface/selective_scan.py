"""selective_scan_fn (Mamba-1) with the mamba_ssm==2.2.2 signature
(mamba_ssm/ops/selective_scan_interface.py upstream; reachable in OmniMamba only with ssm_cfg.layer="Mamba1",
/root/reference/models/stage2/mixer_seq_simple.py:197-201; arithmetic SURVEY.md A.6)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _cabi as abi


def selscan_fwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    out = torch.empty_like(u)
    last = torch.empty(batch, dim, dstate, device=u.device, dtype=torch.float32) if return_last_state else None
    p = abi.SelScanFwd()
    p.u, p.delta, p.A, p.B, p.C = (abi.tdesc(t) for t in (u, delta, A, B, C))
    p.D, p.z, p.delta_bias = abi.tdesc(D), abi.tdesc(z), abi.tdesc(delta_bias)
    p.out, p.last_state = abi.tdesc(out), abi.tdesc(last)
    p.delta_softplus = int(bool(delta_softplus))
    abi.call("omni_selective_scan_fwd", p, u.device)
    return out, last


def selscan_bwd_raw(dout, u, delta, A, B, C, D, z, delta_bias, delta_softplus):
    """Returns du, ddelta, dA (D, N), dB, dC (fp32, (B, G, N, L)), dD, dz, ddelta_bias."""
    batch, dim, seqlen = u.shape
    dstate, ngroups = A.shape[1], B.shape[1]
    dev = u.device
    du, ddelta = torch.empty_like(u), torch.empty_like(delta)
    dz = torch.empty_like(z) if z is not None else None
    dB = torch.zeros(batch, ngroups, dstate, seqlen, device=dev, dtype=torch.float32)
    dC = torch.zeros_like(dB)
    dA_part = torch.empty(batch, dim, dstate, device=dev, dtype=torch.float32)
    dD_part = torch.empty(batch, dim, device=dev, dtype=torch.float32) if D is not None else None
    ddb_part = torch.empty(batch, dim, device=dev, dtype=torch.float32) if delta_bias is not None else None
    ws = torch.empty(max(1, abi.selscan_bwd_workspace_elems(batch, dim, seqlen, dstate)), device=dev, dtype=torch.float32)
    p = abi.SelScanBwd()
    p.u, p.delta, p.A, p.B, p.C = (abi.tdesc(t) for t in (u, delta, A, B, C))
    p.D, p.z, p.delta_bias, p.dout = abi.tdesc(D), abi.tdesc(z), abi.tdesc(delta_bias), abi.tdesc(dout)
    p.du, p.ddelta, p.dB, p.dC, p.dz = (abi.tdesc(t) for t in (du, ddelta, dB, dC, dz))
    p.dA_part, p.dD_part, p.ddelta_bias_part = abi.tdesc(dA_part), abi.tdesc(dD_part), abi.tdesc(ddb_part)
    p.workspace = abi.tdesc(ws)
    p.delta_softplus = int(bool(delta_softplus))
    abi.call("omni_selective_scan_bwd", p, dev)
    return (du, ddelta, dA_part.sum(0), dB, dC, dD_part.sum(0) if dD_part is not None else None, dz,
            ddb_part.sum(0) if ddb_part is not None else None)


class SelectiveScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        if A.is_complex():
            raise NotImplementedError("complex A is not on the OmniMamba path")
        if u.stride(-1) != 1:
            u = u.contiguous()
        if delta.stride(-1) != 1:
            delta = delta.contiguous()
        if z is not None and z.stride(-1) != 1:
            z = z.contiguous()
        ctx.squeeze_B, ctx.squeeze_C = B.dim() == 3, C.dim() == 3
        if B.dim() == 3:
            B = B.unsqueeze(1)
        if C.dim() == 3:
            C = C.unsqueeze(1)
        if B.stride(-1) != 1:
            B = B.contiguous()
        if C.stride(-1) != 1:
            C = C.contiguous()
        ctx.param_dtypes = (A.dtype, D.dtype if D is not None else None, delta_bias.dtype if delta_bias is not None else None)
        A = A.float()
        D = D.float().contiguous() if D is not None else None
        delta_bias = delta_bias.float().contiguous() if delta_bias is not None else None
        out, last = selscan_fwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias)
        ctx.delta_softplus = delta_softplus
        if return_last_state:
            ctx.mark_non_differentiable(last)
        return out if not return_last_state else (out, last)

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, z, delta_bias = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, dz, ddb = selscan_bwd_raw(dout.to(u.dtype), u, delta, A, B, C, D, z, delta_bias,
                                                              ctx.delta_softplus)
        dB, dC = dB.to(B.dtype), dC.to(C.dtype)
        if ctx.squeeze_B:
            dB = dB.squeeze(1)
        if ctx.squeeze_C:
            dC = dC.squeeze(1)
        tA, tD, tb = ctx.param_dtypes
        return (du, ddelta, dA.to(tA), dB, dC, dD.to(tD) if dD is not None else None, dz,
                ddb.to(tb) if ddb is not None else None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """u, delta, z: (B, D, L); A: (D, N); B, C: (B, N, L) | (B, G, N, L); D, delta_bias: (D).
    Returns out (B, D, L) [, last_state (B, D, N) fp32]."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """The API's own pure-PyTorch definition of the op (upstream exports one under this name for its tests).
    Token-by-token recurrence in fp32; differentiable; NOT used by any module of this package."""
    dtype_in = u.dtype
    u, delta = u.float(), delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    G = B.shape[1]
    Bf = B.float().repeat_interleave(dim // G, dim=1)  # (B, D, N, L)
    Cf = C.float().repeat_interleave(dim // G, dim=1)
    state = u.new_zeros(batch, dim, dstate)
    ys = []
    for t in range(seqlen):
        state = torch.exp(delta[:, :, t, None] * A.float()) * state + (delta[:, :, t] * u[:, :, t])[..., None] * Bf[..., t]
        ys.append((state * Cf[..., t]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u * D.float()[:, None]
    if z is not None:
        y = y * F.silu(z.float())
    y = y.to(dtype_in)
    return y if not return_last_state else (y, state)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias,
                   A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """The Mamba-1 block after in_proj, upstream signature (mamba_ssm/ops/selective_scan_interface.py: mamba_inner_fn):
    xz (B, 2*d_inner, L) -> causal conv1d + SiLU -> x_proj -> dt_proj -> selective scan (gated by z) -> out_proj.
    Upstream fuses this into one autograd node to save activation memory; here it is the composition of the same kernels
    (causal_conv1d_fn, selective_scan_fn, both differentiable) and the three small GEMMs - same values, same gradients."""
    from .causal_conv1d import causal_conv1d_fn
    L = xz.shape[-1]
    delta_rank = delta_proj_weight.shape[1]
    d_state = A.shape[-1]
    x, z = xz.chunk(2, dim=1)
    w = conv1d_weight.squeeze(1) if conv1d_weight.dim() == 3 else conv1d_weight
    x = causal_conv1d_fn(x, w, conv1d_bias, activation="silu")
    batch, d_inner = x.shape[0], x.shape[1]
    x_dbl = F.linear(x.transpose(1, 2).reshape(batch * L, d_inner), x_proj_weight)          # (B*L, rank + 2N)
    delta = (delta_proj_weight @ x_dbl[:, :delta_rank].t()).view(d_inner, batch, L).transpose(0, 1)
    if B is None:
        B = x_dbl[:, delta_rank:delta_rank + d_state]
        if B_proj_bias is not None:
            B = B + B_proj_bias.to(dtype=B.dtype)
        B = B.view(batch, L, d_state).transpose(1, 2).contiguous()
    if C is None:
        C = x_dbl[:, -d_state:]
        if C_proj_bias is not None:
            C = C + C_proj_bias.to(dtype=C.dtype)
        C = C.view(batch, L, d_state).transpose(1, 2).contiguous()
    y = selective_scan_fn(x.contiguous(), delta.contiguous(), A, B, C, D, z=z.contiguous(), delta_bias=delta_bias,
                          delta_softplus=delta_softplus)
    return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)
