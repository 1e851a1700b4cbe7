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
        if B.dim() == 3:
            B = B.unsqueeze(1)
        if C.dim() == 3:
            C = C.unsqueeze(1)
        if B.stride(-1) != 1:
            B = B.contiguous()
        if C.stride(-1) != 1:
            C = C.contiguous()
        A = A.float()
        D = D.float().contiguous() if D is not None else None
        delta_bias = delta_bias.float().contiguous() if delta_bias is not None else None
        out, last = selscan_fwd_raw(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
        if return_last_state:
            ctx.mark_non_differentiable(last)
        return out if not return_last_state else (out, last)

    @staticmethod
    def backward(ctx, dout, *args):
        raise NotImplementedError(
            "selective_scan_fn backward (Mamba-1) is not implemented in libomnissm: OmniMamba's default "
            "ssm_cfg.layer is Mamba2 (models/stage2/config_mamba.py:16)")


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


def mamba_inner_fn(*args, **kwargs):
    raise NotImplementedError("mamba_inner_fn (fused Mamba-1 block) is not on the OmniMamba path; "
                              "Mamba (v1) in this package runs conv1d + selective_scan_fn separately")
