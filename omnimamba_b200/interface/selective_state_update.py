"""selective_state_update with the mamba_ssm==2.2.2 signature
(mamba_ssm/ops/triton/selective_state_update.py upstream; Mamba2.step)."""
from __future__ import annotations

import torch

from .. import _cabi as abi


def selective_state_update(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False):
    """state: (batch, dim, dstate) or (batch, nheads, dim, dstate) - updated IN PLACE;
    x, dt, z: (batch, dim) | (batch, nheads, dim); A: (dim, dstate) | (nheads, dim, dstate);
    B, C: (batch, dstate) | (batch, ngroups, dstate); D, dt_bias: (dim,) | (nheads, dim).
    Returns out shaped like x."""
    has_heads = state.dim() > 3
    if state.dim() == 3:
        state = state.unsqueeze(1)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if dt.dim() == 2:
        dt = dt.unsqueeze(1)
    if A.dim() == 2:
        A = A.unsqueeze(0)
    if B.dim() == 2:
        B = B.unsqueeze(1)
    if C.dim() == 2:
        C = C.unsqueeze(1)
    if D is not None and D.dim() == 1:
        D = D.unsqueeze(0)
    if z is not None and z.dim() == 2:
        z = z.unsqueeze(1)
    if dt_bias is not None and dt_bias.dim() == 1:
        dt_bias = dt_bias.unsqueeze(0)
    batch, nheads, dim, dstate = state.shape
    assert x.shape == (batch, nheads, dim)
    assert dt.shape == x.shape
    assert A.shape == (nheads, dim, dstate)
    ngroups = B.shape[1]
    assert nheads % ngroups == 0, "nheads must be divisible by ngroups"
    assert B.shape == (batch, ngroups, dstate)
    assert C.shape == B.shape
    if D is not None:
        assert D.shape == (nheads, dim)
    if z is not None:
        assert z.shape == x.shape
    if dt_bias is not None:
        assert dt_bias.shape == (nheads, dim)
    if B.stride(-1) != 1:
        B = B.contiguous()
    if C.stride(-1) != 1:
        C = C.contiguous()
    out = torch.empty_like(x)
    p = abi.Ssu()
    p.state, p.x, p.dt, p.A, p.B, p.C = (abi.tdesc(t) for t in (state, x, dt, A, B, C))
    p.D, p.z, p.dt_bias, p.out = abi.tdesc(D), abi.tdesc(z), abi.tdesc(dt_bias), abi.tdesc(out)
    p.dt_softplus = int(bool(dt_softplus))
    abi.call("omni_selective_state_update", p, x.device)
    if not has_heads:
        out = out.squeeze(1)
    return out
