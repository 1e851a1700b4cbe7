"""Gated (group) RMSNorm / LayerNorm with the mamba_ssm==2.2.2 signatures
(mamba_ssm/ops/triton/layernorm_gated.py upstream; Mamba2.norm)."""
from __future__ import annotations

import torch

from .. import _cabi as abi


def _nparts(device, rows):
    sm = torch.cuda.get_device_properties(device).multi_processor_count
    return max(1, min(rows, 2 * sm))


def norm_gated_fwd_raw(x2, weight, bias, z2, eps, group_size, norm_before_gate, is_rms_norm, out=None):
    """x2, z2: (M, D) row-major views (row stride free).  Returns out (M, D), mean, rstd."""
    M, D = x2.shape
    gs = D if group_size is None else group_size
    if out is None:
        out = torch.empty(M, D, device=x2.device, dtype=x2.dtype)
    ng = D // gs
    rstd = torch.empty(M * ng, device=x2.device, dtype=torch.float32)
    mean = torch.empty(M * ng, device=x2.device, dtype=torch.float32) if not is_rms_norm else None
    p = abi.NormGatedFwd()
    p.x, p.weight, p.bias, p.z = abi.tdesc(x2), abi.tdesc(weight), abi.tdesc(bias), abi.tdesc(z2)
    p.out, p.rstd, p.mean = abi.tdesc(out), abi.tdesc(rstd), abi.tdesc(mean)
    p.eps, p.group_size, p.norm_before_gate, p.is_rms_norm = eps, gs, int(norm_before_gate), int(is_rms_norm)
    abi.call("omni_norm_gated_fwd", p, x2.device)
    return out, mean, rstd


def norm_gated_bwd_raw(dy2, x2, weight, bias, z2, mean, rstd, eps, group_size, norm_before_gate, is_rms_norm,
                       dx=None, dz=None, recompute_output=False):
    M, D = x2.shape
    gs = D if group_size is None else group_size
    if dx is None:
        dx = torch.empty(M, D, device=x2.device, dtype=x2.dtype)
    if z2 is not None and dz is None:
        dz = torch.empty(M, D, device=x2.device, dtype=z2.dtype)
    nparts = _nparts(x2.device, M)
    dw_part = torch.empty(nparts, D, device=x2.device, dtype=torch.float32)
    db_part = torch.empty(nparts, D, device=x2.device, dtype=torch.float32) if bias is not None else None
    yrec = torch.empty(M, D, device=x2.device, dtype=x2.dtype) if recompute_output else None
    p = abi.NormGatedBwd()
    p.x, p.weight, p.bias, p.z, p.dout = (abi.tdesc(t) for t in (x2, weight, bias, z2, dy2))
    p.rstd, p.mean = abi.tdesc(rstd), abi.tdesc(mean)
    p.dx, p.dz, p.dw_part, p.db_part, p.out_recompute = (abi.tdesc(t) for t in (dx, dz, dw_part, db_part, yrec))
    p.eps, p.group_size, p.norm_before_gate, p.is_rms_norm = eps, gs, int(norm_before_gate), int(is_rms_norm)
    abi.call("omni_norm_gated_bwd", p, x2.device)
    dw = dw_part.sum(0).to(weight.dtype)
    db = db_part.sum(0).to(bias.dtype) if bias is not None else None
    return dx, dw, db, dz, yrec


def _rows(t):
    """(..., D) -> (M, D) view with contiguous last dim (copies only if it must)."""
    if t is None:
        return None
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, z=None, eps=1e-6, group_size=None, norm_before_gate=True, is_rms_norm=False):
        x_shape_og = x.shape
        x2, z2 = _rows(x), _rows(z)
        if z is not None:
            assert z.shape == x_shape_og
        weight = weight.contiguous()
        bias = bias.contiguous() if bias is not None else None
        y, mean, rstd = norm_gated_fwd_raw(x2, weight, bias, z2, eps, group_size, norm_before_gate, is_rms_norm)
        ctx.save_for_backward(x2, weight, bias, mean, rstd, z2)
        ctx.x_shape_og, ctx.eps, ctx.group_size = x_shape_og, eps, group_size
        ctx.norm_before_gate, ctx.is_rms_norm = norm_before_gate, is_rms_norm
        return y.reshape(x_shape_og)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias, mean, rstd, z2 = ctx.saved_tensors
        dy2 = _rows(dy)
        dx, dw, db, dz, _ = norm_gated_bwd_raw(dy2, x2, weight, bias, z2, mean, rstd, ctx.eps, ctx.group_size,
                                               ctx.norm_before_gate, ctx.is_rms_norm)
        return (dx.reshape(ctx.x_shape_og), dw, db, dz.reshape(ctx.x_shape_og) if dz is not None else None,
                None, None, None, None)


def layernorm_fn(x, weight, bias, z=None, eps=1e-6, group_size=None, norm_before_gate=True, is_rms_norm=False):
    return LayerNormFn.apply(x, weight, bias, z, eps, group_size, norm_before_gate, is_rms_norm)


def rmsnorm_fn(x, weight, bias, z=None, eps=1e-6, group_size=None, norm_before_gate=True):
    return LayerNormFn.apply(x, weight, bias, z, eps, group_size, norm_before_gate, True)


class LayerNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, group_size=None, norm_before_gate=True, device=None, dtype=None):
        """If group_size is not None, we do GroupNorm with each group having group_size elements.
        group_size=None is equivalent to group_size=hidden_size (i.e. there's only 1 group)."""
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(hidden_size, **factory_kwargs))
        self.bias = torch.nn.Parameter(torch.empty(hidden_size, **factory_kwargs))
        self.group_size = group_size
        self.norm_before_gate = norm_before_gate
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)
        torch.nn.init.zeros_(self.bias)

    def forward(self, x, z=None):
        """If z is not None, we do norm(x) * silu(z) if norm_before_gate, else norm(x * silu(z))"""
        return layernorm_fn(x, self.weight, self.bias, z=z, group_size=self.group_size, eps=self.eps,
                            norm_before_gate=self.norm_before_gate)


class RMSNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, group_size=None, norm_before_gate=True, device=None, dtype=None):
        """If group_size is not None, we do GroupNorm with each group having group_size elements.
        group_size=None is equivalent to group_size=hidden_size (i.e. there's only 1 group)."""
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(hidden_size, **factory_kwargs))
        self.register_parameter("bias", None)
        self.group_size = group_size
        self.norm_before_gate = norm_before_gate
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)

    def forward(self, x, z=None):
        """If z is not None, we do norm(x) * silu(z) if norm_before_gate, else norm(x * silu(z))"""
        return rmsnorm_fn(x, self.weight, self.bias, z=z, eps=self.eps, group_size=self.group_size,
                          norm_before_gate=self.norm_before_gate)
