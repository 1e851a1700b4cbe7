"""Fused residual-add + LayerNorm/RMSNorm with the mamba_ssm==2.2.2 signatures
(mamba_ssm/ops/triton/layer_norm.py upstream).  Reference call sites:
/root/reference/models/stage2/block.py:86-95 and mixer_seq_simple.py:428-437."""
from __future__ import annotations

import torch

from .. import _cabi as abi
from .layernorm_gated import _nparts, _rows


def add_norm_fwd_raw(x2, weight, bias, residual2, eps, is_rms_norm, out_dtype=None, residual_dtype=None):
    M, D = x2.shape
    y = torch.empty(M, D, device=x2.device, dtype=x2.dtype if out_dtype is None else out_dtype)
    if residual2 is not None:
        residual_dtype = residual2.dtype
    if residual2 is not None or (residual_dtype is not None and residual_dtype != x2.dtype):
        residual_out = torch.empty(M, D, device=x2.device, dtype=residual_dtype)
    else:
        residual_out = None
    rstd = torch.empty(M, device=x2.device, dtype=torch.float32)
    mean = torch.empty(M, device=x2.device, dtype=torch.float32) if not is_rms_norm else None
    p = abi.AddNormFwd()
    p.x, p.residual, p.weight, p.bias = abi.tdesc(x2), abi.tdesc(residual2), abi.tdesc(weight), abi.tdesc(bias)
    p.y, p.residual_out, p.rstd, p.mean = abi.tdesc(y), abi.tdesc(residual_out), abi.tdesc(rstd), abi.tdesc(mean)
    p.eps, p.is_rms_norm = eps, int(is_rms_norm)
    abi.call("omni_add_norm_fwd", p, x2.device)
    # residual_out is None <=> no add happened and dtypes match: the "residual" is x itself
    return y, mean, rstd, residual_out if residual_out is not None else x2


def add_norm_bwd_raw(dy2, xres2, weight, bias, eps, mean, rstd, dresidual2, has_residual, is_rms_norm, x_dtype):
    M, D = xres2.shape
    dx = torch.empty(M, D, device=xres2.device, dtype=x_dtype)
    dresidual_out = None
    if has_residual and dx.dtype != xres2.dtype:
        dresidual_out = torch.empty(M, D, device=xres2.device, dtype=xres2.dtype)
    nparts = _nparts(xres2.device, M)
    dw_part = torch.empty(nparts, D, device=xres2.device, dtype=torch.float32)
    db_part = torch.empty(nparts, D, device=xres2.device, dtype=torch.float32) if bias is not None else None
    p = abi.AddNormBwd()
    p.xres, p.weight, p.bias, p.dy = abi.tdesc(xres2), abi.tdesc(weight), abi.tdesc(bias), abi.tdesc(dy2)
    p.dresidual_in, p.rstd, p.mean = abi.tdesc(dresidual2), abi.tdesc(rstd), abi.tdesc(mean)
    p.dx, p.dresidual, p.dw_part, p.db_part = (abi.tdesc(t) for t in (dx, dresidual_out, dw_part, db_part))
    p.eps, p.is_rms_norm = eps, int(is_rms_norm)
    abi.call("omni_add_norm_bwd", p, xres2.device)
    dw = dw_part.sum(0).to(weight.dtype)
    db = db_part.sum(0).to(bias.dtype) if bias is not None else None
    if has_residual and dresidual_out is None:
        dresidual_out = dx
    return dx, dw, db, dresidual_out


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                is_rms_norm=False):
        x_shape_og = x.shape
        x2 = _rows(x)
        residual2 = None
        if residual is not None:
            assert residual.shape == x_shape_og
            residual2 = _rows(residual)
        weight = weight.contiguous()
        bias = bias.contiguous() if bias is not None else None
        residual_dtype = residual.dtype if residual is not None else (torch.float32 if residual_in_fp32 else None)
        y, mean, rstd, residual_out = add_norm_fwd_raw(x2, weight, bias, residual2, eps, is_rms_norm,
                                                       residual_dtype=residual_dtype)
        ctx.save_for_backward(residual_out, weight, bias, mean, rstd)
        ctx.x_shape_og, ctx.eps, ctx.is_rms_norm = x_shape_og, eps, is_rms_norm
        ctx.has_residual, ctx.prenorm, ctx.x_dtype = residual is not None, prenorm, x.dtype
        y = y.reshape(x_shape_og)
        return y if not prenorm else (y, residual_out.reshape(x_shape_og))

    @staticmethod
    def backward(ctx, dy, *args):
        xres, weight, bias, mean, rstd = ctx.saved_tensors
        dy2 = _rows(dy)
        dresidual2 = None
        if ctx.prenorm:
            dresidual = args[0]
            if dresidual is not None:
                dresidual2 = _rows(dresidual)
        dx, dw, db, dres = add_norm_bwd_raw(dy2, xres, weight, bias, ctx.eps, mean, rstd, dresidual2, ctx.has_residual,
                                            ctx.is_rms_norm, ctx.x_dtype)
        return (dx.reshape(ctx.x_shape_og), dw, db, dres.reshape(ctx.x_shape_og) if ctx.has_residual else None,
                None, None, None, None)


def _check_unsupported(x1, weight1, bias1, dropout_p, rowscale, return_dropout_mask):
    if x1 is not None or weight1 is not None or bias1 is not None:
        raise NotImplementedError("parallel-residual (x1/weight1/bias1) layer norm is not on the OmniMamba path")
    if dropout_p != 0.0 or rowscale is not None or return_dropout_mask:
        raise NotImplementedError("dropout / rowscale in layer_norm_fn is not on the OmniMamba path")


def layer_norm_fn(x, weight, bias, residual=None, x1=None, weight1=None, bias1=None, eps=1e-6, dropout_p=0.0,
                  rowscale=None, prenorm=False, residual_in_fp32=False, is_rms_norm=False,
                  return_dropout_mask=False):
    _check_unsupported(x1, weight1, bias1, dropout_p, rowscale, return_dropout_mask)
    return LayerNormFn.apply(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm)


def rms_norm_fn(x, weight, bias, residual=None, x1=None, weight1=None, bias1=None, eps=1e-6, dropout_p=0.0,
                rowscale=None, prenorm=False, residual_in_fp32=False, return_dropout_mask=False):
    _check_unsupported(x1, weight1, bias1, dropout_p, rowscale, return_dropout_mask)
    return LayerNormFn.apply(x, weight, bias, residual, eps, prenorm, residual_in_fp32, True)


class RMSNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, dropout_p=0.0, device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.eps = eps
        if dropout_p > 0.0:
            self.drop = torch.nn.Dropout(dropout_p)
        else:
            self.drop = None
        self.weight = torch.nn.Parameter(torch.empty(hidden_size, **factory_kwargs))
        self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps,
                           dropout_p=self.drop.p if self.drop is not None and self.training else 0.0,
                           prenorm=prenorm, residual_in_fp32=residual_in_fp32)
