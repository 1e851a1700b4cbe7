"""Head GEMM + cross-entropy: loss = CE(h W^T, labels) as the reference computes it with
img_head / lm_head followed by nn.CrossEntropyLoss on the shifted logits (/root/reference/models/mamba_vlm.py:96-100,
models/omnimamba.py:276-279) - evaluated in row blocks so that the (B L, vocab) logits never exist as one tensor in HBM."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _cabi as abi
from .gemm import _autocast_dtype, mm_nt


def _ce_stats(logits, labels, ignore_index):
    """(lse, per-row loss) of an fp32 logits block: one libomnissm kernel on the GPU."""
    if logits.is_cuda:
        M = logits.shape[0]
        lse = torch.empty(M, device=logits.device, dtype=torch.float32)
        loss = torch.empty(M, device=logits.device, dtype=torch.float32)
        p = abi.SoftmaxCe()
        p.logits, p.labels, p.lse, p.loss = abi.tdesc(logits), abi.tdesc(labels), abi.tdesc(lse), abi.tdesc(loss)
        p.ignore_index = ignore_index
        abi.call("omni_softmax_ce_fwd", p, logits.device)
        return lse, loss
    lse = torch.logsumexp(logits, dim=-1)
    ok = labels != ignore_index
    tgt = logits.gather(1, labels.clamp_min(0).unsqueeze(1)).squeeze(1)
    return lse, (lse - tgt) * ok


def _ce_grad(logits, labels, lse, scale, ignore_index, dtype):
    """(softmax - onehot) * scale as `dtype` (bf16 on the GPU: one libomnissm kernel)."""
    if logits.is_cuda and dtype == torch.bfloat16:
        g = torch.empty(logits.shape, device=logits.device, dtype=torch.bfloat16)
        p = abi.SoftmaxCe()
        p.logits, p.labels, p.lse, p.scale, p.grad = (abi.tdesc(t) for t in (logits, labels, lse, scale.reshape(1), g))
        p.ignore_index = ignore_index
        abi.call("omni_softmax_ce_bwd", p, logits.device)
        return g
    pr = torch.exp(logits - lse[:, None])
    ok = labels != ignore_index
    pr.scatter_add_(1, labels.clamp_min(0).unsqueeze(1), -ok.float().unsqueeze(1))
    pr *= (ok.float() * scale).unsqueeze(1)
    return pr.to(dtype)


class _LinearCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, w, labels, ignore_index, block):
        M = h.shape[0]
        n_valid = (labels != ignore_index).sum().clamp_min(1)
        loss = torch.zeros((), device=h.device, dtype=torch.float32)
        lse = torch.empty(M, device=h.device, dtype=torch.float32)
        for i in range(0, M, block):
            logits = mm_nt(h[i:i + block], w, torch.float32) if h.is_cuda else (h[i:i + block] @ w.t()).float()
            lse_b, loss_b = _ce_stats(logits, labels[i:i + block], ignore_index)
            lse[i:i + block] = lse_b
            loss += loss_b.sum()
        ctx.save_for_backward(h, w, labels, lse, n_valid)
        ctx.ignore_index, ctx.block = ignore_index, block
        return loss / n_valid

    @staticmethod
    def backward(ctx, dloss):
        h, w, labels, lse, n_valid = ctx.saved_tensors
        M, block = h.shape[0], ctx.block
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dw = torch.zeros(w.shape, device=w.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        scale = (dloss / n_valid).float()
        for i in range(0, M, block):
            hb = h[i:i + block]
            logits = mm_nt(hb, w, torch.float32) if h.is_cuda else (hb @ w.t()).float()
            g = _ce_grad(logits, labels[i:i + block], lse[i:i + block], scale, ctx.ignore_index, h.dtype)
            if dh is not None:
                dh[i:i + block] = mm_nt(g, w.t())
            if dw is not None:
                dw += mm_nt(g.t(), hb.t(), torch.float32)
        return dh, dw.to(w.dtype) if dw is not None else None, None, None, None


def linear_cross_entropy(h, weight, labels, ignore_index=-100, block=8192):
    """h (M, d), weight (V, d), labels (M,) -> mean CE over the labels != ignore_index (nn.CrossEntropyLoss default)."""
    ac = _autocast_dtype(h.device.type) if h.is_cuda else None
    if ac is not None:
        h, weight = h.to(ac), weight.to(ac)
    with torch.autocast(h.device.type, enabled=False):
        return _LinearCE.apply(h, weight, labels, ignore_index, block)
