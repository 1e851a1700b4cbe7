"""in_proj (+LoRA) / out_proj GEMMs of the Mamba-2 block.

`linear(x, w, bias)` is F.linear's contract (y = x w^T + bias) and `lora_linear` the reference's LoRA-wrapped in_proj
(/root/reference/models/stage2/lora.py:263-279: y = x W^T + B(A(dropout(x))) * alpha / r).  bf16 CUDA operands run on the
hand-written tcgen05 GEMM of libomnissm (csrc/gemm_tc.cu) - forward, dgrad and wgrad are the same kernel with different
operand majors; anything else (fp32 parameters outside autocast, CPU tensors in the host-logic tests) goes to torch."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from .. import _cabi as abi

_BACKEND = os.environ.get("OMNI_GEMM", "tc")  # "tc" (libomnissm tcgen05 kernel) | "torch" (cuBLAS, for A/B comparisons)


def _autocast_dtype(device_type="cuda"):
    return torch.get_autocast_dtype(device_type) if torch.is_autocast_enabled(device_type) else None


_CAST_CACHE = {}


def cast_param(w, dtype):
    """w.to(dtype) for a GEMM operand.  FROZEN parameters (stage "align": every base weight of the backbone) keep their
    low-precision copy across calls - autocast's own weight cache does the same inside one context, but the custom GEMM path
    casts by hand, and re-casting 26 M weights per layer and step is 4 launches and ~100 MB of traffic for nothing.
    Keyed on the parameter object and its version counter, so an in-place update (optimizer, load_state_dict) refreshes it."""
    if w is None or w.dtype == dtype:
        return w
    if w.requires_grad or not isinstance(w, torch.nn.Parameter):
        return w.to(dtype)
    key = (id(w), dtype)
    hit = _CAST_CACHE.get(key)
    if hit is not None and hit[0] is w and hit[1] == w._version and hit[2].device == w.device:
        return hit[2]
    c = w.detach().to(dtype)
    _CAST_CACHE[key] = (w, w._version, c)
    return c


def gemm_tc_available() -> bool:
    return _BACKEND == "tc" and hasattr(abi, "gemm") and abi.gemm_supported()


def _eligible(*ts):
    return gemm_tc_available() and all(t.is_cuda and t.dtype == torch.bfloat16 for t in ts)


def mm_nt(a, b, out_dtype=None, a2=None, b2=None):
    """a (M, K) @ b (N, K)^T [+ a2 (M, K2) @ b2 (N, K2)^T] -> (M, N).  Operands may be transposed views (either dim
    contiguous): the kernel takes K-major and MN-major tiles alike."""
    od = out_dtype or a.dtype
    if _eligible(a, b) and abi.gemm_operands_ok(a, b, a2, b2) and (b.shape[0] * (4 if od == torch.float32 else 2)) % 16 == 0 \
            and od in (torch.bfloat16, torch.float32):
        return abi.gemm(a, b, od, a2, b2)
    r = a @ b.t()
    if a2 is not None:
        r = r + a2 @ b2.t()
    return r if out_dtype is None else r.to(out_dtype)


def _small_t(t2):
    """(M, r) -> its transpose (r, M) as a K-major operand whose row pitch is a multiple of 8 elements (TMA: 16-byte strides)."""
    M, r = t2.shape
    buf = t2.new_empty(r, (M + 7) // 8 * 8)
    buf[:, :M] = t2.t()
    return buf[:, :M]


class _LinearFn(torch.autograd.Function):
    """y = x w^T (+ x2 w2^T) (+ bias); forward, dgrad and wgrad all run on the same GEMM kernel."""

    @staticmethod
    def forward(ctx, x, w, bias, x2=None, w2=None):
        ctx.save_for_backward(x, w, x2, w2)
        ctx.has_bias = bias is not None
        x2f = x2.reshape(-1, x2.shape[-1]) if x2 is not None else None
        y = mm_nt(x.reshape(-1, x.shape[-1]), w, None, x2f, w2)
        if bias is not None:
            y = y + bias
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, w, x2, w2 = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.stride(-1) != 1 and dy2.stride(0) != 1:
            dy2 = dy2.contiguous()
        need = ctx.needs_input_grad
        dx = dw = db = dx2 = dw2 = None
        if need[0]:
            dx = mm_nt(dy2, w.t()).view(x.shape)                                   # (M, N) @ (N, K)
        if need[1]:
            dw = mm_nt(dy2.t(), x.reshape(-1, x.shape[-1]).t())                    # (N, M) @ (M, K)
        if ctx.has_bias and need[2]:
            db = dy2.sum(0)
        # (rank-r operands: their transposed views have 16-byte rows; contiguous copies are a few hundred KB and let the
        # GEMM use its 128 x 16 tile instead of a 256-wide one)
        if x2 is not None and need[3]:
            dx2 = mm_nt(dy2, w2.t().contiguous()).view(x2.shape)
        if w2 is not None and need[4]:
            dw2 = mm_nt(dy2.t(), _small_t(x2.reshape(-1, x2.shape[-1])))
        return dx, dw, db, dx2, dw2


def _f32_decode(x, w, x2=None, w2=None):
    """fp32 operands at decode shapes (a handful of rows, no gradient: Mamba2.step of a model that runs without autocast, as
    inference_t2i.py does) -> the 3xTF32 weight-streaming kernel; None when it does not apply."""
    if _BACKEND != "tc" or not x.is_cuda or x.dtype != torch.float32 or w.dtype != torch.float32:
        return None
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (x2 is not None and (x2.requires_grad or w2.requires_grad))):
        return None
    x2d = x.reshape(-1, x.shape[-1])
    x22 = x2.reshape(-1, x2.shape[-1]) if x2 is not None else None
    if not abi.gemm_f32_decode_ok(x2d, w, x22, w2):
        return None
    return abi.gemm_f32_decode(x2d, w, x22, w2).view(*x.shape[:-1], w.shape[0])


def linear(x, w, bias=None):
    """F.linear with autocast semantics; bf16 CUDA operands run on the tcgen05 GEMM."""
    ac = _autocast_dtype(x.device.type) if x.is_cuda else None
    if ac is not None:
        x, w = x.to(ac), cast_param(w, ac)
        bias = bias.to(ac) if bias is not None else None
    if not _eligible(x, w):
        y = _f32_decode(x, w)
        if y is not None:
            return y if bias is None else y + bias
        return F.linear(x, w, bias)
    with torch.autocast(x.device.type, enabled=False):
        return _LinearFn.apply(x, w, bias)


def lora_linear(x, w, bias, lora_a, lora_b, scaling, dropout=None):
    """x W^T + (dropout(x) A^T) B^T * scaling  (lora.py:263-279).  The rank-r down projection is a skinny GEMM; the up
    projection rides in the accumulator of the base GEMM as a second operand pair (A2 = s x A^T, B2 = B)."""
    xd = dropout(x) if dropout is not None else x
    ac = _autocast_dtype(x.device.type) if x.is_cuda else None
    if ac is not None:
        x, xd = x.to(ac), xd.to(ac)
        w, lora_a, lora_b = cast_param(w, ac), cast_param(lora_a, ac), cast_param(lora_b, ac)
        bias = bias.to(ac) if bias is not None else None
    if not _eligible(x, w, lora_a, lora_b):
        if x.is_cuda and x.dtype == torch.float32 and lora_a.shape[0] % 4 == 0:
            t = F.linear(xd, lora_a) * scaling                       # (.., r): a few hundred kiloflops
            y = _f32_decode(x, w, t, lora_b)
            if y is not None:
                return y if bias is None else y + bias
        return F.linear(x, w, bias) + F.linear(F.linear(xd, lora_a), lora_b) * scaling
    with torch.autocast(x.device.type, enabled=False):
        t = _LinearFn.apply(xd, lora_a, None) * scaling         # (.., r)
        return _LinearFn.apply(x, w, bias, t, lora_b)


class Linear(torch.nn.Linear):
    """nn.Linear whose forward runs on the tcgen05 GEMM (still an nn.Linear: the reference's LoRA wrapper finds and
    replaces `in_proj` by isinstance, /root/reference/models/stage2/lora.py:76-106)."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)
