"""causal_conv1d_fn / causal_conv1d_update with the causal-conv1d==1.4.0 signatures
(causal_conv1d/causal_conv1d_interface.py upstream; pinned by /root/reference/requirements.txt:12)."""
from __future__ import annotations

import torch

from .. import _cabi as abi


def _act(activation):
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    return abi.ACT_NONE if activation is None else abi.ACT_SILU


def conv1d_fwd_raw(x, weight, bias, seq_idx, initial_states, out, final_states, activation):
    p = abi.Conv1dFwd()
    p.x, p.weight, p.bias = abi.tdesc(x), abi.tdesc(weight), abi.tdesc(bias)
    p.seq_idx, p.initial_states = abi.tdesc(seq_idx), abi.tdesc(initial_states)
    p.out, p.final_states = abi.tdesc(out), abi.tdesc(final_states)
    p.activation = activation
    abi.call("omni_causal_conv1d_fwd", p, x.device)


def conv1d_bwd_raw(x, weight, bias, dout, seq_idx, initial_states, dx, dweight, dbias, dinitial_states, activation):
    p = abi.Conv1dBwd()
    p.x, p.weight, p.bias, p.dout = abi.tdesc(x), abi.tdesc(weight), abi.tdesc(bias), abi.tdesc(dout)
    p.seq_idx, p.initial_states = abi.tdesc(seq_idx), abi.tdesc(initial_states)
    p.dx, p.dweight, p.dbias = abi.tdesc(dx), abi.tdesc(dweight), abi.tdesc(dbias)
    p.dinitial_states = abi.tdesc(dinitial_states)
    p.activation = activation
    abi.call("omni_causal_conv1d_bwd", p, x.device)


class CausalConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                final_states_out=None, activation=None):
        act = _act(activation)
        if x.dim() != 3 or weight.dim() != 2:
            raise ValueError("x must be (batch, dim, seqlen) and weight (dim, width)")
        batch, dim, seqlen = x.shape
        width = weight.shape[1]
        if seq_idx is not None:
            assert initial_states is None, "initial_states must be None if seq_idx is not None"
            assert not return_final_states, "If seq_idx is not None, we don't return final_states_out"
            seq_idx = seq_idx.to(torch.int32)
        if return_final_states:
            assert x.stride(1) == 1, "Only channel-last layout support returning final_states_out"
            if final_states_out is not None:
                assert final_states_out.shape == (batch, dim, width - 1)
            else:
                final_states_out = torch.empty(batch, width - 1, dim, device=x.device, dtype=x.dtype).transpose(1, 2)
        else:
            final_states_out = None
        out = torch.empty_like(x)
        conv1d_fwd_raw(x, weight, bias, seq_idx, initial_states, out, final_states_out, act)
        ctx.save_for_backward(x, weight, bias, seq_idx, initial_states)
        ctx.act = act
        ctx.return_final_states = return_final_states
        ctx.return_dinitial_states = initial_states is not None and initial_states.requires_grad
        if return_final_states:
            ctx.mark_non_differentiable(final_states_out)
            return out, final_states_out
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        x, weight, bias, seq_idx, initial_states = ctx.saved_tensors
        if dout.stride(2) != 1 and dout.stride(1) != 1:
            dout = dout.contiguous()
        dx = torch.empty_like(x)
        dweight = torch.zeros(weight.shape, device=x.device, dtype=torch.float32)
        dbias = torch.zeros(weight.shape[0], device=x.device, dtype=torch.float32) if bias is not None else None
        dinit = torch.empty_like(initial_states) if ctx.return_dinitial_states else None
        conv1d_bwd_raw(x, weight, bias, dout, seq_idx, initial_states, dx, dweight, dbias, dinit, ctx.act)
        return (dx, dweight.to(weight.dtype), dbias.to(bias.dtype) if bias is not None else None, None, dinit,
                None, None, None)


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    """x: (batch, dim, seqlen); weight: (dim, width); bias: (dim,); seq_idx: (batch, seqlen);
    initial_states / final_states_out: (batch, dim, width - 1); activation: None | "silu" | "swish".
    Returns out (batch, dim, seqlen) [, final_states_out]."""
    return CausalConv1dFn.apply(x, weight, bias, seq_idx, initial_states, return_final_states, final_states_out,
                                activation)


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None, cache_seqlens=None):
    """x: (batch, dim) or (batch, dim, seqlen); conv_state: (batch, dim, state_len >= width-1), updated in place;
    cache_seqlens: (batch,) int32 -> conv_state is a ring buffer written at cache_seqlens % state_len."""
    act = _act(activation)
    unsqueeze = x.dim() == 2
    if unsqueeze:
        x = x.unsqueeze(-1)
    out = torch.empty_like(x)
    p = abi.Conv1dUpdate()
    p.x, p.conv_state, p.weight, p.bias = abi.tdesc(x), abi.tdesc(conv_state), abi.tdesc(weight), abi.tdesc(bias)
    if cache_seqlens is not None:
        cache_seqlens = cache_seqlens.to(torch.int32).contiguous()
    p.cache_seqlens = abi.tdesc(cache_seqlens)
    p.out = abi.tdesc(out)
    p.activation = act
    abi.call("omni_causal_conv1d_update", p, x.device)
    return out.squeeze(-1) if unsqueeze else out
