"""Single-token layer core of Mamba2.step (upstream mamba_ssm/modules/mamba2.py: step): causal_conv1d_update +
selective_state_update + gated RMSNorm as one libomnissm kernel (csrc/decode_core.cu), for the
OmniMamba geometry.  Reference call path: /root/reference/models/stage2/generation.py:383-431 -> Block -> Mamba2.step."""
from __future__ import annotations

import torch

from .. import _cabi as abi


def decode_core_supported(nheads, headdim, d_state, ngroups, d_conv) -> bool:
    return nheads == 64 and headdim == 64 and d_state == 128 and ngroups == 1 and 2 <= d_conv <= 4


def mamba2_decode_core(zxbcdt, conv_state, conv_weight, conv_bias, ssm_state, A, D, dt_bias, norm_weight, eps):
    """zxbcdt (B, 2 dim + 2 N + H); conv_state (B, dim + 2 N, W) and ssm_state (B, H, 64, 128) are updated in place.
    Returns rmsnorm(y * silu(z)) * norm_weight, (B, dim), dtype of zxbcdt."""
    B = zxbcdt.shape[0]
    dim = ssm_state.shape[1] * ssm_state.shape[2]
    out = torch.empty(B, dim, device=zxbcdt.device, dtype=zxbcdt.dtype)
    p = abi.DecodeCore()
    p.zxbcdt, p.conv_state, p.conv_weight, p.conv_bias = (abi.tdesc(t) for t in (zxbcdt, conv_state, conv_weight, conv_bias))
    p.ssm_state, p.A, p.D, p.dt_bias, p.norm_weight = (abi.tdesc(t) for t in (ssm_state, A, D, dt_bias, norm_weight))
    p.out = abi.tdesc(out)
    p.eps = float(eps)
    abi.call("omni_mamba2_decode_core", p, zxbcdt.device)
    return out
