"""Mamba2 mixer with the mamba_ssm==2.2.2 constructor, parameter names/shapes and forward/step/cache API
(mamba_ssm/modules/mamba2.py upstream).  The reference builds it in
/root/reference/models/stage2/mixer_seq_simple.py:194-205 and calls it from models/stage2/block.py:117,149-150;
pretrained `state-spaces/mamba2-1.3b` weights load by name (models/omnimamba.py:88-103), so the state_dict
contract of SURVEY.md Appendix C is kept exactly.

Three code paths:
  A  training / no cache  -> mamba_split_conv1d_scan_combined (one autograd node)
  B  prefill with cache   -> causal_conv1d_fn + mamba_chunk_scan_combined(return_final_states) + gated norm
  C  single-token decode  -> causal_conv1d_update + selective_state_update + gated norm (CUDA-graph capturable)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..interface.causal_conv1d import causal_conv1d_fn, causal_conv1d_update
from ..interface.decode import decode_core_supported, mamba2_decode_core
from ..interface.gemm import Linear
from ..interface.layernorm_gated import RMSNorm as RMSNormGated
from ..interface.selective_state_update import selective_state_update
from ..interface.ssd_combined import mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined


class Mamba2(nn.Module):
    def __init__(self, d_model, d_state=128, d_conv=4, conv_init=None, expand=2, headdim=64, d_ssm=None, ngroups=1,
                 A_init_range=(1, 16), D_has_hdim=False, rmsnorm=True, norm_before_gate=False, dt_min=0.001,
                 dt_max=0.1, dt_init_floor=1e-4, dt_limit=(0.0, float("inf")), bias=False, conv_bias=True,
                 chunk_size=256, use_mem_eff_path=True, layer_idx=None, process_group=None, sequence_parallel=True,
                 device=None, dtype=None):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        if process_group is not None:
            raise NotImplementedError("tensor/sequence-parallel Mamba2 is not used by OmniMamba (SURVEY.md 2.3)")
        self.d_model, self.d_state, self.d_conv, self.conv_init, self.expand = d_model, d_state, d_conv, conv_init, expand
        self.process_group, self.sequence_parallel, self.world_size, self.local_rank = None, sequence_parallel, 1, 0
        self.d_inner = self.expand * self.d_model
        self.headdim = headdim
        self.d_ssm = self.d_inner if d_ssm is None else d_ssm
        self.ngroups = ngroups
        assert self.d_ssm % self.headdim == 0
        self.nheads = self.d_ssm // self.headdim
        self.D_has_hdim, self.rmsnorm, self.norm_before_gate = D_has_hdim, rmsnorm, norm_before_gate
        self.dt_limit, self.activation, self.chunk_size = dt_limit, "silu", chunk_size
        self.use_mem_eff_path, self.layer_idx = use_mem_eff_path, layer_idx

        # order of the projection: [z, x, B, C, dt]
        d_in_proj = 2 * self.d_inner + 2 * self.ngroups * self.d_state + self.nheads
        self.in_proj = Linear(self.d_model, d_in_proj, bias=bias, **factory_kwargs)

        conv_dim = self.d_ssm + 2 * self.ngroups * self.d_state
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, bias=conv_bias, kernel_size=d_conv, groups=conv_dim,
                                padding=d_conv - 1, **factory_kwargs)
        if self.conv_init is not None:
            nn.init.uniform_(self.conv1d.weight, -self.conv_init, self.conv_init)
        self.act = nn.SiLU()

        # dt_bias = inverse-softplus of a log-uniform dt in [dt_min, dt_max]
        dt = torch.exp(torch.rand(self.nheads, **factory_kwargs) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))
        self.dt_bias = nn.Parameter(inv_dt)
        self.dt_bias._no_weight_decay = True

        assert A_init_range[0] > 0 and A_init_range[1] >= A_init_range[0]
        A = torch.empty(self.nheads, dtype=torch.float32, device=device).uniform_(*A_init_range)
        self.A_log = nn.Parameter(torch.log(A).to(dtype=dtype))
        self.A_log._no_weight_decay = True

        self.D = nn.Parameter(torch.ones(self.d_ssm if self.D_has_hdim else self.nheads, device=device))
        self.D._no_weight_decay = True

        if self.rmsnorm:
            self.norm = RMSNormGated(self.d_ssm, eps=1e-5, norm_before_gate=self.norm_before_gate,
                                     group_size=self.d_ssm // ngroups, **factory_kwargs)
        self.out_proj = Linear(self.d_inner, self.d_model, bias=bias, **factory_kwargs)

    # -- helpers -------------------------------------------------------------------------------------------
    def _A(self):
        """A = -exp(A_log) in fp32.  Outside autograd (decode, prefill) it is cached per version of A_log: recomputing it is
        three elementwise launches per layer and step inside the captured decode graph (144 of ~430 launches at 48 layers)."""
        if torch.is_grad_enabled() and self.A_log.requires_grad:
            return -torch.exp(self.A_log.float())
        c = getattr(self, "_A_cache", None)
        if c is None or c[0] != self.A_log._version or c[1].device != self.A_log.device or c[2] is not self.A_log:
            with torch.no_grad():
                c = (self.A_log._version, -torch.exp(self.A_log.float()), self.A_log)
            self._A_cache = c
        return c[1]

    def _D(self):
        return self.D.view(self.nheads, self.headdim) if self.D_has_hdim else self.D

    def _conv_w(self):
        return self.conv1d.weight.squeeze(1)

    # -- full-sequence forward ------------------------------------------------------------------------------
    def forward(self, u, seqlen=None, seq_idx=None, cu_seqlens=None, inference_params=None):
        """u: (batch, seqlen, d_model), or (batch*seqlen, d_model) with `seqlen` given.  Returns the same shape."""
        if cu_seqlens is not None:
            raise NotImplementedError("varlen (cu_seqlens) inputs are not on the OmniMamba path")
        seqlen_og = seqlen
        if seqlen is None:
            batch, seqlen, _ = u.shape
        else:
            batch = u.shape[0] // seqlen

        conv_state = ssm_state = None
        if inference_params is not None:
            conv_state, ssm_state = self._get_states_from_cache(inference_params, batch)
            if inference_params.seqlen_offset > 0:
                out, _, _ = self.step(u, conv_state, ssm_state)
                return out

        zxbcdt = self.in_proj(u)
        if seqlen_og is not None:
            zxbcdt = zxbcdt.view(batch, seqlen, zxbcdt.shape[-1])
        A = self._A()
        dt_limit_kwargs = {} if self.dt_limit == (0.0, float("inf")) else dict(dt_limit=self.dt_limit)
        d_mlp = (zxbcdt.shape[-1] - 2 * self.d_ssm - 2 * self.ngroups * self.d_state - self.nheads) // 2

        if self.use_mem_eff_path and inference_params is None and d_mlp == 0:
            out = mamba_split_conv1d_scan_combined(
                zxbcdt, self._conv_w(), self.conv1d.bias, self.dt_bias, A, D=self._D(), chunk_size=self.chunk_size,
                seq_idx=seq_idx, activation=self.activation,
                rmsnorm_weight=self.norm.weight if self.rmsnorm else None,
                rmsnorm_eps=self.norm.eps if self.rmsnorm else 1e-6,
                outproj_weight=self.out_proj.weight, outproj_bias=self.out_proj.bias,
                headdim=None if self.D_has_hdim else self.headdim, ngroups=self.ngroups,
                norm_before_gate=self.norm_before_gate, **dt_limit_kwargs)
            if seqlen_og is not None:
                out = out.reshape(batch * seqlen, out.shape[-1])
            return out

        z0, x0, z, xBC, dt = torch.split(
            zxbcdt, [d_mlp, d_mlp, self.d_ssm, self.d_ssm + 2 * self.ngroups * self.d_state, self.nheads], dim=-1)
        if conv_state is not None:
            # keep the last d_conv inputs (left-padded with zeros for short prompts)
            xBC_t = xBC.transpose(1, 2)
            conv_state.copy_(F.pad(xBC_t, (self.d_conv - xBC_t.shape[-1], 0)))
        xBC = causal_conv1d_fn(xBC.transpose(1, 2), self._conv_w(), bias=self.conv1d.bias, activation=self.activation,
                               seq_idx=seq_idx).transpose(1, 2)
        x, B, C = torch.split(xBC, [self.d_ssm, self.ngroups * self.d_state, self.ngroups * self.d_state], dim=-1)
        y = mamba_chunk_scan_combined(
            x.reshape(batch, seqlen, self.nheads, self.headdim), dt, A,
            B.reshape(batch, seqlen, self.ngroups, self.d_state), C.reshape(batch, seqlen, self.ngroups, self.d_state),
            chunk_size=self.chunk_size, D=self._D(),
            z=z.reshape(batch, seqlen, self.nheads, self.headdim) if not self.rmsnorm else None,
            dt_bias=self.dt_bias, dt_softplus=True, seq_idx=seq_idx, return_final_states=ssm_state is not None,
            **dt_limit_kwargs)
        if ssm_state is not None:
            y, last_state = y
            ssm_state.copy_(last_state)
        y = y.reshape(batch, seqlen, self.d_ssm)
        if self.rmsnorm:
            y = self.norm(y, z)
        if d_mlp > 0:
            y = torch.cat([F.silu(z0) * x0, y], dim=-1)
        if seqlen_og is not None:
            y = y.reshape(batch * seqlen, y.shape[-1])
        return self.out_proj(y)

    # -- single-token decode --------------------------------------------------------------------------------
    def step(self, hidden_states, conv_state, ssm_state):
        """hidden_states: (batch, 1, d_model).  conv_state (batch, conv_dim, d_conv) and ssm_state
        (batch, nheads, headdim, d_state) are updated in place.  Allocation pattern is static (graph capture)."""
        dtype = hidden_states.dtype
        assert hidden_states.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        zxbcdt = self.in_proj(hidden_states.squeeze(1))
        d_mlp = (zxbcdt.shape[-1] - 2 * self.d_ssm - 2 * self.ngroups * self.d_state - self.nheads) // 2
        if (d_mlp == 0 and self.rmsnorm and not self.norm_before_gate and not self.D_has_hdim and self.d_ssm == self.d_inner
                and decode_core_supported(self.nheads, self.headdim, self.d_state, self.ngroups, self.d_conv)
                and zxbcdt.stride(-1) == 1 and ssm_state.is_contiguous()):
            # conv update + state update + gated norm in ONE kernel (the OmniMamba geometry; csrc/decode_core.cu)
            y = mamba2_decode_core(zxbcdt, conv_state, self._conv_w(), self.conv1d.bias, ssm_state, self._A(), self.D,
                                   self.dt_bias, self.norm.weight, self.norm.eps)
            return self.out_proj(y).unsqueeze(1), conv_state, ssm_state
        z0, x0, z, xBC, dt = torch.split(
            zxbcdt, [d_mlp, d_mlp, self.d_ssm, self.d_ssm + 2 * self.ngroups * self.d_state, self.nheads], dim=-1)
        xBC = causal_conv1d_update(xBC, conv_state, self._conv_w(), self.conv1d.bias, self.activation)
        x, B, C = torch.split(xBC, [self.d_ssm, self.ngroups * self.d_state, self.ngroups * self.d_state], dim=-1)
        A = self._A()
        batch = x.shape[0]
        H, P, N, G = self.nheads, self.headdim, self.d_state, self.ngroups
        # stride-0 broadcasts, exactly what upstream passes (einops.repeat): the kernel sees tied A/dt/D
        A3 = A.view(H, 1, 1).expand(H, P, N).to(dtype=torch.float32)
        dt3 = dt.view(batch, H, 1).expand(batch, H, P)
        dt_bias2 = self.dt_bias.view(H, 1).expand(H, P)
        D2 = self.D.view(H, P) if self.D_has_hdim else self.D.view(H, 1).expand(H, P)
        y = selective_state_update(
            ssm_state, x.reshape(batch, H, P), dt3, A3, B.reshape(batch, G, N), C.reshape(batch, G, N), D2,
            z=z.reshape(batch, H, P) if not self.rmsnorm else None, dt_bias=dt_bias2, dt_softplus=True)
        y = y.reshape(batch, H * P)
        if self.rmsnorm:
            y = self.norm(y, z)
        if d_mlp > 0:
            y = torch.cat([F.silu(z0) * x0, y], dim=-1)
        out = self.out_proj(y)
        return out.unsqueeze(1), conv_state, ssm_state

    # -- caches ---------------------------------------------------------------------------------------------
    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        device = self.out_proj.weight.device
        conv_dtype = self.conv1d.weight.dtype if dtype is None else dtype
        conv_state = torch.zeros(batch_size, self.d_conv, self.conv1d.weight.shape[0], device=device,
                                 dtype=conv_dtype).transpose(1, 2)
        ssm_dtype = self.in_proj.weight.dtype if dtype is None else dtype
        ssm_state = torch.zeros(batch_size, self.nheads, self.headdim, self.d_state, device=device, dtype=ssm_dtype)
        return conv_state, ssm_state

    def _get_states_from_cache(self, inference_params, batch_size, initialize_states=False):
        assert self.layer_idx is not None
        if self.layer_idx not in inference_params.key_value_memory_dict:
            conv_state, ssm_state = self.allocate_inference_cache(batch_size, 0)
            inference_params.key_value_memory_dict[self.layer_idx] = (conv_state, ssm_state)
        else:
            conv_state, ssm_state = inference_params.key_value_memory_dict[self.layer_idx]
            if initialize_states:
                conv_state.zero_()
                ssm_state.zero_()
        return conv_state, ssm_state
