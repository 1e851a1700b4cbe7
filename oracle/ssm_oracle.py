"""Pure-PyTorch CPU restatement of the Mamba-2 / Mamba-1 hot-path arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function names the
upstream reference function whose published algorithm it restates; the
call-sites in the reference that reach it are given as /root/reference
file:line.  All math runs in ``compute_dtype`` (fp32 by default, fp64 for
gradient checks); only the final result is cast back to the input dtype, which
is where upstream casts too.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "softplus_thresholded",
    "dt_transform",
    "causal_conv1d_ref",
    "causal_conv1d_update_ref",
    "ssd_recurrent_ref",
    "ssd_chunked_ref",
    "mamba_chunk_scan_combined_ref",
    "selective_state_update_ref",
    "rmsnorm_gated_ref",
    "layer_norm_ref",
    "selective_scan_ref",
    "mamba_split_conv1d_scan_combined_ref",
    "Mamba2Params",
    "mamba2_init_params",
    "mamba2_forward_ref",
    "mamba2_step_ref",
    "mixer_stack_ref",
    "cpu_recurrent_baseline",
]


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def softplus_thresholded(v: torch.Tensor) -> torch.Tensor:
    """softplus with the upstream kernels' cut-over: log1p(exp(v)) for v <= 20, v above.

    Restates the `dt = where(dt <= 20, softplus(dt), dt)` step of mamba_ssm's
    `_chunk_cumsum_fwd_kernel` / `selective_scan_fwd_kernel`.
    """
    return torch.where(v <= 20.0, torch.log1p(torch.exp(torch.clamp(v, max=20.0))), v)


def dt_transform(dt, dt_bias=None, dt_softplus=False, dt_limit=(0.0, float("inf"))):
    """dt' = clamp(softplus(dt + dt_bias), dt_min, dt_max) (SURVEY.md A.3 line 1).

    ``dt``: (..., H); ``dt_bias``: (H,).
    """
    if dt_bias is not None:
        dt = dt + dt_bias.to(dt.dtype)
    if dt_softplus:
        dt = softplus_thresholded(dt)
    lo, hi = dt_limit
    if lo != 0.0 or hi != float("inf"):
        dt = torch.clamp(dt, min=lo, max=hi)
    return dt


def _silu(v):
    return v * torch.sigmoid(v)


# --------------------------------------------------------------------------- #
# causal depthwise conv  (causal-conv1d 1.4.0: causal_conv1d_ref / _update_ref)
# --------------------------------------------------------------------------- #
def causal_conv1d_ref(
    x: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    initial_states: Optional[torch.Tensor] = None,
    return_final_states: bool = False,
    activation: Optional[str] = None,
    compute_dtype: Optional[torch.dtype] = None,
):
    """out[b,d,t] = act(bias[d] + sum_w weight[d,w] * xpad[b,d,t+w]), xpad = [init | x].

    x: (B, D, L); weight: (D, W); bias: (D,); initial_states: (B, D, W-1).
    Reached from Mamba2.forward paths A/B (caller
    /root/reference/models/stage2/block.py:117).  Written as an explicit tap sum
    (not F.conv1d) so the restatement is independent of the library conv.
    """
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    dtype_in = x.dtype
    cd = compute_dtype or weight.dtype
    B_, D_, L_ = x.shape
    W_ = weight.shape[1]
    xf = x.to(cd)
    if initial_states is None:
        left = xf.new_zeros(B_, D_, W_ - 1)
    else:
        left = initial_states.to(cd)
    xpad = torch.cat([left, xf], dim=-1)  # (B, D, L + W - 1)
    wf = weight.to(cd)
    out = xf.new_zeros(B_, D_, L_)
    for w in range(W_):
        out = out + wf[:, w].view(1, D_, 1) * xpad[:, :, w : w + L_]
    if bias is not None:
        out = out + bias.to(cd).view(1, D_, 1)
    if activation is not None:
        out = _silu(out)
    out = out.to(dtype_in)
    if return_final_states:
        # last W-1 inputs (including initial states when L < W-1), zero-padded on the left
        final = xpad[:, :, xpad.shape[-1] - (W_ - 1) :].to(dtype_in)
        return out, final
    return out


def causal_conv1d_update_ref(
    x: torch.Tensor,
    conv_state: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    activation: Optional[str] = None,
    cache_seqlens: Optional[torch.Tensor] = None,
):
    """Single/multi-token decode update; mutates ``conv_state`` in place.

    x: (B, D) or (B, D, T); conv_state: (B, D, S) with S >= W-1; weight: (D, W).
    Without ``cache_seqlens`` the state is a shift register holding the last S
    inputs; with it the state is a ring buffer written at cache_seqlens % S.
    (causal_conv1d_update_ref; Mamba2.step, SURVEY.md 3.2 path C.)
    """
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    dtype_in = x.dtype
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(-1)
    B_, D_, T_ = x.shape
    W_ = weight.shape[1]
    S_ = conv_state.shape[-1]
    assert conv_state.shape == (B_, D_, S_) and S_ >= W_ - 1
    cd = weight.dtype
    if cache_seqlens is None:
        window = torch.cat([conv_state, x.to(conv_state.dtype)], dim=-1).to(cd)  # (B, D, S+T)
        conv_state.copy_(window[:, :, -S_:])
    else:
        pos = cache_seqlens.long().view(B_, 1)
        back = torch.arange(-(W_ - 1), 0, device=x.device).view(1, -1) + pos  # (B, W-1)
        back = torch.remainder(back, S_).unsqueeze(1).expand(-1, D_, -1)
        hist = conv_state.gather(2, back)
        window = torch.cat([hist, x.to(conv_state.dtype)], dim=-1).to(cd)
        dst = torch.remainder(torch.arange(T_, device=x.device).view(1, -1) + pos, S_)
        conv_state.scatter_(2, dst.unsqueeze(1).expand(-1, D_, -1), x.to(conv_state.dtype))
    wf = weight.to(cd)
    Lw = window.shape[-1]
    out = window.new_zeros(B_, D_, T_)
    for t in range(T_):
        end = Lw - (T_ - 1 - t)
        out[:, :, t] = (window[:, :, end - W_ : end] * wf.unsqueeze(0)).sum(-1)
    if bias is not None:
        out = out + bias.to(cd).view(1, D_, 1)
    if activation is not None:
        out = _silu(out)
    out = out.to(dtype_in)
    return out.squeeze(-1) if squeeze else out


# --------------------------------------------------------------------------- #
# SSD scan (Mamba-2 head form)
# --------------------------------------------------------------------------- #
def _expand_groups(t: torch.Tensor, H: int) -> torch.Tensor:
    """(B, L, G, N) -> (B, L, H, N) with head h using group h // (H/G)."""
    G = t.shape[2]
    return t.repeat_interleave(H // G, dim=2)


def ssd_recurrent_ref(x, dt, A, B, C, D=None, initial_states=None, seq_idx=None, compute_dtype=torch.float32):
    """Ground-truth token-by-token recurrence:

        S_t = exp(dt_t A) S_{t-1} + dt_t x_t (x) B_t ;  y_t = S_t C_t + D x_t

    x: (B, L, H, P); dt: (B, L, H) *already transformed*; A: (H,); B, C: (B, L, G, N);
    D: (H,) or (H, P).  Returns (y fp, final_states (B, H, P, N)) in compute_dtype.
    This loop is also "the reference's pure-PyTorch recurrent fallback"
    (selective_scan_ref-style) that bench.py times on the host cores.
    """
    cd = compute_dtype
    Bsz, L, H, P = x.shape
    N = B.shape[-1]
    xf, dtf, Af = x.to(cd), dt.to(cd), A.to(cd)
    Bf, Cf = _expand_groups(B.to(cd), H), _expand_groups(C.to(cd), H)
    S = xf.new_zeros(Bsz, H, P, N) if initial_states is None else initial_states.to(cd).clone()
    ys = []
    for t in range(L):
        decay = torch.exp(dtf[:, t] * Af)  # (B, H)
        if seq_idx is not None and t > 0:
            same = (seq_idx[:, t] == seq_idx[:, t - 1]).to(cd).view(Bsz, 1)
            decay = decay * same
        S = S * decay[:, :, None, None] + (dtf[:, t, :, None] * xf[:, t])[..., None] * Bf[:, t, :, None, :]
        ys.append(torch.einsum("bhpn,bhn->bhp", S, Cf[:, t]))
    y = torch.stack(ys, dim=1) if L > 0 else xf.new_zeros(Bsz, 0, H, P)
    if D is not None:
        Df = D.to(cd)
        y = y + xf * (Df.view(1, 1, H, -1) if Df.dim() == 2 else Df.view(1, 1, H, 1))
    return y, S


def _segsum(a: torch.Tensor) -> torch.Tensor:
    """out[..., i, j] = sum_{j < k <= i} a[..., k] for i >= j, -inf above the diagonal."""
    T = a.shape[-1]
    cs = torch.cumsum(a, dim=-1)
    diff = cs[..., :, None] - cs[..., None, :]
    mask = torch.tril(torch.ones(T, T, dtype=torch.bool, device=a.device))
    return diff.masked_fill(~mask, float("-inf"))


def ssd_chunked_ref(x, dt, A, B, C, chunk_size, D=None, initial_states=None, compute_dtype=torch.float32):
    """Chunked (segsum) evaluation of the same recurrence - the identity every fast
    kernel uses (ssd_minimal_discrete; SURVEY.md A.3 "chunked identity").  Same
    argument conventions as :func:`ssd_recurrent_ref`.
    """
    cd = compute_dtype
    Bsz, L, H, P = x.shape
    N = B.shape[-1]
    Q = chunk_size
    pad = (Q - L % Q) % Q
    xf, dtf = x.to(cd), dt.to(cd)
    Bf, Cf = _expand_groups(B.to(cd), H), _expand_groups(C.to(cd), H)
    if pad:
        xf = F.pad(xf, (0, 0, 0, 0, 0, pad))
        dtf = F.pad(dtf, (0, 0, 0, pad))  # dt = 0 => decay 1, no input: tail is inert
        Bf = F.pad(Bf, (0, 0, 0, 0, 0, pad))
        Cf = F.pad(Cf, (0, 0, 0, 0, 0, pad))
    nC = xf.shape[1] // Q
    xc = xf.view(Bsz, nC, Q, H, P)
    dtc = dtf.view(Bsz, nC, Q, H)
    Bc = Bf.view(Bsz, nC, Q, H, N)
    Cc = Cf.view(Bsz, nC, Q, H, N)
    a = (dtc * A.to(cd).view(1, 1, 1, H)).permute(0, 3, 1, 2)  # (B, H, nC, Q)
    lam = torch.cumsum(a, dim=-1)
    # intra-chunk
    Lmat = torch.exp(_segsum(a))  # (B, H, nC, Q, Q)
    xdt = xc * dtc[..., None]
    y_diag = torch.einsum("bcihn,bcjhn,bhcij,bcjhp->bcihp", Cc, Bc, Lmat, xdt)
    # per-chunk states
    decay_to_end = torch.exp(lam[..., -1:] - lam)  # (B, H, nC, Q)
    chunk_states = torch.einsum("bcjhn,bhcj,bcjhp->bchpn", Bc, decay_to_end, xdt)
    # inter-chunk pass
    S = xf.new_zeros(Bsz, H, P, N) if initial_states is None else initial_states.to(cd).clone()
    entering = []
    for c in range(nC):
        entering.append(S)
        S = S * torch.exp(lam[:, :, c, -1])[:, :, None, None] + chunk_states[:, c]
    S_in = torch.stack(entering, dim=1) if nC > 0 else xf.new_zeros(Bsz, 0, H, P, N)
    y_off = torch.einsum("bcihn,bchpn,bhci->bcihp", Cc, S_in, torch.exp(lam))
    y = (y_diag + y_off).reshape(Bsz, nC * Q, H, P)[:, :L]
    if D is not None:
        Df = D.to(cd)
        y = y + x.to(cd) * (Df.view(1, 1, H, -1) if Df.dim() == 2 else Df.view(1, 1, H, 1))
    return y, S


def mamba_chunk_scan_combined_ref(
    x, dt, A, B, C, chunk_size=None, D=None, z=None, dt_bias=None, initial_states=None,
    seq_idx=None, dt_softplus=False, dt_limit=(0.0, float("inf")), return_final_states=False,
    compute_dtype=torch.float32,
):
    """Oracle for `mamba_chunk_scan_combined` (mamba_ssm/ops/triton/ssd_combined.py;
    SURVEY.md 8(a) row a4).  Uses the *recurrent* definition (chunk_size does not
    change the mathematical result).  y cast to x.dtype; final_states stay in
    compute_dtype (upstream: fp32)."""
    dtt = dt_transform(dt.to(compute_dtype), None if dt_bias is None else dt_bias.to(compute_dtype),
                       dt_softplus, dt_limit)
    y, S = ssd_recurrent_ref(x, dtt, A, B, C, D=D, initial_states=initial_states, seq_idx=seq_idx,
                             compute_dtype=compute_dtype)
    if z is not None:
        y = y * _silu(z.to(compute_dtype))
    y = y.to(x.dtype)
    return (y, S) if return_final_states else y


# --------------------------------------------------------------------------- #
# single-token state update (selective_state_update_ref)
# --------------------------------------------------------------------------- #
def selective_state_update_ref(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False):
    """One recurrence step on a persistent state, mutated in place.

    state: (B, dim, N) or (B, H, P, N); x, dt: (B, dim) | (B, H, P); A: (dim, N) | (H, P, N);
    B, C: (B, N) | (B, G, N); D, dt_bias: (dim,) | (H, P); z like x.  Returns out like x.
    """
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
    batch, H, P, N = state.shape
    G = B.shape[1]
    dtf = dt.float()
    if dt_bias is not None:
        dtf = dtf + dt_bias.float()
    if dt_softplus:
        dtf = softplus_thresholded(dtf)
    dA = torch.exp(dtf.unsqueeze(-1) * A.float())  # (B, H, P, N)
    Bh = B.float().repeat_interleave(H // G, dim=1)  # (B, H, N)
    Ch = C.float().repeat_interleave(H // G, dim=1)
    dB = dtf.unsqueeze(-1) * Bh.unsqueeze(2)  # (B, H, P, N)
    new_state = state.float() * dA + dB * x.float().unsqueeze(-1)
    state.copy_(new_state.to(state.dtype))
    out = torch.einsum("bhpn,bhn->bhp", new_state, Ch)
    if D is not None:
        out = out + x.float() * D.float()
    if z is not None:
        out = out * _silu(z.float())
    out = out.to(x.dtype)
    return out if has_heads else out.squeeze(1)


# --------------------------------------------------------------------------- #
# norms
# --------------------------------------------------------------------------- #
def rmsnorm_gated_ref(x, weight, bias=None, z=None, eps=1e-6, group_size=None, norm_before_gate=True,
                      is_rms_norm=True, upcast=True):
    """Gated (group) RMSNorm / LayerNorm (rms_norm_ref in layernorm_gated.py)."""
    dtype = x.dtype
    wf = weight.float()
    bf = bias.float() if bias is not None else None
    xf = x.float() if upcast else x
    zf = z.float() if (z is not None and upcast) else z
    if zf is not None and not norm_before_gate:
        xf = xf * _silu(zf)
    Dm = xf.shape[-1]
    gs = Dm if group_size is None else group_size
    xg = xf.reshape(*xf.shape[:-1], Dm // gs, gs)
    if is_rms_norm:
        rstd = torch.rsqrt(xg.square().mean(dim=-1, keepdim=True) + eps)
        xn = (xg * rstd).reshape(xf.shape)
    else:
        mu = xg.mean(dim=-1, keepdim=True)
        var = (xg - mu).square().mean(dim=-1, keepdim=True)
        xn = ((xg - mu) * torch.rsqrt(var + eps)).reshape(xf.shape)
    out = xn * wf
    if bf is not None:
        out = out + bf
    if zf is not None and norm_before_gate:
        out = out * _silu(zf)
    return out.to(dtype)


def layer_norm_ref(x, weight, bias=None, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                   is_rms_norm=False):
    """Fused residual-add + (RMS|Layer)Norm (layer_norm_ref / rms_norm_ref in
    mamba_ssm/ops/triton/layer_norm.py; call sites /root/reference/models/stage2/block.py:86-95,
    mixer_seq_simple.py:428-437).  Output dtype = x.dtype; the returned
    residual is fp32 if ``residual_in_fp32`` (or if the incoming residual is fp32)."""
    dtype = x.dtype
    res = x.float()
    if residual is not None:
        res = res + residual.float()
    if is_rms_norm:
        y = res * torch.rsqrt(res.square().mean(dim=-1, keepdim=True) + eps) * weight.float()
    else:
        y = F.layer_norm(res, res.shape[-1:], weight.float(), None, eps)
    if bias is not None:
        y = y + bias.float()
    y = y.to(dtype)
    if not prenorm:
        return y
    if residual_in_fp32 or (residual is not None and residual.dtype == torch.float32):
        res_out = res
    else:
        res_out = res.to(residual.dtype if residual is not None else dtype)
    return y, res_out


# --------------------------------------------------------------------------- #
# Mamba-1 selective scan (selective_scan_ref)
# --------------------------------------------------------------------------- #
def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False, compute_dtype=torch.float32):
    """u, delta, z: (B, D, L); A: (D, N) real; B, C: (D, N) | (B, N, L) | (B, G, N, L); D, delta_bias: (D,).
    Sequential fp32 recurrence x_t = exp(d_t A) x_{t-1} + d_t B_t u_t; y_t = <x_t, C_t> (+ D u) (* silu(z)).
    (mamba_ssm/ops/selective_scan_interface.py::selective_scan_ref)"""
    cd = compute_dtype
    dtype_in = u.dtype
    uf, df = u.to(cd), delta.to(cd)
    if delta_bias is not None:
        df = df + delta_bias.to(cd).view(1, -1, 1)
    if delta_softplus:
        df = softplus_thresholded(df)
    Bsz, Dm, L = uf.shape
    N = A.shape[1]
    Af = A.to(cd)

    def per_token(M):  # -> (B, D, N, L)
        Mf = M.to(cd)
        if Mf.dim() == 2:
            return Mf.view(1, Dm, N, 1).expand(Bsz, Dm, N, L)
        if Mf.dim() == 3:
            return Mf.view(Bsz, 1, N, L).expand(Bsz, Dm, N, L)
        G = Mf.shape[1]
        return Mf.repeat_interleave(Dm // G, dim=1)

    Bt, Ct = per_token(B), per_token(C)
    xs = uf.new_zeros(Bsz, Dm, N)
    ys = []
    for t in range(L):
        xs = torch.exp(df[:, :, t, None] * Af) * xs + df[:, :, t, None] * Bt[:, :, :, t] * uf[:, :, t, None]
        ys.append((xs * Ct[:, :, :, t]).sum(-1))
    y = torch.stack(ys, dim=2) if L > 0 else uf.new_zeros(Bsz, Dm, 0)
    if D is not None:
        y = y + uf * D.to(cd).view(1, -1, 1)
    if z is not None:
        y = y * _silu(z.to(cd))
    y = y.to(dtype_in)
    return (y, xs) if return_last_state else y


# --------------------------------------------------------------------------- #
# fused path-A op and the Mamba2 block (paths A / B / C)
# --------------------------------------------------------------------------- #
def mamba_split_conv1d_scan_combined_ref(
    zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size, initial_states=None, seq_idx=None,
    dt_limit=(0.0, float("inf")), return_final_states=False, activation="silu", rmsnorm_weight=None,
    rmsnorm_eps=1e-6, outproj_weight=None, outproj_bias=None, headdim=None, ngroups=1, norm_before_gate=True,
    compute_dtype=torch.float32,
):
    """Oracle for the fused training op (mamba_split_conv1d_scan_combined; SURVEY.md 3.2 path A,
    8(a) row a2): split -> causal conv (+SiLU) -> SSD -> gated RMSNorm -> out_proj.
    conv1d_weight: (conv_dim, W)."""
    if D.dim() == 1:
        assert headdim is not None
        (H,) = D.shape
    else:
        H, headdim = D.shape
    Bsz, L, _ = zxbcdt.shape
    dim = H * headdim
    N = (conv1d_weight.shape[0] - dim) // ngroups // 2
    d_nonssm = (zxbcdt.shape[-1] - 2 * dim - 2 * ngroups * N - H) // 2
    assert d_nonssm == 0, "MLP-in-mixer split is not on the OmniMamba path"
    z, xBC, dt = torch.split(zxbcdt, [dim, dim + 2 * ngroups * N, H], dim=-1)
    xBC = causal_conv1d_ref(xBC.transpose(1, 2), conv1d_weight, conv1d_bias, activation=activation,
                            compute_dtype=compute_dtype).transpose(1, 2)
    x, Bm, Cm = torch.split(xBC, [dim, ngroups * N, ngroups * N], dim=-1)
    x = x.reshape(Bsz, L, H, headdim)
    Bm = Bm.reshape(Bsz, L, ngroups, N)
    Cm = Cm.reshape(Bsz, L, ngroups, N)
    zh = z.reshape(Bsz, L, H, headdim)
    y, S = mamba_chunk_scan_combined_ref(
        x, dt, A, Bm, Cm, chunk_size, D=D, z=zh if rmsnorm_weight is None else None, dt_bias=dt_bias,
        initial_states=initial_states, seq_idx=seq_idx, dt_softplus=True, dt_limit=dt_limit,
        return_final_states=True, compute_dtype=compute_dtype)
    y = y.reshape(Bsz, L, dim)
    if rmsnorm_weight is not None:
        y = rmsnorm_gated_ref(y, rmsnorm_weight, None, z=z, eps=rmsnorm_eps, group_size=dim // ngroups,
                              norm_before_gate=norm_before_gate)
    if outproj_weight is not None:
        y = F.linear(y.to(outproj_weight.dtype), outproj_weight, outproj_bias)
    return (y, S) if return_final_states else y


class Mamba2Params:
    """Plain container with the upstream `Mamba2` parameter names/shapes."""

    def __init__(self, d_model, d_state=128, d_conv=4, expand=2, headdim=64, ngroups=1, chunk_size=256,
                 rmsnorm_eps=1e-5):
        self.d_model, self.d_state, self.d_conv, self.headdim = d_model, d_state, d_conv, headdim
        self.ngroups, self.chunk_size, self.eps = ngroups, chunk_size, rmsnorm_eps
        self.d_inner = expand * d_model
        self.nheads = self.d_inner // headdim
        self.conv_dim = self.d_inner + 2 * ngroups * d_state
        self.d_in_proj = 2 * self.d_inner + 2 * ngroups * d_state + self.nheads
        self.in_proj_weight = None  # (d_in_proj, d_model)
        self.conv1d_weight = None   # (conv_dim, 1, d_conv)
        self.conv1d_bias = None     # (conv_dim,)
        self.dt_bias = None         # (nheads,)
        self.A_log = None           # (nheads,)
        self.D = None               # (nheads,)
        self.norm_weight = None     # (d_inner,)
        self.out_proj_weight = None  # (d_model, d_inner)


def mamba2_init_params(d_model, seed=0, dtype=torch.float32, **kw) -> Mamba2Params:
    """Mamba2.__init__ default initialisation (dt in [1e-3, 1e-1], A in [1, 16], D = 1)."""
    p = Mamba2Params(d_model, **kw)
    g = torch.Generator().manual_seed(seed)
    k_in = 1.0 / math.sqrt(d_model)
    p.in_proj_weight = ((torch.rand(p.d_in_proj, d_model, generator=g) * 2 - 1) * k_in).to(dtype)
    k_c = 1.0 / math.sqrt(p.d_conv)
    p.conv1d_weight = ((torch.rand(p.conv_dim, 1, p.d_conv, generator=g) * 2 - 1) * k_c).to(dtype)
    p.conv1d_bias = ((torch.rand(p.conv_dim, generator=g) * 2 - 1) * k_c).to(dtype)
    dt = torch.exp(torch.rand(p.nheads, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
    dt = torch.clamp(dt, min=1e-4)
    p.dt_bias = (dt + torch.log(-torch.expm1(-dt))).float()
    p.A_log = torch.log(torch.empty(p.nheads).uniform_(1, 16, generator=g)).float()
    p.D = torch.ones(p.nheads)
    p.norm_weight = torch.ones(p.d_inner, dtype=dtype)
    k_o = 1.0 / math.sqrt(p.d_inner)
    p.out_proj_weight = ((torch.rand(d_model, p.d_inner, generator=g) * 2 - 1) * k_o).to(dtype)
    return p


def mamba2_forward_ref(p: Mamba2Params, u: torch.Tensor, conv_state=None, ssm_state=None,
                       compute_dtype=torch.float32):
    """Mamba2.forward, paths A (no cache) and B (prefill: also fills conv_state/ssm_state in place).
    u: (B, L, d_model).  SURVEY.md Appendix A.1."""
    Bsz, L, _ = u.shape
    zxbcdt = F.linear(u, p.in_proj_weight.to(u.dtype))
    A = -torch.exp(p.A_log.float())
    if conv_state is not None:
        xBC = zxbcdt[..., p.d_inner : p.d_inner + p.conv_dim].transpose(1, 2)  # (B, conv_dim, L)
        Wd = conv_state.shape[-1]
        conv_state.copy_(F.pad(xBC, (Wd - L, 0)) if L < Wd else xBC[..., L - Wd :])
    out, S = mamba_split_conv1d_scan_combined_ref(
        zxbcdt, p.conv1d_weight.squeeze(1), p.conv1d_bias, p.dt_bias, A, p.D, p.chunk_size,
        return_final_states=True, activation="silu", rmsnorm_weight=p.norm_weight, rmsnorm_eps=p.eps,
        outproj_weight=p.out_proj_weight.to(u.dtype), headdim=p.headdim, ngroups=p.ngroups,
        norm_before_gate=False, compute_dtype=compute_dtype)
    if ssm_state is not None:
        ssm_state.copy_(S.to(ssm_state.dtype))
    return out


def mamba2_step_ref(p: Mamba2Params, u: torch.Tensor, conv_state: torch.Tensor, ssm_state: torch.Tensor):
    """Mamba2.step (path C): one token, u: (B, 1, d_model); caches mutated in place."""
    Bsz = u.shape[0]
    zxbcdt = F.linear(u.squeeze(1), p.in_proj_weight.to(u.dtype))
    z, xBC, dt = torch.split(zxbcdt, [p.d_inner, p.conv_dim, p.nheads], dim=-1)
    xBC = causal_conv1d_update_ref(xBC, conv_state, p.conv1d_weight.squeeze(1), p.conv1d_bias, "silu")
    x, Bm, Cm = torch.split(xBC, [p.d_inner, p.ngroups * p.d_state, p.ngroups * p.d_state], dim=-1)
    A = -torch.exp(p.A_log.float())
    H, P, N = p.nheads, p.headdim, p.d_state
    y = selective_state_update_ref(
        ssm_state, x.reshape(Bsz, H, P), dt.unsqueeze(-1).expand(Bsz, H, P),
        A.view(H, 1, 1).expand(H, P, N), Bm.reshape(Bsz, p.ngroups, N), Cm.reshape(Bsz, p.ngroups, N),
        D=p.D.view(H, 1).expand(H, P), z=None, dt_bias=p.dt_bias.view(H, 1).expand(H, P), dt_softplus=True)
    y = rmsnorm_gated_ref(y.reshape(Bsz, p.d_inner), p.norm_weight, None, z=z, eps=p.eps,
                          group_size=p.d_inner // p.ngroups, norm_before_gate=False)
    return F.linear(y, p.out_proj_weight.to(u.dtype)).unsqueeze(1)


def mixer_stack_ref(layers, norm_weights, norm_f_weight, hidden_states, eps=1e-5, residual_in_fp32=True,
                    lora=None, compute_dtype=torch.float32):
    """The layer loop of MixerModel.forward (/root/reference/models/stage2/mixer_seq_simple.py:404-437) over Block.forward
    (block.py:86-117, cond=None): per layer `hidden, residual = add_norm(hidden, residual); hidden = mixer(hidden)`, then the
    final add + norm (prenorm=False).  `layers`: Mamba2Params per layer; `norm_weights[i]`: the layer's RMSNorm weight;
    `lora`: optional per-layer (A (r, d), B (d_in_proj, r), scaling) added to in_proj as lora.py:263-279 does."""
    residual = None
    h = hidden_states
    for i, p in enumerate(layers):
        h, residual = layer_norm_ref(h, norm_weights[i], None, residual=residual, eps=eps, prenorm=True,
                                     residual_in_fp32=residual_in_fp32, is_rms_norm=True)
        if lora is not None and lora[i] is not None:
            la, lb, sc = lora[i]
            q = Mamba2Params.__new__(Mamba2Params)
            q.__dict__.update(p.__dict__)
            q.in_proj_weight = p.in_proj_weight + sc * (lb.to(p.in_proj_weight.dtype) @ la.to(p.in_proj_weight.dtype))
            p = q
        h = mamba2_forward_ref(p, h, compute_dtype=compute_dtype)
    return layer_norm_ref(h, norm_f_weight, None, residual=residual, eps=eps, prenorm=False,
                          residual_in_fp32=residual_in_fp32, is_rms_norm=True)


# --------------------------------------------------------------------------- #
# the CPU arm bench.py times ("pure-PyTorch recurrent fallback")
# --------------------------------------------------------------------------- #
def cpu_recurrent_baseline(x, dt, A, B, C, D, dt_bias):
    """fp32 token loop with all host threads; returns y.  Shapes as mamba_chunk_scan_combined."""
    y, _ = ssd_recurrent_ref(x, dt_transform(dt.float(), dt_bias, True), A, B, C, D=D)
    return y
