"""Deterministic inputs/parameters for the golden fixtures.

Everything is drawn from numpy's frozen legacy ``RandomState`` so the same
arrays are rebuilt bit-for-bit on any box; only the *outputs* of the
independent implementation are stored in the ``.npz`` fixtures.
"""
import math

import numpy as np
import torch

# name -> (d_model, batch, seqlen, chunk_size, seed)
BLOCK_CASES = {
    "cfg1_d256_L128": (256, 2, 128, 256, 0),      # BASELINE.json configs[0]
    "d64_L70_chunk32": (64, 3, 70, 32, 1),        # ragged multi-chunk tail
    "d128_L329_chunk64": (128, 1, 329, 64, 2),    # the stage-1 training length (72+256+1)
}
HEADDIM, D_STATE, D_CONV, NGROUPS, EPS = 64, 128, 4, 1, 1e-5


def block_params(d_model, seed):
    rs = np.random.RandomState(seed)
    d_inner = 2 * d_model
    H = d_inner // HEADDIM
    conv_dim = d_inner + 2 * NGROUPS * D_STATE
    d_in_proj = 2 * d_inner + 2 * NGROUPS * D_STATE + H
    f = lambda *s: torch.from_numpy(rs.uniform(-1.0, 1.0, size=s).astype(np.float32))
    dt = np.exp(rs.uniform(math.log(1e-3), math.log(1e-1), size=H)).clip(min=1e-4)
    return {
        "in_proj.weight": f(d_in_proj, d_model) / math.sqrt(d_model),
        "conv1d.weight": f(conv_dim, 1, D_CONV) / math.sqrt(D_CONV),
        "conv1d.bias": f(conv_dim) / math.sqrt(D_CONV),
        "dt_bias": torch.from_numpy((dt + np.log(-np.expm1(-dt))).astype(np.float32)),
        "A_log": torch.from_numpy(np.log(rs.uniform(1.0, 16.0, size=H)).astype(np.float32)),
        "D": torch.from_numpy(rs.uniform(0.5, 1.5, size=H).astype(np.float32)),
        "norm.weight": torch.from_numpy(rs.uniform(0.5, 1.5, size=d_inner).astype(np.float32)),
        "out_proj.weight": f(d_model, d_inner) / math.sqrt(d_inner),
    }


def block_input(d_model, batch, seqlen, seed):
    rs = np.random.RandomState(seed + 1000)
    return torch.from_numpy(rs.standard_normal((batch, seqlen, d_model)).astype(np.float32))


def scan_inputs(batch, seqlen, nheads, headdim, ngroups, dstate, seed, dtype=torch.float32):
    """Synthetic (x, dt_raw, A, B, C, D, dt_bias) as SURVEY.md 8(d) prescribes."""
    rs = np.random.RandomState(seed)
    n = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    x = n(batch, seqlen, nheads, headdim).to(dtype)
    dt = n(batch, seqlen, nheads).to(dtype)
    Bm = n(batch, seqlen, ngroups, dstate).to(dtype)
    Cm = n(batch, seqlen, ngroups, dstate).to(dtype)
    A = -torch.from_numpy(rs.uniform(1.0, 16.0, size=nheads).astype(np.float32))
    D = torch.from_numpy(rs.uniform(0.5, 1.5, size=nheads).astype(np.float32))
    dt0 = np.exp(rs.uniform(math.log(1e-3), math.log(1e-1), size=nheads)).clip(min=1e-4)
    dt_bias = torch.from_numpy((dt0 + np.log(-np.expm1(-dt0))).astype(np.float32))
    return x, dt, A, Bm, Cm, D, dt_bias
