"""Regenerate tests/golden/*.npz from INDEPENDENT in-container implementations.

Run from the repo root:  python tests/golden/make_golden.py

* mamba2_block_hf.npz  - transformers' pure-PyTorch ``Mamba2Mixer.torch_forward``
  (an independent restatement of mamba_ssm's Mamba2 block) on the cases in
  ``cases.BLOCK_CASES`` with parameters from ``cases.block_params``.
* ops_lib.npz - per-op answers from library primitives (F.conv1d, F.softplus,
  torch.cumsum/einsum in fp64) that share no code with oracle/.
The reference repo itself (hustvl/OmniMamba) has no tests or vectors to import.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402


def hf_block(d_model, batch, seqlen, chunk, seed):
    from transformers.models.mamba2.configuration_mamba2 import Mamba2Config
    from transformers.models.mamba2.modeling_mamba2 import Mamba2Mixer

    d_inner = 2 * d_model
    cfg = Mamba2Config(
        num_heads=d_inner // cases.HEADDIM, head_dim=cases.HEADDIM, hidden_size=d_model,
        state_size=cases.D_STATE, expand=2, conv_kernel=cases.D_CONV, n_groups=cases.NGROUPS,
        use_bias=False, use_conv_bias=True, chunk_size=chunk, layer_norm_epsilon=cases.EPS,
        rms_norm=True, num_hidden_layers=1, vocab_size=16)
    mixer = Mamba2Mixer(cfg, layer_idx=0).float().eval()
    sd = cases.block_params(d_model, seed)
    missing = mixer.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    u = cases.block_input(d_model, batch, seqlen, seed)
    with torch.no_grad():
        y = mixer.torch_forward(u)
    return y.numpy()


def lib_ops():
    out = {}
    rs = np.random.RandomState(7)
    # causal conv via the library grouped conv (fp64)
    x = torch.from_numpy(rs.standard_normal((2, 24, 37)))
    w = torch.from_numpy(rs.standard_normal((24, 4)))
    b = torch.from_numpy(rs.standard_normal(24))
    y = F.conv1d(x, w.unsqueeze(1), b, padding=3, groups=24)[..., :37]
    out["conv_x"], out["conv_w"], out["conv_b"] = x.numpy(), w.numpy(), b.numpy()
    out["conv_y"], out["conv_y_silu"] = y.numpy(), F.silu(y).numpy()
    # SSD scan via a dense fp64 "attention matrix" evaluation: y_i = sum_{j<=i} (C_i.B_j) exp(sum_{j<k<=i} a_k) dt_j x_j
    Bsz, L, H, P, N = 2, 45, 3, 8, 16
    xs = torch.from_numpy(rs.standard_normal((Bsz, L, H, P)))
    dtr = torch.from_numpy(rs.standard_normal((Bsz, L, H)))
    A = -torch.from_numpy(rs.uniform(1, 16, size=H))
    Bm = torch.from_numpy(rs.standard_normal((Bsz, L, 1, N)))
    Cm = torch.from_numpy(rs.standard_normal((Bsz, L, 1, N)))
    D = torch.from_numpy(rs.uniform(0.5, 1.5, size=H))
    dtb = torch.from_numpy(rs.standard_normal(H))
    dt = F.softplus(dtr + dtb)
    a = dt * A                                   # (B, L, H)
    cs = torch.cumsum(a, dim=1)
    seg = cs[:, :, None, :] - cs[:, None, :, :]  # (B, i, j, H)
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool))[None, :, :, None]
    Lmat = torch.where(causal, torch.exp(seg), torch.zeros_like(seg))
    G = torch.einsum("bin,bjn->bij", Cm[:, :, 0], Bm[:, :, 0])
    y = torch.einsum("bij,bijh,bjh,bjhp->bihp", G, Lmat, dt, xs) + D.view(1, 1, H, 1) * xs
    final = torch.einsum("bjh,bjh,bjhp,bjn->bhpn", torch.exp(cs[:, -1:, :] - cs), dt, xs, Bm[:, :, 0])
    for k, v in dict(ssd_x=xs, ssd_dt=dtr, ssd_A=A, ssd_B=Bm, ssd_C=Cm, ssd_D=D, ssd_dt_bias=dtb,
                     ssd_y=y, ssd_final=final).items():
        out[k] = v.numpy()
    return out


def main():
    blocks = {}
    for name, (d_model, batch, seqlen, chunk, seed) in cases.BLOCK_CASES.items():
        blocks[name] = hf_block(d_model, batch, seqlen, chunk, seed)
        print(name, blocks[name].shape, float(np.abs(blocks[name]).mean()))
    np.savez_compressed(os.path.join(HERE, "mamba2_block_hf.npz"), **blocks)
    np.savez_compressed(os.path.join(HERE, "ops_lib.npz"), **lib_ops())


if __name__ == "__main__":
    main()
