"""causal_conv1d backward alone (the C-ABI call, no autograd glue): time, HBM fraction and a check against fp32 autograd of
F.conv1d + SiLU, at the bench shape (16, 4096) and the stage-1 training shape (90, 329); channel-last slice of zxbcdt.

    OMNI_LIB_PATH=.ab/lib_x.so python scripts/bench_conv_bwd.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.causal_conv1d import causal_conv1d_fn, conv1d_bwd_raw  # noqa: E402
from omnimamba_b200 import _cabi as abi  # noqa: E402


def main():
    dev = "cuda"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0
    g = torch.Generator(device=dev).manual_seed(0)
    D, d_inner = 4352, 4096
    for (B, L) in ((16, 4096), (90, 329), (2, 70)):
        zx = torch.randn(B, L, 8512, device=dev, generator=g).bfloat16()
        x = zx[..., d_inner:d_inner + D].transpose(1, 2)
        dzx = torch.randn(B, L, 8512, device=dev, generator=g).bfloat16()
        dout = dzx[..., d_inner:d_inner + D].transpose(1, 2)
        w = torch.randn(D, 4, device=dev, generator=g) * 0.5
        b = torch.randn(D, device=dev, generator=g) * 0.1
        dx = torch.empty(B, L, D, device=dev, dtype=torch.bfloat16).transpose(1, 2)
        dw, db = torch.zeros(D, 4, device=dev), torch.zeros(D, device=dev)
        act = abi.ACT_SILU if hasattr(abi, "ACT_SILU") else 1
        run = lambda: conv1d_bwd_raw(x, w, b, dout, None, None, dx, dw, db, None, act)
        run()
        # reference: fp32 autograd of conv1d + silu on the same (bf16-rounded) inputs
        xr = x.float().detach().requires_grad_()
        wr, br = w.clone().requires_grad_(), b.clone().requires_grad_()
        y = F.silu(F.conv1d(xr, wr.unsqueeze(1), br, padding=3, groups=D)[..., :L])
        y.backward(dout.float())
        rel = lambda a, r: ((a.double() - r.double()).norm() / r.double().norm()).item()
        errs = dict(dx=rel(dx, xr.grad), dw=rel(dw, wr.grad), db=rel(db, br.grad))
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        steps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        by = 3 * B * L * D * 2
        print(json.dumps(dict(op="causal_conv1d bwd kernel", B=B, L=L, ms=ms, bytes=by, gbs=by / ms / 1e6, frac=by / ms / 1e6 / hbm, **errs)))
        # forward through the operator surface (no autograd), same layout
        with torch.no_grad():
            fwd = lambda: causal_conv1d_fn(x, w, b, activation="silu")
            yk = fwd()
            ferr = rel(yk, y.detach())
            for _ in range(5):
                fwd()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fwd()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        by = 2 * B * L * D * 2
        print(json.dumps(dict(op="causal_conv1d fwd", B=B, L=L, ms=ms, bytes=by, gbs=by / ms / 1e6, frac=by / ms / 1e6 / hbm, y=ferr)))


if __name__ == "__main__":
    main()
