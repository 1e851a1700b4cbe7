"""What keeping the chunk states costs the forward and saves the backward (omnissm.h: chunk_states), same process / same box:
forward plain vs forward + state stores, backward with its own forward sweep vs backward on the kept states."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.ssd_combined import _alloc_chunk_states, ssd_bwd_raw, ssd_fwd_raw  # noqa: E402


def timeit(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = "cuda"
    H, P, N = 64, 64, 128
    g = torch.Generator(device=dev).manual_seed(0)
    for (B, L) in ((16, 4096), (90, 329), (64, 1024)):
        rn = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
        x, dt, Bm, Cm, dy = rn(B, L, H, P), rn(B, L, H), rn(B, L, 1, N), rn(B, L, 1, N), rn(B, L, H, P)
        A = -(torch.rand(H, device=dev, generator=g) * 15 + 1)
        dt_bias = torch.rand(H, device=dev, generator=g) * 4 - 6
        D = torch.ones(H, device=dev)
        out = torch.empty_like(x)
        cs = _alloc_chunk_states(B, L, H, P, N, dev, torch.bfloat16)
        kw = dict(D=D, dt_bias=dt_bias, dt_softplus=True)
        f0 = timeit(lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, out=out, **kw))
        f1 = timeit(lambda: ssd_fwd_raw(x, dt, A, Bm, Cm, 256, out=out, chunk_states=cs, **kw))
        kept = ssd_fwd_raw(x, dt, A, Bm, Cm, 256, out=out, chunk_states=cs, **kw)[2]
        b0 = timeit(lambda: ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, **kw), steps=8, warmup=3)
        b1 = timeit(lambda: ssd_bwd_raw(dy, x, dt, A, Bm, Cm, 256, chunk_states=kept, **kw), steps=8, warmup=3)
        print(json.dumps(dict(B=B, L=L, kept=kept is not None, fwd_ms=f0, fwd_keep_ms=f1, bwd_ms=b0, bwd_kept_ms=b1,
                              fwd_bwd_ms=f0 + b0, fwd_bwd_kept_ms=f1 + b1, state_mb=(cs.numel() * 2 / 1e6) if cs is not None else 0)))


if __name__ == "__main__":
    main()
