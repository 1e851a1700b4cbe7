"""The fused single-token layer core (csrc/decode_core.cu) alone, at batch 64 / d_model 2048, states streaming from HBM:
twelve layers' worth of state (12 x 67 MB bf16) rotate inside one CUDA graph.  python scripts/bench_decode_core.py [bf16|fp32]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.decode import mamba2_decode_core  # noqa: E402


def main():
    dt = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else torch.bfloat16
    dev = "cuda"
    B, H, P, N, W = 64, 64, 64, 128, 4
    dim, conv_dim = H * P, H * P + 2 * N
    L = 12
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    states = [(rn(B, H, P, N) * 0.1).to(dt) for _ in range(L)]
    convs = [rn(B, W, conv_dim).to(dt).transpose(1, 2) for _ in range(L)]
    zx = rn(B, 2 * dim + 2 * N + H).to(dt)
    cw, cb = rn(conv_dim, W).to(dt), rn(conv_dim).to(dt)
    A, D, dtb, nw = -torch.rand(H, device=dev) * 8 - 0.5, torch.ones(H, device=dev), rn(H), torch.ones(dim, device=dev, dtype=dt)
    fns = [lambda s=s, c=c: mamba2_decode_core(zx, c, cw, cb, s, A, D, dtb, nw, 1e-5) for s, c in zip(states, convs)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for f in fns:
            f()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (10 * L) * 1e3
    by = 2 * B * H * P * N * states[0].element_size()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6548.5
    print(json.dumps({"op": "mamba2_decode_core", "dtype": str(dt), "batch": B, "us_per_call": us, "state_bytes": by,
                      "gbs": by / us / 1e3, "frac_of_hbm": by / us / 1e3 / hbm}))


if __name__ == "__main__":
    main()
