"""Host-side cost of the public calls at a shape where the kernels are short (the 72-token caption prefill of
inference_t2i.py, batch 2, d_model 2048): wall time per eager call (host + device, back to back) against the device time of
the same call replayed from a CUDA graph.  python scripts/bench_host_overhead.py > gpurun_out/host_overhead.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.causal_conv1d import causal_conv1d_fn  # noqa: E402
from omnimamba_b200.interface.gemm import linear  # noqa: E402
from omnimamba_b200.interface.layernorm_gated import rmsnorm_fn  # noqa: E402
from omnimamba_b200.interface.ssd_combined import mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined  # noqa: E402


def wall_us(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def graph_us(fn, n=200):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n // 10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n // 10 * 10) * 1e3


def main():
    dev = "cuda"
    B, L, H, P, N, d = 2, 72, 64, 64, 128, 2048
    bf = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
    x, dt, Bm, Cm = bf(B, L, H, P), bf(B, L, H), bf(B, L, 1, N), bf(B, L, 1, N)
    A, D, dtb = -torch.rand(H, device=dev) * 8 - 0.5, torch.ones(H, device=dev), torch.randn(H, device=dev)
    zx = bf(B, L, 2 * H * P + 2 * N + H)
    cw, cb = bf(H * P + 2 * N, 4), bf(H * P + 2 * N)
    nw = torch.ones(H * P, device=dev, dtype=torch.bfloat16)
    u, w_in, w_out = bf(B, L, d), bf(2 * H * P + 2 * N + H, d) * 0.02, bf(d, H * P) * 0.02
    y = bf(B, L, H * P)
    xbc = zx[..., H * P:2 * H * P + 2 * N].transpose(1, 2)
    cases = {
        "mamba_chunk_scan_combined fwd": lambda: mamba_chunk_scan_combined(x, dt, A, Bm, Cm, 256, D=D, dt_bias=dtb, dt_softplus=True),
        "causal_conv1d_fn fwd": lambda: causal_conv1d_fn(xbc, cw, cb, activation="silu"),
        "rmsnorm_fn gated fwd": lambda: rmsnorm_fn(y, nw, None, z=y, eps=1e-5, group_size=None, norm_before_gate=False),
        "in_proj linear": lambda: linear(u, w_in),
        "out_proj linear": lambda: linear(y, w_out),
    }
    out = []
    with torch.no_grad():
        for name, fn in cases.items():
            r = {"call": name, "shape": f"B={B} L={L} d_model={d}", "eager_wall_us": wall_us(fn), "graph_device_us": graph_us(fn)}
            r["host_bound"] = r["eager_wall_us"] > 1.5 * r["graph_device_us"]
            out.append(r)
            print(r, file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
