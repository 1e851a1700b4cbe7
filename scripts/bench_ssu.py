"""selective_state_update alone at batch 64 / d_model 2048 geometry (H=64, P=64, N=128), states streaming from HBM (twelve
state sets rotate inside one CUDA graph).  python scripts/bench_ssu.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.selective_state_update import selective_state_update  # noqa: E402


def main():
    dev = "cuda"
    B, H, P, N = 64, 64, 64, 128
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6548.5
    for sdt in (torch.bfloat16, torch.float32):
        g = torch.Generator(device=dev).manual_seed(0)
        rn = lambda *s: torch.randn(*s, device=dev, generator=g)
        states = [(rn(B, H, P, N) * 0.1).to(sdt) for _ in range(12)]
        x, dt = rn(B, H * P).bfloat16().view(B, H, P), rn(B, H).bfloat16().view(B, H, 1).expand(B, H, P)
        A = (-torch.rand(H, device=dev) * 8 - 0.5).view(H, 1, 1).expand(H, P, N)
        Bm, Cm = rn(B, 1, N).bfloat16(), rn(B, 1, N).bfloat16()
        D = torch.ones(H, device=dev).view(H, 1).expand(H, P)
        dtb = rn(H).view(H, 1).expand(H, P)
        fns = [lambda s=s: selective_state_update(s, x, dt, A, Bm, Cm, D, dt_bias=dtb, dt_softplus=True) for s in states]
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
        us = e0.elapsed_time(e1) / 120 * 1e3
        by = 2 * B * H * P * N * states[0].element_size()
        print(json.dumps({"op": "selective_state_update", "state": str(sdt), "us": us, "gbs": by / us / 1e3, "frac_of_hbm": by / us / 1e3 / hbm}))


if __name__ == "__main__":
    main()
