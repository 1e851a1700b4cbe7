"""BASELINE config 4: single-token autoregressive decode of a 48-layer, d_model=2048 Mamba-2 backbone at batch 64, the way
`models/stage2/generation.py:383-431` [R] runs it - one full forward of one token per layer stack, captured in a CUDA graph
after a prefill, replayed 255 times (run on the GPU box):

    python scripts/bench_decode.py [--layers 48] [--batch 64] [--dtype fp32|bf16] > gpurun_out/bench_decode.json

Per layer: fused residual-add + RMSNorm (layer_norm_fn) -> Mamba2.step = in_proj GEMM (torch / cuBLAS: SURVEY 8 f-2) ->
causal_conv1d_update -> selective_state_update -> gated RMSNorm -> out_proj GEMM.  Random weights, synthetic embeddings.
Reports tokens/s = batch * steps / time and the algorithmic bytes per step (ssm state read + write, conv state, weights).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnimamba_b200.interface.layer_norm import layer_norm_fn  # noqa: E402
from omnimamba_b200.modules import Mamba2  # noqa: E402


class IP:
    """generation.py:19-36 [R] InferenceParams, duck-typed"""
    def __init__(self):
        self.seqlen_offset, self.key_value_memory_dict = 0, {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=48)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--steps", type=int, default=255)
    ap.add_argument("--prefill", type=int, default=72)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dt = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    d_model = 2048
    torch.manual_seed(0)
    mixers = [Mamba2(d_model, layer_idx=i, device=dev, dtype=dt) for i in range(args.layers)]
    norms = [torch.ones(d_model, device=dev, dtype=dt) for _ in range(args.layers)]
    for m in mixers:
        m.eval()

    def stack(h, ip):
        res = None
        for m, w in zip(mixers, norms):
            h, res = layer_norm_fn(h, w, None, residual=res, prenorm=True, residual_in_fp32=True, eps=1e-5, is_rms_norm=True)
            h = m(h, inference_params=ip)
        return h

    ip = IP()
    with torch.no_grad():
        stack(torch.randn(args.batch, args.prefill, d_model, device=dev, dtype=dt), ip)   # prefill: fills the caches
        ip.seqlen_offset = args.prefill
        tok = torch.randn(args.batch, 1, d_model, device=dev, dtype=dt)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                stack(tok, ip)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = stack(tok, ip)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / args.steps
    es = 4 if args.dtype == "fp32" else 2
    H, P, N, conv_dim, d_in_proj, d_inner = 64, 64, 128, 4352, 8512, 4096
    state = 2 * args.batch * H * P * N * es + 2 * args.batch * conv_dim * 4 * es
    weights = (d_in_proj * d_model + d_model * d_inner) * es
    by = args.layers * (state + weights)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    print(json.dumps({"metric": "decode tokens/s, 48-layer d_model=2048 Mamba-2 stack, CUDA graph", "value": args.batch / t,
                      "unit": "tokens/s", "ms_per_step": t * 1e3, "batch": args.batch, "layers": args.layers, "dtype": args.dtype,
                      "bytes_per_step": by, "achieved_gbs": by / t / 1e9, "peak_gbs": hbm, "frac": by / t / 1e9 / hbm,
                      "finite": bool(torch.isfinite(out.float()).all().item()),
                      "note": "in_proj / out_proj are torch GEMMs; every other kernel is libomnissm"}))


if __name__ == "__main__":
    main()
