"""Whole-model workloads of BASELINE.json (configs 2, 3, 4) on the 48-layer d_model=2048 stack, timed like bench.py times
the scan (CUDA events on the launch stream, >= 3 warm-ups, max over ranks).  Random-init weights of the reference
architecture (omnimamba_b200/backbone.py), synthetic token ids / embeddings.

  model_fwd  config 2: forward only, (B, L) = (8, 1024) embeddings -> 48 x [add+norm -> Mamba2 (LoRA in_proj)] -> norm; bf16 autocast
  train      config 3: the stage-1 ("align") t2i step of train_stage2.py [R]: per rank 90 sequences of 73 caption + 256 image
             tokens (L = 329), bf16 autocast over fp32 master weights, shifted CE on the image tokens, AdamW(0.9, 0.95);
             gradients of the trainable parameters all-reduced in flat buckets on a side stream while backward runs (DDP)
  decode     config 4: 255 single-token steps at batch 64 under one CUDA graph after a 72-token prefill, fp32 caches/weights
             (as scripts/inference_t2i.py runs it) and bf16
"""
from __future__ import annotations

import json
import os
import time

import torch

D_MODEL, N_LAYER = 2048, 48
H, P, N, CONV_DIM, D_IN_PROJ, D_INNER = 64, 64, 128, 4352, 8512, 4096


def _events_ms(fn, steps, warmup, dist_on):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    return e0.elapsed_time(e1) / steps


def _max(v, dist_on, device):
    from omnimamba_b200.dist import max_over_ranks
    return max_over_ranks(v, device=device) if dist_on else v


def _gemm_flops_per_token(lora=True, train=False, head_vocab=0):
    f = 2 * D_MODEL * D_IN_PROJ + 2 * D_INNER * D_MODEL + (2 * 8 * (D_MODEL + D_IN_PROJ) if lora else 0)
    f *= N_LAYER
    f += 2 * D_MODEL * head_vocab
    return f * (2 if train else 1)   # stage "align": frozen base weights -> forward + dgrad, no wgrad of W


def model_fwd(device, steps=5, warmup=3, batch=8, seqlen=1024, n_layer=N_LAYER):
    from omnimamba_b200 import _cabi
    from omnimamba_b200.backbone import MixerStack
    torch.manual_seed(0)
    stack = MixerStack(D_MODEL, n_layer, device=device).eval()
    x = torch.randn(batch, seqlen, D_MODEL, device=device, dtype=torch.bfloat16)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return stack(x)

    _cabi.reset_launch_count()
    y = step()
    launches = _cabi.launch_count()
    ms_eager = _events_ms(step, steps, warmup, False)
    # the same forward replayed from a CUDA graph (static input; every libomnissm call is capturable): removes the host-side
    # launch gaps between ~670 kernels
    ms = ms_eager
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            yg = step()
        ms_graph = _events_ms(graph.replay, steps, warmup, False)
        same = bool(torch.equal(yg, y))
        if same and ms_graph < ms_eager:
            ms = ms_graph
    except Exception as e:  # noqa: BLE001
        ms_graph, same = None, f"{type(e).__name__}: {e}"[:200]
    toks = batch * seqlen
    flops = _gemm_flops_per_token() * toks * n_layer / N_LAYER
    del stack
    return {"workload": f"config 2: {n_layer}-layer d_model={D_MODEL} forward, B={batch} L={seqlen}, bf16 autocast, LoRA in_proj",
            "value": toks / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms, "ms_per_step_eager": ms_eager,
            "ms_per_step_cuda_graph": ms_graph, "graph_output_bit_equal": same, "libomnissm_launches_per_step": launches,
            "gemm_tflops_achieved": flops / ms / 1e9, "finite": bool(torch.isfinite(y.float()).all().item())}


def train(device, rank, world, steps=5, warmup=3, batch=90, n_layer=N_LAYER, overlap=True):
    from omnimamba_b200 import _cabi
    from omnimamba_b200.backbone import T2IModel
    from omnimamba_b200.dist import BucketedGradReducer
    dist_on = world > 1
    torch.manual_seed(0)                       # same weights on every rank (DDP broadcasts rank 0's)
    model = T2IModel(D_MODEL, n_layer, device=device).freeze_backbones("align")
    model.train()
    for layer in model.backbone.layers:        # only the active task's adapters receive gradients (DDP with
        for nme, p in layer.mixer.in_proj.named_parameters():   # find_unused_parameters=False needs every trainable one used)
            if "mmu_lora" in nme:
                p.requires_grad_(False)
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = BucketedGradReducer(params, bucket_bytes=256 << 20)
    opt = torch.optim.AdamW(params, lr=8e-4, betas=(0.9, 0.95), weight_decay=0.0, fused=True)
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    image_ids = torch.randint(0, 16384, (batch, 256), device=device, generator=g)
    caption_ids = torch.randint(0, 50277, (batch, 73), device=device, generator=g)
    loss_box = [None]
    timed = [False]

    def step():
        reducer.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(image_ids, caption_ids)
        loss.backward()
        if not overlap and dist_on:
            torch.cuda.current_stream().synchronize()
        reducer.finish(timed=timed[0])
        opt.step()
        loss_box[0] = loss

    _cabi.reset_launch_count()
    step()
    launches = _cabi.launch_count()
    ms = _max(_events_ms(step, steps, warmup, dist_on), dist_on, device)
    timed[0] = True
    step()
    torch.cuda.synchronize()
    exposed = _max(reducer.exposed(), dist_on, device)
    L = 73 + 256
    toks = batch * L
    flops = _gemm_flops_per_token(train=True, head_vocab=16384) * toks * n_layer / N_LAYER
    out = {"workload": f"config 3: stage-1 (align) t2i train step, {n_layer}-layer d_model={D_MODEL}, per-rank B={batch} L={L}, bf16 autocast, "
                       "AdamW, LoRA r=8 + embeddings + img_head trainable",
           "value": world * toks / ms * 1e3, "unit": "tokens/s", "tokens_per_s_per_gpu": toks / ms * 1e3, "ms_per_step": ms,
           "n_gpus": world, "loss": float(loss_box[0].item()), "libomnissm_launches_per_step": launches,
           "gemm_tflops_per_gpu": flops / ms / 1e9,
           "allreduce": {"bytes": reducer.total_bytes(), "buckets": len(reducer.buckets), "exposed_ms": exposed,
                         "overlapped_with_backward": overlap,
                         "nccl_algo": os.environ.get("NCCL_ALGO", "NCCL's own choice (NCCL_ALGO unset; the reference forces Tree)")},
           "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2**30}
    del model, opt, reducer
    return out


def decode(device, dtype=torch.float32, batch=64, steps=255, prefill=72, n_layer=N_LAYER, hbm_gbs=6650.0):
    from omnimamba_b200 import _cabi
    from omnimamba_b200.backbone import InferenceParams, MixerStack
    torch.manual_seed(0)
    stack = MixerStack(D_MODEL, n_layer, device=device, dtype=dtype, lora=False).eval()
    ip = InferenceParams(max_seqlen=prefill + steps + 1, max_batch_size=batch)
    with torch.no_grad():
        stack(torch.randn(batch, prefill, D_MODEL, device=device, dtype=dtype), ip)
        ip.seqlen_offset = prefill
        tok = torch.randn(batch, 1, D_MODEL, device=device, dtype=dtype)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                stack(tok, ip)
        torch.cuda.current_stream().wait_stream(side)
        _cabi.reset_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = stack(tok, ip)
        launches = _cabi.launch_count()
        ms = _events_ms(graph.replay, steps, 3, False)
    es = 4 if dtype == torch.float32 else 2
    state = 2 * batch * H * P * N * es + 2 * batch * CONV_DIM * 4 * es
    weights = (D_IN_PROJ * D_MODEL + D_MODEL * D_INNER) * es
    by = n_layer * (state + weights)
    res = {"workload": f"config 4: single-token decode, {n_layer}-layer d_model={D_MODEL}, batch {batch}, {steps} graph replays, "
                       f"{'fp32' if es == 4 else 'bf16'} weights and caches",
           "value": batch / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms, "libomnissm_launches_per_step": launches,
           "bytes_per_step": by, "achieved_gbs": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / hbm_gbs,
           "finite": bool(torch.isfinite(out.float()).all().item())}
    del stack, graph
    return res


def run_all(device, rank, world, which, hbm_gbs, steps=5):
    """-> dict name -> result (or {"error": ...}); whole-model legs never take the bench line down with them."""
    out = {}
    for name in which:
        t0 = time.time()
        try:
            if name == "model_fwd":
                if world == 1:
                    out[name] = model_fwd(device, steps=steps)
            elif name == "train":
                out[name] = train(device, rank, world, steps=steps)
            elif name == "decode":
                if world == 1:
                    out["decode_fp32"] = decode(device, torch.float32, hbm_gbs=hbm_gbs)
                    out["decode_bf16"] = decode(device, torch.bfloat16, hbm_gbs=hbm_gbs)
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
        torch.cuda.empty_cache()
        if name in out and isinstance(out[name], dict):
            out[name]["wall_s"] = round(time.time() - t0, 1)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["model_fwd", "train", "decode"])
    ap.add_argument("--layers", type=int, default=N_LAYER)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    N_LAYER_RUN = a.layers
    print(json.dumps(run_all(dev, 0, 1, a.which, 6548.5)))
