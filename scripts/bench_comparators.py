"""GPU comparators for the headline op: the SSD forward of this repo beside the library kernels that ship in
the image, on the bench workload (B=16, L=4096, H=64, P=64, G=1, N=128, bf16, D, dt_bias, softplus), same inputs.

  vllm       vllm.model_executor.layers.mamba.ops.ssd_combined.mamba_chunk_scan_combined_varlen - the Triton port of
             mamba_ssm's five forward kernels (chunk 256, fp32 states); the stand-in for "the reference mamba-ssm build",
             which cannot be installed here.  Its kernels are autotuned: the first call compiles ~60 configurations.
  flashinfer flashinfer.mamba.SSDCombined - CuTe-DSL tcgen05 kernel (chunk 128, bf16 states; y returned as (B, L, H, P))

Timed like bench.py (5 warm-up + 20 steps, CUDA events).  Prints one JSON line; a comparator that fails to import, compile or
run is reported with its error instead of a number.   python scripts/bench_comparators.py [vllm] [flashinfer]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from omnimamba_b200.interface.ssd_combined import ssd_fwd_raw  # noqa: E402

B, L, H, P, G, N = int(os.environ.get("OMNI_CMP_BATCH", "16")), 4096, 64, 64, 1, 128


def timed(fn, steps=20, warmup=5):
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


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def kernel_times(fn, steps=5):
    """Device time per call split by kernel name (CUPTI via torch.profiler): {name: ms per call}, total ms per call."""
    from torch.profiler import ProfilerActivity, profile
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
    out = {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if t and t > 0:
            out[ev.key[:80]] = t / 1e3 / steps
    return dict(sorted(out.items(), key=lambda kv: -kv[1])[:8]), sum(out.values())


def main():
    which = sys.argv[1:] or ["vllm", "flashinfer", "fla"]
    host = bench.make_inputs(B, L)
    d = {k: v.cuda() for k, v in host.items()}
    out = torch.empty(B, L, H, P, device="cuda", dtype=torch.bfloat16)
    ours = lambda: ssd_fwd_raw(d["x"], d["dt"], d["A"], d["B"], d["C"], 256, D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True,
                               out=out)
    res = {"workload": f"B={B} L={L} H={H} P={P} G={G} N={N} bf16", "ours_ms": timed(ours)}
    y_ours = out.clone()
    print("ours", res["ours_ms"], "ms", file=sys.stderr, flush=True)
    try:
        res["ours_kernels_ms"], res["ours_kernel_sum_ms"] = kernel_times(ours)
    except Exception as e:  # noqa: BLE001
        res["ours_kernels_error"] = f"{type(e).__name__}: {e}"[:200]
    # forward + backward of this repo (ssd_fwd_raw + ssd_bwd_raw, as bench.py's fwd_bwd leg)
    from omnimamba_b200.interface.ssd_combined import ssd_bwd_raw
    dy = torch.randn(B, L, H, P, device="cuda", dtype=torch.bfloat16)

    def ours_fb():
        ours()
        ssd_bwd_raw(dy, d["x"], d["dt"], d["A"], d["B"], d["C"], 256, D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
    res["ours_fwd_bwd_ms"] = timed(ours_fb, steps=10, warmup=3)

    if "vllm" in which:
        try:
            t0 = time.time()
            from vllm.model_executor.layers.mamba.ops.ssd_combined import mamba_chunk_scan_combined_varlen
            Q = 256
            nch = B * L // Q
            i32 = dict(device="cuda", dtype=torch.int32)
            cu_seqlens = torch.arange(0, B * L + 1, L, **i32)
            cu_chunks = torch.arange(0, B * L + 1, Q, **i32)
            last_chunk = torch.arange(1, B + 1, **i32) * (L // Q) - 1
            seq_idx = torch.arange(nch, **i32) // (L // Q)
            xf, dtf = d["x"].reshape(B * L, H, P), d["dt"].reshape(B * L, H)
            Bf, Cf = d["B"].reshape(B * L, G, N), d["C"].reshape(B * L, G, N)
            yv = torch.empty(B * L, H, P, device="cuda", dtype=torch.bfloat16)
            run = lambda: mamba_chunk_scan_combined_varlen(xf, dtf, d["A"], Bf, Cf, Q, cu_seqlens, cu_chunks, last_chunk, seq_idx,
                                                           yv, D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True,
                                                           state_dtype=torch.float32)
            run()
            torch.cuda.synchronize()
            res["vllm_triton_compile_s"] = time.time() - t0
            res["vllm_triton_ms"] = timed(run)
            res["vllm_kernels_ms"], res["vllm_kernel_sum_ms"] = kernel_times(run)
            res["vllm_vs_ours_rel_l2"] = rel_l2(yv.view(B, L, H, P), y_ours)
        except Exception as e:  # noqa: BLE001
            res["vllm_triton_error"] = f"{type(e).__name__}: {e}"[:300]

    if "flashinfer" in which:
        try:
            t0 = time.time()
            from flashinfer.mamba import SSDCombined
            ssd = SSDCombined(chunk_size=128, nheads=H, headdim=P, dstate=N, ngroups=G)
            Db = d["D"].to(torch.bfloat16)
            run = lambda: ssd.run(d["x"], d["dt"], d["A"], d["B"], d["C"], D=Db, dt_bias=d["dt_bias"], dt_softplus=True)
            yf, _ = run()
            torch.cuda.synchronize()
            res["flashinfer_compile_s"] = time.time() - t0
            res["flashinfer_ms"] = timed(run)
            res["flashinfer_kernels_ms"], res["flashinfer_kernel_sum_ms"] = kernel_times(run)
            # (run() returns y as (B, L, H, P): its kernel writes (B, H, P, chunks, Q) and the wrapper copies - that copy and
            # the Triton chunk-cumsum pre-kernel are part of its public call and of the time above)
            res["flashinfer_vs_ours_rel_l2"] = rel_l2(yf.reshape(B, L, H, P), y_ours)
        except Exception as e:  # noqa: BLE001
            res["flashinfer_error"] = f"{type(e).__name__}: {e}"[:300]

    if "fla" in which:
        # fla.ops.simple_gla.chunk_simple_gla: the forward+backward comparator SURVEY.md 8(d) names (q = C, k = B, v = x dt,
        # g = dt A per head; B / C have to be materialised per head for it - its own layout cost is part of its public call)
        try:
            t0 = time.time()
            import torch.nn.functional as F
            from fla.ops.simple_gla import chunk_simple_gla
            dtt = F.softplus(d["dt"].float() + d["dt_bias"])
            q = d["C"].expand(B, L, H, N).contiguous().requires_grad_()
            k = d["B"].expand(B, L, H, N).contiguous().requires_grad_()
            v = (d["x"].float() * dtt[..., None]).to(torch.bfloat16).requires_grad_()
            g = (dtt * d["A"]).float().requires_grad_()

            def fla_fwd():
                return chunk_simple_gla(q, k, v, g=g, scale=1.0)[0]

            def fla_fb():
                o = fla_fwd()
                o.backward(dy)
                q.grad = k.grad = v.grad = g.grad = None
            yf = fla_fwd()
            torch.cuda.synchronize()
            res["fla_compile_s"] = time.time() - t0
            with torch.no_grad():
                res["fla_fwd_ms"] = timed(fla_fwd)
            res["fla_fwd_bwd_ms"] = timed(fla_fb, steps=10, warmup=3)
            y_nod = (y_ours.float() - d["x"].float() * d["D"][None, None, :, None])
            res["fla_vs_ours_rel_l2 (y - D x)"] = rel_l2(yf, y_nod)
        except Exception as e:  # noqa: BLE001
            res["fla_error"] = f"{type(e).__name__}: {e}"[:300]

    print(json.dumps(res))


if __name__ == "__main__":
    main()
