"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank runs the oracle's scan + autograd on its batch shard, parameter
gradients are all-reduced as bench.py / DDP do, and the result must equal the single-process run on the whole batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grads(x, dt, A, Bm, Cm, D, dt_bias):
    import oracle
    A, D, dt_bias = (t.clone().requires_grad_() for t in (A, D, dt_bias))
    y = oracle.mamba_chunk_scan_combined_ref(x, dt, A, Bm, Cm, 64, D=D, dt_bias=dt_bias, dt_softplus=True)
    y.square().sum().backward()
    return [A.grad, D.grad, dt_bias.grad], y.detach()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cases import scan_inputs
        from omnimamba_b200.dist import allreduce_param_grads, max_over_ranks, shard_batch
        torch.set_num_threads(1)
        x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(3, 40, 4, 16, 1, 16, 0, torch.float32)  # odd batch: uneven shards
        xs, dts, Bs, Cs = shard_batch([x, dt, Bm, Cm], rank, world)
        g, y = _grads(xs, dts, A, Bs, Cs, D, dt_bias)
        allreduce_param_grads(g)
        slow = max_over_ranks(1.0 + rank)
        q.put((rank, [t.clone() for t in g], y, slow))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from omnimamba_b200.dist import shard_range
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_gloo_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from cases import scan_inputs
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, dt, A, Bm, Cm, D, dt_bias = scan_inputs(3, 40, 4, 16, 1, 16, 0, torch.float32)
    g_ref, y_ref = _grads(x, dt, A, Bm, Cm, D, dt_bias)
    for rank, g, y, slow in res:
        assert slow == 2.0                                   # max over ranks
        for a, b in zip(g, g_ref):                           # every rank holds the full-batch parameter gradients
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    y_cat = torch.cat([r[2] for r in res], 0)                # no data-path collective: outputs are just the shards
    assert torch.allclose(y_cat, y_ref, rtol=1e-6, atol=1e-6)


def test_bench_host_locality_is_harmless_without_a_gpu():
    """bench.py allocates its pinned e2e buffers on the CPUs local to the GPU.  Without a GPU (or without sysfs PCI data, as in
    the GPU box's VM) the lookup must report "unknown" and the context manager must leave the CPU affinity untouched."""
    sys.path.insert(0, ROOT)
    import bench
    node, cpus, link = bench.gpu_host_locality(0)
    assert isinstance(node, int) and isinstance(link, str)
    before = os.sched_getaffinity(0)
    with bench.near_gpu(0) as loc:
        assert loc.node == node
    assert os.sched_getaffinity(0) == before


def _reducer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from omnimamba_b200.dist import BucketedGradReducer
        torch.set_num_threads(1)
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
        net[2].weight.requires_grad_(False)                      # a frozen parameter in the middle (stage "align" pattern)
        red = BucketedGradReducer(list(net.parameters()), bucket_bytes=40)   # tiny buckets: three of them
        data = torch.randn(4, 6, generator=torch.Generator().manual_seed(1))
        out = []
        for step in range(2):                                    # two steps: zero_grad() must re-arm the buckets
            red.zero_grad()
            net(data[rank * 2:(rank + 1) * 2] * (step + 1)).square().mean().backward()
            red.finish()
            out.append([p.grad.clone() for p in net.parameters() if p.requires_grad])
        q.put((rank, out, len(red.buckets), red.total_bytes()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_bucketed_grad_reducer_two_ranks_equals_full_batch_mean():
    """DDP semantics: after backward every rank holds the gradient of the mean loss over the GLOBAL batch."""
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reducer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
    net[2].weight.requires_grad_(False)
    data = torch.randn(4, 6, generator=torch.Generator().manual_seed(1))
    for step in range(2):
        net.zero_grad()
        net(data * (step + 1)).square().mean().backward()
        ref = [p.grad for p in net.parameters() if p.requires_grad]
        for rank, out, nb, nbytes in res:
            assert nb >= 3 and nbytes == 4 * sum(p.numel() for p in net.parameters() if p.requires_grad)
            for a, b in zip(out[step], ref):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
