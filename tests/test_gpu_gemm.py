"""The tcgen05 GEMM behind in_proj / out_proj (csrc/gemm_tc.cu) against a plain PyTorch fp32
reference of the same op (F.linear on the same bf16-valued operands, fp32 math), through the C-ABI.

Tolerances: fp32 output <= 1e-5 relative L2 (fp32 accumulation of exact bf16 products); bf16 output: the error beyond the
output rounding itself (tests/parity_metric.py) <= 1e-4, and bit-equality with the rounded fp32 output."""
import pytest
import torch
import torch.nn.functional as F

from parity_metric import excess_over_rounding, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


def _operand(rows, cols, seed, major, pad=0):
    """(rows, cols) bf16 operand on the device, K-major (cols contiguous, optional row padding) or MN-major (a transposed view)."""
    t = _rand((rows, cols), seed)
    pad += (-(rows if major == "mn" else cols) - pad) % 8   # row pitch stays a multiple of 8 elements (16 bytes)
    if major == "k":
        buf = torch.zeros(rows, cols + pad, dtype=torch.bfloat16, device=DEV)
        buf[:, :cols] = t.to(DEV)
        return t, buf[:, :cols]
    buf = torch.zeros(cols, rows + pad, dtype=torch.bfloat16, device=DEV)
    buf[:, :rows] = t.t().to(DEV)
    return t, buf[:, :rows].t()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 2048), (329, 8512, 2048), (1000, 2048, 4096), (90 * 329, 264, 200),
                                   (1, 8, 8), (130, 8, 2048), (64, 16384, 2048), (64, 8512, 2048), (3, 2048, 4096)])
@pytest.mark.parametrize("amaj,bmaj", [("k", "k"), ("k", "mn"), ("mn", "mn"), ("mn", "k")])
@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_matches_fp32_reference(M, N, K, amaj, bmaj, mode):
    """mode 1: single-CTA 128 x 256 (or 128 x 64) tiles; mode 2: CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles) for
    every shape they are legal for - ragged M / N / K, all four operand-major combinations."""
    from omnimamba_b200 import _cabi
    a, ad = _operand(M, K, 1, amaj, pad=8)
    b, bd = _operand(N, K, 2, bmaj, pad=16)
    ref = a.float() @ b.float().t()
    assert _cabi.gemm_operands_ok(ad, bd)
    lib = _cabi.lib()
    try:
        lib.omni_debug_set_gemm_mode(mode)
        o32 = _cabi.gemm(ad, bd, torch.float32)
        o16 = _cabi.gemm(ad, bd, torch.bfloat16)
        torch.cuda.synchronize()
    finally:
        lib.omni_debug_set_gemm_mode(0)
    e32 = rel_l2(o32, ref)
    ex = excess_over_rounding(o16, ref)
    print(f"gemm M={M} N={N} K={K} {amaj}/{bmaj}: fp32-out rel_l2 {e32:.2e}, bf16-out excess {ex:.2e}, plain {rel_l2(o16, ref):.2e}")
    assert e32 <= 1e-5, e32
    assert ex <= 1e-4, ex
    assert torch.equal(o16.cpu(), o32.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("M", [1, 3, 64, 100, 128])
@pytest.mark.parametrize("N,K,K2", [(8512, 2048, 0), (2048, 4096, 0), (8512, 2048, 8), (264, 1000, 0), (16384, 2048, 0), (1000, 520, 24)])
def test_gemm_decode_shapes_stream_the_weights(M, N, K, K2):
    """M = batch <= 128 rows of K-major operands take the weight-streaming kernel (csrc/gemm_skinny.cu: weight rows on the MMA
    M dimension, K split over a thread-block cluster, partial tiles reduced through DSMEM): in_proj / out_proj of Mamba2.step
    at d_model = 2048, ragged N / K, every cluster size (1, 2, 4, 8), with and without the LoRA pair - against fp32 PyTorch and
    against the tile kernel of gemm_tc.cu on the same operands."""
    from omnimamba_b200 import _cabi
    x, w = _rand((M, K), 31), _rand((N, K), 32, 0.05)
    ref = x.float() @ w.float().t()
    extra = ()
    if K2:
        t, bl = _rand((M, K2), 33), _rand((N, K2), 34, 0.1)
        ref = ref + t.float() @ bl.float().t()
        extra = (t.to(DEV), bl.to(DEV))
    xd, wd = x.to(DEV), w.to(DEV)
    lib = _cabi.lib()
    _cabi.reset_launch_count()
    o32 = _cabi.gemm(xd, wd, torch.float32, *extra)
    o16 = _cabi.gemm(xd, wd, torch.bfloat16, *extra)
    try:
        lib.omni_debug_set_gemm_mode(3)       # the same GEMM on the tile kernel
        t32 = _cabi.gemm(xd, wd, torch.float32, *extra)
    finally:
        lib.omni_debug_set_gemm_mode(0)
    torch.cuda.synchronize()
    e32, ex = rel_l2(o32, ref), excess_over_rounding(o16, ref)
    print(f"skinny gemm M={M} N={N} K={K} K2={K2}: fp32-out {e32:.2e}, bf16-out excess {ex:.2e}, vs tile kernel {rel_l2(o32, t32):.2e}")
    assert e32 <= 1e-5 and ex <= 1e-4
    assert rel_l2(o32, t32) <= 1e-5
    assert torch.equal(o16.cpu(), o32.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("M", [1, 5, 64, 128])
@pytest.mark.parametrize("N,K,K2", [(8512, 2048, 0), (2048, 4096, 0), (8512, 2048, 8), (260, 132, 0), (1000, 1000, 4)])
def test_gemm_f32_decode_is_fp32_accurate(M, N, K, K2):
    """omni_gemm_f32_decode: fp32 operands on the tensor cores through the error-compensated 3xTF32 split (csrc/gemm_skinny.cu)
    - the projections of Mamba2.step when the model runs in fp32, as inference_t2i.py does.  Against an fp64 reference the
    result must be as good as an fp32 GEMM (north-star tolerance 1e-5; asserted 2e-6), far from plain TF32 (~5e-4)."""
    from omnimamba_b200 import _cabi
    g = torch.Generator().manual_seed(N + K + M)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.05
    ref = x.double() @ w.double().t()
    extra = ()
    if K2:
        t, bl = torch.randn(M, K2, generator=g), torch.randn(N, K2, generator=g) * 0.1
        ref = ref + t.double() @ bl.double().t()
        extra = (t.to(DEV), bl.to(DEV))
    xd, wd = x.to(DEV), w.to(DEV)
    assert _cabi.gemm_f32_decode_ok(xd, wd, *extra)
    _cabi.reset_launch_count()
    out = _cabi.gemm_f32_decode(xd, wd, *extra)
    torch.cuda.synchronize()
    assert _cabi.launch_count() == 1
    e = rel_l2(out.double().cpu(), ref)
    e_torch = rel_l2((xd @ wd.t() + (extra[0] @ extra[1].t() if K2 else 0)).double().cpu(), ref)
    print(f"gemm_f32_decode M={M} N={N} K={K} K2={K2}: rel_l2 vs fp64 {e:.2e} (torch fp32 matmul: {e_torch:.2e})")
    assert e <= 2e-6, e


def test_linear_fp32_decode_uses_the_tf32x3_kernel():
    """interface.gemm.linear / lora_linear on fp32 CUDA tensors without gradients at a decode shape run on libomnissm, and
    agree with F.linear to fp32 accuracy; with gradients enabled the call stays on torch (no backward for this kernel)."""
    from omnimamba_b200 import _cabi
    from omnimamba_b200.interface.gemm import linear, lora_linear
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(64, 1, 2048, generator=g).to(DEV), (torch.randn(8512, 2048, generator=g) * 0.02).to(DEV)
    la, lb = (torch.randn(8, 2048, generator=g) * 0.02).to(DEV), (torch.randn(8512, 8, generator=g) * 0.05).to(DEV)
    with torch.no_grad():
        _cabi.reset_launch_count()
        y = linear(x, w)
        y2 = lora_linear(x, w, None, la, lb, 4.0)
        n = _cabi.launch_count()
    assert n == 2
    ref = F.linear(x.double(), w.double())
    ref2 = ref + F.linear(F.linear(x.double(), la.double()), lb.double()) * 4.0
    assert rel_l2(y.double(), ref) <= 2e-6 and rel_l2(y2.double(), ref2) <= 2e-6
    _cabi.reset_launch_count()
    xg = x.clone().requires_grad_()
    linear(xg, w).sum().backward()
    assert _cabi.launch_count() == 0 and xg.grad is not None


def test_gemm_second_operand_pair_is_lora():
    """out = x W^T + t B^T in one accumulator (the LoRA branch of in_proj, lora.py:263-279): d_model=2048 -> 8512, r = 8."""
    from omnimamba_b200 import _cabi
    M, K, N, r = 700, 2048, 8512, 8
    x, w = _rand((M, K), 3), _rand((N, K), 4, 0.02)
    t, bl = _rand((M, r), 5), _rand((N, r), 6, 0.1)
    ref = x.float() @ w.float().t() + t.float() @ bl.float().t()
    lib = _cabi.lib()
    for mode in (1, 2):
        try:
            lib.omni_debug_set_gemm_mode(mode)
            out = _cabi.gemm(x.to(DEV), w.to(DEV), torch.float32, t.to(DEV), bl.to(DEV))
            torch.cuda.synchronize()
        finally:
            lib.omni_debug_set_gemm_mode(0)
        e = rel_l2(out, ref)
        print(f"gemm + LoRA pair (tile mode {mode}): {e:.2e}")
        assert e <= 1e-5, e


def test_linear_autograd_matches_torch():
    """interface.gemm.linear / lora_linear: forward, dgrad, wgrad (all on the tcgen05 kernel) against F.linear in fp32."""
    from omnimamba_b200.interface.gemm import linear, lora_linear
    from omnimamba_b200 import _cabi
    M, K, N, r, s = 2 * 329, 2048, 8512, 8, 4.0
    x, w = _rand((2, 329, K), 7), _rand((N, K), 8, 0.02)
    la, lb = _rand((r, K), 9, 0.02), _rand((N, r), 10, 0.05)
    dy = _rand((2, 329, N), 11)
    leaves = [t.float().requires_grad_() for t in (x, w, la, lb)]
    yr = F.linear(leaves[0], leaves[1]) + F.linear(F.linear(leaves[0], leaves[2]), leaves[3]) * s
    yr.backward(dy.float())
    d = [t.to(DEV).requires_grad_() for t in (x, w, la, lb)]
    _cabi.reset_launch_count()
    y = lora_linear(d[0], d[1], None, d[2], d[3], s)
    y.backward(dy.to(DEV))
    torch.cuda.synchronize()
    assert _cabi.launch_count() >= 6, "forward (2) + backward (>= 4) GEMMs must run on libomnissm"
    # the rank-r intermediate t = s x A^T is rounded to bf16 before the up-projection (as under autocast upstream)
    assert excess_over_rounding(y, yr) <= 2e-3
    for name, a, b in zip(("dx", "dW", "dA", "dB"), d, leaves):
        e = excess_over_rounding(a.grad, b.grad)
        print(f"lora_linear {name}: excess {e:.2e} plain {rel_l2(a.grad, b.grad):.2e}")
        assert e <= 5e-3, (name, e)
    y2 = linear(d[0].detach(), d[1].detach())
    assert excess_over_rounding(y2, F.linear(x.float(), w.float())) <= 1e-4


@pytest.mark.parametrize("M,V,d,block", [(700, 16384, 2048, 256), (333, 1000, 256, 128), (64, 50288, 512, 64)])
def test_linear_cross_entropy_matches_torch(M, V, d, block):
    """Head GEMM + shifted-label cross-entropy in row blocks (interface.linear_ce): loss, dh and dW against
    nn.CrossEntropyLoss on fp32 logits of the same bf16-valued operands; ignored labels included; the logits GEMMs, the
    softmax statistics and the gradient (softmax - onehot) all run in libomnissm."""
    from omnimamba_b200 import _cabi
    from omnimamba_b200.interface.linear_ce import linear_cross_entropy
    g = torch.Generator().manual_seed(M + V)
    h, w = _rand((M, d), 21), _rand((V, d), 22, 0.05)
    labels = torch.randint(0, V, (M,), generator=g)
    labels[::5] = -100
    hr, wr = h.float().requires_grad_(), w.float().requires_grad_()
    ref = F.cross_entropy(hr @ wr.t(), labels, ignore_index=-100)
    ref.backward()
    hd, wd = h.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    _cabi.reset_launch_count()
    loss = linear_cross_entropy(hd, wd, labels.to(DEV), block=block)
    loss.backward()
    torch.cuda.synchronize()
    nb = (M + block - 1) // block
    assert _cabi.launch_count() >= 6 * nb - 2, "per row block: 2 logits GEMMs, 2 CE kernels, dgrad and wgrad GEMMs"
    e = abs(loss.item() - ref.item()) / abs(ref.item())
    edh, edw = excess_over_rounding(hd.grad, hr.grad), rel_l2(wd.grad, wr.grad)
    print(f"linear_cross_entropy M={M} V={V} d={d}: loss rel err {e:.2e}, dh excess {edh:.2e} (plain {rel_l2(hd.grad, hr.grad):.2e}), dW {edw:.2e}")
    assert e <= 1e-5
    assert edh <= 5e-3 and edw <= 5e-3     # the bf16 rounding of (softmax - onehot) before the two gradient GEMMs
