// Row-wise softmax cross-entropy over a block of fp32 logits: the loss half of the fused head GEMM + loss
// (the reference computes img_head / lm_head logits for every position and hands them to
// nn.CrossEntropyLoss, /root/reference/models/mamba_vlm.py:96-100, models/omnimamba.py:276-279).  The logits of a row block
// come from the tcgen05 GEMM (fp32 out); these two kernels turn them into (lse, loss) and, in the backward, into the bf16
// gradient (softmax - onehot) * scale that feeds the dgrad / wgrad GEMMs - one pass over the block each, instead of
// logsumexp + gather + exp + scatter + mul + cast as separate elementwise passes.
#include "common.cuh"

namespace omni {
namespace {

constexpr int kCeThreads = 256;

__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

// one CTA per row: lse[r] = logsumexp(logits[r, :]); loss[r] = lse - logits[r, label] (0 for ignored rows)
__global__ void __launch_bounds__(kCeThreads) ce_fwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                                            float* __restrict__ lse, float* __restrict__ loss, int V, int64_t ignore) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float* row = logits + r * ld;
  const bool vec = (V % 4 == 0) && (ld % 4 == 0);
  float m = -INFINITY;
  if (vec) {
    for (int i = threadIdx.x; i < V / 4; i += kCeThreads) {
      const float4 v = reinterpret_cast<const float4*>(row)[i];
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
  } else {
    for (int i = threadIdx.x; i < V; i += kCeThreads) m = fmaxf(m, row[i]);
  }
  m = block_max(m, red);
  float s = 0.f;
  if (vec) {
    for (int i = threadIdx.x; i < V / 4; i += kCeThreads) {   // (second pass: the 64 KB row is an L1 / L2 hit)
      const float4 v = reinterpret_cast<const float4*>(row)[i];
      s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    }
  } else {
    for (int i = threadIdx.x; i < V; i += kCeThreads) s += __expf(row[i] - m);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float l = m + __logf(s);
    lse[r] = l;
    const int64_t lab = labels[r];
    loss[r] = (lab == ignore || lab < 0 || lab >= V) ? 0.f : l - row[lab];
  }
}

// grad[r, v] = (exp(logits[r, v] - lse[r]) - [v == label_r]) * scale   (bf16; zero rows for ignored labels)
__global__ void __launch_bounds__(kCeThreads) ce_bwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                                            const float* __restrict__ lse, const float* __restrict__ scale,
                                                            __nv_bfloat16* __restrict__ grad, int64_t gld, int V, int64_t ignore) {
  const int64_t r = blockIdx.x;
  const float* row = logits + r * ld;
  __nv_bfloat16* g = grad + r * gld;
  const int64_t lab = labels[r];
  const bool live = !(lab == ignore || lab < 0 || lab >= V);
  const float sc = live ? scale[0] : 0.f, l = lse[r];
  if ((V % 4 == 0) && (ld % 4 == 0) && (gld % 4 == 0)) {
    for (int i = threadIdx.x; i < V / 4; i += kCeThreads) {
      const float4 v = reinterpret_cast<const float4*>(row)[i];
      float p[4] = {__expf(v.x - l), __expf(v.y - l), __expf(v.z - l), __expf(v.w - l)};
      if (live && lab >= 4 * i && lab < 4 * i + 4) p[lab - 4 * i] -= 1.f;
      uint2 o;
      __nv_bfloat162 a = __floats2bfloat162_rn(p[0] * sc, p[1] * sc), b = __floats2bfloat162_rn(p[2] * sc, p[3] * sc);
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(g)[i] = o;
    }
  } else {
    for (int i = threadIdx.x; i < V; i += kCeThreads) g[i] = __float2bfloat16_rn((__expf(row[i] - l) - (live && i == lab ? 1.f : 0.f)) * sc);
  }
}

int check_ce(const omni_softmax_ce_params_t* p, int64_t& M, int64_t& V) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t& lg = p->logits;
  OMNI_CHECK(present(lg) && lg.ndim == 2 && lg.dtype == OMNI_F32 && (lg.shape[1] <= 1 || lg.stride[1] == 1), OMNI_BAD_SHAPE,
             "softmax_ce: logits must be fp32 (M, V) with contiguous rows");
  M = lg.shape[0]; V = lg.shape[1];
  OMNI_CHECK(V >= 1 && V < (1ll << 31), OMNI_BAD_SHAPE, "softmax_ce: bad vocabulary size");
  OMNI_CHECK(present(p->labels) && shape_is(p->labels, 1, M) && p->labels.dtype == OMNI_I64 && (M <= 1 || p->labels.stride[0] == 1),
             OMNI_BAD_SHAPE, "softmax_ce: labels must be contiguous int64 (M)");
  OMNI_CHECK(present(p->lse) && shape_is(p->lse, 1, M) && p->lse.dtype == OMNI_F32 && (M <= 1 || p->lse.stride[0] == 1), OMNI_BAD_SHAPE,
             "softmax_ce: lse must be contiguous fp32 (M)");
  return OMNI_OK;
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_softmax_ce_fwd(const omni_softmax_ce_params_t* p, void* stream) {
  int64_t M = 0, V = 0;
  if (int rc = check_ce(p, M, V)) return rc;
  OMNI_CHECK(present(p->loss) && shape_is(p->loss, 1, M) && p->loss.dtype == OMNI_F32 && (M <= 1 || p->loss.stride[0] == 1), OMNI_BAD_SHAPE,
             "softmax_ce: loss must be contiguous fp32 (M)");
  if (M == 0) return OMNI_OK;
  ce_fwd_kernel<<<(unsigned)M, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(p->logits.data), M > 1 ? p->logits.stride[0] : V, static_cast<const int64_t*>(p->labels.data),
      static_cast<float*>(p->lse.data), static_cast<float*>(p->loss.data), (int)V, p->ignore_index);
  OMNI_CUDA_LAUNCH_CHECK("ce_fwd_kernel");
  return OMNI_OK;
}

extern "C" int omni_softmax_ce_bwd(const omni_softmax_ce_params_t* p, void* stream) {
  int64_t M = 0, V = 0;
  if (int rc = check_ce(p, M, V)) return rc;
  OMNI_CHECK(present(p->grad) && shape_is(p->grad, 2, M, V) && p->grad.dtype == OMNI_BF16 && (V <= 1 || p->grad.stride[1] == 1), OMNI_BAD_SHAPE,
             "softmax_ce: grad must be bf16 (M, V) with contiguous rows");
  OMNI_CHECK(present(p->scale) && p->scale.dtype == OMNI_F32, OMNI_BAD_SHAPE, "softmax_ce: scale must be a device fp32 scalar");
  if (M == 0) return OMNI_OK;
  ce_bwd_kernel<<<(unsigned)M, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(p->logits.data), M > 1 ? p->logits.stride[0] : V, static_cast<const int64_t*>(p->labels.data),
      static_cast<const float*>(p->lse.data), static_cast<const float*>(p->scale.data), static_cast<__nv_bfloat16*>(p->grad.data),
      M > 1 ? p->grad.stride[0] : V, (int)V, p->ignore_index);
  OMNI_CUDA_LAUNCH_CHECK("ce_bwd_kernel");
  return OMNI_OK;
}
