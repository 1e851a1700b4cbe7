// Single-token Mamba-2 layer core for decode: conv-state update + SiLU, selective state update and gated RMSNorm in ONE
// kernel (the three launches Mamba2.step makes between in_proj and out_proj - causal_conv1d_update,
// selective_state_update, RMSNormGated - inside the CUDA graph of /root/reference/models/stage2/generation.py:383-431).
//
//   zxbcdt (B, 2 dim + 2 N + H) = [z | xBC | dt]   (dim = H * 64, ngroups = 1, d_state N = 128)
//   conv_state[b, c, :] <- shift left, append xBC[b, c];  u = silu(bias + <w[c, :], conv_state[b, c, :]>)      (in place)
//   dt' = softplus(dt + dt_bias);  S[b,h,p,:] <- S exp(dt' A_h) + dt' x_p B;  y_p = <S, C> + D_h x_p          (in place)
//   out[b, :] = rmsnorm(y * silu(z)) * w     (one group over the whole row: the reduction crosses all heads of a sequence)
//
// One thread-block CLUSTER of 8 CTAs per sequence, one CTA per 8 heads, one warp per head.  The conv outputs x, B, C and
// the scan outputs y never leave shared memory; the row-wide sum of squares is reduced over the cluster through
// distributed shared memory.  B and C are shared by all heads: every CTA computes them from the OLD conv state, and rank 0
// writes their shifted state back only after a cluster barrier.  HBM traffic = the state (read + write) + O(d_in_proj).
#include <cooperative_groups.h>

#include <mutex>

#include "umma.cuh"

namespace cg = cooperative_groups;

namespace omni {
namespace {

#ifndef OMNI_DEC_HEADS
#define OMNI_DEC_HEADS 8   // heads (= warps) per CTA; 64 / OMNI_DEC_HEADS CTAs form the cluster of a sequence (4 -> 16, non-portable size)
#endif
constexpr int kP = 64, kN = 128, kHeadsPerCta = OMNI_DEC_HEADS, kCluster = 64 / OMNI_DEC_HEADS, kDecThreads = 32 * OMNI_DEC_HEADS;
constexpr int kBcPerThread = 2 * kN / kDecThreads;   // B / C conv channels per thread (every CTA computes all 256)
constexpr int kMaxW = 4;

struct DecArgs {
  const void* zx; void* conv_state; const void* conv_w; const void* conv_b; void* state;
  const float* A; const void* D; const void* dt_bias; const void* norm_w; void* out;
  int64_t zx_b, cs_b, cs_c, cs_k, cw_c, cw_k, st_b, st_h, st_p, o_b;
  int B, H, W;
  int io_dtype, cs_dtype, cw_dtype, cb_dtype, D_dtype, db_dtype, nw_dtype;
  float eps;
};

template <typename TS> struct Raw4;
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
template <> struct Raw4<__half> { using type = uint2; };
template <typename TS> __device__ __forceinline__ void unpack4(const typename Raw4<TS>::type& r, float (&o)[4]);
template <> __device__ __forceinline__ void unpack4<float>(const float4& r, float (&o)[4]) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
template <> __device__ __forceinline__ void unpack4<__nv_bfloat16>(const uint2& r, float (&o)[4]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack4<__half>(const uint2& r, float (&o)[4]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
  const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename TS> __device__ __forceinline__ typename Raw4<TS>::type pack4(const float (&o)[4]);
template <> __device__ __forceinline__ float4 pack4<float>(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
template <> __device__ __forceinline__ uint2 pack4<__nv_bfloat16>(const float (&o)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
template <> __device__ __forceinline__ uint2 pack4<__half>(const float (&o)[4]) {
  __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

#ifndef OMNI_DEC_OCC
#define OMNI_DEC_OCC 4     // CTAs per SM the register budget is cut for (4 x 256 threads x 64 registers)
#endif
#ifndef OMNI_DEC_ROWS16
#define OMNI_DEC_ROWS16 8   // state rows per warp step for a 16-bit state (fp32: half of it)
#endif
// Occupancy: 64 sequences x 8 CTAs = 512 CTAs.  At 2 CTAs per SM (the 32-row steps of the first version: 128 registers) they
// run as 1.73 waves on 296 slots - two wave times for 1.73 waves of work - and every CTA's conv phase / cluster barriers /
// norm phase leave its slot's share of the HBM stream idle.  At 64 registers four CTAs fit per SM: ALL 512 CTAs are resident
// at once (592 slots), there is no second wave, and the phases of one CTA hide behind the state streams of the three others.
// Measured at batch 64, bf16 state (scripts/bench_decode_core.py, same box): 2 CTAs x 32-row steps 38.0 us, 3 x 16 47.0, 4 x 16
// 41.2 (spills), 4 x 8 33.1, 4 x 4 33.6, 5 x 8 38.2 (spills), 5 x 4 34.4, 6 x 4 35.6;  fp32 state: 59.4 -> 55.7 us.
#ifndef OMNI_DEC_BULK
#define OMNI_DEC_BULK 0    // 1: the state rows arrive through bulk async copies into a per-warp ring in shared memory (no registers held
#endif                     // by loads in flight); 0: plain vector loads into registers.  Measured: 34.7 vs 33.4 us (bf16), 57.7 vs 56.1
                           // (fp32) - bytes in flight are not what limits the kernel, the 3-or-4-CTAs-per-SM granularity is
constexpr int kDecRing = 3;                  // ring slots per warp
constexpr int kDecSlotBytes = 2048;          // one step of R rows (8 x 256 B for a 16-bit state, 4 x 512 B for fp32)
constexpr int kDecDynSmem = OMNI_DEC_BULK ? kDecThreads / 32 * kDecRing * kDecSlotBytes : 0;   // 48 KB per CTA: four CTAs per SM
template <typename TS>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kDecThreads, OMNI_DEC_OCC) mamba2_decode_core_kernel(DecArgs a) {
  constexpr int R = sizeof(TS) == 4 ? OMNI_DEC_ROWS16 / 2 : OMNI_DEC_ROWS16;   // state rows per warp step
  static_assert(!OMNI_DEC_BULK || R * kN * (int)sizeof(TS) == kDecSlotBytes, "a ring slot holds one step");
  using Raw = typename Raw4<TS>::type;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();           // 8 heads of the sequence
  const int b = blockIdx.x / kCluster;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dim = a.H * kP, conv_dim = dim + 2 * kN;
  __shared__ float xs[kHeadsPerCta * kP], Bs[kN], Cs[kN], ys[kHeadsPerCta * kP], red[32];
  __shared__ float part;

  pdl_trigger();   // (PDL, common.cuh: out_proj may start streaming its weights while this kernel runs)
  // (the first step's state rows do not depend on anything computed here - nor on in_proj, the previous kernel: the state
  // was last written by this layer's own launch of the previous token - so they are requested before pdl_wait(), and arrive
  // while the conv phase and the first cluster barrier run)
  const int h_w = cr * kHeadsPerCta + warp;
  TS* const sbase = static_cast<TS*>(a.state) + (int64_t)b * a.st_b + (int64_t)h_w * a.st_h + lane * 4;
#if OMNI_DEC_BULK
  // State stream, bulk version: the 64 rows of this warp's head are contiguous (host-checked), 2 KB steps of R rows travel
  // global -> shared memory as cp.async.bulk copies into a ring of kDecRing slots per warp, each completing on its own
  // mbarrier.  Bytes in flight cost no registers: 4 CTAs x 8 warps x 3 slots x 2 KB = 192 KB per SM can be outstanding (the
  // register version holds 64 KB), which is what the ~1.5 us HBM latency needs at this bandwidth.  The updated rows go back
  // with ordinary vector stores, so a slot is free again as soon as the warp has read it.
  extern __shared__ __align__(128) uint8_t dec_ring[];
  __shared__ __align__(8) uint64_t ring_bar[kDecThreads / 32][kDecRing];
  uint8_t* const my_ring = dec_ring + warp * (kDecRing * kDecSlotBytes);
  const TS* const gslice = static_cast<const TS*>(a.state) + (int64_t)b * a.st_b + (int64_t)h_w * a.st_h;   // 64 x 128 contiguous
  constexpr int kSteps = kP / R;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kDecRing; ++k) umma::mbar_init(&ring_bar[warp][k], 1);
    umma::mbar_fence_init();
#pragma unroll
    for (int k = 0; k < kDecRing; ++k) {   // the state does not depend on the previous kernel: requested before pdl_wait()
      umma::mbar_expect_tx(&ring_bar[warp][k], kDecSlotBytes);
      umma::bulk_load_1d(my_ring + k * kDecSlotBytes, gslice + (int64_t)k * R * kN, kDecSlotBytes, &ring_bar[warp][k]);
    }
  }
  __syncwarp();
#else
  // (measured and not kept: a different starting step per warp, against DRAM channel camping - 32.7 vs 33.2 us, noise)
  constexpr int rot = 0;
  Raw raw[R];
#pragma unroll
  for (int r = 0; r < R; ++r) raw[r] = *reinterpret_cast<const Raw*>(sbase + (int64_t)(rot + r) * a.st_p);
#endif

  pdl_wait();      // zxbcdt comes from the previous kernel (in_proj)
  // ---- 1. conv-state update + SiLU: 2 x channels per thread (this CTA's heads) + 1 B/C channel per thread (all CTAs) ----
  const char* zrow = static_cast<const char*>(a.zx);
  auto zx_at = [&](int col) { return ld_any(a.zx, a.io_dtype, (int64_t)b * a.zx_b + col); };
  float bc_keep[kBcPerThread][kMaxW];  // the shifted B / C state of this thread's channels (written back by rank 0 after the barrier)
  (void)zrow;
#pragma unroll
  for (int which = 0; which < 2 + kBcPerThread; ++which) {
    // which 0, 1: x channels (cr * 2 T + tid * 2 + which);  2 ..: B / C channels dim + tid + T (which - 2)
    const int bci = tid + kDecThreads * (which - 2);
    const int ch = which < 2 ? cr * (kHeadsPerCta * kP) + tid * 2 + which : dim + bci;
    float st[kMaxW];
#pragma unroll
    for (int k = 0; k < kMaxW; ++k)
      st[k] = (k + 1 < a.W) ? ld_any(a.conv_state, a.cs_dtype, (int64_t)b * a.cs_b + (int64_t)ch * a.cs_c + (int64_t)(k + 1) * a.cs_k) : 0.f;
    st[a.W - 1 < kMaxW ? a.W - 1 : kMaxW - 1] = zx_at(dim + ch);   // (W <= 4: the appended sample)
    float acc = a.conv_b ? ld_any(a.conv_b, a.cb_dtype, ch) : 0.f;
#pragma unroll
    for (int k = 0; k < kMaxW; ++k)
      if (k < a.W) acc = fmaf(ld_any(a.conv_w, a.cw_dtype, (int64_t)ch * a.cw_c + (int64_t)k * a.cw_k), st[k], acc);
    const float u = silu_f(acc);
    if (which < 2) {
      xs[tid * 2 + which] = u;
#pragma unroll
      for (int k = 0; k < kMaxW; ++k)
        if (k < a.W) st_any(a.conv_state, a.cs_dtype, (int64_t)b * a.cs_b + (int64_t)ch * a.cs_c + (int64_t)k * a.cs_k, st[k]);
    } else {
      if (bci < kN) Bs[bci] = u; else Cs[bci - kN] = u;
#pragma unroll
      for (int k = 0; k < kMaxW; ++k) bc_keep[which - 2][k] = st[k];
    }
  }
  __syncthreads();
  cluster.sync();   // every CTA of the sequence has read the old B / C conv state
  if (cr == 0) {
#pragma unroll
    for (int j = 0; j < kBcPerThread; ++j) {
      const int ch = dim + tid + kDecThreads * j;
#pragma unroll
      for (int k = 0; k < kMaxW; ++k)
        if (k < a.W) st_any(a.conv_state, a.cs_dtype, (int64_t)b * a.cs_b + (int64_t)ch * a.cs_c + (int64_t)k * a.cs_k, bc_keep[j][k]);
    }
  }

  // ---- 2. selective state update: warp = head, R rows per step, lane = 4 consecutive n ----------------------------------
  {
    const int h = cr * kHeadsPerCta + warp, n = lane * 4;
    float dtv = zx_at(dim + conv_dim + h) + (a.dt_bias ? ld_any(a.dt_bias, a.db_dtype, h) : 0.f);
    dtv = softplus_f(dtv);
    const float dA = __expf(dtv * a.A[h]);
    const float Dh = a.D ? ld_any(a.D, a.D_dtype, h) : 0.f;
    float Bv[4], Cv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { Bv[e] = Bs[n + e] * dtv; Cv[e] = Cs[n + e]; }   // (dt folded into B: S += x_p (dt B))
#pragma unroll 1
    for (int pstep = 0; pstep < kP; pstep += R) {
#if OMNI_DEC_BULK
      const int p0 = pstep;
      const int step = p0 / R, slot = step % kDecRing;
      while (!umma::mbar_try_wait(&ring_bar[warp][slot], (uint32_t)((step / kDecRing) & 1))) {
      }
      Raw raw[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        raw[r] = *reinterpret_cast<const Raw*>(my_ring + slot * kDecSlotBytes + (r * kN + lane * 4) * (int)sizeof(TS));
      __syncwarp();   // every lane has read the slot: refill it with the step kDecRing ahead
      if (lane == 0 && step + kDecRing < kSteps) {
        umma::mbar_expect_tx(&ring_bar[warp][slot], kDecSlotBytes);
        umma::bulk_load_1d(my_ring + slot * kDecSlotBytes, gslice + (int64_t)(step + kDecRing) * R * kN, kDecSlotBytes, &ring_bar[warp][slot]);
      }
#else
      const int p0 = (pstep + rot) & (kP - 1);
      if (pstep > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) raw[r] = *reinterpret_cast<const Raw*>(sbase + (int64_t)(p0 + r) * a.st_p);
      }
#endif
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float xv = xs[warp * kP + p0 + r];
        float S[4];
        unpack4<TS>(raw[r], S);
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          S[e] = fmaf(S[e], dA, xv * Bv[e]);
          sum = fmaf(S[e], Cv[e], sum);
        }
        *reinterpret_cast<Raw*>(sbase + (int64_t)(p0 + r) * a.st_p) = pack4<TS>(S);
        acc[r] = sum;
      }
      // transposing butterfly (as ssu_tied_kernel): lane l ends with the complete sum of row l (R = 32) or l >> 1 (R = 16)
#pragma unroll
      for (int half = R / 2, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
          if (j < half) {
            const float send = up ? acc[j] : acc[j + half];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            acc[j] = (up ? acc[j + half] : acc[j]) + recv;
          }
        }
      }
      // (R < 32: the butterfly stops with 32 / R lanes holding partial sums of the same row)
      float y = acc[0];
#pragma unroll
      for (int off = 1; off < 32 / R; off <<= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
      const int row = lane / (32 / R);
      if (lane % (32 / R) == 0) ys[warp * kP + p0 + row] = fmaf(xs[warp * kP + p0 + row], Dh, y);
    }
  }
  __syncthreads();

  // ---- 3. gate, row-wide sum of squares over the cluster, norm weight --------------------------------------------------
  float g[2];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int col = tid + kDecThreads * k;                       // column inside this CTA's 512
    const float z = zx_at(cr * (kHeadsPerCta * kP) + col);
    g[k] = ys[col] * silu_f(z);
    ss = fmaf(g[k], g[k], ss);
  }
  ss = block_sum(ss, red);
  if (tid == 0) part = ss;
  cluster.sync();
  float total = 0.f;
#pragma unroll
  for (int r = 0; r < kCluster; ++r) total += *cluster.map_shared_rank(&part, r);
  cluster.sync();   // nobody leaves (and frees its shared memory) while a peer may still read `part`
  const float rstd = rsqrtf(total / (float)dim + a.eps);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int col = cr * (kHeadsPerCta * kP) + tid + kDecThreads * k;
    st_any(a.out, a.io_dtype, (int64_t)b * a.o_b + col, g[k] * rstd * ld_any(a.norm_w, a.nw_dtype, col));
  }
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_mamba2_decode_core(const omni_mamba2_decode_core_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t &zx = p->zxbcdt, &cs = p->conv_state, &st = p->ssm_state, &o = p->out;
  OMNI_CHECK(present(st) && st.ndim == 4 && is_float_dtype(st.dtype) && st.shape[2] == kP && st.shape[3] == kN && st.stride[3] == 1 &&
                 st.stride[2] % 4 == 0 && st.stride[1] % 4 == 0 && st.stride[0] % 4 == 0 && aligned16(st.data),
             OMNI_UNSUPPORTED, "decode_core: ssm_state must be (B, H, 64, 128) with contiguous, 16-byte aligned rows");
  OMNI_CHECK(!OMNI_DEC_BULK || st.stride[2] == kN, OMNI_UNSUPPORTED,
             "decode_core: the (64, 128) state of a head must be contiguous (it is streamed with bulk copies)");
  const int64_t Bsz = st.shape[0], H = st.shape[1], dim = H * kP, conv_dim = dim + 2 * kN;
  OMNI_CHECK(H % kHeadsPerCta == 0 && H / kHeadsPerCta == kCluster, OMNI_UNSUPPORTED,
             "decode_core: built for nheads = 64 (d_model = 2048): one cluster of 8 CTAs x 8 heads per sequence");
  OMNI_CHECK(present(zx) && shape_is(zx, 2, Bsz, 2 * dim + 2 * kN + H) && is_float_dtype(zx.dtype) && zx.stride[1] == 1, OMNI_BAD_SHAPE,
             "decode_core: zxbcdt must be (B, 2 dim + 2 N + H) with contiguous rows (ngroups = 1)");
  OMNI_CHECK(present(cs) && cs.ndim == 3 && cs.shape[0] == Bsz && cs.shape[1] == conv_dim && cs.shape[2] >= 2 && cs.shape[2] <= kMaxW &&
                 is_float_dtype(cs.dtype), OMNI_BAD_SHAPE, "decode_core: conv_state must be (B, dim + 2 N, W), W <= 4");
  const int64_t W = cs.shape[2];
  OMNI_CHECK(present(p->conv_weight) && shape_is(p->conv_weight, 2, conv_dim, W) && is_float_dtype(p->conv_weight.dtype), OMNI_BAD_SHAPE,
             "decode_core: conv_weight must be (dim + 2 N, W)");
  OMNI_CHECK(!present(p->conv_bias) || (shape_is(p->conv_bias, 1, conv_dim) && p->conv_bias.stride[0] == 1), OMNI_BAD_SHAPE,
             "decode_core: conv_bias must be contiguous (dim + 2 N)");
  OMNI_CHECK(present(p->A) && shape_is(p->A, 1, H) && p->A.dtype == OMNI_F32 && p->A.stride[0] == 1, OMNI_BAD_SHAPE,
             "decode_core: A must be contiguous fp32 (H)");
  auto vec = [&](const omni_tensor_t& t, int64_t n) { return !present(t) || (shape_is(t, 1, n) && is_float_dtype(t.dtype) && t.stride[0] == 1); };
  OMNI_CHECK(vec(p->D, H) && vec(p->dt_bias, H) && present(p->norm_weight) && vec(p->norm_weight, dim), OMNI_BAD_SHAPE,
             "decode_core: D, dt_bias (H) and norm_weight (dim) must be contiguous");
  OMNI_CHECK(present(o) && shape_is(o, 2, Bsz, dim) && o.dtype == zx.dtype && o.stride[1] == 1, OMNI_BAD_SHAPE,
             "decode_core: out must be (B, dim), dtype of zxbcdt");
  if (Bsz == 0) return OMNI_OK;
  DecArgs a{};
  a.zx = zx.data; a.conv_state = cs.data; a.conv_w = p->conv_weight.data; a.conv_b = p->conv_bias.data; a.state = st.data;
  a.A = static_cast<const float*>(p->A.data); a.D = p->D.data; a.dt_bias = p->dt_bias.data; a.norm_w = p->norm_weight.data; a.out = o.data;
  a.zx_b = zx.stride[0]; a.cs_b = cs.stride[0]; a.cs_c = cs.stride[1]; a.cs_k = cs.stride[2];
  a.cw_c = p->conv_weight.stride[0]; a.cw_k = p->conv_weight.stride[1];
  a.st_b = st.stride[0]; a.st_h = st.stride[1]; a.st_p = st.stride[2]; a.o_b = o.stride[0];
  a.B = (int)Bsz; a.H = (int)H; a.W = (int)W;
  a.io_dtype = zx.dtype; a.cs_dtype = cs.dtype; a.cw_dtype = p->conv_weight.dtype; a.cb_dtype = p->conv_bias.dtype;
  a.D_dtype = p->D.dtype; a.db_dtype = p->dt_bias.dtype; a.nw_dtype = p->norm_weight.dtype;
  a.eps = p->eps;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (kDecDynSmem > 0) {  // static + dynamic shared memory exceed the 48 KB default: opt in once
    static std::once_flag once_smem;
    std::call_once(once_smem, [] {
      cudaFuncSetAttribute(mamba2_decode_core_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecDynSmem);
      cudaFuncSetAttribute(mamba2_decode_core_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecDynSmem);
      cudaFuncSetAttribute(mamba2_decode_core_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecDynSmem);
    });
  }
  if (kCluster > 8) {  // 16-CTA clusters are a non-portable size: opt in once
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(mamba2_decode_core_kernel<float>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(mamba2_decode_core_kernel<__nv_bfloat16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(mamba2_decode_core_kernel<__half>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    });
  }
  const unsigned grid = (unsigned)(Bsz * kCluster);
  switch (st.dtype) {   // (cluster dims are compile-time: __cluster_dims__)
    case OMNI_F32: launch_pdl(kPdlCore, mamba2_decode_core_kernel<float>, dim3(grid), dim3(kDecThreads), kDecDynSmem, s, 1, a); break;
    case OMNI_BF16: launch_pdl(kPdlCore, mamba2_decode_core_kernel<__nv_bfloat16>, dim3(grid), dim3(kDecThreads), kDecDynSmem, s, 1, a); break;
    default: launch_pdl(kPdlCore, mamba2_decode_core_kernel<__half>, dim3(grid), dim3(kDecThreads), kDecDynSmem, s, 1, a); break;
  }
  OMNI_CUDA_LAUNCH_CHECK("mamba2_decode_core_kernel");
  return OMNI_OK;
}
