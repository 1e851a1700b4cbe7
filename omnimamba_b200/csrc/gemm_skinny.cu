// Weight-streaming bf16 GEMM for decode-shaped projections (M = batch <= 128 rows) on tcgen05 (sm_100a).
//
//   C[M, N] = X[M, K] W[N, K]^T (+ X2[M, K2] W2[N, K2]^T),  K-major operands, fp32 accumulation, bf16 or fp32 C.
//
// The single-token step of the reference (Mamba2.step behind /root/reference/models/stage2/generation.py:383-431) runs
// in_proj (2048 -> 8512) and out_proj (4096 -> 2048) once per layer and token at M = batch (64 in inference_t2i.py): 35 MB
// and 17 MB of weights are read for 2.2 / 1.1 GFLOP - the GEMM is a WEIGHT STREAM, bound by HBM, not by the tensor pipe.
// The tile kernels of gemm_tc.cu cover such a shape with 67 (or 16) CTAs that each walk the whole K dimension, a third of
// the DRAM peak.  This kernel turns the problem around:
//   * the WEIGHT rows are the MMA M dimension (128 rows of W per CTA), the batch is the MMA N dimension (64 or 128);
//   * the K dimension is split over a thread-block CLUSTER (1, 2 or 4 CTAs, chosen so that as many SMs as possible stream
//     weights): CTA r of a cluster accumulates K slice r in TMEM, parks its fp32 partial tile in shared memory, and after
//     one cluster barrier every CTA reduces 1 / ksplit of the batch rows over the cluster's partials through DSMEM
//     (ld.shared::cluster) and writes them, bf16 or fp32, straight into C (rows of 128 consecutive n: coalesced);
//   * per CTA: warp 0 TMA producer (8 stages of W 128 x 64 + X MT x 64, 128B swizzle: ~190 KB in flight per SM),
//     warp 1 MMA issuer, warps 2-5 epilogue / reduction.
// The optional second operand pair (the LoRA branch of the reference's in_proj, lora.py:263-279) is appended as extra K
// steps of cluster rank 0.
#include <mutex>

#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

constexpr int SK_THREADS = 192;
constexpr int SK_BK = 64;
constexpr uint32_t SK_W_BYTES = 128 * SK_BK * 2;           // 16 KB
constexpr uint32_t SK_RING = 196608;                       // stage ring (8 x 24 KB or 6 x 32 KB); reused for the fp32 partial tile
enum { SB_FULL = 0, SB_EMPTY = 8, SB_ACC = 16, SB_COUNT = 17 };
constexpr uint32_t SK_BAR = SK_RING;
constexpr uint32_t SK_TMEMPTR = SK_BAR + SB_COUNT * 8;
constexpr uint32_t SK_SMEM = SK_TMEMPTR + 16;

struct SkinnyArgs {
  int M, N, K1, K2;
  int ksplit;        // cluster size along K
  int ksteps1;       // K steps (of 64) of the first operand pair per cluster rank
  int out_f32;
  void* C; int64_t ldc;
};

__device__ __forceinline__ uint32_t sk_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 sk_ld_dsmem4(uint32_t saddr, uint32_t rank) {  // 16 bytes of CTA `rank`'s shared memory
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}
// Reduce rows [m_lo, m_lo + rows) of the cluster's KS partial tiles ([m][128 n] fp32 at `part` in every CTA) and write them to
// C.  128 threads; a thread owns 4 consecutive n of one row per step, and all KS loads of up to four steps are in flight
// together (the loads are DSMEM round trips: issued one by one they took 15 of the kernel's 20 microseconds).
template <int KS>
__device__ __forceinline__ void sk_reduce_rows(const float* part, int t, int rank, int m_lo, int m_hi, int n0, const SkinnyArgs& a) {
  const int n4 = t & 31, mrow = t >> 5;   // 4 rows per step
  if (n0 + 4 * n4 >= a.N) return;
#pragma unroll 1
  for (int mb = m_lo; mb < m_hi; mb += 16) {
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = mb + 4 * j + mrow;
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_hi) {
        const uint32_t sa = smem_u32(part + m * 128 + 4 * n4);
        float4 v[KS];
#pragma unroll
        for (int r = 0; r < KS; ++r) v[r] = KS > 1 ? sk_ld_dsmem4(sa, (uint32_t)r) : *reinterpret_cast<const float4*>(part + m * 128 + 4 * n4);
#pragma unroll
        for (int r = 0; r < KS; ++r) { acc[j].x += v[r].x; acc[j].y += v[r].y; acc[j].z += v[r].z; acc[j].w += v[r].w; }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = mb + 4 * j + mrow;
      if (m < m_hi) {
        if (a.out_f32) {
          *reinterpret_cast<float4*>(static_cast<float*>(a.C) + (int64_t)m * a.ldc + n0 + 4 * n4) = acc[j];
        } else {
          uint2 o;
          o.x = pack_bf16(acc[j].x, acc[j].y);
          o.y = pack_bf16(acc[j].z, acc[j].w);
          *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(a.C) + (int64_t)m * a.ldc + n0 + 4 * n4) = o;
        }
      }
    }
  }
}
__device__ __forceinline__ bool sk_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// MT = MMA N = padded batch (64 or 128); kStages * (16 KB + MT * 128 B) <= SK_RING
template <int MT, int kStages>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapX1,
                   const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapX2, SkinnyArgs a) {
  constexpr uint32_t X_BYTES = MT * SK_BK * 2, STAGE = SK_W_BYTES + X_BYTES;
  static_assert(kStages * STAGE <= SK_RING && kStages <= 8, "stage ring");
  static_assert(MT * 128 * 4 <= SK_RING, "partial tile");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SK_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SK_TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = a.ksplit > 1 ? sk_cluster_rank() : 0u;
  const int tile = blockIdx.x / a.ksplit, n0 = tile * 128;
  // K steps of this CTA: its slice of the first pair, then (rank 0) all steps of the second pair
  const int nk1 = (a.K1 + SK_BK - 1) / SK_BK;
  const int kb = min((int)rank * a.ksteps1, nk1), ke = min(kb + a.ksteps1, nk1);
  const int n1 = ke - kb, n2 = rank == 0 ? (a.K2 + SK_BK - 1) / SK_BK : 0, nsteps = n1 + n2;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars[SB_FULL + i], 1); mbar_init(&bars[SB_EMPTY + i], 1); }
    mbar_init(&bars[SB_ACC], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, MT < 32 ? 32 : MT);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapX1);
    if (a.K2 > 0) { tma_prefetch_desc(&mapW2); tma_prefetch_desc(&mapX2); }
  }
  __syncwarp();   // (warp 0 ran the single-thread set-up above: converge before the aligned barrier)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;
  pdl_trigger();

  if (warp == 0) {
    // ============ TMA producer ==========================================================================================
    if (lane == 0) {
      auto load_w = [&](int it, int s) {
        uint8_t* st = smem + s * STAGE;
        mbar_expect_tx(&bars[SB_FULL + s], STAGE);
        if (it < n1) tma_load_2d(st, &mapW1, &bars[SB_FULL + s], (kb + it) * SK_BK, n0);
        else tma_load_2d(st, &mapW2, &bars[SB_FULL + s], (it - n1) * SK_BK, n0);
      };
      auto load_x = [&](int it, int s) {
        uint8_t* st = smem + s * STAGE + SK_W_BYTES;
        if (it < n1) tma_load_2d(st, &mapX1, &bars[SB_FULL + s], (kb + it) * SK_BK, 0);
        else tma_load_2d(st, &mapX2, &bars[SB_FULL + s], (it - n1) * SK_BK, 0);
      };
      // PDL (common.cuh): the weights are constants - the first ring of W tiles is requested while the kernel that produces
      // X (add + norm before in_proj, the layer core before out_proj) may still be running; X is only read after pdl_wait()
      const int first = min(nsteps, kStages);
      for (int it = 0; it < first; ++it) load_w(it, it);
      pdl_wait();
      for (int it = 0; it < first; ++it) load_x(it, it);
      for (int it = first; it < nsteps; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&bars[SB_EMPTY + s], ph ^ 1);
        load_w(it, s);
        load_x(it, s);
      }
    }
    __syncwarp();   // (the cluster barrier below is .aligned: the warp must arrive converged)
  } else if (warp == 1) {
    // ============ MMA issuer: D[n 128][m MT] += W[n][k] X[m][k] =========================================================
    const bool leader = sk_elect();
    const uint32_t idesc = make_idesc(128, MT, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      mbar_wait(&bars[SB_FULL + s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * STAGE);
      const uint64_t dW = make_sdesc(sa, 16, 1024), dX = make_sdesc(sa + SK_W_BYTES, 16, 1024);
#pragma unroll
      for (uint32_t k = 0; k < SK_BK / 16; ++k)
        if (leader) mma_ss(tb, dW + k * 2, dX + k * 2, idesc, (it | (int)k) != 0);
      if (leader) mma_commit(&bars[SB_EMPTY + s]);
      __syncwarp();
    }
    if (leader) mma_commit(&bars[SB_ACC]);
    __syncwarp();
  }
  // ============ epilogue (warps 2-5; TMEM lane = weight row n): partial tile -> shared memory, [m][n] so that lanes are
  //              consecutive words ==================================================================================
  float* part = reinterpret_cast<float*>(smem);   // (the stage ring is idle once the accumulator is complete)
  if (warp >= 2) {
    const int q = warp & 3, n = q * 32 + lane;
    if (nsteps > 0) {
      if (lane == 0) mbar_wait(&bars[SB_ACC], 0);
      __syncwarp();
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < MT; c += 32) {
      uint32_t v[32];
      if (nsteps > 0) {
        tmem_ld32(tmem_addr(tb, q * 32, c), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) part[(c + e) * 128 + n] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before();
  if (a.ksplit > 1) sk_cluster_sync(); else __syncthreads();
  pdl_wait();   // (C is written below: nothing of the previous kernel may still be reading the buffer it occupies)
  if (warp >= 2) {
    // reduce 1 / ksplit of the batch rows over the cluster's partial tiles -> C[m][n0 ..]
    const int t = tid - 64;
    const int rows = (MT + a.ksplit - 1) / a.ksplit, m_lo = (int)rank * rows, m_hi = min(min(m_lo + rows, MT), a.M);
    if (a.ksplit == 1) sk_reduce_rows<1>(part, t, (int)rank, m_lo, m_hi, n0, a);
    else if (a.ksplit == 2) sk_reduce_rows<2>(part, t, (int)rank, m_lo, m_hi, n0, a);
    else if (a.ksplit == 4) sk_reduce_rows<4>(part, t, (int)rank, m_lo, m_hi, n0, a);
    else sk_reduce_rows<8>(part, t, (int)rank, m_lo, m_hi, n0, a);
  }
  // nobody leaves while a peer may still read its partial tile
  if (a.ksplit > 1) sk_cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc(tb, MT < 32 ? 32 : MT);
}


// ---- fp32 operands: 3xTF32 on the tensor cores ---------------------------------------------------------------------------
// inference_t2i.py runs the model in fp32 (no autocast), so the decode step of the reference multiplies fp32 weights: 70 MB +
// 34 MB per layer, which cuBLAS's SIMT SGEMM turns into a COMPUTE-bound 110 + 55 us at M = 64 (10 % of the HBM roofline).  Here
// the same weight stream feeds kind::tf32 MMAs with the usual error-compensated split: every operand tile is rewritten in
// shared memory as hi = tf32-rounded value (exactly representable, so the tensor core's own conversion changes nothing) and
// lo = a - hi (a second tile), and  A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo  accumulates in fp32 in TMEM; the dropped lo x lo
// term and the tf32 rounding of lo are ~2^-22 relative, i.e. the result is fp32-accurate (tests: <= 2e-6 against fp64).
// Roles: warp 0 TMA (W_hi / X_hi tiles = the raw fp32 data, 128-byte rows of 32 k), warps 2-5 split the tiles of a stage in
// place (hi) and into the stage's lo buffers, warp 1 issues 3 x 4 MMAs (K = 8 each) per stage; then the K-split reduction of
// the bf16 kernel.  The weight stream stays the bound: 12 MMAs of 128 x 64 x 8 per 16 KB of weights are ~6 us per GEMM.
// The tensor core adds into its fp32 accumulator with truncation, a bias that grows linearly with the number of additions
// (measured: 1e-6 at K = 132, 7e-6 at K = 1024 per accumulator).  The accumulation is therefore cut into groups of kStages
// K steps (K = 128): each group starts a fresh TMEM accumulator (two, alternating), and the splitter warps - which own the
// TMEM lanes - add the finished group into fp32 registers with ordinary round-to-nearest additions.
constexpr int SF_BK = 32;                                   // fp32 elements per K step (128 B rows)
constexpr uint32_t SF_W_BYTES = 128 * SF_BK * 4;            // 16 KB
enum { SFB_FULL = 0, SFB_SPLIT = 4, SFB_EMPTY = 8, SFB_ACCF = 12, SFB_ACCE = 14, SFB_COUNT = 16 };
constexpr uint32_t kTf32Mask = 0xFFFFE000u;
#ifndef OMNI_SF_SPLIT_WARPS
#define OMNI_SF_SPLIT_WARPS 8
#endif
constexpr int SF_SPLIT_WARPS = OMNI_SF_SPLIT_WARPS;    // warps that split tiles (the first four also own the TMEM lanes)
constexpr int SF_THREADS = 64 + 32 * SF_SPLIT_WARPS;
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// hi = fp32 rounded to the nearest tf32 (10 explicit mantissa bits), lo = a - hi (exact in fp32)
__device__ __forceinline__ void tf32_split(uint32_t a, uint32_t& hi, uint32_t& lo) {
  hi = (a + 0x1000u) & kTf32Mask;
  lo = __float_as_uint(__uint_as_float(a) - __uint_as_float(hi));
}

template <int MT, int kStages>
__global__ void __launch_bounds__(SF_THREADS, 1)
gemm_skinny_f32_kernel(const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapX1,
                       const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapX2, SkinnyArgs a) {
  constexpr uint32_t X_BYTES = MT * SF_BK * 4, STAGE = 2 * SF_W_BYTES + 2 * X_BYTES;
  constexpr uint32_t OFF_WLO = SF_W_BYTES, OFF_XHI = 2 * SF_W_BYTES, OFF_XLO = 2 * SF_W_BYTES + X_BYTES;
  static_assert(kStages * STAGE <= SK_RING && kStages <= 4, "stage ring");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SK_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SK_TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = a.ksplit > 1 ? sk_cluster_rank() : 0u;
  const int tile = blockIdx.x / a.ksplit, n0 = tile * 128;
  const int nk1 = (a.K1 + SF_BK - 1) / SF_BK;
  const int kb = min((int)rank * a.ksteps1, nk1), ke = min(kb + a.ksteps1, nk1);
  const int n1 = ke - kb, n2 = rank == 0 ? (a.K2 + SF_BK - 1) / SF_BK : 0, nsteps = n1 + n2;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars[SFB_FULL + i], 1);
      mbar_init(&bars[SFB_SPLIT + i], SF_SPLIT_WARPS);
      mbar_init(&bars[SFB_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[SFB_ACCF + i], 1); mbar_init(&bars[SFB_ACCE + i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 4 * MT);   // two accumulation buffers of 2 MT columns (hi x hi + lo x hi | hi x lo)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapX1);
    if (a.K2 > 0) { tma_prefetch_desc(&mapW2); tma_prefetch_desc(&mapX2); }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;
  pdl_trigger();
  constexpr int G = kStages;                       // K steps per accumulation group
  const int ngroups = (nsteps + G - 1) / G;
  float sum[MT];                                   // (splitter / epilogue warps: this thread's weight row n, all batch columns)
#pragma unroll
  for (int e = 0; e < MT; ++e) sum[e] = 0.f;

  if (warp == 0) {
    // ============ TMA producer (the weights are constants: first ring of W tiles before pdl_wait) =========================
    if (lane == 0) {
      auto load_w = [&](int it, int s) {
        uint8_t* st = smem + s * STAGE;
        mbar_expect_tx(&bars[SFB_FULL + s], SF_W_BYTES + X_BYTES);
        if (it < n1) tma_load_2d(st, &mapW1, &bars[SFB_FULL + s], (kb + it) * SF_BK, n0);
        else tma_load_2d(st, &mapW2, &bars[SFB_FULL + s], (it - n1) * SF_BK, n0);
      };
      auto load_x = [&](int it, int s) {
        uint8_t* st = smem + s * STAGE + OFF_XHI;
        if (it < n1) tma_load_2d(st, &mapX1, &bars[SFB_FULL + s], (kb + it) * SF_BK, 0);
        else tma_load_2d(st, &mapX2, &bars[SFB_FULL + s], (it - n1) * SF_BK, 0);
      };
      const int first = min(nsteps, kStages);
      for (int it = 0; it < first; ++it) load_w(it, it);
      pdl_wait();
      for (int it = 0; it < first; ++it) load_x(it, it);
      for (int it = first; it < nsteps; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&bars[SFB_EMPTY + s], ph ^ 1);
        load_w(it, s);
        load_x(it, s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============ MMA issuer: D[n][m] += W_hi X_hi + W_lo X_hi + W_hi X_lo ====================================================
    const bool leader = sk_elect();
    const uint32_t idesc = make_idesc(128, MT, 2, 2, kMajorK, kMajorK);   // (a / b format 2 = TF32)
    const uint32_t idesc2 = make_idesc(128, 2 * MT, 2, 2, kMajorK, kMajorK);
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      const int grp = it / G, buf = grp & 1;
      const bool first = it % G == 0;
      if (first && grp >= 2) {                       // the splitters have drained the group that used this accumulator
        mbar_wait(&bars[SFB_ACCE + buf], (uint32_t)(((grp >> 1) - 1) & 1));
        tc_fence_after();
      }
      mbar_wait(&bars[SFB_SPLIT + s], ph);
      tc_fence_after();
      const uint32_t acc = tb + (uint32_t)buf * 2 * MT;
      const uint32_t sa = smem_u32(smem + s * STAGE);
      const uint64_t dWh = make_sdesc(sa, 16, 1024), dWl = make_sdesc(sa + OFF_WLO, 16, 1024);
      const uint64_t dXh = make_sdesc(sa + OFF_XHI, 16, 1024), dXl = make_sdesc(sa + OFF_XLO, 16, 1024);
      // Per K = 8 step (32 bytes): W_hi [X_hi ; X_lo]^T as ONE MMA of N = 2 MT (the X_hi and X_lo tiles are adjacent rows of one
      // swizzled tile) into columns [0, 2 MT), then W_lo X_hi^T into columns [0, MT): two instructions instead of three - these
      // small MMAs cost ~their issue overhead, not their flops (measured: 12 per stage paced the kernel, not the weight stream).
#pragma unroll
      for (uint32_t k = 0; k < SF_BK / 8; ++k) {
        if (leader) mma_ss_tf32(acc, dWh + k * 2, dXh + k * 2, idesc2, !(first && k == 0));
        if (leader) mma_ss_tf32(acc, dWl + k * 2, dXh + k * 2, idesc, true);
      }
      (void)dXl;
      if (leader) mma_commit(&bars[SFB_EMPTY + s]);
      if (leader && (it % G == G - 1 || it == nsteps - 1)) mma_commit(&bars[SFB_ACCF + buf]);
      __syncwarp();
    }
  } else {
    // ============ splitters (warps 2-5): hi in place, lo into the stage's second buffers ==================================
    const int t = tid - 64;                 // 0 .. 32 SF_SPLIT_WARPS - 1
    const int q = warp & 3;
    const bool owner = warp < 6;            // warps 2-5 own the four TMEM lane quadrants: they drain and write the partial tile
    int drained = 0;
    auto drain = [&](int g) {   // finished group g: its accumulator -> the fp32 register sums (round-to-nearest adds)
      const int buf = g & 1;
      mbar_wait(&bars[SFB_ACCF + buf], (uint32_t)((g >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < MT; c += 32) {
        uint32_t v[32], v2[32];
        tmem_ld32(tmem_addr(tb, q * 32, buf * 2 * MT + c), v);
        tmem_ld32(tmem_addr(tb, q * 32, buf * 2 * MT + MT + c), v2);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) sum[c + e] += __uint_as_float(v[e]) + __uint_as_float(v2[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[SFB_ACCE + buf]);
    };
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      mbar_wait(&bars[SFB_FULL + s], ph);
      uint8_t* st = smem + s * STAGE;
#pragma unroll
      for (int k = 0; k < (int)(SF_W_BYTES / 16 / (32 * SF_SPLIT_WARPS)); ++k) {   // 1024 16-byte units of W
        uint4* ph4 = reinterpret_cast<uint4*>(st) + t + 32 * SF_SPLIT_WARPS * k;
        uint4 v = *ph4, lo;
        tf32_split(v.x, v.x, lo.x); tf32_split(v.y, v.y, lo.y); tf32_split(v.z, v.z, lo.z); tf32_split(v.w, v.w, lo.w);
        *ph4 = v;
        *(reinterpret_cast<uint4*>(st + OFF_WLO) + t + 32 * SF_SPLIT_WARPS * k) = lo;
      }
#pragma unroll
      for (int k = 0; k < (int)(X_BYTES / 16 / (32 * SF_SPLIT_WARPS)); ++k) {
        uint4* ph4 = reinterpret_cast<uint4*>(st + OFF_XHI) + t + 32 * SF_SPLIT_WARPS * k;
        uint4 v = *ph4, lo;
        tf32_split(v.x, v.x, lo.x); tf32_split(v.y, v.y, lo.y); tf32_split(v.z, v.z, lo.z); tf32_split(v.w, v.w, lo.w);
        *ph4 = v;
        *(reinterpret_cast<uint4*>(st + OFF_XLO) + t + 32 * SF_SPLIT_WARPS * k) = lo;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[SFB_SPLIT + s]);
      // (the groups before the one this stage belongs to have all their MMAs issued: drain them while the stream goes on)
      if (owner) while (drained < it / G) drain(drained++);
    }
    if (owner) while (drained < ngroups) drain(drained++);
  }
  // ============ epilogue (warps 2-5): partial tile -> shared memory [m][n], cluster reduction -> C (as the bf16 kernel) =====
  float* part = reinterpret_cast<float*>(smem);
  if (warp >= 2 && warp < 6) {
    // (every MMA of this CTA is complete - the last group has been drained - so the stage ring is free for the partial tile)
    const int n = (warp & 3) * 32 + lane;
#pragma unroll
    for (int e = 0; e < MT; ++e) part[e * 128 + n] = sum[e];
  }
  tc_fence_before();
  if (a.ksplit > 1) sk_cluster_sync(); else __syncthreads();
  pdl_wait();
  if (warp >= 2 && warp < 6) {
    const int t = tid - 64;
    const int rows = (MT + a.ksplit - 1) / a.ksplit, m_lo = (int)rank * rows, m_hi = min(min(m_lo + rows, MT), a.M);
    if (a.ksplit == 1) sk_reduce_rows<1>(part, t, (int)rank, m_lo, m_hi, n0, a);
    else if (a.ksplit == 2) sk_reduce_rows<2>(part, t, (int)rank, m_lo, m_hi, n0, a);
    else sk_reduce_rows<4>(part, t, (int)rank, m_lo, m_hi, n0, a);
  }
  if (a.ksplit > 1) sk_cluster_sync(); else __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 4 * MT);
}
}  // namespace
int g_skinny_ksplit = 0;
namespace {
int skinny_map(CUtensorMap* m, const omni_tensor_t& t, int rows_box) {  // K-major (rows, K): dims (K, rows), box {64, rows_box}
  const uint64_t dims[2] = {(uint64_t)t.shape[1], (uint64_t)t.shape[0]};
  const uint64_t str[1] = {(uint64_t)(t.shape[0] > 1 ? t.stride[0] : t.shape[1]) * 2};
  const uint32_t box[2] = {64, (uint32_t)rows_box};
  return make_tmap(m, t.data, 2, dims, str, box, OMNI_BF16);
}

}  // namespace

// Does the decode-shaped path take this GEMM?  (K-major operands, M <= 128, enough weight rows to be worth a stream.)
bool gemm_skinny_eligible(int64_t M, int64_t N, int64_t K1, int64_t K2, int amaj, int bmaj) {
  // (N % 8: the reduction writes groups of four consecutive n with one vector store)
  return M >= 1 && M <= 128 && N >= 256 && N % 8 == 0 && K1 >= 256 && amaj == 0 && bmaj == 0 && K2 >= 0;
}

int gemm_skinny(const omni_gemm_params_t* p, cudaStream_t s) {
  const omni_tensor_t &X = p->a, &W = p->b, &C = p->out;
  const int64_t M = X.shape[0], K1 = X.shape[1], N = W.shape[0];
  const int64_t K2 = present(p->a2) ? p->a2.shape[1] : 0;
  const int MT = M <= 64 ? 64 : 128;
  const int tiles_n = (int)((N + 127) / 128), nk1 = (int)((K1 + SK_BK - 1) / SK_BK);
  // K split: clusters of 1 / 2 / 4 CTAs; take the largest split that keeps one wave (<= SM count) and >= 4 K steps per CTA.
  // (Measured at M = 64, weights from HBM, CUDA graph: in_proj 11.7 / 11.3 / 21.0 us with 1 / 2 / 4 CTAs - four would be two
  // waves; out_proj 16.7 / 11.6 / 9.4 / 14.7 us with 1 / 2 / 4 / 8 - clusters of eight schedule and synchronise too slowly.)
  int ksplit = 1;
  for (int c = 2; c <= 4; c *= 2)
    if ((int64_t)tiles_n * c <= sm_count() && nk1 / c >= 4) ksplit = c;
  // debug override (omni_debug_set_gemm_mode(10 + ksplit)), honoured while the grid stays one wave
  if (g_skinny_ksplit > 0 && (int64_t)tiles_n * g_skinny_ksplit <= sm_count()) ksplit = g_skinny_ksplit;
  SkinnyArgs a{};
  a.M = (int)M; a.N = (int)N; a.K1 = (int)K1; a.K2 = (int)K2;
  a.ksplit = ksplit; a.ksteps1 = (nk1 + ksplit - 1) / ksplit;
  a.out_f32 = C.dtype == OMNI_F32; a.C = C.data; a.ldc = M > 1 ? C.stride[0] : N;
  CUtensorMap mW1, mX1, mW2, mX2;
  if (int rc = skinny_map(&mW1, W, 128)) return rc;
  if (int rc = skinny_map(&mX1, X, MT)) return rc;
  if (K2 > 0) {
    if (int rc = skinny_map(&mW2, p->b2, 128)) return rc;
    if (int rc = skinny_map(&mX2, p->a2, MT)) return rc;
  } else {
    mW2 = mW1; mX2 = mX1;
  }
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(gemm_skinny_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM);
    cudaFuncSetAttribute(gemm_skinny_kernel<128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM);
  });
  const dim3 grid((unsigned)(tiles_n * ksplit)), block(SK_THREADS);
  cudaError_t e = MT == 64 ? launch_pdl(kPdlGemm, gemm_skinny_kernel<64, 8>, grid, block, SK_SMEM, s, (unsigned)ksplit, mW1, mX1, mW2, mX2, a)
                           : launch_pdl(kPdlGemm, gemm_skinny_kernel<128, 6>, grid, block, SK_SMEM, s, (unsigned)ksplit, mW1, mX1, mW2, mX2, a);
  if (e != cudaSuccess) return set_error(OMNI_CUDA_ERROR, "gemm_skinny_kernel launch: %s", cudaGetErrorString(e));
  count_launch();
  return OMNI_OK;
}

// fp32 operands (3xTF32): same shapes as the bf16 path, fp32 C
bool gemm_skinny_f32_eligible(int64_t M, int64_t N, int64_t K1, int64_t K2) {
  return M >= 1 && M <= 128 && N >= 256 && N % 4 == 0 && K1 >= 128 && K1 % 4 == 0 && K2 >= 0 && K2 % 4 == 0;
}

int gemm_skinny_f32(const omni_gemm_params_t* p, cudaStream_t s) {
  const omni_tensor_t &X = p->a, &W = p->b, &C = p->out;
  const int64_t M = X.shape[0], K1 = X.shape[1], N = W.shape[0];
  const int64_t K2 = present(p->a2) ? p->a2.shape[1] : 0;
  const int MT = M <= 64 ? 64 : 128;
  const int tiles_n = (int)((N + 127) / 128), nk1 = (int)((K1 + SF_BK - 1) / SF_BK);
  int ksplit = 1;
  for (int c = 2; c <= 4; c *= 2)
    if ((int64_t)tiles_n * c <= sm_count() && nk1 / c >= 8) ksplit = c;
  if (g_skinny_ksplit > 0 && g_skinny_ksplit <= 4 && (int64_t)tiles_n * g_skinny_ksplit <= sm_count()) ksplit = g_skinny_ksplit;
  SkinnyArgs a{};
  a.M = (int)M; a.N = (int)N; a.K1 = (int)K1; a.K2 = (int)K2;
  a.ksplit = ksplit; a.ksteps1 = (nk1 + ksplit - 1) / ksplit;
  a.out_f32 = 1; a.C = C.data; a.ldc = M > 1 ? C.stride[0] : N;
  auto map32 = [](CUtensorMap* m, const omni_tensor_t& t, int rows_box) {  // K-major fp32 (rows, K): box {32, rows_box}
    const uint64_t dims[2] = {(uint64_t)t.shape[1], (uint64_t)t.shape[0]};
    const uint64_t str[1] = {(uint64_t)(t.shape[0] > 1 ? t.stride[0] : t.shape[1]) * 4};
    const uint32_t box[2] = {(uint32_t)SF_BK, (uint32_t)rows_box};
    return make_tmap(m, t.data, 2, dims, str, box, OMNI_F32);
  };
  CUtensorMap mW1, mX1, mW2, mX2;
  if (int rc = map32(&mW1, W, 128)) return rc;
  if (int rc = map32(&mX1, X, MT)) return rc;
  if (K2 > 0) {
    if (int rc = map32(&mW2, p->b2, 128)) return rc;
    if (int rc = map32(&mX2, p->a2, MT)) return rc;
  } else {
    mW2 = mW1; mX2 = mX1;
  }
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(gemm_skinny_f32_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM);
    cudaFuncSetAttribute(gemm_skinny_f32_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM);
  });
  const dim3 grid((unsigned)(tiles_n * ksplit)), block(SF_THREADS);
  cudaError_t e = MT == 64 ? launch_pdl(kPdlGemm, gemm_skinny_f32_kernel<64, 4>, grid, block, SK_SMEM, s, (unsigned)ksplit, mW1, mX1, mW2, mX2, a)
                           : launch_pdl(kPdlGemm, gemm_skinny_f32_kernel<128, 3>, grid, block, SK_SMEM, s, (unsigned)ksplit, mW1, mX1, mW2, mX2, a);
  if (e != cudaSuccess) return set_error(OMNI_CUDA_ERROR, "gemm_skinny_f32_kernel launch: %s", cudaGetErrorString(e));
  count_launch();
  return OMNI_OK;
}

}  // namespace omni

using namespace omni;

extern "C" int omni_gemm_f32_decode(const omni_gemm_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t &A = p->a, &B = p->b, &C = p->out;
  auto kmajor32 = [](const omni_tensor_t& t) {
    return present(t) && t.ndim == 2 && t.dtype == OMNI_F32 && aligned16(t.data) && (t.shape[1] <= 1 || t.stride[1] == 1) &&
           (t.shape[0] <= 1 || (t.stride[0] % 4 == 0 && t.stride[0] >= t.shape[1]));
  };
  OMNI_CHECK(kmajor32(A) && kmajor32(B), OMNI_BAD_STRIDE, "gemm_f32: a (M, K) and b (N, K) must be fp32, K contiguous, 16-byte aligned rows");
  const int64_t M = A.shape[0], K1 = A.shape[1], N = B.shape[0];
  OMNI_CHECK(B.shape[1] == K1, OMNI_BAD_SHAPE, "gemm_f32: a is (M, K), b must be (N, K)");
  OMNI_CHECK(present(C) && shape_is(C, 2, M, N) && C.dtype == OMNI_F32 && (N <= 1 || C.stride[1] == 1) && aligned16(C.data) &&
                 (M <= 1 || C.stride[0] % 4 == 0),
             OMNI_BAD_SHAPE, "gemm_f32: out must be (M, N) fp32 with contiguous, 16-byte aligned rows");
  int64_t K2 = 0;
  if (present(p->a2) || present(p->b2)) {
    OMNI_CHECK(kmajor32(p->a2) && kmajor32(p->b2), OMNI_BAD_STRIDE, "gemm_f32: the second operand pair must have the layout of the first");
    K2 = p->a2.shape[1];
    OMNI_CHECK(p->a2.shape[0] == M && p->b2.shape[0] == N && p->b2.shape[1] == K2, OMNI_BAD_SHAPE, "gemm_f32: a2 (M, K2), b2 (N, K2)");
  }
  OMNI_CHECK(get_encode_tiled() != nullptr && gemm_skinny_f32_eligible(M, N, K1, K2), OMNI_UNSUPPORTED,
             "gemm_f32: only decode shapes (M <= 128, N >= 256, K >= 128, N and K multiples of 4) run on the 3xTF32 kernel");
  return gemm_skinny_f32(p, static_cast<cudaStream_t>(stream));
}
