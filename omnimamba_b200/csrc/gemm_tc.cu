// bf16 GEMM on tcgen05 tensor cores for the two projections of the Mamba-2 block (sm_100a).
//
//   C[M, N] = A1[M, K1] B1[N, K1]^T  (+ A2[M, K2] B2[N, K2]^T),  bf16 operands, fp32 accumulation in TMEM, bf16 or fp32 C.
//
// Replaces the cuBLAS calls behind `in_proj` / `out_proj` (F.linear inside mamba_ssm's Mamba2.forward and
// mamba_split_conv1d_scan_combined) and their backward GEMMs.  The optional second operand
// pair is the LoRA branch of the reference's in_proj (/root/reference/models/stage2/lora.py:263-279): with A2 = s (x A^T)
// (M x r) and B2 = the adapter's up-projection (N x r) the rank-r update rides in the SAME accumulator as x W^T - one pass
// over the 8512-wide output instead of a GEMM, a skinny GEMM and an elementwise add over (M, 8512).
//
// Either operand may be K-major (rows of K contiguous: x, W in the forward) or MN-major (given as its transpose, MN
// contiguous: W in dgrad, dy^T and x^T in wgrad); the shared-memory descriptors take both, so forward, dgrad and wgrad are
// one kernel.  Ragged M / N / K are handled by TMA (out-of-bounds elements load as zero, stores are clipped).
//
// One persistent CTA per SM, 192 threads:
//   warp 0      TMA producer: 4-stage ring of (A 128 x 64, B 256 x 64) bf16 tiles, 128B swizzle, 48 KB per stage
//   warp 1      MMA issuer: tcgen05.mma 128 x 256 x 16 (4 per stage), accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-5   epilogue: TMEM -> registers -> bf16 / fp32 -> swizzled 16 KB staging slab -> TMA store (two slabs in flight),
//               overlapping the main loop of the next tile
// Tiles are walked in groups of kGroupM row blocks x all column blocks so that the A panel of a group and all of B stay in
// the 126 MB L2 while 148 CTAs stream through them.
#include <mutex>

#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

// Two tile shapes: 128 x 256 (4 stages) for the training / prefill GEMMs, 128 x 64 (8 stages) when there are too few
// 256-wide column blocks to occupy the machine (decode: M = batch <= 128 rows, the weights stream once).
constexpr int BM = 128, BK = 64, kGroupM = 16;
constexpr int kGemmThreads = 192;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t SLAB_BYTES = 128 * 128;  // 128 rows x 128 B (64 bf16 or 32 fp32 columns)
constexpr uint32_t G_EPI = 4 * (A_BYTES + 256 * BK * 2);   // = 8 * (A_BYTES + 64 * BK * 2): both shapes use 192 KB of stages
constexpr uint32_t G_BAR = G_EPI + 2 * SLAB_BYTES;
constexpr int kMaxStages = 8;
int g_gemm_mode = 0;   // debug (omni_debug_set_gemm_mode): 0 auto, 1 single-CTA tiles only, 2 CTA pairs whenever the shape allows,
                       // 3 auto without the weight-streaming kernel of gemm_skinny.cu
enum { GB_FULL = 0, GB_EMPTY = kMaxStages, GB_ACC_FULL = 2 * kMaxStages, GB_ACC_EMPTY = 2 * kMaxStages + 2, GB_COUNT = 2 * kMaxStages + 4 };
constexpr uint32_t G_TMEMPTR = G_BAR + GB_COUNT * 8;
constexpr uint32_t G_SMEM = G_TMEMPTR + 16;
static_assert(G_SMEM <= 232448, "shared memory budget");

struct GemmArgs {
  int M, N, K1, K2;
  int a_mn, b_mn;   // 1: the operand is MN-major (its transpose is what lies row-major in memory)
  int out_f32;
  int tiles_m, tiles_n;
};

__device__ __forceinline__ bool elect_one_g() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int BN>
__device__ __forceinline__ void tile_coords(int t, const GemmArgs& a, int& m0, int& n0) {
  const int per_group = kGroupM * a.tiles_n;
  const int g = t / per_group, r = t - g * per_group;
  const int first_m = g * kGroupM;
  const int gm = min(kGroupM, a.tiles_m - first_m);
  m0 = (first_m + r % gm) * BM;
  n0 = (r / gm) * BN;
}

template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
               const __grid_constant__ CUtensorMap mapC, GemmArgs a) {
  constexpr uint32_t B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;   // (the epilogue reads 32 columns at a time)
  constexpr uint32_t kTmemCols = 2 * kAccStride;
  static_assert(kStages * STAGE_BYTES <= G_EPI && kStages <= kMaxStages, "stage ring");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + G_TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars[GB_FULL + i], 1);
      mbar_init(&bars[GB_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[GB_ACC_FULL + i], 1);
      mbar_init(&bars[GB_ACC_EMPTY + i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB1); tma_prefetch_desc(&mapC);
    if (a.K2 > 0) { tma_prefetch_desc(&mapA2); tma_prefetch_desc(&mapB2); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;

  const int ntiles = a.tiles_m * a.tiles_n;
  const int nk1 = (a.K1 + BK - 1) / BK, nk2 = (a.K2 + BK - 1) / BK, nk = nk1 + nk2;

  if (warp == 0) {
    // ============ TMA producer ==========================================================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0;
        tile_coords<BN>(t, a, m0, n0);
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&bars[GB_EMPTY + s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          uint64_t* full = &bars[GB_FULL + s];
          mbar_expect_tx(full, STAGE_BYTES);
          const bool second = kt >= nk1;
          const CUtensorMap* mA = second ? &mapA2 : &mapA1;
          const CUtensorMap* mB = second ? &mapB2 : &mapB1;
          const int k0 = (second ? kt - nk1 : kt) * BK;
          if (!a.a_mn) {
            tma_load_2d(sa, mA, full, k0, m0);                 // box {64 k, 128 rows}
          } else {
            tma_load_2d(sa, mA, full, m0, k0);                 // two boxes {64 m, 64 k rows}
            tma_load_2d(sa + 8192, mA, full, m0 + 64, k0);
          }
          if (!a.b_mn) {
            tma_load_2d(sb, mB, full, k0, n0);                 // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, mB, full, n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============ MMA issuer (warp-uniform control flow, one elected lane issues) =====================================
    const bool leader = elect_one_g();
    const uint32_t idesc = make_idesc(BM, BN, kFmtBF16, kFmtBF16, a.a_mn ? kMajorMN : kMajorK, a.b_mn ? kMajorMN : kMajorK);
    // K-major tiles advance by 32 B per 16-wide K step; MN-major tiles (64-element blocks 8 KB apart) by 16 rows = 2 KB
    const uint32_t a_lbo = a.a_mn ? 8192u : 16u, b_lbo = a.b_mn ? 8192u : 16u;
    const uint32_t a_step = a.a_mn ? 128u : 2u, b_step = a.b_mn ? 128u : 2u;
    uint32_t it = 0, tile_no = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tile_no) {
      const uint32_t buf = tile_no & 1, aph = (tile_no >> 1) & 1;
      mbar_wait(&bars[GB_ACC_EMPTY + buf], aph ^ 1);
      tc_fence_after();
      const uint32_t acc = tb + buf * kAccStride;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&bars[GB_FULL + s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t dA = make_sdesc(sa, a_lbo, 1024), dB = make_sdesc(sa + A_BYTES, b_lbo, 1024);
#pragma unroll
        for (uint32_t k = 0; k < BK / 16; ++k)
          if (leader) mma_ss(acc, dA + k * a_step, dB + k * b_step, idesc, (kt | (int)k) != 0);
        if (leader) mma_commit(&bars[GB_EMPTY + s]);
        __syncwarp();
      }
      if (leader) mma_commit(&bars[GB_ACC_FULL + buf]);
      __syncwarp();
    }
  } else {
    // ============ epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = rows of the tile ===============================
    const int q = warp & 3, row = q * 32 + lane;
    const bool issuer = warp == 2 && lane == 0;
    const uint32_t rsw = (uint32_t)(row & 7);
    uint32_t tile_no = 0, slab_no = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tile_no) {
      int m0, n0;
      tile_coords<BN>(t, a, m0, n0);
      const uint32_t buf = tile_no & 1, aph = (tile_no >> 1) & 1;
      if (lane == 0) mbar_wait(&bars[GB_ACC_FULL + buf], aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = tmem_addr(tb, q * 32, buf * kAccStride);
      const int ncols = min(BN, a.N - n0);
      const int slab_cols = a.out_f32 ? 32 : 64;
      const int nslabs = (ncols + slab_cols - 1) / slab_cols;
#pragma unroll 1
      for (int sl = 0; sl < nslabs; ++sl, ++slab_no) {
        uint8_t* slab = smem + G_EPI + (slab_no & 1) * SLAB_BYTES;
        uint32_t v0[32], v1[32];
        tmem_ld32(acc + sl * slab_cols, v0);
        if (BN > 32 && !a.out_f32) tmem_ld32(acc + sl * slab_cols + 32, v1);   // (BN <= 32: columns 32.. of the slab are clipped by the store)
        else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v1[e] = 0u;
        }
        if (issuer) tma_store_wait_read<1>();   // the store issued two slabs ago has read this buffer
        tmem_ld_wait();
        if (sl == nslabs - 1) {                 // accumulator buffer fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[GB_ACC_EMPTY + buf]);
        }
        named_bar(1, 128);
        uint8_t* rowp = slab + row * 128;
        if (a.out_f32) {
#pragma unroll
          for (uint32_t c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rowp + ((c ^ rsw) << 4)) = make_uint4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
        } else {
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v0[8 * c + 0]), __uint_as_float(v0[8 * c + 1]));
            o.y = pack_bf16(__uint_as_float(v0[8 * c + 2]), __uint_as_float(v0[8 * c + 3]));
            o.z = pack_bf16(__uint_as_float(v0[8 * c + 4]), __uint_as_float(v0[8 * c + 5]));
            o.w = pack_bf16(__uint_as_float(v0[8 * c + 6]), __uint_as_float(v0[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + ((c ^ rsw) << 4)) = o;
            o.x = pack_bf16(__uint_as_float(v1[8 * c + 0]), __uint_as_float(v1[8 * c + 1]));
            o.y = pack_bf16(__uint_as_float(v1[8 * c + 2]), __uint_as_float(v1[8 * c + 3]));
            o.z = pack_bf16(__uint_as_float(v1[8 * c + 4]), __uint_as_float(v1[8 * c + 5]));
            o.w = pack_bf16(__uint_as_float(v1[8 * c + 6]), __uint_as_float(v1[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + (((c + 4) ^ rsw) << 4)) = o;
          }
        }
        fence_proxy_async_smem();
        named_bar(2, 128);
        if (issuer) {
          tma_store_2d(&mapC, slab, n0 + sl * slab_cols, m0);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, kTmemCols);
}

// ---- 2-CTA variant (cta_group::2): a pair of CTAs on the two SMs of a TPC works on ONE 256 x 256 tile ------------------
// CTA r of the pair stages rows [128 r, +128) of A and rows [128 r, +128) of B; the leader's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256): each SM multiplies its own half of A with BOTH halves of B (the peer's half comes out
// of the peer's shared memory) into its own TMEM.  Per 256 x 256 x 64 step a CTA loads 32 KB instead of 48 KB for a
// 128 x 256 tile: a third less L2 -> SM traffic per flop (the 1-CTA kernel moves ~16 TB/s out of L2 at 1.4 PFLOP/s, which is
// where it saturates) and half the shared-memory reads of B per SM.  Six 32 KB stages.
//   full[s]   (leader's) : both CTAs' TMA loads of stage s (complete_tx on the leader's barrier) + the leader's expect_tx
//   empty[s]  (each CTA) : tcgen05.commit multicast to both CTAs - the MMAs have read stage s
//   acc_full  (each CTA) : commit multicast - the tile's accumulators are complete;   acc_empty (leader's): 4 + 4 epilogue warps
constexpr int kStages2 = 6;
constexpr uint32_t STAGE2_BYTES = 2 * A_BYTES;   // A 128 x 64 + B 128 x 64 per CTA
static_assert(kStages2 * STAGE2_BYTES <= G_EPI, "stage ring (2-CTA)");

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {  // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a tile into THIS CTA's shared memory whose bytes are counted on the barrier at cluster address `bar_cluster`
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {  // arrives on `bar` (same offset) in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
                const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
                const __grid_constant__ CUtensorMap mapC, GemmArgs a) {
  constexpr int BN = 256;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + G_TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (tid == 0) {
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&bars[GB_FULL + i], 1);
      mbar_init(&bars[GB_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[GB_ACC_FULL + i], 1);
      mbar_init(&bars[GB_ACC_EMPTY + i], 8);   // four epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // (the same warp of both CTAs)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapB1); tma_prefetch_desc(&mapC);
    if (a.K2 > 0) { tma_prefetch_desc(&mapA2); tma_prefetch_desc(&mapB2); }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // both CTAs' barriers are initialised before either touches the other's
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;

  const int tiles_m2 = (a.M + 255) / 256;
  const int ntiles = tiles_m2 * a.tiles_n;
  const int nk1 = (a.K1 + BK - 1) / BK, nk2 = (a.K2 + BK - 1) / BK, nk = nk1 + nk2;
  auto coords = [&](int t, int& m0, int& n0) {   // groups of kGroupM / 2 row blocks of 256 x all column blocks (L2 reuse)
    constexpr int GM = kGroupM / 2;
    const int per_group = GM * a.tiles_n;
    const int g = t / per_group, r = t - g * per_group;
    const int first_m = g * GM;
    const int gm = min(GM, tiles_m2 - first_m);
    m0 = (first_m + r % gm) * 256;
    n0 = (r / gm) * BN;
  };

  if (warp == 0) {
    // ============ TMA producer (both CTAs): own halves of A and B, bytes counted on the LEADER's full barrier ==============
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int m0, n0;
        coords(t, m0, n0);
        m0 += 128 * (int)rank;
        n0 += 128 * (int)rank;
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const uint32_t s = it % kStages2, ph = (it / kStages2) & 1;
          mbar_wait(&bars[GB_EMPTY + s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE2_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (rank == 0) mbar_expect_tx(&bars[GB_FULL + s], 2 * STAGE2_BYTES);
          const uint32_t full = mapa_rank(smem_u32(&bars[GB_FULL + s]), 0);
          const bool second = kt >= nk1;
          const CUtensorMap* mA = second ? &mapA2 : &mapA1;
          const CUtensorMap* mB = second ? &mapB2 : &mapB1;
          const int k0 = (second ? kt - nk1 : kt) * BK;
          if (!a.a_mn) {
            tma_load_2d_2sm(sa, mA, full, k0, m0);
          } else {
            tma_load_2d_2sm(sa, mA, full, m0, k0);
            tma_load_2d_2sm(sa + 8192, mA, full, m0 + 64, k0);
          }
          if (!a.b_mn) {
            tma_load_2d_2sm(sb, mB, full, k0, n0);
          } else {
            tma_load_2d_2sm(sb, mB, full, n0, k0);
            tma_load_2d_2sm(sb + 8192, mB, full, n0 + 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============ MMA issuer: the leader CTA's warp 1 only ================================================================
    if (rank == 0) {
      const bool leader = elect_one_g();
      const uint32_t idesc = make_idesc(256, BN, kFmtBF16, kFmtBF16, a.a_mn ? kMajorMN : kMajorK, a.b_mn ? kMajorMN : kMajorK);
      const uint32_t a_lbo = a.a_mn ? 8192u : 16u, b_lbo = a.b_mn ? 8192u : 16u;
      const uint32_t a_step = a.a_mn ? 128u : 2u, b_step = a.b_mn ? 128u : 2u;
      uint32_t it = 0, tile_no = 0;
      for (int t = pair; t < ntiles; t += npairs, ++tile_no) {
        const uint32_t buf = tile_no & 1, aph = (tile_no >> 1) & 1;
        mbar_wait(&bars[GB_ACC_EMPTY + buf], aph ^ 1);
        tc_fence_after();
        const uint32_t acc = tb + buf * BN;
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const uint32_t s = it % kStages2, ph = (it / kStages2) & 1;
          mbar_wait(&bars[GB_FULL + s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE2_BYTES);
          const uint64_t dA = make_sdesc(sa, a_lbo, 1024), dB = make_sdesc(sa + A_BYTES, b_lbo, 1024);
#pragma unroll
          for (uint32_t k = 0; k < BK / 16; ++k)
            if (leader) mma_ss_2sm(acc, dA + k * a_step, dB + k * b_step, idesc, (kt | (int)k) != 0);
          if (leader) mma_commit_2sm(&bars[GB_EMPTY + s]);
          __syncwarp();
        }
        if (leader) mma_commit_2sm(&bars[GB_ACC_FULL + buf]);
        __syncwarp();
      }
    }
  } else {
    // ============ epilogue (both CTAs): own 128 rows of the 256-row tile ==================================================
    const int q = warp & 3, row = q * 32 + lane;
    const bool issuer = warp == 2 && lane == 0;
    const uint32_t rsw = (uint32_t)(row & 7);
    uint32_t tile_no = 0, slab_no = 0;
    for (int t = pair; t < ntiles; t += npairs, ++tile_no) {
      int m0, n0;
      coords(t, m0, n0);
      m0 += 128 * (int)rank;
      const uint32_t buf = tile_no & 1, aph = (tile_no >> 1) & 1;
      if (lane == 0) mbar_wait(&bars[GB_ACC_FULL + buf], aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = tmem_addr(tb, q * 32, buf * BN);
      const int ncols = min(BN, a.N - n0);
      const int slab_cols = a.out_f32 ? 32 : 64;
      const int nslabs = (ncols + slab_cols - 1) / slab_cols;
#pragma unroll 1
      for (int sl = 0; sl < nslabs; ++sl, ++slab_no) {
        uint8_t* slab = smem + G_EPI + (slab_no & 1) * SLAB_BYTES;
        uint32_t v0[32], v1[32];
        tmem_ld32(acc + sl * slab_cols, v0);
        if (!a.out_f32) tmem_ld32(acc + sl * slab_cols + 32, v1);
        if (issuer) tma_store_wait_read<1>();
        tmem_ld_wait();
        if (sl == nslabs - 1) {   // accumulator buffer fully read: tell the leader's MMA warp (remote arrive from the peer CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&bars[GB_ACC_EMPTY + buf]), 0));
        }
        named_bar(1, 128);
        uint8_t* rowp = slab + row * 128;
        if (a.out_f32) {
#pragma unroll
          for (uint32_t c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rowp + ((c ^ rsw) << 4)) = make_uint4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
        } else {
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v0[8 * c + 0]), __uint_as_float(v0[8 * c + 1]));
            o.y = pack_bf16(__uint_as_float(v0[8 * c + 2]), __uint_as_float(v0[8 * c + 3]));
            o.z = pack_bf16(__uint_as_float(v0[8 * c + 4]), __uint_as_float(v0[8 * c + 5]));
            o.w = pack_bf16(__uint_as_float(v0[8 * c + 6]), __uint_as_float(v0[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + ((c ^ rsw) << 4)) = o;
            o.x = pack_bf16(__uint_as_float(v1[8 * c + 0]), __uint_as_float(v1[8 * c + 1]));
            o.y = pack_bf16(__uint_as_float(v1[8 * c + 2]), __uint_as_float(v1[8 * c + 3]));
            o.z = pack_bf16(__uint_as_float(v1[8 * c + 4]), __uint_as_float(v1[8 * c + 5]));
            o.w = pack_bf16(__uint_as_float(v1[8 * c + 6]), __uint_as_float(v1[8 * c + 7]));
            *reinterpret_cast<uint4*>(rowp + (((c + 4) ^ rsw) << 4)) = o;
          }
        }
        fence_proxy_async_smem();
        named_bar(2, 128);
        if (issuer && m0 < a.M) {
          tma_store_2d(&mapC, slab, n0 + sl * slab_cols, m0);
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair may still use its shared memory / barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

// (M, K) operand: which dim is contiguous?  0 = K-major, 1 = MN-major, -1 = neither / misaligned
int operand_major(const omni_tensor_t& t) {
  if (!present(t) || t.ndim != 2 || t.dtype != OMNI_BF16 || !aligned16(t.data)) return -1;
  if (t.stride[1] == 1 && (t.shape[0] == 1 || (t.stride[0] % 8 == 0 && t.stride[0] >= t.shape[1]))) return 0;
  if (t.stride[0] == 1 && (t.shape[1] == 1 || (t.stride[1] % 8 == 0 && t.stride[1] >= t.shape[0]))) return 1;
  return -1;
}

int operand_map(CUtensorMap* m, const omni_tensor_t& t, int major, int rows_box) {
  // K-major: dims (K, MN), box {64, rows_box};  MN-major: dims (MN, K), box {64, 64}
  if (major == 0) {
    const uint64_t dims[2] = {(uint64_t)t.shape[1], (uint64_t)t.shape[0]};
    const uint64_t str[1] = {(uint64_t)(t.shape[0] > 1 ? t.stride[0] : t.shape[1]) * 2};
    const uint32_t box[2] = {64, (uint32_t)rows_box};
    return make_tmap(m, t.data, 2, dims, str, box, OMNI_BF16);
  }
  const uint64_t dims[2] = {(uint64_t)t.shape[0], (uint64_t)t.shape[1]};
  const uint64_t str[1] = {(uint64_t)(t.shape[1] > 1 ? t.stride[1] : t.shape[0]) * 2};
  const uint32_t box[2] = {64, 64};
  return make_tmap(m, t.data, 2, dims, str, box, OMNI_BF16);
}

}  // namespace
}  // namespace omni

namespace omni {
// gemm_skinny.cu: weight-streaming kernel for decode-shaped GEMMs (M <= 128 rows, K split over a cluster)
bool gemm_skinny_eligible(int64_t M, int64_t N, int64_t K1, int64_t K2, int amaj, int bmaj);
int gemm_skinny(const omni_gemm_params_t* p, cudaStream_t s);
}  // namespace omni

using namespace omni;

extern "C" int omni_gemm_bf16_supported(void) { return get_encode_tiled() != nullptr ? 1 : 0; }

extern "C" int omni_gemm_bf16(const omni_gemm_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t &A = p->a, &B = p->b, &C = p->out;
  const int amaj = operand_major(A), bmaj = operand_major(B);
  OMNI_CHECK(amaj >= 0 && bmaj >= 0, OMNI_BAD_STRIDE,
             "gemm: a (M, K) and b (N, K) must be bf16, 16-byte aligned, with one contiguous dim and the other stride a multiple of 8");
  const int64_t M = A.shape[0], K1 = A.shape[1], N = B.shape[0];
  OMNI_CHECK(B.shape[1] == K1, OMNI_BAD_SHAPE, "gemm: a is (M, K), b must be (N, K)");
  OMNI_CHECK(present(C) && shape_is(C, 2, M, N) && (C.dtype == OMNI_BF16 || C.dtype == OMNI_F32) && (N <= 1 || C.stride[1] == 1) &&
                 aligned16(C.data) && (M <= 1 || (C.stride[0] * dtype_size(C.dtype)) % 16 == 0),
             OMNI_BAD_SHAPE, "gemm: out must be (M, N) bf16 / fp32 with contiguous, 16-byte aligned rows");
  int64_t K2 = 0;
  if (present(p->a2) || present(p->b2)) {
    OMNI_CHECK(operand_major(p->a2) == amaj && operand_major(p->b2) == bmaj, OMNI_BAD_STRIDE,
               "gemm: the second operand pair must have the layout of the first");
    K2 = p->a2.shape[1];
    OMNI_CHECK(p->a2.shape[0] == M && p->b2.shape[0] == N && p->b2.shape[1] == K2, OMNI_BAD_SHAPE, "gemm: a2 (M, K2), b2 (N, K2)");
  }
  if (M == 0 || N == 0) return OMNI_OK;
  OMNI_CHECK(K1 > 0, OMNI_BAD_SHAPE, "gemm: K must be positive");
  OMNI_CHECK(M < (1ll << 31) && N < (1ll << 31) && K1 < (1ll << 31), OMNI_BAD_SHAPE, "gemm: dims must fit in 31 bits");
  // decode-shaped (M = batch <= 128 rows of K-major operands): the weights are streamed once, K split over a cluster
  // (g_gemm_mode 3, debug: keep such shapes on the tile kernels below)
  if (g_gemm_mode != 3 && g_gemm_mode != 1 && gemm_skinny_eligible(M, N, K1, K2, amaj, bmaj))
    return gemm_skinny(p, static_cast<cudaStream_t>(stream));
  // too few 128 x 256 tiles to occupy the SMs (decode: M = batch): 128 x 64 tiles, four times as many CTAs stream the weights
  const int64_t tiles_m = (M + BM - 1) / BM;
  const bool narrow = tiles_m * ((N + 255) / 256) * 2 <= sm_count() && N > 64;
  // N <= 16 (the rank-r LoRA GEMMs): 128 x 16 tiles - a 256-wide tile would spend 16 x the tensor time and stage 32 KB of
  // zeros per K step; needs a K-major b (its MN-major form would be 16-byte rows)
  const bool tiny_n = N <= 16 && bmaj == 0 && (K2 == 0);
  const int BN = tiny_n ? 16 : (narrow ? 64 : 256);
  CUtensorMap mA1, mB1, mA2, mB2, mC;
  if (int rc = operand_map(&mA1, A, amaj, BM)) return rc;
  if (int rc = operand_map(&mB1, B, bmaj, BN)) return rc;
  if (K2 > 0) {
    if (int rc = operand_map(&mA2, p->a2, amaj, BM)) return rc;
    if (int rc = operand_map(&mB2, p->b2, bmaj, BN)) return rc;
  } else {
    mA2 = mA1; mB2 = mB1;
  }
  {
    const bool f32 = C.dtype == OMNI_F32;
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)(M > 1 ? C.stride[0] : N) * (f32 ? 4u : 2u)};
    const uint32_t box[2] = {f32 ? 32u : 64u, 128};
    if (int rc = make_tmap(&mC, C.data, 2, dims, str, box, C.dtype)) return rc;
  }
  GemmArgs a{};
  a.M = (int)M; a.N = (int)N; a.K1 = (int)K1; a.K2 = (int)K2;
  a.a_mn = amaj; a.b_mn = bmaj; a.out_f32 = C.dtype == OMNI_F32;
  a.tiles_m = (int)tiles_m; a.tiles_n = (int)((N + BN - 1) / BN);
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(gemm_tc_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    cudaFuncSetAttribute(gemm_tc_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    cudaFuncSetAttribute(gemm_tc_kernel<16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
  });
  // CTA pairs (256 x 256 tiles) when there are enough of them to fill the machine; g_gemm_mode (debug): 1 forces the
  // 1-CTA kernel, 2 the 2-CTA kernel
  const int64_t tiles2 = ((M + 255) / 256) * ((N + 255) / 256);
  const bool pairs = !narrow && !tiny_n && (g_gemm_mode == 2 || (g_gemm_mode == 0 && tiles2 >= sm_count()));
  if (pairs) {
    // both CTAs of a pair load 128-row boxes of B
    if (int rc = operand_map(&mB1, B, bmaj, 128)) return rc;
    if (K2 > 0) { if (int rc = operand_map(&mB2, p->b2, bmaj, 128)) return rc; } else { mB2 = mB1; }
    a.tiles_n = (int)((N + 255) / 256);
    const int grid2 = (int)std::min<int64_t>(2 * tiles2, (int64_t)(sm_count() & ~1));
    gemm_tc2_kernel<<<grid2, kGemmThreads, G_SMEM, static_cast<cudaStream_t>(stream)>>>(mA1, mB1, mA2, mB2, mC, a);
    OMNI_CUDA_LAUNCH_CHECK("gemm_tc2_kernel");
    return OMNI_OK;
  }
  const int64_t ntiles = (int64_t)a.tiles_m * a.tiles_n;
  const int grid = (int)std::min<int64_t>(ntiles, sm_count());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tiny_n) gemm_tc_kernel<16, 8><<<grid, kGemmThreads, G_SMEM, s>>>(mA1, mB1, mA2, mB2, mC, a);
  else if (narrow) gemm_tc_kernel<64, 8><<<grid, kGemmThreads, G_SMEM, s>>>(mA1, mB1, mA2, mB2, mC, a);
  else gemm_tc_kernel<256, 4><<<grid, kGemmThreads, G_SMEM, s>>>(mA1, mB1, mA2, mB2, mC, a);
  OMNI_CUDA_LAUNCH_CHECK("gemm_tc_kernel");
  return OMNI_OK;
}

namespace omni { extern int g_skinny_ksplit; }
extern "C" void omni_debug_set_gemm_mode(int mode) {
  if (mode >= 10) { omni::g_skinny_ksplit = mode - 10; return; }   // 10: automatic K split again; 11 / 12 / 14 / 18: forced
  omni::g_gemm_mode = mode;
}
