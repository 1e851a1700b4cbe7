// Chunked SSD backward on tcgen05 tensor cores (sm_100a): the per-chunk gradient kernel.
//
// The backward of mamba_chunk_scan_combined (upstream's _chunk_scan_bwd_* / _chunk_state_bwd_* /
// _bmm_chunk_bwd Triton kernels, K9) is split into two state sweeps and one chunk-parallel kernel:
//   1. ssd_tc_fwd_kernel mode 1: forward sweep, stores the state S_c entering every chunk      (fp16, workspace)
//   2. ssd_tc_fwd_kernel mode 2: reverse sweep, stores the state gradient dS_{c+1} leaving it  (fp16, workspace)
//   3. THIS kernel, one work item per (batch, chunk, pair of heads), no dependence between items.
// With Lambda the in-chunk cumsum of dt*A, L_ij = exp(Lambda_i - Lambda_j) (i >= j), es_j = exp(Lambda_last - Lambda_j):
//   w_j   = sum_{i>=j} (B_j.C_i) L_ij dy_i + es_j (dS_{c+1} B_j)          dx_j = dt_j w_j + D dy_j,  ddt_j = x_j.w_j + A da_j
//   dC_i  = sum_h [ sum_{j<=i} (dy_i.x_j) L_ij dt_j B_j + exp(Lambda_i) dy_i S_c ]
//   dB_j  = sum_h [ sum_{i>=j} (dy_i.x_j) L_ij dt_j C_i + es_j dt_j x_j dS_{c+1} ]
//   da_k  = sum_{i>=k} (r_i - x_i.dxdiag_i) + sum_{j<k} x_j.dxstate_j + <dS_{c+1}, exp(Lambda_last) S_c>
// (dxdiag / dxstate = the two parts of dt_j w_j;  r_i = dy_i.(y_i - D x_i) is formed in fp32 WITHOUT the forward's bf16
// output - it cancels against x.dxdiag and the rounding of y would dominate dA: since y - D x is linear in C_i,
// r_i = <C_i, dC_i^(h)>, the head's own contribution to dC_i: the part through M_h B is read from the dC accumulator
// between the two heads, the part through the incoming state is dy_i . (exp(Lambda_i) C_i S_c), one more GEMM (G10).)
// Ten GEMM groups per item, all fp16 operands / fp32 accumulation in TMEM (R0..R3 = four 128-column regions):
//   G1 CB^T = B C^T -> R0      G3 ws = B dS16^T -> R1            PT_h = CB^T o L (SIMT, in place)   G2 wd_h = PT_h dy_h -> R2
//   G4 G_h = dy_h x_h^T -> R0  G4' G_h^T -> R1   M_h = G_h o L o dt, MT_h (SIMT, in place)
//   G5 dC += M_h B -> R2       G7 dB += MT_h C -> R3   G6 dC += (exp(Lambda) dy) S_c -> R2   G8 dB += (es dt x) dS -> R3
//   G10 Yoff = C S16^T -> R0 (for r_i)
// The kernel is deliberately phase-synchronous (every warp works on every SIMT phase, __syncthreads between phases, one
// thread issues the MMAs): 16 warps per phase hide the dependent-issue latency that a warp-specialised pipeline exposes
// (profiles/, forward kernel), at the price of not overlapping the tensor pipe with the SIMT phases inside a CTA.
#include <algorithm>
#include <mutex>

#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

constexpr int Q = 128, HD = 64, NS = 128;
constexpr int kThreads = 512;
constexpr int kTmaThread = 480;  // lane 0 of warp 15 issues every TMA load / store (thread 0 issues the MMAs)

constexpr uint32_t SM_X = 0;         // [head 2][Q rows x 128 B]  x bf16 -> fp16 -> X' = es dt x        32 KB
constexpr uint32_t SM_DY = 32768;    // [head 2][Q rows x 128 B]  dy bf16 -> fp16 -> exp(Lambda_i) dy   32 KB
constexpr uint32_t SM_B = 65536;     // [n-half 2][Q rows x 128 B] fp16                                 32 KB
constexpr uint32_t SM_C = 98304;     //                                                                 32 KB
constexpr uint32_t SM_S = 131072;    // [n-half 2][128 (h,p) rows x 128 B] S_c fp16                     32 KB
constexpr uint32_t SM_DS = 163840;   // dS_{c+1} fp16                                                   32 KB
constexpr uint32_t SM_STG = 196608;  // dx staging, one head: [Q rows x 128 B]                          16 KB
constexpr uint32_t SM_TAB = 212992;
struct BTab {
  float lam[2][Q];    // log2(e) * inclusive cumsum of dt*A
  float dtv[2][Q];    // transformed dt
  float ci[2][Q];     // exp(lam_i - ref(block of i))          (<= 1; also the row factor u_i of the M build)
  float v[2][3][96];  // v[h][w-1][j] = exp(ref(w) - lam_j) dt_j, j < 32 w
  float vd[2][Q];     // exp(ref(block of j) - lam_j) dt_j
  float eL[2][Q];     // exp(lam_i)
  float es[2][Q];     // exp(lam_last - lam_j)
  float vpre[2][Q];   // raw dt + bias (for the softplus derivative)
  float ddtd[2][Q];   // x_j . w_j            (accumulated by the dx epilogue)
  float cdx[2][Q];    // x_j . wd_j
  float rr[2][Q];     // <C_i, (M_0 B)_i> and <C_i, ((M_0 + M_1) B)_i>: the within-chunk part of r_i
  float roff[2][Q];   // dy_i . (exp(lam_i) C_i S_c): the incoming-state part of r_i
  float rbase[1][Q];  // <C_i, acc_i> of what the dC accumulator held before this item (earlier head pairs of the group)
  float gii[2][Q];    // dy_i . x_i
  float zc[2];        // <dS_{c+1}, S_c>
  float lam_last[2];
  int bsafe[2][4];    // per 32-token block: decays by < 2^100, so the factorised diagonal block cannot overflow
  float wsum[2][4];   // scan scratch
  float wsum2[2][4];
};
static_assert(sizeof(BTab) % 16 == 0, "BTab alignment");
constexpr uint32_t SM_BAR = SM_TAB + sizeof(BTab);
enum { BB_TMA = 0, BB_C1, BB_C2, BB_C3, BB_C4, BB_C5, BB_C6, BB_C8, BB_COUNT };
constexpr uint32_t SM_TMEMPTR = SM_BAR + BB_COUNT * 8;
constexpr uint32_t SMEM_BYTES = SM_TMEMPTR + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr uint32_t R0 = 0, R1 = 128, R2 = 256, R3 = 384;

struct BwdArgs {
  const void* dt; const float* A; const void* D; const void* dt_bias;
  void* ddt; float* dB; float* dC; float* dA_part; float* ddtb_part; float* dD_part;
  int64_t dt_b, dt_l, dt_h, ddt_b, ddt_l, ddt_h, dB_b, dB_l, dB_g, dC_b, dC_l, dC_g;
  int B, L, H, G, nchunks;
  int dt_dtype, D_dtype, dtb_dtype, ddt_dtype;
  int dt_softplus;
  float dt_min, dt_max;
  long long* trace; int trace_items;  // debug: clock64 of CTA 0 at phase boundaries (omni_debug_set_bwd_trace)
};
#define BTR(ev)                                                                                   \
  do {                                                                                            \
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && it_count < a.trace_items)           \
      a.trace[it_count * 32 + (ev)] = clock64();                                                  \
  } while (0)

__device__ __forceinline__ float ex2f(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 h2f2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float2 bf2f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t bf16x2_to_f16x2(uint32_t v) {
  return pack_f16_sat(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float softplus_fast(float v) {
  if (v > 20.f) return v;
  const float u = __expf(-fabsf(v));
  const float l = u < 0.03125f ? u * (1.f - u * (0.5f - u * (0.33333334f - u * (0.25f - u * 0.2f)))) : __logf(1.f + u);
  return fmaxf(v, 0.f) + l;
}

// One 32 x 32 block of a decay-weighted matrix: out[c] = val[c] * colf[c] * rowf, optionally masked, packed to fp16.
//   mask: 0 none, 1 keep column <= lane (lower triangle, rows i / columns j), 2 keep column >= lane (rows j / columns i)
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {  // FMUL2: two fp32 multiplies per issue slot
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ void scale_block(const uint32_t (&val)[32], const float* colf, float rowf, int mask, int lane,
                                            uint32_t (&pk)[16]) {
  const float4* cf = reinterpret_cast<const float4*>(colf);
  const float2 rr = make_float2(rowf, rowf);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 f = cf[e];
    const float2 a01 = mul2(mul2(make_float2(__uint_as_float(val[4 * e + 0]), __uint_as_float(val[4 * e + 1])), rr), make_float2(f.x, f.y));
    const float2 a23 = mul2(mul2(make_float2(__uint_as_float(val[4 * e + 2]), __uint_as_float(val[4 * e + 3])), rr), make_float2(f.z, f.w));
    float p0 = a01.x, p1 = a01.y, p2 = a23.x, p3 = a23.y;
    if (mask == 1) {
      p0 = 4 * e + 0 <= lane ? p0 : 0.f; p1 = 4 * e + 1 <= lane ? p1 : 0.f;
      p2 = 4 * e + 2 <= lane ? p2 : 0.f; p3 = 4 * e + 3 <= lane ? p3 : 0.f;
    } else if (mask == 2) {
      p0 = 4 * e + 0 >= lane ? p0 : 0.f; p1 = 4 * e + 1 >= lane ? p1 : 0.f;
      p2 = 4 * e + 2 >= lane ? p2 : 0.f; p3 = 4 * e + 3 >= lane ? p3 : 0.f;
    }
    pk[2 * e] = pack_f16_sat(p0, p1);
    pk[2 * e + 1] = pack_f16_sat(p2, p3);
  }
}
// Diagonal block with extreme decay: the decay of every element directly, exp2(min(sgn (lam_col - lam_row), 0)).
__device__ __forceinline__ void scale_block_direct(const uint32_t (&val)[32], const float* lamc, const float* colmul, float lamr,
                                                float rowmul, float sgn, int mask, int lane, uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float p[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 2 * e + k;
      const float dec = ex2f(fminf(sgn * (lamc[c] - lamr), 0.f));
      const bool keep = mask == 1 ? c <= lane : c >= lane;
      p[k] = keep ? __uint_as_float(val[c]) * dec * rowmul * (colmul ? colmul[c] : 1.f) : 0.f;
    }
    pk[e] = pack_f16_sat(p[0], p[1]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
ssd_tc_bwd_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY,
                  const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
                  const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapDS,
                  const __grid_constant__ CUtensorMap mapDX, BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SM_TMEMPTR);
  BTab* tab = reinterpret_cast<BTab*>(smem + SM_TAB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, wq = warp >> 2;  // TMEM lane quadrant / 32-column block of this warp
  const int row = q * 32 + lane;

  if (tid == 0) {
    for (int i = 0; i < BB_COUNT; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapDY); tma_prefetch_desc(&mapB); tma_prefetch_desc(&mapC);
    tma_prefetch_desc(&mapS); tma_prefetch_desc(&mapDS); tma_prefetch_desc(&mapDX);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;

  const int HP = a.H >> 1, hpg = a.H / a.G;
  const int nitems = a.B * a.nchunks * HP;
  // instruction / shared-memory descriptors are rebuilt where the single issuing thread needs them (keeping fourteen 64-bit
  // descriptors live in every thread's registers across the item loop caused spills)
#define id_kk make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorK)
#define id_ts64 make_idesc(128, 64, kFmtF16, kFmtF16, kMajorK, kMajorMN)
#define id_ts128 make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorMN)
#define dXk make_sdesc(smem_u32(smem + SM_X), 16, 1024)
#define dDYk make_sdesc(smem_u32(smem + SM_DY), 16, 1024)
#define dDYm make_sdesc(smem_u32(smem + SM_DY), 16384, 1024)
#define dBk make_sdesc(smem_u32(smem + SM_B), 16, 1024)
#define dBm make_sdesc(smem_u32(smem + SM_B), 16384, 1024)
#define dCk make_sdesc(smem_u32(smem + SM_C), 16, 1024)
#define dCm make_sdesc(smem_u32(smem + SM_C), 16384, 1024)
#define dSm make_sdesc(smem_u32(smem + SM_S), 16384, 1024)
#define dSk make_sdesc(smem_u32(smem + SM_S), 16, 1024)
#define dDSk make_sdesc(smem_u32(smem + SM_DS), 16, 1024)
#define dDSm make_sdesc(smem_u32(smem + SM_DS), 16384, 1024)
  auto koff = [](uint32_t k) { return ((k >> 2) << 10) + ((k & 3) << 1); };  // k-major 128-wide K: 16-byte units

  // Work split: every CTA takes a CONTIGUOUS range of items, ordered (batch, chunk, head pair): consecutive items then
  // belong to the same (batch, chunk) group, whose dC / dB (sums over the heads of the group) stay in the TMEM accumulators
  // R2 / R3 across items and reach HBM once per group (vector reductions only because a group can straddle two CTAs).
  const int item_lo = (int)(((int64_t)nitems * blockIdx.x) / gridDim.x), item_hi = (int)(((int64_t)nitems * (blockIdx.x + 1)) / gridDim.x);
  uint32_t ph = 0;  // mbarrier phase parity: every barrier completes exactly once per item
  int it_count = 0;
  bool fresh = true;  // R2 / R3 hold nothing yet for the current group
  // raw dt of (head tid >> 7, token tid & 127) of an item, for the table threads; loaded one item ahead
  // (raw bits: converting at the load site makes the thread - and through the phase barrier every warp - wait ~1000 cycles
  // for the strided gather right there; measured with the phase trace)
  auto load_dt = [&](int item_) -> uint32_t {
    if (tid >= 256 || item_ >= item_hi) return 0u;
    const int hp_ = item_ % HP, bc_ = item_ / HP, c_ = bc_ % a.nchunks, b_ = bc_ / a.nchunks;
    const int t_ = c_ * Q + (tid & 127), h_ = hp_ * 2 + ((tid >> 7) & 1);
    if (t_ >= a.L) return 0u;
    const int64_t off = b_ * a.dt_b + (int64_t)t_ * a.dt_l + (int64_t)h_ * a.dt_h;
    return a.dt_dtype == OMNI_F32 ? __ldg(static_cast<const uint32_t*>(a.dt) + off)
                                  : (uint32_t)__ldg(static_cast<const unsigned short*>(a.dt) + off);
  };
  auto dt_to_f = [&](uint32_t bits) -> float {
    if (a.dt_dtype == OMNI_F32) return __uint_as_float(bits);
    if (a.dt_dtype == OMNI_BF16) return __uint_as_float(bits << 16);
    return __half2float(__ushort_as_half((unsigned short)bits));
  };
  uint32_t dt_next = load_dt(item_lo);
  // The six tiles of an item (TMA thread only).  They are requested as soon as the PREVIOUS item's last MMA group has read its
  // tiles (BB_C8) - ~1000 cycles before the item starts - unless that item ends a (batch, chunk) group, whose dC / dB flush
  // stages through the x and B tiles.
  auto issue_loads = [&](int item_) {
    const int hp_ = item_ % HP, bc_ = item_ / HP, c_ = bc_ % a.nchunks, b_ = bc_ / a.nchunks;
    const int h0_ = hp_ * 2, grp_ = h0_ / hpg, t0_ = c_ * Q;
    mbar_expect_tx(&bars[BB_TMA], 6 * 32768);
    tma_load_4d(smem + SM_X, &mapX, &bars[BB_TMA], 0, h0_, t0_, b_);
    tma_load_4d(smem + SM_X + 16384, &mapX, &bars[BB_TMA], 0, h0_ + 1, t0_, b_);
    tma_load_4d(smem + SM_DY, &mapDY, &bars[BB_TMA], 0, h0_, t0_, b_);
    tma_load_4d(smem + SM_DY + 16384, &mapDY, &bars[BB_TMA], 0, h0_ + 1, t0_, b_);
    tma_load_4d(smem + SM_B, &mapB, &bars[BB_TMA], 0, grp_, t0_, b_);
    tma_load_4d(smem + SM_B + 16384, &mapB, &bars[BB_TMA], 64, grp_, t0_, b_);
    tma_load_4d(smem + SM_C, &mapC, &bars[BB_TMA], 0, grp_, t0_, b_);
    tma_load_4d(smem + SM_C + 16384, &mapC, &bars[BB_TMA], 64, grp_, t0_, b_);
    tma_load_4d(smem + SM_S, &mapS, &bars[BB_TMA], 0, h0_ * HD, c_, b_);
    tma_load_4d(smem + SM_S + 16384, &mapS, &bars[BB_TMA], 64, h0_ * HD, c_, b_);
    tma_load_4d(smem + SM_DS, &mapDS, &bars[BB_TMA], 0, h0_ * HD, c_, b_);
    tma_load_4d(smem + SM_DS + 16384, &mapDS, &bars[BB_TMA], 64, h0_ * HD, c_, b_);
  };
  bool loads_issued = false;   // (meaningful in the TMA thread only)
#pragma unroll 1
  for (int item = item_lo; item < item_hi; ++item, ph ^= 1, ++it_count) {
    const int hp = item % HP, bc = item / HP, c = bc % a.nchunks, b = bc / a.nchunks;
    const int h0 = hp * 2, grp = h0 / hpg, t0 = c * Q;
    const bool last_of_group = item + 1 == item_hi || (item + 1) / HP != bc || (h0 + 2) / hpg != grp;

    BTR(0);
    // ---- A. tile loads (one thread; normally already issued at the end of the previous item, see `issue_loads`) and the
    //         decay tables (threads 0..255: head tid >> 7, token tid & 127) ---------------------------------------------
    if (tid == kTmaThread && !loads_issued) issue_loads(item);
    loads_issued = false;
    float my_dt = 0.f, my_lam = 0.f;
    const int th = (tid >> 7) & 1, tj = tid & 127;  // (head, token) of the table threads
    if (tid < 256) {
      const int t = t0 + tj, h = h0 + th;
      float vpre = 0.f;
      if (t < a.L) {
        vpre = dt_to_f(dt_next);
        if (a.dt_bias) vpre += ld_any(a.dt_bias, a.dtb_dtype, h);
        float v = a.dt_softplus ? softplus_fast(vpre) : vpre;
        my_dt = fminf(fmaxf(v, a.dt_min), a.dt_max);
      }
      tab->vpre[th][tj] = vpre;
      float incl = my_dt * a.A[h] * 1.4426950408889634f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      my_lam = incl;
      if (lane == 31) tab->wsum[th][(tid >> 5) & 3] = incl;
      tab->ddtd[th][tj] = 0.f;
      tab->cdx[th][tj] = 0.f;
      tab->rr[th][tj] = 0.f;
      tab->roff[th][tj] = 0.f;
      tab->rbase[th & 0][tj] = 0.f;
      if (tj < 2 && th == 0) tab->zc[tj] = 0.f;
    }
    if (tid < 256) {  // (the upper half of the CTA needs lam and dt of its token too)
      tab->lam[th][tj] = my_lam;   // still relative to the start of the token's 32-block
      tab->dtv[th][tj] = my_dt;
    }
    __syncthreads();
    {
      const int wj = tj >> 5;  // 32-token block of this token
      float ref[4];            // lam just before block 0..3 (= running sum of the earlier warps' totals)
      ref[0] = 0.f;
      ref[1] = tab->wsum[th][0];
      ref[2] = ref[1] + tab->wsum[th][1];
      ref[3] = ref[2] + tab->wsum[th][2];
      const float lam_last = ref[3] + tab->wsum[th][3];
      const float myref = wj == 0 ? ref[0] : (wj == 1 ? ref[1] : (wj == 2 ? ref[2] : ref[3]));
      const float rel = tid < 256 ? my_lam : tab->lam[th][tj];   // lam_j - ref(block of j)
      const float dtj = tid < 256 ? my_dt : tab->dtv[th][tj];
      const float lamj = rel + myref;
      if (tid < 256) {  // lower half: the tables indexed by the token itself
        tab->ci[th][tj] = ex2f(rel);
        tab->vd[th][tj] = ex2f(-rel) * dtj;
        tab->eL[th][tj] = ex2f(lamj);
        const bool ok = __all_sync(0xffffffffu, -rel < 100.f);
        if (lane == 0) tab->bsafe[th][wj] = ok ? 1 : 0;
        if (tj == 0) tab->lam_last[th] = lam_last;
      } else {          // upper half: the tables that refer to later blocks / the end of the chunk
        tab->es[th][tj] = ex2f(lam_last - lamj);
#pragma unroll
        for (int w = 1; w < 4; ++w)
          if (wj < w) tab->v[th][w - 1][tj] = ex2f(ref[w] - lamj) * dtj;
      }
      __syncthreads();  // every thread has read the block-relative lam before it is replaced by the chunk-wide value
      if (tid < 256) tab->lam[th][tj] = lamj;
    }
    BTR(1);
    mbar_wait(&bars[BB_TMA], ph);
    BTR(2);
    if (tid == kTmaThread && item + 1 < item_hi) {  // next item's x / dy / S / dS tiles -> L2 while this item computes
      const int hp1 = (item + 1) % HP, bc1 = (item + 1) / HP, c1 = bc1 % a.nchunks, b1 = bc1 / a.nchunks, g1 = hp1 * 2;
      tma_prefetch_4d(&mapX, 0, g1, c1 * Q, b1);
      tma_prefetch_4d(&mapX, 0, g1 + 1, c1 * Q, b1);
      tma_prefetch_4d(&mapDY, 0, g1, c1 * Q, b1);
      tma_prefetch_4d(&mapDY, 0, g1 + 1, c1 * Q, b1);
      tma_prefetch_4d(&mapS, 0, g1 * HD, c1, b1);
      tma_prefetch_4d(&mapS, 64, g1 * HD, c1, b1);
      tma_prefetch_4d(&mapDS, 0, g1 * HD, c1, b1);
      tma_prefetch_4d(&mapDS, 64, g1 * HD, c1, b1);
    }
    // ---- C. G1: CB^T = B C^T -> R0;  G3: ws = B dS16^T -> R1  (neither needs x / dy: issued before their conversion) ----
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) mma_ss(tb + R0, dBk + koff(k), dCk + koff(k), id_kk, k > 0);
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) mma_ss(tb + R1, dBk + koff(k), dDSk + koff(k), id_kk, k > 0);
      mma_commit(&bars[BB_C1]);
    }
    // ---- B. x, dy: bf16 -> fp16 in place (while G1 / G3 run) --------------------------------------------------------------
    {
      uint4* xt = reinterpret_cast<uint4*>(smem + SM_X);  // SM_X and SM_DY are adjacent: 4096 16-byte slots
#pragma unroll 2
      for (int k = 0; k < 8; ++k) {
        uint4 v = xt[tid + 512 * k];
        v.x = bf16x2_to_f16x2(v.x); v.y = bf16x2_to_f16x2(v.y); v.z = bf16x2_to_f16x2(v.z); v.w = bf16x2_to_f16x2(v.w);
        xt[tid + 512 * k] = v;
      }
    }
    fence_proxy_async_smem();
    BTR(3);
    // ---- D. PT_h[j][i] = CB^T[j][i] L_ij (i >= j), fp16, in place (rows j, 32-column block ib = wq) ---------------------
    __syncthreads();               // tables (and the fp16 tiles) published
    mbar_wait(&bars[BB_C1], ph);
    BTR(4);
    bool safe = true;
#pragma unroll
    for (int e = 0; e < 8; ++e) safe = safe && (&tab->bsafe[0][0])[e] != 0;
    tc_fence_after();
    uint32_t wsr[2][16];  // ws of this thread's (row j, 16 columns p = 16 wq ..) for both heads: R1 is re-used for wd
    {
      uint32_t cbt[32];
      tmem_ld32(tmem_addr(tb, q * 32, R0 + 32 * wq), cbt);  // (unconditional: a predicated load would put cbt in local memory)
      tmem_ld16(tmem_addr(tb, q * 32, R1 + 16 * wq), wsr[0]);
      tmem_ld16(tmem_addr(tb, q * 32, R1 + 64 + 16 * wq), wsr[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncthreads();  // every block is in registers before any warp overwrites the region with fp16
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[16];
        if (wq < q) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        } else if (wq > q || safe) {
          const float refb = wq == 0 ? 0.f : tab->lam[h][32 * wq - 1];
          scale_block(cbt, &tab->ci[h][32 * wq], ex2f(refb - tab->lam[h][row]), wq == q ? 2 : 0, lane, pk);
        } else {
          scale_block_direct(cbt, &tab->lam[h][32 * wq], nullptr, tab->lam[h][row], 1.f, 1.f, 2, lane, pk);
        }
        tmem_st16(tmem_addr(tb, q * 32, R0 + 64 * h + 16 * wq), pk);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    BTR(5);
    // ---- E. G2: wd_h = PT_h dy_h -> R1 + 64 h ----------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (uint32_t hk = 0; hk < 16; ++hk) {
        const uint32_t h = hk >> 3, k = hk & 7;
        mma_ts(tb + R1 + 64 * h, tb + R0 + 64 * h + 8 * k, dDYm + h * 1024 + k * 128, id_ts64, k > 0);
      }
      mma_commit(&bars[BB_C2]);
    }
    // <C_i, acc_i> over this warp's 32 columns n of the dC accumulator (row i = `row`), added to dst[row]
    auto dot_c_acc = [&](float* dst) {
      uint32_t v[32];
      tmem_ld32(tmem_addr(tb, q * 32, R2 + 32 * wq), v);
      float part = 0.f;
      const uint8_t* crow = smem + SM_C + (wq >> 1) * 16384;
      uint4 cv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cv[k] = *reinterpret_cast<const uint4*>(crow + sw128(row, (wq & 1) * 4 + k));
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t cw[4] = {cv[k].x, cv[k].y, cv[k].z, cv[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 cf = h2f2(cw[e]);
          part += cf.x * __uint_as_float(v[8 * k + 2 * e]) + cf.y * __uint_as_float(v[8 * k + 2 * e + 1]);
        }
      }
      atomicAdd(&dst[row], part);
    };
    if (!fresh) dot_c_acc(tab->rbase[0]);  // what the dC accumulator already holds from earlier items of the group
    // (both of these only need what is already there: they run while G2 executes)
    {  // zc_h = <dS_{c+1}, S_c> over the head's 64 rows: the two fp16 tiles share their (swizzled) layout, so it is an
       // elementwise product of equal offsets - 32 values per thread (it used to be the trace of a 128 x 128 x 128 GEMM)
      const uint4* st = reinterpret_cast<const uint4*>(smem + SM_S);
      const uint4* dst = reinterpret_cast<const uint4*>(smem + SM_DS);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int slot = warp * 128 + k * 32 + lane;      // 16-byte slot; row = (slot >> 3) & 127: a warp stays inside one head
        const uint4 sv = st[slot], dv = dst[slot];
        const uint32_t sw4[4] = {sv.x, sv.y, sv.z, sv.w}, dw4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a2 = h2f2(sw4[e]), b2 = h2f2(dw4[e]);
          acc += a2.x * b2.x + a2.y * b2.y;
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(&tab->zc[(warp >> 2) & 1], acc);   // slots [warp * 128, +128): rows 16 (warp & 7) .. +15 of n-half warp >> 3
    }
    // ---- F. dx_j = dt_j (wd_j + es_j ws_j) + D dy_j; x.w and x.wd row sums; one head at a time through the staging tile --
    mbar_wait(&bars[BB_C2], ph);
    BTR(6);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t wd[16];
      tmem_ld16(tmem_addr(tb, q * 32, R1 + 64 * h + 16 * wq), wd);
      tmem_ld_wait();
      const float esj = tab->es[h][row], dtj = tab->dtv[h][row];
      const float Dh = a.D ? ld_any(a.D, a.D_dtype, h0 + h) : 0.f;
      float sw = 0.f, swd = 0.f;
      if (h == 1) {  // the staging tile is reused: the TMA store of head 0 must have read it
        if (tid == kTmaThread) tma_store_wait_read<0>();
        __syncthreads();
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t off = h * 16384 + sw128(row, 2 * wq + k);
        const uint4 xv = *reinterpret_cast<const uint4*>(smem + SM_X + off);
        const uint4 dv = *reinterpret_cast<const uint4*>(smem + SM_DY + off);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 xf = h2f2(xw[e]), df = h2f2(dw[e]);
          const float d0 = __uint_as_float(wd[8 * k + 2 * e]), d1 = __uint_as_float(wd[8 * k + 2 * e + 1]);
          const float w0 = d0 + esj * __uint_as_float(wsr[h][8 * k + 2 * e]), w1 = d1 + esj * __uint_as_float(wsr[h][8 * k + 2 * e + 1]);
          sw += xf.x * w0 + xf.y * w1;
          swd += xf.x * d0 + xf.y * d1;
          o[e] = pack_bf16(dtj * w0 + Dh * df.x, dtj * w1 + Dh * df.y);
        }
        *reinterpret_cast<uint4*>(smem + SM_STG + sw128(row, 2 * wq + k)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      atomicAdd(&tab->ddtd[h][row], sw);
      atomicAdd(&tab->cdx[h][row], swd);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == kTmaThread) {  // rows beyond L are clipped by the tensor map
        tma_store_4d(&mapDX, smem + SM_STG, 0, h0 + h, t0, b);
        tma_store_commit();
      }
    }
    BTR(7);
    dt_next = load_dt(item + 1);  // in flight during the rest of this item
    // ---- G. G4: G_h0 = dy_h0 x_h0^T -> R0;  G4': G_h0^T -> R1 ---------------------------------------------------------
    auto issue_g = [&](uint32_t h) {
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) mma_ss(tb + R0, dDYk + h * 1024 + 2 * k, dXk + h * 1024 + 2 * k, id_kk, k > 0);
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) mma_ss(tb + R1, dXk + h * 1024 + 2 * k, dDYk + h * 1024 + 2 * k, id_kk, k > 0);
    };
    if (tid == 0) {
      tc_fence_after();
      issue_g(0);
      mma_commit(&bars[BB_C3]);
    }
    // ---- H/J. M_h = G_h o L o dt (rows i, cols j <= i) in R0, MT_h = G_h^T o L o dt (rows j, cols i >= j) in R1 ---------
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      mbar_wait(&bars[h == 0 ? BB_C3 : BB_C4], ph);
      BTR(8 + 2 * h);
      tc_fence_after();
      if (h == 1) dot_c_acc(tab->rr[0]);  // base + M_0 B: head 0's within-chunk part of r_i
      uint32_t pk[16];
      {  // M block (row i = `row`, column block jb = wq): the source block is in registers before any warp stores fp16
        uint32_t g[32];
        tmem_ld32(tmem_addr(tb, q * 32, R0 + 32 * wq), g);
        tmem_ld_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (wq > q) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        } else if (wq < q) {
          scale_block(g, &tab->v[h][q - 1][32 * wq], tab->ci[h][row], 0, lane, pk);
        } else {
          if (safe) scale_block(g, &tab->vd[h][32 * wq], tab->ci[h][row], 1, lane, pk);
          else scale_block_direct(g, &tab->lam[h][32 * wq], &tab->dtv[h][32 * wq], tab->lam[h][row], 1.f, -1.f, 1, lane, pk);
          float gd = 0.f;  // G_ii = dy_i . x_i sits on the diagonal of this block
#pragma unroll
          for (int e = 0; e < 32; ++e) gd = e == lane ? __uint_as_float(g[e]) : gd;
          tab->gii[h][row] = gd;
        }
        tmem_st16(tmem_addr(tb, q * 32, R0 + 16 * wq), pk);
      }
      {  // MT block (row j = `row`, column block ib = wq), in R1
        uint32_t gt[32];
        tmem_ld32(tmem_addr(tb, q * 32, R1 + 32 * wq), gt);
        tmem_ld_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (wq < q) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        } else if (wq > q || safe) {
          const float refb = wq == 0 ? 0.f : tab->lam[h][32 * wq - 1];
          scale_block(gt, &tab->ci[h][32 * wq], ex2f(refb - tab->lam[h][row]) * tab->dtv[h][row], wq == q ? 2 : 0, lane, pk);
        } else {
          scale_block_direct(gt, &tab->lam[h][32 * wq], nullptr, tab->lam[h][row], tab->dtv[h][row], 1.f, 2, lane, pk);
        }
      }
      tmem_st16(tmem_addr(tb, q * 32, R1 + 16 * wq), pk);
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      BTR(9 + 2 * h);
      // ---- I/K. G5: dC (+)= M_h B -> R2;  G7: dB (+)= MT_h C -> R3;  then the next head's G / G^T ----------------------
      if (tid == 0) {
        tc_fence_after();
        const bool first = fresh && h == 0;  // the accumulators start a new (batch, chunk) group
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) mma_ts(tb + R2, tb + R0 + 8 * k, dBm + k * 128, id_ts128, !(first && k == 0));
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) mma_ts(tb + R3, tb + R1 + 8 * k, dCm + k * 128, id_ts128, !(first && k == 0));
        if (h == 0) {
          issue_g(1);
          mma_commit(&bars[BB_C4]);
        } else {
          mma_commit(&bars[BB_C5]);
        }
      }
    }
    // ---- L. x16 -> X' = es dt x, dy16 -> exp(lam_i) dy, in place (packed fp16 multiplies) - BEFORE waiting for G5 / G7 of head 1:
    //         those read M_1 / MT_1 (TMEM) and the B / C tiles, not x / dy, so the scaling runs under them ---------------
    {
      uint4* xt = reinterpret_cast<uint4*>(smem + SM_X);
#pragma unroll 2
      for (int k = 0; k < 8; ++k) {
        const int slot = tid + 512 * k;
        const int r = (slot >> 3) & 127, hh = (slot >> 10) & 1;
        const float s = slot < 2048 ? tab->es[hh][r] * tab->dtv[hh][r] : tab->eL[hh][r];
        const __half2 s2 = __float2half2_rn(s);
        uint4 v = xt[slot];
        __half2* hv = reinterpret_cast<__half2*>(&v);
        hv[0] = __hmul2(hv[0], s2); hv[1] = __hmul2(hv[1], s2); hv[2] = __hmul2(hv[2], s2); hv[3] = __hmul2(hv[3], s2);
        xt[slot] = v;
      }
    }
    mbar_wait(&bars[BB_C5], ph);
    BTR(12);
    tc_fence_after();
    dot_c_acc(tab->rr[1]);  // base + (M_0 + M_1) B: both heads' within-chunk parts
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    BTR(13);
    // ---- M. G10: C S16^T -> R0;  then G6: dC += (exp(lam) dy) S_c -> R2;  G8: dB += X' dS_{c+1} -> R3 ----
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id_kmn = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorMN);
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) mma_ss(tb + R0, dCk + koff(k), dSk + koff(k), id_kk, k > 0);    // G10 first: roff waits for it
      mma_commit(&bars[BB_C6]);
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) mma_ss(tb + R2, dDYk + koff(k), dSm + k * 128, id_kmn, true);   // G6, G8 only feed the accumulators:
#pragma unroll
      for (uint32_t k = 0; k < 8; ++k) mma_ss(tb + R3, dXk + koff(k), dDSm + k * 128, id_kmn, true);   // they run under roff / da / ddt
      mma_commit(&bars[BB_C8]);
    }
    // ---- N0. roff_i = (exp(lam_i) dy_i) . (C_i S_c): rows i, this warp's 32 columns (h, p) of G10 ---------------------------
    mbar_wait(&bars[BB_C6], ph);
    BTR(14);
    tc_fence_after();
    {
      uint32_t v[32];
      tmem_ld32(tmem_addr(tb, q * 32, R0 + 32 * wq), v);
      const uint8_t* drow = smem + SM_DY + (wq >> 1) * 16384;
      uint4 dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dv[k] = *reinterpret_cast<const uint4*>(drow + sw128(row, (wq & 1) * 4 + k));
      tmem_ld_wait();
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t dw[4] = {dv[k].x, dv[k].y, dv[k].z, dv[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 df = h2f2(dw[e]);
          part += df.x * __uint_as_float(v[8 * k + 2 * e]) + df.y * __uint_as_float(v[8 * k + 2 * e + 1]);
        }
      }
      atomicAdd(&tab->roff[wq >> 1][row], part);
    }
    tc_fence_before();
    __syncthreads();
    // ---- N1. da, ddt and the per-(batch, head) parameter sums: warps 0-3 head 0, warps 4-7 head 1, one token per lane ----
    {
      const int h = (warp >> 2) & 1, hg = h0 + h, j = (warp & 3) * 32 + lane;
      float e1 = 0.f, e2 = 0.f, dtk = 0.f, i1 = 0.f, i2 = 0.f;
      if (warp < 8) {
        dtk = tab->dtv[h][j];
        const float cd = dtk * tab->cdx[h][j];                       // x_j . dxdiag_j
        e2 = dtk * (tab->ddtd[h][j] - tab->cdx[h][j]);               // x_j . dxstate_j
        const float r = (h == 0 ? tab->rr[0][j] - tab->rbase[0][j] : tab->rr[1][j] - tab->rr[0][j]) + tab->roff[h][j];  // dy_j . (y_j - D x_j)
        e1 = r - cd;
        i1 = e1; i2 = e2;  // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float u1 = __shfl_up_sync(0xffffffffu, i1, o), u2 = __shfl_up_sync(0xffffffffu, i2, o);
          if (lane >= o) { i1 += u1; i2 += u2; }
        }
        if (lane == 31) { tab->wsum[h][warp & 3] = i1; tab->wsum2[h][warp & 3] = i2; }
      }
      __syncthreads();
      if (warp < 8) {
        const int wj = warp & 3;
        float before1 = 0.f, before2 = 0.f, tot1 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float t1 = tab->wsum[h][w], t2 = tab->wsum2[h][w];
          tot1 += t1;
          if (w < wj) { before1 += t1; before2 += t2; }
        }
        const float rev = tot1 - (before1 + i1 - e1);   // sum of e1 over tokens >= j
        const float fwd = before2 + i2 - e2;            // sum of e2 over tokens < j
        const float da = rev + fwd + ex2f(tab->lam_last[h]) * tab->zc[h];
        float ddt = tab->ddtd[h][j] + a.A[hg] * da;
        float sA = dtk * da, sB = 0.f, sD = tab->gii[h][j];
        // back through clamp and softplus
        const float vpre = tab->vpre[h][j];
        const float vact = a.dt_softplus ? softplus_fast(vpre) : vpre;
        if (vact < a.dt_min || vact > a.dt_max) ddt = 0.f;
        if (a.dt_softplus && vpre <= 20.f) ddt *= 1.f / (1.f + __expf(-vpre));
        const int t = t0 + j;
        if (t < a.L) {
          st_any(a.ddt, a.ddt_dtype, b * a.ddt_b + (int64_t)t * a.ddt_l + (int64_t)hg * a.ddt_h, ddt);
          sB = ddt;
        }
        sA = warp_sum(sA); sB = warp_sum(sB); sD = warp_sum(sD);
        if (lane == 0) {
          atomicAdd(a.dA_part + b * a.H + hg, sA);
          atomicAdd(a.ddtb_part + b * a.H + hg, sB);
          atomicAdd(a.dD_part + b * a.H + hg, sD);
        }
      }
    }
    BTR(15);
    fresh = false;
    mbar_wait(&bars[BB_C8], ph);   // G6 / G8 have read the S / dS / x / dy tiles (next item's loads) and completed dC / dB (flush)
    tc_fence_after();
    if (tid == kTmaThread && !last_of_group && item + 1 < item_hi) {
      // every MMA group of this item is complete and no SIMT phase below reads a tile: the next item's loads start now
      issue_loads(item + 1);
      loads_issued = true;
    }
    // ---- N2. end of a (batch, chunk) group: dC (R2), dB (R3): TMEM -> fp32 staging over the dead x/dy and B/C tiles ->
    //          coalesced vector reductions ----------------------------------------------------------------------------------
    if (last_of_group) {
      __syncthreads();
      BTR(16);
      float* stC = reinterpret_cast<float*>(smem + SM_X);   // [128 rows][32 float4 chunks], chunk slot = chunk ^ (row & 31)
      float* stB = reinterpret_cast<float*>(smem + SM_B);
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        uint32_t v[32];
        tmem_ld32(tmem_addr(tb, q * 32, (which ? R3 : R2) + 32 * wq), v);
        tmem_ld_wait();
        float* dstp = (which ? stB : stC) + row * 128;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          *reinterpret_cast<uint4*>(dstp + (((8 * wq + e) ^ (row & 31)) << 2)) = make_uint4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      }
      tc_fence_before();
      __syncthreads();
#pragma unroll 1
      for (int k = 0; k < 16; ++k) {  // 2 tiles x 128 rows x 32 chunks = 8192 float4 reductions, 16 per thread
        const int idx = tid + 512 * k, which = idx >> 12, r = (idx >> 5) & 127, ch = idx & 31;
        if (t0 + r < a.L) {
          const float4 val = *reinterpret_cast<const float4*>((which ? stB : stC) + r * 128 + ((ch ^ (r & 31)) << 2));
          float* gp = which ? a.dB + b * a.dB_b + (int64_t)(t0 + r) * a.dB_l + (int64_t)grp * a.dB_g + 4 * ch
                            : a.dC + b * a.dC_b + (int64_t)(t0 + r) * a.dC_l + (int64_t)grp * a.dC_g + 4 * ch;
          red_add_v4(gp, val.x, val.y, val.z, val.w);
        }
      }
      fresh = true;
    }
    BTR(17);
    if (tid == kTmaThread) tma_store_wait_read<0>();  // the dx staging tile and the next item's loads
    tc_fence_before();
    __syncthreads();
  }
  if (tid == kTmaThread) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

long long* g_btrace = nullptr;
int g_btrace_items = 0;
}  // namespace

// ---- host ---------------------------------------------------------------------------------------------------------
int ssd_tc_state_sweep(int mode, const omni_tensor_t& xlike, const omni_tensor_t& dt, const omni_tensor_t& A,
                       const omni_tensor_t& dt_bias, const omni_tensor_t& init, const omni_tensor_t& fin, const void* ws_bslot,
                       void* ws_states, int64_t G, int dt_softplus, float dt_min, float dt_max, cudaStream_t s,
                       void* hand_slots, void* piece_ws);
int64_t ssd_tc_sweep_piece_bytes(int64_t Bsz, int64_t L, int64_t H);
int64_t ssd_tc_hand_bytes();  // state hand-off slots + flags of the half-item schedule (one set per sweep)
int ssd_tc_prep(const omni_tensor_t& Bm, const omni_tensor_t& Cm, void* wsB, void* wsC, cudaStream_t s);

int64_t ssd_tc_bwd_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t ngroups) {
  const int64_t nchunks = (seqlen + Q - 1) / Q;
  const int64_t bc = 2 * batch * seqlen * ngroups * NS * 2;            // fp16 copies of B and C
  const int64_t st = batch * nchunks * nheads * HD * NS * 2;           // fp16 states, one tensor
  return ((bc + 255) / 256) * 256 + 2 * st + 256 + 2 * ssd_tc_hand_bytes() + 256 + ssd_tc_sweep_piece_bytes(batch, seqlen, nheads);
}

bool ssd_tc_bwd_supported(const omni_ssd_bwd_params_t* p) {
  const omni_tensor_t &x = p->x, &Bm = p->B, &Cm = p->C, &dy = p->dout, &dx = p->dx;
  auto bf16_4d = [](const omni_tensor_t& t, int64_t inner) {
    return present(t) && t.ndim == 4 && t.dtype == OMNI_BF16 && t.shape[3] == inner && t.stride[3] == 1 && aligned16(t.data);
  };
  if (!bf16_4d(x, HD) || !bf16_4d(dy, HD) || !bf16_4d(dx, HD) || !bf16_4d(Bm, NS) || !bf16_4d(Cm, NS)) return false;
  const int64_t H = x.shape[2], G = Bm.shape[2];
  if (G <= 0 || H % G != 0 || (H / G) % 2 != 0) return false;
  if (present(p->z) || present(p->dz) || present(p->seq_idx)) return false;
  if (present(p->D) && p->D.ndim != 1) return false;
  for (const omni_tensor_t* t : {&x, &dy, &dx, &Bm, &Cm})
    for (int d = 0; d < 3; ++d)
      if (t->shape[d] > 1 && (t->stride[d] * 2) % 16 != 0) return false;
  for (const omni_tensor_t* t : {&p->dB, &p->dC}) {
    if (!present(*t) || t->dtype != OMNI_F32 || t->ndim != 4 || t->stride[3] != 1 || !aligned16(t->data)) return false;
    for (int d = 0; d < 3; ++d)
      if (t->shape[d] > 1 && t->stride[d] % 4 != 0) return false;
  }
  if (!present(p->dA_part) || !present(p->ddt_bias_part) || !present(p->dD_part)) return false;
  if (p->dD_part.ndim != 2) return false;
  for (const omni_tensor_t* t : {&p->dA_part, &p->ddt_bias_part, &p->dD_part})
    if (t->dtype != OMNI_F32 || t->ndim != 2 || t->stride[1] != 1 || t->stride[0] != H) return false;
  if (present(p->dinitial_states) && (p->dinitial_states.dtype != OMNI_F32)) return false;
  const omni_tensor_t& ws = p->workspace;
  const int64_t need = ssd_tc_bwd_workspace_bytes(x.shape[0], x.shape[1], H, G);
  if (!present(ws) || ws.ndim != 1 || ws.stride[0] != 1 || ws.shape[0] * dtype_size(ws.dtype) < need ||
      (reinterpret_cast<uintptr_t>(ws.data) & 255) != 0)
    return false;
  return get_encode_tiled() != nullptr;
}

int ssd_tc_bwd(const omni_ssd_bwd_params_t* p, cudaStream_t s) {
  const omni_tensor_t &x = p->x, &dt = p->dt, &Bm = p->B, &Cm = p->C, &dy = p->dout, &dx = p->dx;
  const int64_t Bsz = x.shape[0], L = x.shape[1], H = x.shape[2], G = Bm.shape[2];
  const int64_t nchunks = (L + Q - 1) / Q;
  OMNI_CHECK(present(dt) && shape_is(dt, 3, Bsz, L, H) && is_float_dtype(dt.dtype), OMNI_BAD_SHAPE, "ssd bwd: dt must be (B, L, H)");
  OMNI_CHECK(present(p->ddt) && shape_is(p->ddt, 3, Bsz, L, H) && is_float_dtype(p->ddt.dtype), OMNI_BAD_SHAPE,
             "ssd bwd: ddt must be (B, L, H)");
  OMNI_CHECK(present(p->A) && shape_is(p->A, 1, H) && p->A.dtype == OMNI_F32 && (H <= 1 || p->A.stride[0] == 1), OMNI_BAD_SHAPE,
             "ssd bwd: A must be contiguous fp32 (H)");
  OMNI_CHECK(shape_is(Bm, 4, Bsz, L, G, NS) && shape_is(Cm, 4, Bsz, L, G, NS) && shape_is(p->dB, 4, Bsz, L, G, NS) &&
                 shape_is(p->dC, 4, Bsz, L, G, NS), OMNI_BAD_SHAPE, "ssd bwd: B/C/dB/dC must be (B, L, G, N)");
  OMNI_CHECK(shape_is(dy, 4, Bsz, L, H, HD) && shape_is(dx, 4, Bsz, L, H, HD), OMNI_BAD_SHAPE, "ssd bwd: dout/dx must match x");
  if (present(p->D))
    OMNI_CHECK(shape_is(p->D, 1, H) && is_float_dtype(p->D.dtype) && (H <= 1 || p->D.stride[0] == 1), OMNI_BAD_SHAPE,
               "ssd bwd: D must be contiguous (H)");
  if (present(p->dt_bias))
    OMNI_CHECK(shape_is(p->dt_bias, 1, H) && is_float_dtype(p->dt_bias.dtype) && (H <= 1 || p->dt_bias.stride[0] == 1),
               OMNI_BAD_SHAPE, "ssd bwd: dt_bias must be contiguous (H)");

  // workspace: [fp16 B | fp16 C | pad to 256 | fp16 S_c (B, nchunks, H*64, 128) | fp16 dS_{c+1} (same)]
  uint8_t* ws = static_cast<uint8_t*>(p->workspace.data);
  const int64_t rows = Bsz * L * G;
  __half* wsB = reinterpret_cast<__half*>(ws);
  __half* wsC = wsB + rows * NS;
  const int64_t bc = ((2 * rows * NS * 2 + 255) / 256) * 256;
  const int64_t st_bytes = Bsz * nchunks * H * HD * NS * 2;
  __half* wsS = reinterpret_cast<__half*>(ws + bc);
  __half* wsDS = reinterpret_cast<__half*>(ws + bc + st_bytes);
  uint8_t* hand = ws + bc + 2 * st_bytes;
  hand += (256 - reinterpret_cast<uintptr_t>(hand) % 256) % 256;
  uint8_t* piece_ws = hand + 2 * ssd_tc_hand_bytes();
  piece_ws += (256 - reinterpret_cast<uintptr_t>(piece_ws) % 256) % 256;
  if (ssd_tc_sweep_piece_bytes(Bsz, L, H) == 0) piece_ws = nullptr;

  if (int rc = ssd_tc_prep(Bm, Cm, wsB, wsC, s)) return rc;
  omni_tensor_t none{};
  // forward states S_c (x, B, sj) and reverse state gradients dS_{c+1} (dy, C, exp(lam)); the reverse sweep's final
  // state is dS_0 = the gradient of initial_states
  // (skipped when the forward left its chunk states - omnissm.h: chunk_states - for this call)
  const omni_tensor_t& cs = p->chunk_states;
  const bool have_states = present(cs) && cs.ndim == 1 && cs.stride[0] == 1 && cs.dtype == OMNI_F16 && aligned16(cs.data) &&
                           cs.shape[0] * 2 >= st_bytes;
  if (present(cs)) OMNI_CHECK(have_states, OMNI_BAD_SHAPE, "ssd bwd: chunk_states must be the forward's 1-D fp16 tensor");
  if (have_states) wsS = static_cast<__half*>(cs.data);
  else if (int rc = ssd_tc_state_sweep(1, x, dt, p->A, p->dt_bias, p->initial_states, none, wsB, wsS, G, p->dt_softplus, p->dt_min,
                                       p->dt_max, s, hand, piece_ws))
    return rc;
  if (int rc = ssd_tc_state_sweep(2, dy, dt, p->A, p->dt_bias, p->dfinal_states, p->dinitial_states, wsC, wsDS, G, p->dt_softplus,
                                  p->dt_min, p->dt_max, s, hand + ssd_tc_hand_bytes(), piece_ws))
    return rc;

  BwdArgs a{};
  a.dt = dt.data; a.dt_dtype = dt.dtype; a.dt_b = dt.stride[0]; a.dt_l = dt.stride[1]; a.dt_h = dt.stride[2];
  a.A = static_cast<const float*>(p->A.data);
  if (present(p->D)) { a.D = p->D.data; a.D_dtype = p->D.dtype; }
  if (present(p->dt_bias)) { a.dt_bias = p->dt_bias.data; a.dtb_dtype = p->dt_bias.dtype; }
  a.ddt = p->ddt.data; a.ddt_dtype = p->ddt.dtype; a.ddt_b = p->ddt.stride[0]; a.ddt_l = p->ddt.stride[1]; a.ddt_h = p->ddt.stride[2];
  a.dB = static_cast<float*>(p->dB.data); a.dB_b = p->dB.stride[0]; a.dB_l = p->dB.stride[1]; a.dB_g = p->dB.stride[2];
  a.dC = static_cast<float*>(p->dC.data); a.dC_b = p->dC.stride[0]; a.dC_l = p->dC.stride[1]; a.dC_g = p->dC.stride[2];
  a.dA_part = static_cast<float*>(p->dA_part.data);
  a.ddtb_part = static_cast<float*>(p->ddt_bias_part.data);
  a.dD_part = static_cast<float*>(p->dD_part.data);
  a.B = (int)Bsz; a.L = (int)L; a.H = (int)H; a.G = (int)G; a.nchunks = (int)nchunks;
  a.dt_softplus = p->dt_softplus; a.dt_min = p->dt_min; a.dt_max = p->dt_max;
  a.trace = g_btrace; a.trace_items = g_btrace_items;
  // the three per-(batch, head) sums are accumulated with atomics across the chunks of a sequence
  cudaMemsetAsync(a.dA_part, 0, sizeof(float) * Bsz * H, s);
  cudaMemsetAsync(a.ddtb_part, 0, sizeof(float) * Bsz * H, s);
  cudaMemsetAsync(a.dD_part, 0, sizeof(float) * Bsz * H, s);

  auto tmap4 = [&](CUtensorMap* m, const void* base, const int64_t* shape, const int64_t* stride, bool bf16, int inner,
                   int rows_) -> int {
    const uint64_t dims[4] = {(uint64_t)shape[3], (uint64_t)shape[2], (uint64_t)shape[1], (uint64_t)shape[0]};
    auto st = [&](int d) { return (uint64_t)(shape[d] > 1 ? stride[d] : shape[3]) * 2; };
    const uint64_t strides[3] = {st(2), st(1), st(0)};
    const uint32_t box[4] = {(uint32_t)inner, 1, (uint32_t)rows_, 1};
    return make_tmap_16bit(m, base, 4, dims, strides, box, bf16);
  };
  const int64_t bc_shape[4] = {Bsz, L, G, NS}, bc_stride[4] = {L * G * NS, G * NS, NS, 1};
  CUtensorMap mX, mDY, mB, mC, mS, mDS, mDX;
  if (int rc = tmap4(&mX, x.data, x.shape, x.stride, true, 64, Q)) return rc;
  if (int rc = tmap4(&mDY, dy.data, dy.shape, dy.stride, true, 64, Q)) return rc;
  if (int rc = tmap4(&mDX, dx.data, dx.shape, dx.stride, true, 64, Q)) return rc;
  if (int rc = tmap4(&mB, wsB, bc_shape, bc_stride, false, 64, Q)) return rc;
  if (int rc = tmap4(&mC, wsC, bc_shape, bc_stride, false, 64, Q)) return rc;
  {  // states: (n 128, rows H*64, chunk, batch), box 64 x 128 rows
    const uint64_t dims[4] = {(uint64_t)NS, (uint64_t)(H * HD), (uint64_t)nchunks, (uint64_t)Bsz};
    const uint64_t strides[3] = {(uint64_t)NS * 2, (uint64_t)(H * HD * NS) * 2, (uint64_t)(nchunks * H * HD * NS) * 2};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int rc = make_tmap_16bit(&mS, wsS, 4, dims, strides, box, false)) return rc;
    if (int rc = make_tmap_16bit(&mDS, wsDS, 4, dims, strides, box, false)) return rc;
  }
  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(ssd_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  });
  const int64_t nitems = Bsz * nchunks * (H / 2);
  const int grid = (int)std::min<int64_t>(nitems, sm_count());
  ssd_tc_bwd_kernel<<<grid, kThreads, SMEM_BYTES, s>>>(mX, mDY, mB, mC, mS, mDS, mDX, a);
  OMNI_CUDA_LAUNCH_CHECK("ssd_tc_bwd_kernel");
  return OMNI_OK;
}

}  // namespace omni

// debug: CTA 0 of subsequent tensor-core SSD backward launches records clock64() per (item, phase) into buf[items * 32]
extern "C" void omni_debug_set_bwd_trace(void* buf, int items) {
  omni::g_btrace = static_cast<long long*>(buf);
  omni::g_btrace_items = items;
}
