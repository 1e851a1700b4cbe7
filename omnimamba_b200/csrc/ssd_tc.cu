// Chunked SSD forward on tcgen05 tensor cores with TMA-staged tiles (sm_100a).
//
// Replaces the five Triton kernels behind mamba_ssm's mamba_chunk_scan_combined (_chunk_cumsum, _chunk_state,
// _state_passing, _bmm_chunk, _chunk_scan) with ONE persistent kernel that reads x, dt, B, C
// once, writes y once and keeps the running (P x N) state on chip (fp32, in TMEM) for the whole sequence.
//
// Work item = (batch b, pair of heads h0, h0+1 of one group).  B_t / C_t are shared by the heads of a group, so two
// heads are stacked along the MMA M dimension wherever the contraction allows it.  Per chunk of Q = 128 tokens:
//   CB    [i][j]      = sum_n C[i][n] B[j][n]                    SS  M=128 N=128 K=128   (shared by both heads)
//   Yoff  [i][(h,p)]  = sum_n C[i][n] S16[(h,p)][n]              SS  M=128 N=128 K=128   (state entering the chunk)
//   S     [(h,p)][n] += sum_j X'[j][(h,p)] B[j][n]               SS  M=128 N=128 K=128   (A MN-major; fp32 state lives in TMEM)
//   Ydiag_h[i][p]     = sum_j P_h[i][j] x_h[j][p]                TS  M=128 N=64  K=128   (A = P in TMEM, per head)
//   y_h[i][p] = Ydiag_h + exp(L_i) Yoff + D_h x_h[i][p]
// with L = inclusive cumsum of dt*A inside the chunk, P_h[i][j] = CB[i][j] exp(L_i - L_j) dt_j (j <= i),
// X'[j][(h,p)] = x_h[j][p] dt_j exp(L_last - L_j) (the x tile scaled IN PLACE once Ydiag has consumed it), and
// S <- exp(L_last) S before the accumulation.  Precision: every tensor-core operand is fp16, not bf16: the COMPUTED
// operands (P, S16, X') carry 11 significant bits instead of 8, which keeps y within ~2e-4 of the fp32 reference before
// the output rounding (bf16 operands give 2e-3).  A and B of one tcgen05.mma must share the 16-bit format (an fp16 x
// bf16 kind::f16 MMA faults on sm_100a), so the bf16 inputs are converted too - exactly for 6.1e-5 <= |v| <= 65504
// (saturating above, absolute error <= 3e-8 below): B and C (shared by all heads of a group) once per launch by a
// streaming pre-pass into the caller's workspace, x in place in shared memory by the state-keeper warps.
//
// The only true scan is the scalar decay cumsum (one warp per head, Hillis-Steele shuffle scan: "table warps").
// All accumulators, dt, L and the state are fp32.
//
// Warp roles (512 threads, 1 CTA per SM; TMEM lane quadrant = warp % 4):
//   warps 0,1   per-head dt / cumsum / decay tables (dt gathers prefetched into L2, registers two chunks ahead)
//   warp 2      TMA producer (x and C one stage, B two stages; 128B swizzle)          warp 3   tcgen05.mma issuer
//   warps 4-11  P builders (two per quadrant, split by 32-column blocks), then the x pass (lane = row i of one head:
//               x -> fp16 in place, X' = x dt decay -> second tile, D x -> the head's Ydiag accumulator in TMEM)
//   warps 12-15 epilogue (lane = row i; y staged in two 2 KB 64B-swizzled slots per warp, TMA stores of 32 rows x 32
//               columns) and then the state keepers (S16 copy + decay rescale, lane = (h,p))
// Roles made of several warps let one warp watch the mbarrier and park the others on a named barrier (see `wait1`).
// Tensor-pipe issue order per chunk g:  Yoff(g) | Ydiag_0(g) | Ydiag_1(g) | CB(g+1) | S-update(g).  The epilogue needs the
// first three; CB(g+1) must follow Ydiag(g) (P is built in place over CB) and opens the longest dependent chain
// CB -> P build -> x pass -> Ydiag; the state update has a whole chunk of slack.  Accumulators are released per head.
// Modes 1 / 2 (state sweeps of the backward) run only table -> in-place tile scaling -> S-update -> state keeper, with
// the x-like tile rotating through three stages (XA, XB, the C slot).
// Work schedule: see the comment at `split` in the kernel (half-item hand-off for the left-over items).
// TMEM columns: [0,128) CB then P in place | [128,256) S | [256,384) Yoff | [384,512) Ydiag (64 per head).
#include <algorithm>
#include <mutex>

#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

constexpr int Q = 128;     // tokens per chunk
constexpr int HD = 64;     // headdim
constexpr int NS = 128;    // d_state
// Warps that watch mbarriers for their role sit on schedulers 2 and 3 (with the TMA / MMA threads); the table warps (0, 1),
// whose serial dt -> cumsum -> exp chain feeds every chunk, share their schedulers with no spinning warp.
// In the FORWARD the epilogue leader is warp 12 instead (scheduler 0): there scheduler 3 is the busiest one during the P
// build (quadrant 3 has four of the ten blocks, plus the MMA issuer), which paces the chunk (-0.8 % kernel time).
constexpr int kLeadPX = 2, kLeadE = 3, kLeadEFwd = 0;
// (measured: storing y with 16-byte global stores straight from registers instead of the staged TMA stores makes the
// epilogue AND the concurrent P build slower - 32 scattered rows per store instruction load the LSU: 0.52 vs 0.49 ms)
constexpr bool kDirectY = false;
constexpr int kHandSlots = 128;  // hand-off slots of the half-item schedule (>= grid / 2)
constexpr int kThreads = 512;   // 16 warps: 2 table, TMA, MMA, 8 P build + x pass, 4 epilogue + state

// shared-memory map (bytes, relative to the 1024B-aligned base)
constexpr uint32_t SM_XA = 0;           // [head 2][Q rows x 128 B]  x bf16 (TMA) -> fp16 in place            32 KB
constexpr uint32_t SM_XB = 32768;       // [head 2][Q rows x 128 B]  X' = x dt exp(lam_last - lam_j), fp16     32 KB
constexpr uint32_t SM_B = 65536;        // [stage 2][n-half 2][Q rows x 128 B] fp16                            64 KB
constexpr uint32_t SM_C = 131072;       // [n-half 2][Q rows x 128 B] fp16                                     32 KB
constexpr uint32_t SM_S = 163840;       // [n-half 2][128 (h,p) rows x 128 B] fp16                             32 KB
constexpr uint32_t SM_Y = 196608;       // [epilogue warp 4][32 rows x 128 B] y staging for the TMA stores     16 KB
constexpr uint32_t SM_TAB = 212992;     // [stage 2] Tab
struct Tab {
  float lam[2][Q];      // log2(e) * inclusive cumsum of dt*A  (all exps are ex2)
  float dtv[2][Q];      // transformed dt
  float sj[2][Q];       // exp(lam_last - lam_j) dt_j
  float eL[2][Q];       // exp(lam_i)
  float v[2][3][96];    // v[h][w-1][j] = exp(lam_{32w-1} - lam_j) dt_j, j < 32w
  float vd[2][Q];       // exp(lam_{32(j/32)-1} - lam_j) dt_j  (>= dt_j: reference = start of j's own 32-block)
  float dchunk[2];      // exp(lam_last)
  int safe[2];          // 1: every 32-block decays by < 2^100, so the factorised diagonal block cannot overflow
  float pad[4];
};
static_assert(sizeof(Tab) % 16 == 0, "Tab alignment");
constexpr uint32_t SM_BAR = SM_TAB + 2 * sizeof(Tab);
enum {
  B_FULL_B = 0, B_EMPTY_B = 2, B_TAB_READY = 4, B_TAB_FREE = 6,
  B_FULL_C = 8, B_FULL_X, B_CB_DONE, B_P_READY, B_X16_READY, B_S_READY, B_YOFF_DONE, B_U_DONE, B_YD_DONE,
  B_ACC_FREE = 17,  // [head 2] epilogue has read both accumulators of the head
  B_DX_READY = 19,  // [head 2] x pass part (b): D x is in the head's Ydiag accumulator
  // state sweeps (modes 1 / 2): x-like tiles rotate through three stages (XA, XB and the C slot, all unused otherwise)
  B_SW_FULL = 21,   // [stage 3] TMA tile landed
  B_SW_XR = 24,     // [stage 3] the x pass has scaled the tile in place
  B_COUNT = 27
};
constexpr uint32_t SM_TMEMPTR = SM_BAR + B_COUNT * 8;
constexpr uint32_t SM_TOTAL = SM_TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SM_TOTAL;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

constexpr uint32_t TM_CB = 0, TM_S = 128, TM_YOFF = 256, TM_YD = 384;

struct TcArgs {
  const void* dt; const float* A; const void* D; const void* dt_bias; const void* init; float* fin;
  const __nv_bfloat16* x; void* out;
  int64_t dt_b, dt_l, dt_h, i_b, i_h, i_p, x_b, x_l, x_h, o_b, o_l, o_h;
  int B, L, H, G;
  int dt_dtype, D_dtype, dtb_dtype, init_dtype, out_dtype;
  int dt_softplus;
  float dt_min, dt_max;
  // half-item schedule (forward only, see the kernel): fp32 state hand-off slots [slot][128 (h,p)][128 n] and their flags
  float* hand; int* flags;
  long long* trace; int trace_chunks;  // debug: per-event clock64 of CTA 0 (omni_debug_set_trace)
  // mode 0: the forward.  Modes 1 / 2 run only the state recurrence of the same pipeline and TMA-store the fp16 state
  // ENTERING every chunk (for the backward): 1 = forward states S_c from (x, B, sj); 2 = reverse sweep of the state
  // gradient dS_{c+1} from (dy in the x slot, C in the B slot, exp(lam_i) as the row scale), chunks visited last to first.
  int mode;
  int no_store;  // modes 1 / 2: do not store the per-chunk states (only the final state is wanted: piece schedule)
};

// trace slot layout: trace[g * 32 + event]
// (compiled in only in the TRACE instantiations: the untraced kernels do not carry the ~30 probe sites)
#define TR(ev)                                                                                    \
  do {                                                                                            \
    if constexpr (TRACE) {                                                                        \
      if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && (int)g < a.trace_chunks) {        \
        long long tr_now;  /* volatile asm: keeps its place among the barriers and waits */       \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(tr_now)::"memory");                          \
        a.trace[g * 32 + (ev)] = tr_now;                                                          \
      }                                                                                           \
    }                                                                                             \
  } while (0)

// debug: microseconds the producer of a half-item hand-off holds its flag back (omni_debug_set_handoff_delay)
static __device__ uint32_t g_handoff_delay_us = 0;

__device__ __forceinline__ float ex2f(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- small device helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {  // FMUL2: two fp32 multiplies per issue slot
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {  // FFMA2
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 u2f2(uint32_t lo, uint32_t hi) { return make_float2(__uint_as_float(lo), __uint_as_float(hi)); }
__device__ __forceinline__ float2 h2f2(uint32_t v) {  // packed fp16 pair -> two fp32
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 bf2f2(uint32_t v) {  // packed bf16 pair -> two fp32 (exact)
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// softplus with the upstream cut-over at 20, fast path: max(v,0) + log1p(exp(-|v|)), log1p by series for small arguments
__device__ __forceinline__ float softplus_fast(float v) {
  if (v > 20.f) return v;
  const float u = __expf(-fabsf(v));
  const float l = u < 0.03125f ? u * (1.f - u * (0.5f - u * (0.33333334f - u * (0.25f - u * 0.2f)))) : __logf(1.f + u);
  return fmaxf(v, 0.f) + l;
}

// Walks the chunks of this CTA's work items in processing order (one integer division per item, not per chunk).
// A work UNIT is a contiguous range of chunk steps [c0, cend) of one item (batch, head pair): normally the whole sequence.
struct ChunkIter {
  int u, c, c0, cend, b, h0;
};

// Code size matters here: sixteen warps run five different role loops, and with everything unrolled the kernel was 130 KB
// of SASS - far beyond the 32 KB instruction cache level - so a third of all issue slots were lost to instruction
// fetch (ncu stall_no_inst).  Inner loops are therefore kept rolled (#pragma unroll 1) wherever the body is large.
template <int MODE, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
ssd_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapY,
                  const __grid_constant__ CUtensorMap mapS, TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SM_TMEMPTR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_FULL_B + i], 1);
      mbar_init(&bars[B_EMPTY_B + i], 1);
      mbar_init(&bars[B_TAB_READY + i], 2);
      mbar_init(&bars[B_TAB_FREE + i], 12);  // 8 P / x-pass warps + 4 epilogue / state warps
      mbar_init(&bars[B_ACC_FREE + i], 4);
      mbar_init(&bars[B_DX_READY + i], 4);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars[B_SW_FULL + i], 1);
      mbar_init(&bars[B_SW_XR + i], 8);
    }
    mbar_init(&bars[B_FULL_C], 1);
    mbar_init(&bars[B_FULL_X], 1);
    mbar_init(&bars[B_CB_DONE], 1);
    mbar_init(&bars[B_P_READY], 8);
    mbar_init(&bars[B_X16_READY], 8);
    mbar_init(&bars[B_S_READY], 4);
    mbar_init(&bars[B_YOFF_DONE], 1);
    mbar_init(&bars[B_U_DONE], 1);
    mbar_init(&bars[B_YD_DONE], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    if (lane == 0) {
      tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapB); tma_prefetch_desc(&mapC); tma_prefetch_desc(&mapY);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_ptr;
  // Waits.  try_wait suspends the thread, but a suspended thread is woken by every mbarrier event of the CTA (TMA byte
  // counts included), and each wake-up costs issue slots that the working warps of the same scheduler need (a third of
  // all executed instructions were wait loops).  Roles made of several warps therefore let ONE warp watch the mbarrier
  // and park the others on a hardware named barrier, which costs nothing while waiting.
  const uint32_t hint = g_mbar_hint_ns;
  auto wait1 = [&](uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(smem_u32(&bars[bar])), "r"(parity), "r"(hint) : "memory");
    } while (!ok);
  };

  const int HP = a.H >> 1;                    // head pairs
  const int nitems = a.B * HP;
  const int nchunks = (a.L + Q - 1) / Q;
  const int hpg = a.H / a.G;                  // heads per group
  // Schedule.  CTA `bid` of `grid` takes the items bid, bid + grid, ... of the FR full rounds.  The R = nitems % grid items
  // left over would cost every CTA a whole extra item time for R / grid of the machine (512 items on 148 SMs: 3.46 rounds
  // run as 4).  When 2 R <= grid they are cut in two half sequences instead: CTA r < R runs the FIRST half of left-over item
  // r before anything else and parks the fp32 state in a hand-off slot, CTA R + r runs the SECOND half after its own items -
  // three item times later, so the flag it polls is long set - and the makespan drops from FR + 1 to FR + 1/2 item times.
  const int grid = (int)gridDim.x, bid = (int)blockIdx.x;
  const int FR = nitems / grid, R = nitems % grid;
  const bool split = a.hand != nullptr && R > 0 && 2 * R <= grid && R <= kHandSlots && FR >= 1 && nchunks >= 2;
  const int half = nchunks >> 1;
  const int pre = (split && bid < R) ? 1 : 0;
  const bool tail_half = split && bid >= R && bid < 2 * R, tail_full = !split && bid < R;
  const uint32_t total = (uint32_t)(pre * half + FR * nchunks + (tail_full ? nchunks : tail_half ? nchunks - half : 0));
  // (chunks this CTA processes; g = running chunk counter = barrier phases)
  auto it_set = [&](ChunkIter& it, int u) {
    int item, c0 = 0, c1 = nchunks;
    if (pre && u == 0) { item = FR * grid + bid; c1 = half; }
    else {
      const int k = u - pre;
      if (k < FR) item = k * grid + bid;
      else if (split) { item = FR * grid + bid - R; c0 = half; }
      else item = FR * grid + bid;
    }
    it.u = u; it.c = c0; it.c0 = c0; it.cend = c1; it.b = item / HP; it.h0 = (item - it.b * HP) * 2;
  };
  auto it_next = [&](ChunkIter& it) {
    if (++it.c == it.cend) it_set(it, it.u + 1);
  };
  // (one instantiation per mode: the forward does not carry the sweeps' code and vice versa.)  MODE 3 = the forward of a caller
  // that will run the backward: as mode 0, plus a copy of the fp16 state entering every chunk (the S16 tile the state
  // keepers build for Yoff anyway, each warp TMA-storing its own rows) - what the backward's forward sweep (mode 1) would
  // otherwise recompute.
  constexpr int mode = MODE == 3 ? 0 : MODE;
  constexpr bool save = MODE == 3;
  constexpr int leadE = mode == 0 ? kLeadEFwd : kLeadE;
  auto cphys = [&](int c) { return mode == 2 ? nchunks - 1 - c : c; };  // chunk visited at step c of an item

  if (warp == 2) {
    // ============ TMA producer ========================================================================================
    if (lane == 0 && total > 0) {
      auto load_b = [&](const ChunkIter& it, uint32_t st) {
        mbar_expect_tx(&bars[B_FULL_B + st], 32768);
        tma_load_4d(smem + SM_B + st * 32768, &mapB, &bars[B_FULL_B + st], 0, it.h0 / hpg, cphys(it.c) * Q, it.b);
        tma_load_4d(smem + SM_B + st * 32768 + 16384, &mapB, &bars[B_FULL_B + st], 64, it.h0 / hpg, cphys(it.c) * Q, it.b);
      };
      auto load_c = [&](const ChunkIter& it) {
        mbar_expect_tx(&bars[B_FULL_C], 32768);
        tma_load_4d(smem + SM_C, &mapC, &bars[B_FULL_C], 0, it.h0 / hpg, it.c * Q, it.b);
        tma_load_4d(smem + SM_C + 16384, &mapC, &bars[B_FULL_C], 64, it.h0 / hpg, it.c * Q, it.b);
      };
      auto load_x = [&](const ChunkIter& it) {
        mbar_expect_tx(&bars[B_FULL_X], 32768);
        tma_load_4d(smem + SM_XA, &mapX, &bars[B_FULL_X], 0, it.h0, cphys(it.c) * Q, it.b);
        tma_load_4d(smem + SM_XA + 16384, &mapX, &bars[B_FULL_X], 0, it.h0 + 1, cphys(it.c) * Q, it.b);
      };
      auto prefetch = [&](const ChunkIter& it) {  // pull a later chunk's x / C tiles into L2 (B is loaded two chunks ahead)
        if (mode == 0) {
          tma_prefetch_4d(&mapC, 0, it.h0 / hpg, it.c * Q, it.b);
          tma_prefetch_4d(&mapC, 64, it.h0 / hpg, it.c * Q, it.b);
        }
        tma_prefetch_4d(&mapX, 0, it.h0, cphys(it.c) * Q, it.b);
        tma_prefetch_4d(&mapX, 0, it.h0 + 1, cphys(it.c) * Q, it.b);
      };
      if (mode != 0) {
        // state sweeps: B two chunks ahead (two stages), the x-like tile three chunks ahead (three stages); the state
        // update of chunk g releases B stage g & 1 and tile stage g % 3
        auto load_xs = [&](const ChunkIter& it, uint32_t s3) {
          uint8_t* dst = smem + (s3 == 0 ? SM_XA : s3 == 1 ? SM_XB : SM_C);
          mbar_expect_tx(&bars[B_SW_FULL + s3], 32768);
          tma_load_4d(dst, &mapX, &bars[B_SW_FULL + s3], 0, it.h0, cphys(it.c) * Q, it.b);
          tma_load_4d(dst + 16384, &mapX, &bars[B_SW_FULL + s3], 0, it.h0 + 1, cphys(it.c) * Q, it.b);
        };
        ChunkIter itb, itx;
        it_set(itb, 0);
        itx = itb;
        for (uint32_t k = 0; k < 3 && k < total; ++k) {
          if (k < 2) { load_b(itb, k); it_next(itb); }
          load_xs(itx, k);
          it_next(itx);
        }
        uint32_t s3 = 0;
#pragma unroll 1
        for (uint32_t g = 0; g + 2 < total; ++g) {
          wait1(B_EMPTY_B + (g & 1), (g >> 1) & 1);  // S-update(g) has read B stage g & 1 and tile stage g % 3
          TR(1);
          load_b(itb, g & 1);
          it_next(itb);
          if (g + 3 < total) {
            TR(0);
            load_xs(itx, s3);
            it_next(itx);
          }
          if (++s3 == 3) s3 = 0;
        }
      } else {
      ChunkIter it1, it2;  // chunks g + 1 and g + 2
      it_set(it1, 0);
      if (mode == 0) load_c(it1);
      load_b(it1, 0); load_x(it1);
      it_next(it1);
      it2 = it1;
      if (total > 1) { load_b(it1, 1); prefetch(it1); }
      it_next(it2);
#pragma unroll 1
      for (uint32_t g = 0; g + 1 < total; ++g) {
        if (g + 2 < total) prefetch(it2);
        if (mode == 0) {
          wait1(B_YOFF_DONE, g & 1);            // Yoff(g) has read C(g)
          TR(2);
          load_c(it1);
        }
        if (mode == 0) wait1(B_YD_DONE, g & 1);   // Ydiag(g) has read x(g)
        else wait1(B_X16_READY, g & 1);           // (state sweeps: the x pass has read it)
        TR(0);
        load_x(it1);
        if (g + 2 < total) {
          wait1(B_EMPTY_B + (g & 1), (g >> 1) & 1);  // S-update(g) has read B stage g & 1
          TR(1);
          load_b(it2, g & 1);
        }
        it_next(it1);
        it_next(it2);
      }
      }
    }
  } else if (warp == 3) {
    // ============ MMA issuer ==========================================================================================
    // (warp-uniform control flow: every lane waits and computes the descriptors, one elected lane issues - the compiler then
    // keeps the operands in uniform registers instead of funnelling per-thread registers through R2UR loops)
    if (total > 0) {
      const bool leader = elect_one();
      const uint32_t id_nn = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorK, kMajorK);
      const uint32_t id_u = make_idesc(128, 128, kFmtF16, kFmtF16, kMajorMN, kMajorMN);
      const uint32_t id_yd = make_idesc(128, 64, kFmtF16, kFmtF16, kMajorK, kMajorMN);
      // descriptors of k-step 0; k-major tiles advance by 32 B inside a 64-wide half and by 16 KB between halves,
      // mn-major tiles by 16 rows = 2 KB (the address field counts 16-byte units)
      const uint64_t dC = make_sdesc(smem_u32(smem + SM_C), 16, 1024), dS = make_sdesc(smem_u32(smem + SM_S), 16, 1024);
      const uint64_t dXA = make_sdesc(smem_u32(smem + SM_XA), 16384, 1024), dXB = make_sdesc(smem_u32(smem + SM_XB), 16384, 1024);
      const uint32_t has_D = a.D != nullptr ? 1u : 0u;
#pragma unroll 1
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t st = g & 1, n = g >> 1, ph = g & 1;
        const uint64_t dBk = make_sdesc(smem_u32(smem + SM_B + st * 32768), 16, 1024);
        const uint64_t dBm = make_sdesc(smem_u32(smem + SM_B + st * 32768), 16384, 1024);
        if (mode != 0) {  // state sweeps: only the state update, A operand = tile stage g % 3 (scaled in place)
          const uint32_t s3 = g % 3, k3 = (g / 3) & 1;
          const uint64_t dXs = make_sdesc(smem_u32(smem + (s3 == 0 ? SM_XA : s3 == 1 ? SM_XB : SM_C)), 16384, 1024);
          wait1(B_FULL_B + st, n & 1);
          wait1(B_S_READY, ph);
          wait1(B_SW_XR + s3, k3);
          tc_fence_after();
          TR(5);
#pragma unroll
          for (uint32_t k = 0; k < 8; ++k) if (leader) mma_ss(tb + TM_S, dXs + k * 128, dBm + k * 128, id_u, true);
          if (leader) mma_commit(&bars[B_U_DONE]);
          if (leader) mma_commit(&bars[B_EMPTY_B + st]);
          continue;
        }
        // Issue order per chunk: Yoff(g) | Ydiag(g) | CB(g+1) | S-update(g).  The epilogue only needs the first two, CB(g+1)
        // must follow Ydiag(g) (P lives where CB lands) and feeds the next P build, and the state update has a whole
        // chunk of slack (its consumers are the next chunk's state keepers and x pass).
        if (g == 0) {
          wait1(B_FULL_B + 0, 0);
          wait1(B_FULL_C, 0);
          tc_fence_after();
          TR(3);
#pragma unroll
          for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t off = ((k >> 2) << 10) + ((k & 3) << 1);
            if (leader) mma_ss(tb + TM_CB, dC + off, dBk + off, id_nn, k > 0);
          }
          if (leader) mma_commit(&bars[B_CB_DONE]);
        }
        // Yoff = C S16^T  (state entering the chunk); first, so that C(g) is released early (its reload must land before
        // CB(g+1)) - it only needs the state copy and the accumulators the previous epilogue has read
        // (the state copy is the last of the three to arrive: waiting for it last saves two ~100-cycle already-complete
        // polls between its arrival and the issue, i.e. C(g) is released that much earlier)
        if (g > 0) {
          wait1(B_ACC_FREE + 0, ph ^ 1);
          wait1(B_ACC_FREE + 1, ph ^ 1);
        }
        wait1(B_S_READY, ph);
        tc_fence_after();
        TR(4);
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
          const uint32_t off = ((k >> 2) << 10) + ((k & 3) << 1);
          if (leader) mma_ss(tb + TM_YOFF, dC + off, dS + off, id_nn, k > 0);
        }
        if (leader) mma_commit(&bars[B_YOFF_DONE]);
        // Ydiag_h (+)= P_h x_h   (the accumulator already holds D x)
        wait1(B_P_READY, ph);
#pragma unroll
        for (uint32_t h = 0; h < 2; ++h) {
          wait1(B_DX_READY + h, ph);
          tc_fence_after();
          if (h == 0) TR(6);
#pragma unroll
          for (uint32_t k = 0; k < 8; ++k)
            if (leader) mma_ts(tb + TM_YD + 64 * h, tb + TM_CB + 32 * (k >> 1) + 16 * h + 8 * (k & 1), dXA + h * 1024 + k * 128, id_yd,
                   (has_D | k) != 0);
        }
        if (leader) mma_commit(&bars[B_YD_DONE]);
        // CB(g+1) = C B^T of the next chunk
        if (g + 1 < total) {
          const uint32_t st1 = (g + 1) & 1;
          const uint64_t dBk1 = make_sdesc(smem_u32(smem + SM_B + st1 * 32768), 16, 1024);
          wait1(B_FULL_B + st1, ((g + 1) >> 1) & 1);
          wait1(B_FULL_C, ph ^ 1);
          tc_fence_after();
          TR(3);
#pragma unroll
          for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t off = ((k >> 2) << 10) + ((k & 3) << 1);
            if (leader) mma_ss(tb + TM_CB, dC + off, dBk1 + off, id_nn, k > 0);
          }
          if (leader) mma_commit(&bars[B_CB_DONE]);
        }
        // S += X'^T B   (S was rescaled by exp(lam_last) by the state keepers; X16_READY: part (a) of the x pass)
        wait1(B_X16_READY, ph);
        tc_fence_after();
        TR(5);
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) if (leader) mma_ss(tb + TM_S, dXB + k * 128, dBm + k * 128, id_u, true);
        if (leader) mma_commit(&bars[B_U_DONE]);
        if (leader) mma_commit(&bars[B_EMPTY_B + st]);
      }
    }
  } else if (warp < 2) {
    // ============ table warps (one head each): dt transform, decay cumsum, exp tables ================================
    // One lane = 4 consecutive tokens.  dt rows are 2-byte gathers with a token stride: an HBM miss costs ~2500 cycles, more
    // than a state-sweep chunk takes, so the lines are pulled into L2 three chunks ahead and the register load (one chunk
    // ahead) becomes an L2 hit.  Addresses are byte offsets advanced per chunk; per-head constants are reloaded per item.
    const int hh = warp;
    const int esz = a.dt_dtype == OMNI_F32 ? 4 : 2;
    const int64_t tstride = a.dt_l * esz;  // bytes between consecutive tokens
    const char* dtbase = static_cast<const char*>(a.dt);
    auto chunk_ptr = [&](const ChunkIter& it) -> const char* {  // this lane's first token of the chunk
      return dtbase + (it.b * a.dt_b + (int64_t)(it.h0 + hh) * a.dt_h) * esz + (int64_t)(cphys(it.c) * Q + lane * 4) * tstride;
    };
    uint32_t raw[4];  // raw dt bits of the NEXT chunk: loaded a chunk ahead (an L2 hit thanks to the prefetch), converted when used
    auto load_raw = [&](const ChunkIter& it, bool valid) {
      const char* ptr = chunk_ptr(it);
      const int t0 = cphys(it.c) * Q + lane * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        raw[k] = 0u;
        if (valid && t0 + k < a.L) {
          if (esz == 4) raw[k] = __ldg(reinterpret_cast<const uint32_t*>(ptr + k * tstride));
          else raw[k] = __ldg(reinterpret_cast<const unsigned short*>(ptr + k * tstride));
        }
      }
    };
    auto prefetch_raw = [&](const ChunkIter& it) {
      const char* ptr = chunk_ptr(it);
      const int t0 = cphys(it.c) * Q + lane * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t0 + k < a.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + k * tstride));
    };
    auto raw_to_f = [&](uint32_t bits) -> float {
      if (a.dt_dtype == OMNI_F32) return __uint_as_float(bits);
      if (a.dt_dtype == OMNI_BF16) return __uint_as_float(bits << 16);
      return __half2float(__ushort_as_half((unsigned short)bits));
    };
    ChunkIter it, itn, itp;
    it_set(it, 0);
    itn = it;
    load_raw(itn, total > 0);
    itp = itn;
#pragma unroll 1
    for (uint32_t k = 1; k <= 3 && k < total; ++k) {
      it_next(itp);
      prefetch_raw(itp);
    }
    float Ah2 = 0.f, bias = 0.f;
#pragma unroll 1
    for (uint32_t g = 0; g < total; ++g) {
      const int c = cphys(it.c);
      const uint32_t st = g & 1, n = g >> 1;
      Tab* tab = reinterpret_cast<Tab*>(smem + SM_TAB) + st;
      if (hh == 0) TR(8);
      if (it.c == it.c0) {  // new unit: per-head constants
        Ah2 = a.A[it.h0 + hh] * 1.4426950408889634f;
        bias = a.dt_bias ? ld_any(a.dt_bias, a.dtb_dtype, it.h0 + hh) : 0.f;
      }
      float dtv[4], lam[4];
      float run = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = c * Q + lane * 4 + k;
        float v = 0.f;
        if (t < a.L) {
          v = raw_to_f(raw[k]) + bias;
          if (a.dt_softplus) v = softplus_fast(v);
          v = fminf(fmaxf(v, a.dt_min), a.dt_max);
        }
        dtv[k] = v;
        run += v * Ah2;
        lam[k] = run;
      }
      it_next(itn);
      load_raw(itn, g + 1 < total);  // next chunk's raw dt: in flight while this chunk's tables are built
      if (g + 4 < total) {
        it_next(itp);
        prefetch_raw(itp);
      }
      float incl = run;  // warp inclusive scan of the per-lane totals (Hillis-Steele)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const float excl = incl - run;
#pragma unroll
      for (int k = 0; k < 4; ++k) lam[k] += excl;
      const float lam_last = __shfl_sync(0xffffffffu, lam[3], 31);
      const float dchunk = ex2f(lam_last);
      if (mode != 0) {
        // state sweeps: only the row scale of the x-like tile (sj forward, exp(lam_i) reverse) and the chunk decay
        float4 r4;
        if (mode == 1) {
          r4.x = ex2f(lam_last - lam[0]) * dtv[0]; r4.y = ex2f(lam_last - lam[1]) * dtv[1];
          r4.z = ex2f(lam_last - lam[2]) * dtv[2]; r4.w = ex2f(lam_last - lam[3]) * dtv[3];
        } else {
          r4.x = ex2f(lam[0]); r4.y = ex2f(lam[1]); r4.z = ex2f(lam[2]); r4.w = ex2f(lam[3]);
        }
        if (hh == 0) wait1(B_TAB_FREE + st, (n & 1) ^ 1);
        named_bar_sync(7, 64);
        if (hh == 0) TR(9);
        if (mode == 1) reinterpret_cast<float4*>(tab->sj[hh])[lane] = r4;
        else reinterpret_cast<float4*>(tab->eL[hh])[lane] = r4;
        if (lane == 0) tab->dchunk[hh] = dchunk;
      } else {
      float ref[3];
#pragma unroll
      for (int w = 1; w < 4; ++w) ref[w - 1] = __shfl_sync(0xffffffffu, lam[3], 8 * w - 1);
      // every table entry is computed before waiting for the slot, so that only the stores sit behind the wait
      float4 l4 = make_float4(lam[0], lam[1], lam[2], lam[3]), d4 = make_float4(dtv[0], dtv[1], dtv[2], dtv[3]), s4, e4, vd4;
      float4 v4[3];
      s4.x = ex2f(lam_last - lam[0]) * dtv[0]; s4.y = ex2f(lam_last - lam[1]) * dtv[1];
      s4.z = ex2f(lam_last - lam[2]) * dtv[2]; s4.w = ex2f(lam_last - lam[3]) * dtv[3];
      e4.x = ex2f(lam[0]); e4.y = ex2f(lam[1]); e4.z = ex2f(lam[2]); e4.w = ex2f(lam[3]);
#pragma unroll
      for (int w = 1; w < 4; ++w) {
        const float rf = ref[w - 1];
        v4[w - 1].x = ex2f(rf - lam[0]) * dtv[0]; v4[w - 1].y = ex2f(rf - lam[1]) * dtv[1];
        v4[w - 1].z = ex2f(rf - lam[2]) * dtv[2]; v4[w - 1].w = ex2f(rf - lam[3]) * dtv[3];
      }
      // diagonal blocks: reference = cumsum just before the lane's own 32-token block (0 for the first block)
      const float myref = lane < 8 ? 0.f : (lane < 16 ? ref[0] : (lane < 24 ? ref[1] : ref[2]));
      vd4.x = ex2f(myref - lam[0]) * dtv[0]; vd4.y = ex2f(myref - lam[1]) * dtv[1];
      vd4.z = ex2f(myref - lam[2]) * dtv[2]; vd4.w = ex2f(myref - lam[3]) * dtv[3];
      const bool ok = __all_sync(0xffffffffu, myref - lam[3] < 100.f);
      if (hh == 0) wait1(B_TAB_FREE + st, (n & 1) ^ 1);
      named_bar_sync(7, 64);
      if (hh == 0) TR(9);
      reinterpret_cast<float4*>(tab->lam[hh])[lane] = l4;
      reinterpret_cast<float4*>(tab->dtv[hh])[lane] = d4;
      reinterpret_cast<float4*>(tab->sj[hh])[lane] = s4;
      reinterpret_cast<float4*>(tab->eL[hh])[lane] = e4;
#pragma unroll
      for (int w = 1; w < 4; ++w)
        if (lane < 8 * w) reinterpret_cast<float4*>(tab->v[hh][w - 1])[lane] = v4[w - 1];
      reinterpret_cast<float4*>(tab->vd[hh])[lane] = vd4;
      if (lane == 0) {
        tab->dchunk[hh] = dchunk;
        tab->safe[hh] = ok ? 1 : 0;
      }
      }
      __syncwarp();
      if (hh == 0) TR(10);
      if (lane == 0) mbar_arrive(&bars[B_TAB_READY + st]);
      it_next(it);
    }
  } else if (warp < 12) {
    // ============ P builders (lane = row i, 32-column blocks split between the two warps of a quadrant) ==============
    const int pw = warp - 4, q = pw & 3, sub = pw >> 2, i = q * 32 + lane;
    // blocks of this warp, 4 bits per step (block index | 4 = diagonal block | 8 = zero-fill): (quadrant, sub) ->
    //   q3: {3d, 0} {1, 2}   q2: {2d} {0, 1, z3}   q1: {1d, z2} {0, z3}   q0: {0d} {z1, z2, z3}
    const uint32_t plan = sub == 0 ? (q == 3 ? 0xF07u : q == 2 ? 0xFF6u : q == 1 ? 0xFA5u : 0xFF4u)
                                   : (q == 3 ? 0xF21u : q == 2 ? 0xFB10u : q == 1 ? 0xFB0u : 0xFBA9u);
    const int xw = pw;
    const uint32_t rx = (uint32_t)(i & 7) << 4;
    const uint32_t xrow = sub * 16384 + i * 128;  // byte offset of this lane's x row inside the XA / XB tiles
    ChunkIter it;
    it_set(it, 0);
#pragma unroll 1
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t st = g & 1, n = g >> 1, ph = g & 1;
      const Tab* tab = reinterpret_cast<const Tab*>(smem + SM_TAB) + st;
      if (mode != 0) {
        // state sweeps: scale the x-like tile of stage g % 3 in place (rows x sj, or dy rows x exp(lam_i) in the reverse
        // sweep); the stage belongs to this chunk until its state update has read it, so nothing else to wait for
        const uint32_t s3 = g % 3, k3 = (g / 3) & 1;
        uint8_t* tile = smem + (s3 == 0 ? SM_XA : s3 == 1 ? SM_XB : SM_C);
        if (pw == kLeadPX) {
          wait1(B_TAB_READY + st, n & 1);
          wait1(B_SW_FULL + s3, k3);
        }
        named_bar_sync(2, 256);
        if (pw == 0) TR(17);
        const float sjs = mode == 2 ? tab->eL[sub][i] : tab->sj[sub][i];
        const float2 ss = make_float2(sjs, sjs);
#pragma unroll 4
        for (int k8 = 0; k8 < 8; ++k8) {
          uint4* ptr = reinterpret_cast<uint4*>(tile + xrow + (((uint32_t)k8 << 4) ^ rx));
          const uint4 v = *ptr;
          const float2 s0 = mul2(bf2f2(v.x), ss), s1 = mul2(bf2f2(v.y), ss), s2 = mul2(bf2f2(v.z), ss), s3v = mul2(bf2f2(v.w), ss);
          *ptr = make_uint4(pack_f16_sat(s0.x, s0.y), pack_f16_sat(s1.x, s1.y), pack_f16_sat(s2.x, s2.y), pack_f16_sat(s3v.x, s3v.y));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (pw == 0) TR(18);
        if (lane == 0) {
          mbar_arrive(&bars[B_SW_XR + s3]);
          mbar_arrive(&bars[B_TAB_FREE + st]);
        }
        continue;
      }
      // (D of this warp's head, for the x pass: issued before the waits so that the load never stalls the pass)
      const float Dh = a.D ? ld_any(a.D, a.D_dtype, it.h0 + sub) : 0.f;
      if (pw == kLeadPX) {
        wait1(B_TAB_READY + st, n & 1);
        wait1(B_CB_DONE, ph);
      }
      named_bar_sync(2, 256);
      {
      // ---- P build: P_h = CB o decay o dt (causal), fp16, written over CB in TMEM
      float lam_i[2], u_i[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        lam_i[h] = tab->lam[h][i];
        u_i[h] = ex2f(lam_i[h] - (q > 0 ? tab->lam[h][32 * q - 1] : 0.f));
      }
      const bool safe = tab->safe[0] != 0 && tab->safe[1] != 0;
      tc_fence_after();
      if (pw == 3) TR(12);
#pragma unroll 1
      for (uint32_t pl = plan; (pl & 0xFu) != 0xFu; pl >>= 4) {
        const int jb = pl & 3;
        const uint32_t col = tmem_addr(tb, q * 32, TM_CB + 32 * jb);
        if (pl & 8u) {  // above the diagonal
          uint32_t z[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) z[e] = 0u;
          tmem_st16(col, z);
          tmem_st16(col + 16, z);
          continue;
        }
        const bool diag = (pl & 4u) != 0;
        uint32_t cb[32];
        tmem_ld32(col, cb);
        tmem_ld_wait();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t pk[16];
          const float uh = h == 0 ? u_i[0] : u_i[1];
          const float2 uu = make_float2(uh, uh);
          if (!diag) {  // below the diagonal: no mask, row factor u_i x column table v
            const float4* vv = reinterpret_cast<const float4*>(&tab->v[h][q - 1][32 * jb]);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 f = vv[e];
              const float2 p01 = mul2(mul2(u2f2(cb[4 * e + 0], cb[4 * e + 1]), uu), make_float2(f.x, f.y));
              const float2 p23 = mul2(mul2(u2f2(cb[4 * e + 2], cb[4 * e + 3]), uu), make_float2(f.z, f.w));
              pk[2 * e] = pack_f16_sat(p01.x, p01.y);
              pk[2 * e + 1] = pack_f16_sat(p23.x, p23.y);
            }
          } else if (safe) {  // diagonal block: column <= row
            const float4* vv = reinterpret_cast<const float4*>(&tab->vd[h][32 * jb]);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 f = vv[e];
              const float2 p01 = mul2(mul2(u2f2(cb[4 * e + 0], cb[4 * e + 1]), uu), make_float2(f.x, f.y));
              const float2 p23 = mul2(mul2(u2f2(cb[4 * e + 2], cb[4 * e + 3]), uu), make_float2(f.z, f.w));
              pk[2 * e] = pack_f16_sat(4 * e + 0 <= lane ? p01.x : 0.f, 4 * e + 1 <= lane ? p01.y : 0.f);
              pk[2 * e + 1] = pack_f16_sat(4 * e + 2 <= lane ? p23.x : 0.f, 4 * e + 3 <= lane ? p23.y : 0.f);
            }
          } else {  // diagonal block with extreme decay: direct exp2(lam_i - lam_j) dt_j
            const float4* lj = reinterpret_cast<const float4*>(&tab->lam[h][32 * jb]);
            const float4* dj = reinterpret_cast<const float4*>(&tab->dtv[h][32 * jb]);
            const float li = h == 0 ? lam_i[0] : lam_i[1];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 l4 = lj[e], d4 = dj[e];
              float2 e01 = make_float2(ex2f(fminf(li - l4.x, 0.f)), ex2f(fminf(li - l4.y, 0.f)));
              float2 e23 = make_float2(ex2f(fminf(li - l4.z, 0.f)), ex2f(fminf(li - l4.w, 0.f)));
              e01 = mul2(mul2(e01, make_float2(d4.x, d4.y)), u2f2(cb[4 * e + 0], cb[4 * e + 1]));
              e23 = mul2(mul2(e23, make_float2(d4.z, d4.w)), u2f2(cb[4 * e + 2], cb[4 * e + 3]));
              pk[2 * e] = pack_f16_sat(4 * e + 0 <= lane ? e01.x : 0.f, 4 * e + 1 <= lane ? e01.y : 0.f);
              pk[2 * e + 1] = pack_f16_sat(4 * e + 2 <= lane ? e23.x : 0.f, 4 * e + 3 <= lane ? e23.y : 0.f);
            }
          }
          tmem_st16(col + 16 * h, pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (pw == 3) TR(13);
      if (lane == 0) mbar_arrive(&bars[B_P_READY]);
      }
      // ---- x pass (lane = row i of head `sub`): x row -> fp16 in place (Ydiag operand), X' row -> XB (state operand),
      //      D x -> the head's Ydiag accumulator (released by the epilogue of the previous chunk: ACC_FREE per head)
      if (xw == kLeadPX) {
        if (g > 0) wait1(B_U_DONE, ph ^ 1);     // S-update(g-1) has read XB
        wait1(B_FULL_X, ph);
      }
      if (q == kLeadPX && mode == 0 && g > 0) wait1(B_ACC_FREE + sub, ph ^ 1);
      named_bar_sync(3, 256);
      if (mode == 0 && g > 0) tc_fence_after();
      const float sji = mode == 2 ? tab->eL[sub][i] : tab->sj[sub][i];  // (reverse sweep: dy rows scale by exp(lam_i))
      if (xw == 0) TR(17);
      {
        const float2 ss = make_float2(sji, sji), dd = make_float2(Dh, Dh);
#pragma unroll 1
        for (int k8 = 0; k8 < 8; k8 += 2) {  // two 16-byte chunks (16 columns) per step
          uint32_t dx[16];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t off = xrow + (((uint32_t)(k8 + k) << 4) ^ rx);
            uint4 v = *reinterpret_cast<const uint4*>(smem + SM_XA + off);
            const float2 f0 = bf2f2(v.x), f1 = bf2f2(v.y), f2 = bf2f2(v.z), f3 = bf2f2(v.w);
            uint4 o16, op;
            o16.x = pack_f16_sat(f0.x, f0.y); o16.y = pack_f16_sat(f1.x, f1.y);
            o16.z = pack_f16_sat(f2.x, f2.y); o16.w = pack_f16_sat(f3.x, f3.y);
            const float2 s0 = mul2(f0, ss), s1 = mul2(f1, ss), s2 = mul2(f2, ss), s3 = mul2(f3, ss);
            op.x = pack_f16_sat(s0.x, s0.y); op.y = pack_f16_sat(s1.x, s1.y);
            op.z = pack_f16_sat(s2.x, s2.y); op.w = pack_f16_sat(s3.x, s3.y);
            *reinterpret_cast<uint4*>(smem + SM_XA + off) = o16;
            *reinterpret_cast<uint4*>(smem + SM_XB + off) = op;
            const float2 d0 = mul2(f0, dd), d1 = mul2(f1, dd), d2 = mul2(f2, dd), d3 = mul2(f3, dd);
            dx[8 * k + 0] = __float_as_uint(d0.x); dx[8 * k + 1] = __float_as_uint(d0.y);
            dx[8 * k + 2] = __float_as_uint(d1.x); dx[8 * k + 3] = __float_as_uint(d1.y);
            dx[8 * k + 4] = __float_as_uint(d2.x); dx[8 * k + 5] = __float_as_uint(d2.y);
            dx[8 * k + 6] = __float_as_uint(d3.x); dx[8 * k + 7] = __float_as_uint(d3.y);
          }
          if (a.D != nullptr && mode == 0) tmem_st16(tmem_addr(tb, q * 32, TM_YD + 64 * sub + 8 * k8), dx);
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (xw == 0) TR(18);
      if (lane == 0) {
        mbar_arrive(&bars[B_X16_READY]);
        if (mode == 0) mbar_arrive(&bars[B_DX_READY + sub]);
      }
      if (lane == 0) mbar_arrive(&bars[B_TAB_FREE + st]);
      it_next(it);
    }
  } else {
    // ============ epilogue (lane = row i) + state keepers (TMEM lane r = (head, p)) ==================================
    const int w = warp - 12, r = w * 32 + lane, hh = r >> 6, p = r & 63;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    uint8_t* ybuf = smem + SM_Y + w * 4096;   // two 2 KB slots: [32 rows x 64 B] = 32 columns of one head, 64B swizzle
    ChunkIter sn, ep;  // chunk gg (state step) and chunk gg - 1 (epilogue)
    it_set(sn, 0);
    ep = sn;
    bool keep_pending = false;   // MODE 3: a chunk-state store of this warp may still be reading its rows of the S16 tile
    int epi_groups = 0;          // y-store groups this warp committed in the last epilogue (4 or 0)
    int pend = 0, pend_b = 0, pend_h0 = 0;  // unit that ended with the previous state step: 1 = whole item / second half, 2 = first half
    // state at the end of a unit (all of TM_S, complete once the unit's last S-update has been waited for): the final state
    // of the item, or - first half of a split item - the hand-off slot of this CTA followed by its flag
    auto store_unit_state = [&]() {
      float* dst = pend == 2 ? a.hand + ((int64_t)bid * 128 + r) * NS
                             : (a.fin ? a.fin + ((int64_t)(pend_b * a.H + pend_h0 + hh) * HD + p) * NS : nullptr);
      if (dst != nullptr) {
#pragma unroll 1
        for (int k4 = 0; k4 < 4; ++k4) {
          uint32_t v[32];
          tmem_ld32(tmem_addr(tb, w * 32, TM_S + 32 * k4), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + 32 * k4 + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                        __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
      if (pend == 2) {
        __threadfence();
        named_bar_sync(1, 128);
        if (r == 0) {
          // (debug knob: hold the flag back to exercise the consumer's wait, tests/test_gpu_tc.py)
          for (uint32_t left = g_handoff_delay_us; left > 0; --left) __nanosleep(1000);
          atomicExch(a.flags + bid, 1);
        }
      }
      pend = 0;
    };
    // iteration gg = 0 .. total: [epilogue of chunk gg - 1] then [state step of chunk gg].  The epilogue comes first: it is
    // on the chunk-to-chunk critical path (it frees the accumulators), the state step has slack until Yoff(gg).
#pragma unroll 1
    for (uint32_t gg = 0; gg < total + 1; ++gg) {
      if (gg > 0 && mode == 0) {
        // ---- epilogue(g = gg - 1): y = Ydiag (+ D x) + exp(lam_i) Yoff, row i = r
        const uint32_t g = gg - 1;
        const uint32_t st = g & 1, ph = g & 1;
        const Tab* tab = reinterpret_cast<const Tab*>(smem + SM_TAB) + st;
        const float eL0 = tab->eL[0][r], eL1 = tab->eL[1][r];
        if (w == leadE) wait1(B_YD_DONE, ph);   // Yoff(g) was issued before Ydiag(g): complete as well
        named_bar_sync(6, 128);
        tc_fence_after();
        if (w == 0) TR(19);
        const int eb = ep.b, eh0 = ep.h0, ec = ep.c;
        const int t = ec * Q + r;
        epi_groups = (!kDirectY && a.out_dtype == OMNI_BF16 && ec * Q + w * 32 < a.L) ? 4 : 0;
#pragma unroll 1
        for (int s4 = 0; s4 < 4; ++s4) {  // s4 = head * 2 + 32-column half
          const int hx = s4 >> 1, half = s4 & 1;
          uint8_t* slot = ybuf + half * 2048;
          uint32_t v0[32], v1[32];
          tmem_ld32(tmem_addr(tb, w * 32, TM_YOFF + 32 * s4), v0);
          tmem_ld32(tmem_addr(tb, w * 32, TM_YD + 32 * s4), v1);
          if (!kDirectY && a.out_dtype == OMNI_BF16) {  // the slot is free once the store issued two steps ago has read it
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          tmem_ld_wait();
          const float eh = hx == 0 ? eL0 : eL1;
          // (scalar FFMA or the packed FFMA2 form, -DOMNI_EPI_FFMA2: measured equal, 377.7 vs 376.2 us, although the packed form
          // executes ~900 more instructions per chunk on operand pairing - the kernel is not issue-bound)
          float2 y[16];
#ifdef OMNI_EPI_FFMA2
          const float2 ee = make_float2(eh, eh);
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] = fma2(u2f2(v0[2 * e], v0[2 * e + 1]), ee, u2f2(v1[2 * e], v1[2 * e + 1]));
#else
#pragma unroll
          for (int e = 0; e < 16; ++e)
            y[e] = make_float2(fmaf(__uint_as_float(v0[2 * e]), eh, __uint_as_float(v1[2 * e])),
                               fmaf(__uint_as_float(v0[2 * e + 1]), eh, __uint_as_float(v1[2 * e + 1])));
#endif
          if (half == 1) {  // both accumulators of this head have been read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&bars[B_ACC_FREE + hx]);
              if (hx == 1) mbar_arrive(&bars[B_TAB_FREE + st]);
            }
          }
          if (kDirectY && a.out_dtype == OMNI_BF16) {
            if (t < a.L) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.out) + eb * a.o_b + (int64_t)t * a.o_l +
                                                    (int64_t)(eh0 + hx) * a.o_h + 32 * half);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                dst[k] = make_uint4(pack_bf16(y[4 * k].x, y[4 * k].y), pack_bf16(y[4 * k + 1].x, y[4 * k + 1].y),
                                    pack_bf16(y[4 * k + 2].x, y[4 * k + 2].y), pack_bf16(y[4 * k + 3].x, y[4 * k + 3].y));
            }
          } else if (a.out_dtype == OMNI_BF16) {
            uint8_t* yrow = slot + lane * 64;
            const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(yrow + (((uint32_t)k ^ sw) << 4)) =
                  make_uint4(pack_bf16(y[4 * k].x, y[4 * k].y), pack_bf16(y[4 * k + 1].x, y[4 * k + 1].y),
                             pack_bf16(y[4 * k + 2].x, y[4 * k + 2].y), pack_bf16(y[4 * k + 3].x, y[4 * k + 3].y));
            fence_proxy_async_smem();
            __syncwarp();
            if (ec * Q + w * 32 < a.L) {  // (warp-uniform; rows beyond L are clipped by the tensor map)
              if (elect_one()) {  // elected lane + uniform operands: no R2UR funnel around the store (-1.7 % kernel time)
                tma_store_4d(&mapY, slot, 32 * half, eh0 + hx, ec * Q + w * 32, eb);
                tma_store_commit();
              }
            }
          } else if (t < a.L) {  // fp32 output (parity tests): straight to HBM
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(a.out) + eb * a.o_b + (int64_t)t * a.o_l +
                                                    (int64_t)(eh0 + hx) * a.o_h + 32 * half);
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k] = make_float4(y[2 * k].x, y[2 * k].y, y[2 * k + 1].x, y[2 * k + 1].y);
          }
          if (w == 0 && s4 < 3) TR(23 + s4);
        }
        if (w == 0) TR(22);
      }
      if (gg > 0) it_next(ep);
      if (gg < total) {
        // ---- S16 = fp16(S) for Yoff(gg);  S <- exp(lam_last(gg)) S  (initial state on the first chunk of an item)
        const uint32_t g = gg;  // (for the trace macro)
        const uint32_t st = gg & 1, n = gg >> 1;
        const Tab* tab = reinterpret_cast<const Tab*>(smem + SM_TAB) + st;
        if (w == leadE) {
          wait1(B_TAB_READY + st, n & 1);
          // (forward: Yoff(gg-1) has read the S16 tile - it was issued before Ydiag(gg-1), whose completion the epilogue of
          // chunk gg-1 has just waited for)
          if (gg > 0) wait1(B_U_DONE, (gg - 1) & 1);     // S-update(gg-1) is complete
        }
        named_bar_sync(6, 128);
        const float dch = tab->dchunk[hh];
        if (mode != 0) {  // state sweeps: the previous TMA store must have read the S16 tile
          if (w == 0 && lane == 0) tma_store_wait_read<0>();
          named_bar_sync(1, 128);
        }
        if (save && keep_pending) {
          // this warp's store of the previous step must have read its rows of the tile.  The groups committed since are the y
          // stores of the epilogue just finished: exactly four with bf16 output when the warp's rows are inside the sequence
          // (those may stay pending), none otherwise
          if (lane == 0) {
            if (epi_groups == 4) tma_store_wait_read<4>(); else tma_store_wait_read<0>();
          }
          __syncwarp();
        }
        if (gg > 0) tc_fence_after();
        if (w == 0) TR(15);
        if (sn.c == sn.c0) {
          if (pend) store_unit_state();  // the unit that just ended (its last S-update was waited for above)
          if (sn.c0 > 0) {
            // second half of a split item: the state after the first half comes from hand-off slot bid - R
            const int slot = bid - R;
            if (r == 0) {
              // The producer (CTA bid - R) ran its half item FIRST and is co-resident (the host only enables `split` when
              // the whole grid fits on the device at once), so the flag is normally long set.  The wait is bounded so
              // that a scheduling surprise cannot hang the GPU - but it never falls through: on a timeout (~4 s) the
              // kernel traps, the launch fails with a CUDA error and no stale state is ever consumed.
              int ok = 0;
              for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(ok) : "l"(a.flags + slot) : "memory");
                if (!ok) __nanosleep(1000);
              }
              if (!ok) __trap();
            }
            named_bar_sync(1, 128);
            __threadfence();
            const float* src = a.hand + ((int64_t)slot * 128 + r) * NS;
#pragma unroll 1
            for (int k16 = 0; k16 < 8; ++k16) {
              uint32_t v[16];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 f = __ldcg(reinterpret_cast<const float4*>(src + 16 * k16 + e));
                v[e] = __float_as_uint(f.x); v[e + 1] = __float_as_uint(f.y); v[e + 2] = __float_as_uint(f.z); v[e + 3] = __float_as_uint(f.w);
              }
              tmem_st16(tmem_addr(tb, w * 32, TM_S + 16 * k16), v);
            }
          } else {
          // state entering the item: initial_states or zero, through TMEM so that the hot loop below has one source
          const int64_t ibase = sn.b * a.i_b + (int64_t)(sn.h0 + hh) * a.i_h + (int64_t)p * a.i_p;
#pragma unroll 1
          for (int k16 = 0; k16 < 8; ++k16) {
            uint32_t v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0u;
            if (a.init != nullptr) {  // (cold: once per item; kept rolled so the dtype dispatch is not replicated 16 times)
#pragma unroll 1
              for (int e = 0; e < 16; ++e) {
                const uint32_t bits = __float_as_uint(ld_any(a.init, a.init_dtype, ibase + 16 * k16 + e));
#pragma unroll
                for (int e2 = 0; e2 < 16; ++e2) v[e2] = e2 == e ? bits : v[e2];
              }
            }
            tmem_st16(tmem_addr(tb, w * 32, TM_S + 16 * k16), v);
          }
          }
          tmem_st_wait();
        }
        if (sn.c + 1 == sn.cend) {  // this unit ends with this step
          pend = sn.cend < nchunks ? 2 : 1;
          pend_b = sn.b; pend_h0 = sn.h0;
        }
        const float2 dd = make_float2(dch, dch);
#pragma unroll 1
        for (int k4 = 0; k4 < 4; ++k4) {
          uint32_t v[32];
          tmem_ld32(tmem_addr(tb, w * 32, TM_S + 32 * k4), v);
          tmem_ld_wait();
          uint8_t* srow = smem + SM_S + (k4 >> 1) * 16384 + r * 128;
          const uint32_t c0 = (uint32_t)(k4 & 1) << 6;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16-byte chunks of 8 n
            uint4 o;
            o.x = pack_f16_sat(__uint_as_float(v[8 * k + 0]), __uint_as_float(v[8 * k + 1]));
            o.y = pack_f16_sat(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3]));
            o.z = pack_f16_sat(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5]));
            o.w = pack_f16_sat(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7]));
            *reinterpret_cast<uint4*>(srow + ((c0 + ((uint32_t)k << 4)) ^ rx)) = o;
          }
          uint32_t s0[16], s1[16];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float2 t0 = mul2(u2f2(v[e], v[e + 1]), dd), t1 = mul2(u2f2(v[16 + e], v[17 + e]), dd);
            s0[e] = __float_as_uint(t0.x); s0[e + 1] = __float_as_uint(t0.y);
            s1[e] = __float_as_uint(t1.x); s1[e + 1] = __float_as_uint(t1.y);
          }
          tmem_st16(tmem_addr(tb, w * 32, TM_S + 32 * k4), s0);
          tmem_st16(tmem_addr(tb, w * 32, TM_S + 32 * k4 + 16), s1);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (w == 0) TR(16);
        if (lane == 0) mbar_arrive(&bars[B_S_READY]);
        if (save) {
          // Forward that keeps its chunk states: every warp stores ITS OWN 32 rows of the tile (each thread wrote row r, both
          // n-halves) with two 4 KB TMA stores - no barrier between the four warps, which otherwise run their state step and
          // epilogue independently.  Measured at (16, 4096), plain forward 0.385 ms: this form 0.438 ms, one 32 KB store by one
          // thread behind two 128-thread barriers 0.430 ms, a coalesced LSU copy by the four warps 0.450 ms - the ~12 % are
          // the price of the extra 0.54 GB written beside the stream (+49 % DRAM traffic), not of how the tile leaves.
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&mapS, smem + SM_S + w * 4096, 0, sn.h0 * HD + w * 32, sn.c, sn.b);
            tma_store_4d(&mapS, smem + SM_S + 16384 + w * 4096, 64, sn.h0 * HD + w * 32, sn.c, sn.b);
            tma_store_commit();
          }
          keep_pending = true;
        }
        if (mode != 0 && !a.no_store) {  // state sweeps: the state ENTERING this chunk -> workspace[b][chunk][(h,p)][n] (fp16)
          named_bar_sync(1, 128);
          if (w == 0 && lane == 0) {
            tma_store_4d(&mapS, smem + SM_S, 0, sn.h0 * HD, cphys(sn.c), sn.b);
            tma_store_4d(&mapS, smem + SM_S + 16384, 64, sn.h0 * HD, cphys(sn.c), sn.b);
            tma_store_commit();
          }
        }
        if (mode != 0 && lane == 0) mbar_arrive(&bars[B_TAB_FREE + st]);
        it_next(sn);
      }
    }
    if (total > 0 && pend) {  // state of the last unit
      wait1(B_U_DONE, (total - 1) & 1);
      tc_fence_after();
      store_unit_state();
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tb, 512);
}

// Streaming pre-pass: B and C (B, L, G, N) bf16, any batch/seq/group strides -> contiguous fp16 copies in the workspace.
struct PrepArgs {
  const __nv_bfloat16* src[2];
  __half* dst[2];
  int64_t s_b[2], s_l[2], s_g[2];
  int L, G;
  int64_t rows;  // B * L * G rows of NS elements
};
__global__ void __launch_bounds__(256) ssd_tc_prep_kernel(PrepArgs a) {
  const int which = blockIdx.y;
  const int64_t nvec = a.rows * (NS / 8);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (NS / 8);
    const int c = (int)(i % (NS / 8));
    const int64_t gidx = row % a.G, t = (row / a.G) % a.L, b = row / ((int64_t)a.G * a.L);
    uint4 v = *reinterpret_cast<const uint4*>(a.src[which] + b * a.s_b[which] + t * a.s_l[which] + gidx * a.s_g[which] + c * 8);
    v.x = pack_f16_sat(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
    v.y = pack_f16_sat(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
    v.z = pack_f16_sat(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u));
    v.w = pack_f16_sat(__uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
    *reinterpret_cast<uint4*>(a.dst[which] + row * NS + c * 8) = v;
  }
}

// Fast pre-pass for the usual layouts (rows a constant pitch apart across batch / token / group: contiguous tensors and
// slices of zxbcdt or of the conv output): four 16-byte vectors in flight per thread, no divisions.  The bf16 source is
// read once (evict-first); the fp16 copies stay in L2 for the scan kernel.
__global__ void __launch_bounds__(256) ssd_tc_prep_fast_kernel(PrepArgs a, int64_t pitch0, int64_t pitch1) {
  const int which = blockIdx.y;
  const int64_t nvec = a.rows * (NS / 8), pitch = which == 0 ? pitch0 : pitch1;
  const __nv_bfloat16* src = which == 0 ? a.src[0] : a.src[1];  // (no dynamic indexing of the parameter struct)
  __half* dst = which == 0 ? a.dst[0] : a.dst[1];
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + k * 256;
    if (i < nvec) v[k] = __ldcs(reinterpret_cast<const uint4*>(src + (i >> 4) * pitch + (i & 15) * 8));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = base + k * 256;
    if (i < nvec) {
      uint4 o;
      o.x = pack_f16_sat(__uint_as_float(v[k].x << 16), __uint_as_float(v[k].x & 0xffff0000u));
      o.y = pack_f16_sat(__uint_as_float(v[k].y << 16), __uint_as_float(v[k].y & 0xffff0000u));
      o.z = pack_f16_sat(__uint_as_float(v[k].z << 16), __uint_as_float(v[k].z & 0xffff0000u));
      o.w = pack_f16_sat(__uint_as_float(v[k].w << 16), __uint_as_float(v[k].w & 0xffff0000u));
      *reinterpret_cast<uint4*>(dst + i * 8) = o;
    }
  }
}
static_assert(NS / 8 == 16, "ssd_tc_prep_fast_kernel: 16 vectors per row");

// debug state (omni_debug_set_trace / omni_debug_set_split): plain process globals, read once per launch on the calling
// thread - set them only while no other thread is launching (the tracing scripts and tests are single-threaded)
bool g_no_split = false;
// ---- piece schedule (few long sequences: B H / 2 work items cannot fill 148 SMs) ---------------------------------------
// Every sequence is cut into k equal pieces that run as independent sequences: (1) a state sweep (mode 1, no stores) gives
// the state each piece would END with from a zero start, (2) ssd_piece_decay_kernel the total decay exp(sum dt A) of each
// piece, (3) ssd_piece_combine_kernel chains them, S_enter(p) = decay(p-1) S_enter(p-1) + S_local(p-1), and (4) the forward
// runs on the pieces with S_enter as their initial states.  Costs one extra sweep (~0.55 of a forward) for k-fold parallelism.
struct PieceArgs {
  const void* dt; const float* A; const void* dt_bias;
  int64_t dt_b, dt_l, dt_h;
  int Bk, Lp, H;          // pieces (B * k), tokens per piece, heads
  int dt_dtype, dtb_dtype, dt_softplus;
  float dt_min, dt_max;
  float* decay;           // [Bk][H]
};
__global__ void __launch_bounds__(256) ssd_piece_decay_kernel(PieceArgs a) {
  const int h = blockIdx.x, bp = blockIdx.y;
  const float bias = a.dt_bias ? ld_any(a.dt_bias, a.dtb_dtype, h) : 0.f;
  float sum = 0.f;
  for (int t = threadIdx.x; t < a.Lp; t += blockDim.x) {
    float v = ld_any(a.dt, a.dt_dtype, (int64_t)bp * a.dt_b + (int64_t)t * a.dt_l + (int64_t)h * a.dt_h) + bias;
    if (a.dt_softplus) v = softplus_fast(v);
    sum += fminf(fmaxf(v, a.dt_min), a.dt_max);
  }
  __shared__ float red[32];
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) a.decay[(int64_t)bp * a.H + h] = __expf(sum * a.A[h]);
}
struct CombineArgs {
  const void* init; int init_dtype; int64_t i_b, i_h, i_p;
  const float* local;     // [B * k][H][64][128] state each piece ends with from a zero start
  const float* decay;     // [B * k][H]
  float* enter;           // [B * k][H][64][128] state each piece starts from
  float* fin;             // optional [B][H][64][128]: the state after the last piece of the chain
  int B, k, H;
  int reverse;            // 1: the chain runs from the last piece to the first (reverse sweep of the backward)
};
__global__ void __launch_bounds__(256) ssd_piece_combine_kernel(CombineArgs a) {
  const int64_t per = (int64_t)a.H * HD * NS, total = (int64_t)a.B * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const int64_t r = i - b * per;
    const int h = (int)(r / (HD * NS)), pn = (int)(r - (int64_t)h * HD * NS), pp = pn / NS, n = pn - pp * NS;
    float S = a.init ? ld_any(a.init, a.init_dtype, b * a.i_b + (int64_t)h * a.i_h + (int64_t)pp * a.i_p + n) : 0.f;
    for (int step = 0; step < a.k; ++step) {
      const int p = a.reverse ? a.k - 1 - step : step;
      const int64_t o = ((int64_t)(b * a.k + p)) * per + r;
      a.enter[o] = S;
      if (step + 1 < a.k || a.fin != nullptr) S = a.decay[(int64_t)(b * a.k + p) * a.H + h] * S + a.local[o];
    }
    if (a.fin != nullptr) a.fin[i] = S;
  }
}
__global__ void __launch_bounds__(256) ssd_piece_final_kernel(const float* piece_fin, float* fin, int B, int k, int64_t per) {
  const int64_t total = (int64_t)B * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per;
    fin[i] = piece_fin[(b * k + (k - 1)) * per + (i - b * per)];
  }
}
bool g_no_pieces = false;  // debug (omni_debug_set_handoff): plain schedule only

long long* g_trace = nullptr;
int g_trace_chunks = 0;
int g_trace_mode = 0;  // which launch mode records (0 forward, 1 / 2 state sweeps)

bool tmap_stride_ok(int64_t elems) { return elems >= 0 && (elems * 2) % 16 == 0; }

}  // namespace

bool ssd_tc_fwd_supported(const omni_ssd_fwd_params_t* p) {
  const omni_tensor_t &x = p->x, &Bm = p->B, &Cm = p->C, &o = p->out;
  if (!present(x) || x.ndim != 4 || x.dtype != OMNI_BF16 || x.shape[3] != HD || x.stride[3] != 1) return false;
  if (!present(Bm) || Bm.ndim != 4 || Bm.dtype != OMNI_BF16 || Bm.shape[3] != NS || Bm.stride[3] != 1) return false;
  if (!present(Cm) || Cm.ndim != 4 || Cm.dtype != OMNI_BF16 || Cm.shape[3] != NS || Cm.stride[3] != 1) return false;
  if (!present(o) || o.ndim != 4 || (o.dtype != OMNI_BF16 && o.dtype != OMNI_F32) || o.stride[3] != 1) return false;
  const int64_t H = x.shape[2], G = Bm.shape[2];
  if (G <= 0 || H % G != 0 || (H / G) % 2 != 0) return false;
  if (present(p->z) || present(p->seq_idx)) return false;
  if (present(p->D) && p->D.ndim != 1) return false;
  if (x.shape[1] < 1 || x.shape[0] < 1) return false;
  if (!aligned16(x.data)) return false;  // TMA: 16-byte aligned base and strides
  for (int d = 0; d < 3; ++d)
    if (x.shape[d] > 1 && !tmap_stride_ok(x.stride[d])) return false;
  for (const omni_tensor_t* t : {&Bm, &Cm})  // the pre-pass reads B and C rows with 16-byte vectors
    if (!aligned16(t->data) || (t->shape[2] > 1 && t->stride[2] % 8) || (t->shape[1] > 1 && t->stride[1] % 8) ||
        (t->shape[0] > 1 && t->stride[0] % 8))
      return false;
  const int64_t need = omni_ssd_fwd_workspace_bytes(x.shape[0], x.shape[1], H, HD, G, NS);
  const omni_tensor_t& ws = p->workspace;
  if (!present(ws) || ws.ndim != 1 || ws.stride[0] != 1 || ws.shape[0] * dtype_size(ws.dtype) < need || !aligned16(ws.data))
    return false;
  // y rows are written with 16-byte vector stores
  if (!aligned16(o.data)) return false;
  for (int d = 0; d < 3; ++d)
    if (o.shape[d] > 1 && (o.stride[d] * dtype_size(o.dtype)) % 16 != 0) return false;
  if (present(p->final_states) && p->final_states.dtype != OMNI_F32) return false;
  return get_encode_tiled() != nullptr;
}

// fp16 copies of B and C (any batch / seq / group strides) in wsB / wsC (contiguous (B, L, G, N))
int ssd_tc_prep(const omni_tensor_t& Bm, const omni_tensor_t& Cm, void* wsB, void* wsC, cudaStream_t s) {
  const int64_t Bsz = Bm.shape[0], L = Bm.shape[1], G = Bm.shape[2];
  const int64_t rows = Bsz * L * G;
  PrepArgs pa{};
  pa.src[0] = static_cast<const __nv_bfloat16*>(Bm.data); pa.src[1] = static_cast<const __nv_bfloat16*>(Cm.data);
  pa.dst[0] = static_cast<__half*>(wsB); pa.dst[1] = static_cast<__half*>(wsC);
  pa.s_b[0] = Bm.stride[0]; pa.s_l[0] = Bm.stride[1]; pa.s_g[0] = Bm.stride[2];
  pa.s_b[1] = Cm.stride[0]; pa.s_l[1] = Cm.stride[1]; pa.s_g[1] = Cm.stride[2];
  pa.L = (int)L; pa.G = (int)G; pa.rows = rows;
  const int64_t nvec = rows * (NS / 8);
  // rows of both tensors a constant pitch apart?  (a size-1 dim may carry any stride)
  auto row_pitch = [&](const omni_tensor_t& t) -> int64_t {
    const int64_t pitch = G > 1 ? t.stride[2] : t.stride[1];
    const bool ok = (G == 1 || L == 1 || t.stride[1] == G * t.stride[2]) && (Bsz == 1 || t.stride[0] == L * G * pitch);
    return ok ? pitch : -1;
  };
  const int64_t pB = row_pitch(Bm), pC = row_pitch(Cm);
  if (pB >= 0 && pC >= 0 && (nvec + 1023) / 1024 < (1ll << 31)) {
    ssd_tc_prep_fast_kernel<<<dim3((unsigned)((nvec + 1023) / 1024), 2), 256, 0, s>>>(pa, pB, pC);
    OMNI_CUDA_LAUNCH_CHECK("ssd_tc_prep_fast_kernel");
    return OMNI_OK;
  }
  const unsigned gx = (unsigned)std::min<int64_t>((nvec + 255) / 256, (int64_t)sm_count() * 8);
  ssd_tc_prep_kernel<<<dim3(gx, 2), 256, 0, s>>>(pa);
  OMNI_CUDA_LAUNCH_CHECK("ssd_tc_prep_kernel");
  return OMNI_OK;
}

namespace {
// Common launcher.  mode 0: forward (x, wsB, wsC -> out).  mode 1 / 2: state sweeps (xlike, ws_bslot -> ws_states).
int tc_launch(int mode, const omni_tensor_t& x, const omni_tensor_t& dt, const omni_tensor_t& A, const omni_tensor_t& D,
              const omni_tensor_t& dt_bias, const omni_tensor_t& init, const omni_tensor_t& fin, const omni_tensor_t& o,
              const void* wsB, const void* wsC, void* ws_states, int64_t G, int dt_softplus, float dt_min, float dt_max,
              cudaStream_t s, float* hand = nullptr, int* flags = nullptr, bool no_store = false) {
  const int64_t Bsz = x.shape[0], L = x.shape[1], H = x.shape[2];
  const int64_t nchunks = (L + Q - 1) / Q;
  OMNI_CHECK(present(dt) && shape_is(dt, 3, Bsz, L, H) && is_float_dtype(dt.dtype), OMNI_BAD_SHAPE, "ssd: dt must be (B, L, H)");
  OMNI_CHECK(present(A) && shape_is(A, 1, H) && A.dtype == OMNI_F32 && (H <= 1 || A.stride[0] == 1), OMNI_BAD_SHAPE,
             "ssd: A must be contiguous fp32 (H)");
  TcArgs a{};
  a.mode = mode;
  a.dt = dt.data; a.dt_dtype = dt.dtype; a.dt_b = dt.stride[0]; a.dt_l = dt.stride[1]; a.dt_h = dt.stride[2];
  a.A = static_cast<const float*>(A.data);
  a.x = static_cast<const __nv_bfloat16*>(x.data); a.x_b = x.stride[0]; a.x_l = x.stride[1]; a.x_h = x.stride[2];
  if (mode == 0) {
    a.out = o.data; a.out_dtype = o.dtype; a.o_b = o.stride[0]; a.o_l = o.stride[1]; a.o_h = o.stride[2];
  } else {
    a.out_dtype = OMNI_BF16;
  }
  if (present(D)) {
    OMNI_CHECK(shape_is(D, 1, H) && is_float_dtype(D.dtype) && (H <= 1 || D.stride[0] == 1), OMNI_BAD_SHAPE,
               "ssd: D must be contiguous (H)");
    a.D = D.data; a.D_dtype = D.dtype;
  }
  if (present(dt_bias)) {
    OMNI_CHECK(shape_is(dt_bias, 1, H) && is_float_dtype(dt_bias.dtype) && (H <= 1 || dt_bias.stride[0] == 1),
               OMNI_BAD_SHAPE, "ssd: dt_bias must be contiguous (H)");
    a.dt_bias = dt_bias.data; a.dtb_dtype = dt_bias.dtype;
  }
  if (present(init)) {
    OMNI_CHECK(shape_is(init, 4, Bsz, H, HD, NS) && is_float_dtype(init.dtype) && init.stride[3] == 1, OMNI_BAD_SHAPE,
               "ssd: initial_states must be (B, H, P, N)");
    a.init = init.data; a.init_dtype = init.dtype; a.i_b = init.stride[0]; a.i_h = init.stride[1]; a.i_p = init.stride[2];
  }
  if (present(fin)) {
    OMNI_CHECK(shape_is(fin, 4, Bsz, H, HD, NS) && fin.dtype == OMNI_F32 && fin.stride[3] == 1 && fin.stride[2] == NS &&
                   fin.stride[1] == HD * NS && fin.stride[0] == H * HD * NS && aligned16(fin.data),
               OMNI_BAD_SHAPE, "ssd: final_states must be contiguous fp32 (B, H, P, N)");
    a.fin = static_cast<float*>(fin.data);
  }
  a.B = (int)Bsz; a.L = (int)L; a.H = (int)H; a.G = (int)G;
  a.dt_softplus = dt_softplus; a.dt_min = dt_min; a.dt_max = dt_max;
  a.trace = mode == g_trace_mode ? g_trace : nullptr; a.trace_chunks = g_trace_chunks;
  a.hand = hand; a.flags = flags;
  a.no_store = no_store ? 1 : 0;

  auto tmap4 = [&](CUtensorMap* m, const void* base, const int64_t* shape, const int64_t* stride, bool bf16, int rows) -> int {
    // dims innermost first: (inner, dim2, L, B); a size-1 dim may carry any stride: give TMA a harmless legal one
    const uint64_t dims[4] = {(uint64_t)shape[3], (uint64_t)shape[2], (uint64_t)shape[1], (uint64_t)shape[0]};
    auto st = [&](int d) { return (uint64_t)(shape[d] > 1 ? stride[d] : shape[3]) * 2; };
    const uint64_t strides[3] = {st(2), st(1), st(0)};
    const uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
    return make_tmap_16bit(m, base, 4, dims, strides, box, bf16);
  };
  const int64_t bc_shape[4] = {Bsz, L, G, NS}, bc_stride[4] = {L * G * NS, G * NS, NS, 1};
  CUtensorMap mX, mB, mC, mY, mS;
  if (int rc = tmap4(&mX, x.data, x.shape, x.stride, true, Q)) return rc;
  if (int rc = tmap4(&mB, wsB, bc_shape, bc_stride, false, Q)) return rc;
  if (int rc = tmap4(&mC, wsC ? wsC : wsB, bc_shape, bc_stride, false, Q)) return rc;
  if (mode == 0 && o.dtype == OMNI_BF16) {  // y leaves through per-warp TMA stores of 32 rows
    // y leaves through per-warp TMA stores of 32 rows x 32 columns (64B swizzle)
    const uint64_t dims[4] = {(uint64_t)o.shape[3], (uint64_t)o.shape[2], (uint64_t)o.shape[1], (uint64_t)o.shape[0]};
    auto st = [&](int d) { return (uint64_t)(o.shape[d] > 1 ? o.stride[d] : o.shape[3]) * 2; };
    const uint64_t strides[3] = {st(2), st(1), st(0)};
    const uint32_t box[4] = {32, 1, 32, 1};
    if (int rc = make_tmap_16bit(&mY, o.data, 4, dims, strides, box, true, 64)) return rc;
  } else {
    mY = mX;  // unused
  }
  const bool save = mode == 0 && ws_states != nullptr;   // forward that also stores the chunk states (kernel MODE 3)
  if ((mode != 0 && !no_store) || save) {  // fp16 states: (n 128, rows H*64, chunk, batch), box 64 x 128 rows (MODE 3: 32 rows)
    const uint64_t dims[4] = {(uint64_t)NS, (uint64_t)(H * HD), (uint64_t)nchunks, (uint64_t)Bsz};
    const uint64_t strides[3] = {(uint64_t)NS * 2, (uint64_t)(H * HD * NS) * 2, (uint64_t)(nchunks * H * HD * NS) * 2};
    const uint32_t box[4] = {64, save ? 32u : 128u, 1, 1};
    if (int rc = make_tmap_16bit(&mS, ws_states, 4, dims, strides, box, false)) return rc;
  } else {
    mS = mX;  // unused
  }

  static std::once_flag once[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(once[dev & 63], [] {
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    cudaFuncSetAttribute(ssd_tc_fwd_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  });
  const int nitems = (int)(Bsz * (H / 2));
  const int grid = nitems < sm_count() ? nitems : sm_count();
  // The half-item schedule makes one CTA wait for another: only legal when every CTA of the grid is resident at once.
  // (1 CTA per SM by construction; a context with fewer SMs than sm_count() - MPS partitions, green contexts - reports
  // fewer co-resident blocks here and runs the plain schedule.)
  static int resident[64];
  static std::once_flag occ_once[64];
  std::call_once(occ_once[dev & 63], [dev] {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ssd_tc_fwd_kernel<0, false>, kThreads, SMEM_BYTES) != cudaSuccess) per_sm = 0;
    resident[dev & 63] = per_sm * sm_count();
  });
  if (resident[dev & 63] < grid || g_no_split) { flags = nullptr; a.hand = nullptr; a.flags = nullptr; }
  if (flags != nullptr) {  // hand-off flags of the half-item schedule (a memset node under graph capture)
    if (cudaMemsetAsync(flags, 0, kHandSlots * sizeof(int), s) != cudaSuccess) { flags = nullptr; a.hand = nullptr; a.flags = nullptr; }
  }
  if (a.trace != nullptr && !save) {  // debug instantiations with the probe sites
    if (mode == 0) ssd_tc_fwd_kernel<0, true><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
    else if (mode == 1) ssd_tc_fwd_kernel<1, true><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
    else ssd_tc_fwd_kernel<2, true><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
  } else if (save) ssd_tc_fwd_kernel<3, false><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
  else if (mode == 0) ssd_tc_fwd_kernel<0, false><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
  else if (mode == 1) ssd_tc_fwd_kernel<1, false><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
  else ssd_tc_fwd_kernel<2, false><<<grid, kThreads, SMEM_BYTES, s>>>(mX, mB, mC, mY, mS, a);
  OMNI_CUDA_LAUNCH_CHECK("ssd_tc_fwd_kernel");
  return OMNI_OK;
}
}  // namespace

namespace {
// pieces per sequence of the piece schedule (1 = plain schedule)
int piece_count(int64_t Bsz, int64_t L, int64_t H) {
  const int64_t items = Bsz * (H / 2), sm = sm_count();
  if (g_no_pieces || items <= 0 || items * 2 > sm) return 1;
  // (more, shorter pieces were measured and do not pay: (1, 65 536) with 16 pieces of 32 chunks - the 512-item schedule of the
  // bench shape - 0.705 ms against 0.698 ms with 4 pieces of 128 chunks: the extra sweep has per-item costs of its own)
  int64_t k = sm / items;
  if (k > 16) k = 16;
  while (k > 1 && (L % k != 0 || L / k < 8 * Q)) --k;
  return (int)k;
}
int64_t piece_ws_bytes(int64_t Bsz, int64_t L, int64_t H) {
  const int k = piece_count(Bsz, L, H);
  if (k <= 1) return 0;
  return 3 * Bsz * k * H * HD * NS * (int64_t)sizeof(float) + Bsz * k * H * (int64_t)sizeof(float) + 1024;
}
// (B, L, ...) -> (B k, L / k, ...): a view, when the batch stride is L rows
bool piece_view(const omni_tensor_t& t, int k, omni_tensor_t& v) {
  v = t;
  if (!present(t)) return true;
  if (t.shape[0] > 1 && t.stride[0] != t.shape[1] * t.stride[1]) return false;
  v.shape[0] = t.shape[0] * k;
  v.shape[1] = t.shape[1] / k;
  v.stride[0] = v.shape[1] * t.stride[1];
  return true;
}
}  // namespace

int64_t ssd_tc_chunk_states_bytes(int64_t batch, int64_t seqlen, int64_t nheads) {
  return batch * ((seqlen + Q - 1) / Q) * nheads * HD * NS * 2;
}
namespace {
bool chunk_states_ok(const omni_ssd_fwd_params_t* p) {
  const omni_tensor_t& c = p->chunk_states;
  if (!present(c) || c.ndim != 1 || c.stride[0] != 1 || c.dtype != OMNI_F16 || !aligned16(c.data)) return false;
  return c.shape[0] * 2 >= ssd_tc_chunk_states_bytes(p->x.shape[0], p->x.shape[1], p->x.shape[2]);
}
}  // namespace

int ssd_tc_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s) {
  const omni_tensor_t &x = p->x, &Bm = p->B, &Cm = p->C, &o = p->out;
  const int64_t Bsz = x.shape[0], L = x.shape[1], H = x.shape[2], G = Bm.shape[2];
  OMNI_CHECK(shape_is(Bm, 4, Bsz, L, G, NS) && shape_is(Cm, 4, Bsz, L, G, NS), OMNI_BAD_SHAPE, "ssd: B/C must be (B, L, G, N)");
  OMNI_CHECK(shape_is(o, 4, Bsz, L, H, HD), OMNI_BAD_SHAPE, "ssd: out must match x");
  // pre-pass: fp16 copies of B and C in the caller's workspace
  __half* wsB = static_cast<__half*>(p->workspace.data);
  __half* wsC = wsB + Bsz * L * G * NS;
  if (int rc = ssd_tc_prep(Bm, Cm, wsB, wsC, s)) return rc;
  char* tail = reinterpret_cast<char*>(wsC + Bsz * L * G * NS);
  tail += (256 - reinterpret_cast<uintptr_t>(tail) % 256) % 256;
  float* hand = reinterpret_cast<float*>(tail);
  int* flags = reinterpret_cast<int*>(tail + (size_t)kHandSlots * 128 * NS * sizeof(float));

  // ---- piece schedule: too few (batch, head pair) items for the SMs -> cut every sequence into k independent pieces ----
  int k = piece_count(Bsz, L, H);
  omni_tensor_t xv, dtv, ov;
  if (k > 1 && !(piece_view(x, k, xv) && piece_view(p->dt, k, dtv) && piece_view(o, k, ov))) k = 1;
  if (k > 1) {
    const int64_t Bk = Bsz * k, per = H * HD * NS;
    char* pw = reinterpret_cast<char*>(flags + kHandSlots);
    pw += (256 - reinterpret_cast<uintptr_t>(pw) % 256) % 256;
    float* local = reinterpret_cast<float*>(pw);
    float* enter = local + Bk * per;
    float* pfin = enter + Bk * per;
    float* decay = pfin + Bk * per;
    omni_tensor_t none{};
    auto state_tensor = [&](float* ptr) {
      omni_tensor_t t{};
      t.data = ptr; t.dtype = OMNI_F32; t.ndim = 4;
      t.shape[0] = Bk; t.shape[1] = H; t.shape[2] = HD; t.shape[3] = NS;
      t.stride[3] = 1; t.stride[2] = NS; t.stride[1] = HD * NS; t.stride[0] = per;
      return t;
    };
    const omni_tensor_t t_local = state_tensor(local), t_enter = state_tensor(enter), t_pfin = state_tensor(pfin);
    // (1) state each piece ends with from a zero start: the state sweep of the backward, without its per-chunk stores
    if (int rc = tc_launch(1, xv, dtv, p->A, none, p->dt_bias, none, t_local, none, wsB, nullptr, nullptr, G, p->dt_softplus,
                           p->dt_min, p->dt_max, s, hand, flags, true))
      return rc;
    // (2) total decay of every piece, (3) the chain over the pieces of a sequence
    PieceArgs pa{};
    pa.dt = dtv.data; pa.A = static_cast<const float*>(p->A.data); pa.dt_bias = p->dt_bias.data;
    pa.dt_b = dtv.stride[0]; pa.dt_l = dtv.stride[1]; pa.dt_h = dtv.stride[2];
    pa.Bk = (int)Bk; pa.Lp = (int)(L / k); pa.H = (int)H;
    pa.dt_dtype = dtv.dtype; pa.dtb_dtype = p->dt_bias.dtype; pa.dt_softplus = p->dt_softplus;
    pa.dt_min = p->dt_min; pa.dt_max = p->dt_max; pa.decay = decay;
    ssd_piece_decay_kernel<<<dim3((unsigned)H, (unsigned)Bk), 256, 0, s>>>(pa);
    OMNI_CUDA_LAUNCH_CHECK("ssd_piece_decay_kernel");
    CombineArgs ca{};
    if (present(p->initial_states)) {
      const omni_tensor_t& in = p->initial_states;
      OMNI_CHECK(shape_is(in, 4, Bsz, H, HD, NS) && is_float_dtype(in.dtype) && in.stride[3] == 1, OMNI_BAD_SHAPE,
                 "ssd: initial_states must be (B, H, P, N)");
      ca.init = in.data; ca.init_dtype = in.dtype; ca.i_b = in.stride[0]; ca.i_h = in.stride[1]; ca.i_p = in.stride[2];
    }
    ca.local = local; ca.decay = decay; ca.enter = enter; ca.B = (int)Bsz; ca.k = k; ca.H = (int)H;
    ssd_piece_combine_kernel<<<sm_count() * 4, 256, 0, s>>>(ca);
    OMNI_CUDA_LAUNCH_CHECK("ssd_piece_combine_kernel");
    // (4) the forward over the pieces, each from the state it enters with
    const bool want_fin = present(p->final_states);
    if (int rc = tc_launch(0, xv, dtv, p->A, p->D, p->dt_bias, t_enter, want_fin ? t_pfin : none, ov, wsB, wsC, nullptr, G,
                           p->dt_softplus, p->dt_min, p->dt_max, s, hand, flags))
      return rc;
    if (want_fin) {
      const omni_tensor_t& fin = p->final_states;
      OMNI_CHECK(shape_is(fin, 4, Bsz, H, HD, NS) && fin.dtype == OMNI_F32 && fin.stride[3] == 1 && fin.stride[2] == NS &&
                     fin.stride[1] == HD * NS && fin.stride[0] == per,
                 OMNI_BAD_SHAPE, "ssd: final_states must be contiguous fp32 (B, H, P, N)");
      ssd_piece_final_kernel<<<sm_count() * 2, 256, 0, s>>>(pfin, static_cast<float*>(fin.data), (int)Bsz, k, per);
      OMNI_CUDA_LAUNCH_CHECK("ssd_piece_final_kernel");
    }
    return OMNI_OK;
  }
  return tc_launch(0, x, p->dt, p->A, p->D, p->dt_bias, p->initial_states, p->final_states, o, wsB, wsC,
                   chunk_states_ok(p) ? p->chunk_states.data : nullptr, G, p->dt_softplus, p->dt_min, p->dt_max, s, hand, flags);
}

// 1 when ssd_tc_fwd called with these params fills p->chunk_states (plain schedule only: the piece schedule runs the
// forward on piece views whose chunk grid need not be the backward's)
bool ssd_tc_fwd_saves_states(const omni_ssd_fwd_params_t* p) {
  if (!ssd_tc_fwd_supported(p) || !chunk_states_ok(p)) return false;
  const omni_tensor_t& x = p->x;
  int k = piece_count(x.shape[0], x.shape[1], x.shape[2]);
  omni_tensor_t v;
  if (k > 1 && !(piece_view(x, k, v) && piece_view(p->dt, k, v) && piece_view(p->out, k, v))) k = 1;
  return k == 1;
}

// State sweeps for the backward (ssd_tc_bwd.cu): mode 1 = forward states from (x, fp16 B), mode 2 = reverse sweep of the
// state gradients from (dy, fp16 C).  `init` seeds the recurrence, `fin` (fp32, optional) receives its last state.
int ssd_tc_state_sweep(int mode, const omni_tensor_t& xlike, const omni_tensor_t& dt, const omni_tensor_t& A,
                       const omni_tensor_t& dt_bias, const omni_tensor_t& init, const omni_tensor_t& fin, const void* ws_bslot,
                       void* ws_states, int64_t G, int dt_softplus, float dt_min, float dt_max, cudaStream_t s,
                       void* hand_slots, void* piece_ws) {
  omni_tensor_t none{};
  float* hand = static_cast<float*>(hand_slots);
  int* flags = hand_slots ? reinterpret_cast<int*>(static_cast<char*>(hand_slots) + (size_t)kHandSlots * 128 * NS * sizeof(float))
                          : nullptr;
  // piece schedule (few long sequences), as in the forward: a store-free sweep gives the state every piece ends with from
  // a zero start, the chain over the pieces (last to first in the reverse sweep) gives the state each one enters with,
  // and the real sweep then runs on k times as many independent sequences
  const int64_t Bsz = xlike.shape[0], L = xlike.shape[1], H = xlike.shape[2];
  int k = piece_ws != nullptr ? piece_count(Bsz, L, H) : 1;
  while (k > 1 && (L / k) % Q != 0) --k;     // (the per-chunk state tensor is addressed per piece: pieces are whole chunks)
  if (k > 1 && L % k != 0) k = 1;
  omni_tensor_t xv, dtv;
  if (k > 1 && !(piece_view(xlike, k, xv) && piece_view(dt, k, dtv))) k = 1;
  if (k > 1) {
    const int64_t Bk = Bsz * k, per = H * HD * NS;
    float* local = static_cast<float*>(piece_ws);
    float* enter = local + Bk * per;
    float* decay = enter + Bk * per;
    auto state_tensor = [&](float* ptr) {
      omni_tensor_t t{};
      t.data = ptr; t.dtype = OMNI_F32; t.ndim = 4;
      t.shape[0] = Bk; t.shape[1] = H; t.shape[2] = HD; t.shape[3] = NS;
      t.stride[3] = 1; t.stride[2] = NS; t.stride[1] = HD * NS; t.stride[0] = per;
      return t;
    };
    if (int rc = tc_launch(mode, xv, dtv, A, none, dt_bias, none, state_tensor(local), none, ws_bslot, nullptr, nullptr, G,
                           dt_softplus, dt_min, dt_max, s, hand, flags, true))
      return rc;
    PieceArgs pa{};
    pa.dt = dtv.data; pa.A = static_cast<const float*>(A.data); pa.dt_bias = dt_bias.data;
    pa.dt_b = dtv.stride[0]; pa.dt_l = dtv.stride[1]; pa.dt_h = dtv.stride[2];
    pa.Bk = (int)Bk; pa.Lp = (int)(L / k); pa.H = (int)H;
    pa.dt_dtype = dtv.dtype; pa.dtb_dtype = dt_bias.dtype; pa.dt_softplus = dt_softplus;
    pa.dt_min = dt_min; pa.dt_max = dt_max; pa.decay = decay;
    ssd_piece_decay_kernel<<<dim3((unsigned)H, (unsigned)Bk), 256, 0, s>>>(pa);
    OMNI_CUDA_LAUNCH_CHECK("ssd_piece_decay_kernel");
    CombineArgs ca{};
    if (present(init)) {
      OMNI_CHECK(shape_is(init, 4, Bsz, H, HD, NS) && is_float_dtype(init.dtype) && init.stride[3] == 1, OMNI_BAD_SHAPE,
                 "ssd: initial / final state gradients must be (B, H, P, N)");
      ca.init = init.data; ca.init_dtype = init.dtype; ca.i_b = init.stride[0]; ca.i_h = init.stride[1]; ca.i_p = init.stride[2];
    }
    if (present(fin)) {
      OMNI_CHECK(shape_is(fin, 4, Bsz, H, HD, NS) && fin.dtype == OMNI_F32 && fin.stride[3] == 1 && fin.stride[2] == NS &&
                     fin.stride[1] == HD * NS && fin.stride[0] == per,
                 OMNI_BAD_SHAPE, "ssd: the sweep's last state must be contiguous fp32 (B, H, P, N)");
      ca.fin = static_cast<float*>(fin.data);
    }
    ca.local = local; ca.decay = decay; ca.enter = enter; ca.B = (int)Bsz; ca.k = k; ca.H = (int)H; ca.reverse = mode == 2;
    ssd_piece_combine_kernel<<<sm_count() * 4, 256, 0, s>>>(ca);
    OMNI_CUDA_LAUNCH_CHECK("ssd_piece_combine_kernel");
    return tc_launch(mode, xv, dtv, A, none, dt_bias, state_tensor(enter), none, none, ws_bslot, nullptr, ws_states, G, dt_softplus,
                     dt_min, dt_max, s, hand, flags);
  }
  return tc_launch(mode, xlike, dt, A, none, dt_bias, init, fin, none, ws_bslot, nullptr, ws_states, G, dt_softplus, dt_min,
                   dt_max, s, hand, flags);
}

// scratch of the sweeps' piece schedule (both sweeps share it): two per-piece state tensors + the piece decays
int64_t ssd_tc_sweep_piece_bytes(int64_t Bsz, int64_t L, int64_t H) {
  const int k = piece_count(Bsz, L, H);
  if (k <= 1) return 0;
  return 2 * Bsz * k * H * HD * NS * (int64_t)sizeof(float) + Bsz * k * H * (int64_t)sizeof(float) + 1024;
}

int64_t ssd_tc_hand_bytes() { return (int64_t)kHandSlots * (128 * NS * (int64_t)sizeof(float) + sizeof(int)) + 256; }

}  // namespace omni

// debug: suspend-time hint (ns) of the mbarrier waits inside the tensor-core SSD kernel
extern "C" void omni_debug_set_mbar_hint(unsigned ns) { cudaMemcpyToSymbol(omni::umma::g_mbar_hint_ns, &ns, sizeof(ns)); }

// debug: hold back the hand-off flag of the half-item schedule by `us` microseconds (exercises the consumer's wait);
// `schedules` switches the work schedules on (default 3): bit 0 half-item hand-off, bit 1 piece schedule
extern "C" void omni_debug_set_handoff(unsigned delay_us, int schedules) {
  cudaMemcpyToSymbol(omni::g_handoff_delay_us, &delay_us, sizeof(delay_us));
  omni::g_no_split = (schedules & 1) == 0;    // bit 0: half-item hand-off schedule
  omni::g_no_pieces = (schedules & 2) == 0;   // bit 1: piece schedule (few long sequences)
}

// debug: CTA 0 of the next ssd_tc launches records clock64() per (chunk, event) into buf[chunks * 32] (device int64)
extern "C" void omni_debug_set_trace(void* buf, int chunks) {
  omni::g_trace = static_cast<long long*>(buf);
  omni::g_trace_mode = chunks / 1000;  // (debug convention: chunks = 1000 * mode + number of chunks)
  omni::g_trace_chunks = chunks % 1000;
}

// bytes of caller-provided workspace the tensor-core forward needs (fp16 copies of B and C)
extern "C" int64_t omni_ssd_fwd_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim, int64_t ngroups,
                                                int64_t dstate) {
  (void)nheads; (void)headdim;
  // fp16 copies of B and C + the hand-off slots (fp32 state of a head pair) and flags of the half-item schedule + the
  // per-piece states of the piece schedule (few long sequences)
  return 2 * batch * seqlen * ngroups * dstate * 2 + 256 + (int64_t)omni::kHandSlots * (128 * 128 * 4 + 4) + 256 +
         omni::piece_ws_bytes(batch, seqlen, nheads);
}
