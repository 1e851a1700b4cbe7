// omni_selftest: exercises, on one CTA, exactly the four tcgen05 GEMM forms the chunked SSD kernel is built from,
// so the descriptor / TMEM / TMA conventions of umma.cuh are pinned by a test before the big kernel relies on them.
//   D1[i][j]      = sum_n C[i][n] B[j][n]            A: smem K-major (TMA)        B: smem K-major (TMA)        N=128
//   D2[i][p]      = sum_j f16(P[i][j]) X[j][p]       A: TMEM fp16 (tcgen05.st)    B: smem MN-major bf16 (TMA)  N=64
//   D3[r][n]      = sum_j bf16(Xs[r][j]) B[j][n]     A: TMEM bf16 (tcgen05.st)    B: smem MN-major (TMA, LBO)  N=128
//   D4[i][r]      = sum_n C[i][n] bf16(S[r][n])      A: smem K-major (TMA)        B: smem K-major (st.shared)  N=128
//   bit6: D3 with A = Xs^T as an MN-major SMEM operand (SS form); bit7: D3 with A = Xs as a K-major SMEM operand x MN-major B.  bit8: every operand fp16 instead of bf16.
//   A and B of one tcgen05.mma must share the 16-bit format: kind::f16 with an fp16 A and a bf16 B raises an illegal-
//   instruction fault on sm_100a (measured), which is why the SSD kernel converts B, C and x to fp16.
#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

struct SelfArgs {
  const float* P;   // [128][128]
  const float* Xs;  // [128][128]
  const float* S;   // [128][128]
  float* D1; float* D2; float* D3; float* D4;
  int which;  // bit0 D1, bit1 D2, bit2 D4, bit3 D3 (TS), bit6 D3 (SS, MN-major smem A), bit8 fp16 mode (else bf16)
};

__global__ void __launch_bounds__(128) selftest_kernel(const __grid_constant__ CUtensorMap mapC,
                                                       const __grid_constant__ CUtensorMap mapB,
                                                       const __grid_constant__ CUtensorMap mapX, SelfArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sC = smem;            // 2 x [128 rows][64 n]  (K halves)  32 KB
  uint8_t* sB = smem + 32768;    // 2 x [128 rows][64 n]              32 KB
  uint8_t* sX = smem + 65536;    // [128 j][64 p]                     16 KB
  uint8_t* sS = smem + 81920;    // 2 x [128 r][64 n] K-major, thread-written   32 KB
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool f16m = (a.which & 256) != 0;  // fp16 mode: Cm, Bm, X hold fp16 and every computed operand is packed as fp16
  const int fmt = f16m ? kFmtF16 : kFmtBF16;

  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, 32768 + 32768 + 16384);
    tma_load_2d(sC, &mapC, &bar_tma, 0, 0);
    tma_load_2d(sC + 16384, &mapC, &bar_tma, 64, 0);
    tma_load_2d(sB, &mapB, &bar_tma, 0, 0);
    tma_load_2d(sB + 16384, &mapB, &bar_tma, 64, 0);
    tma_load_2d(sX, &mapX, &bar_tma, 0, 0);
  }
  // thread = row.  TMEM columns: D1 0..127, D2 128..191, D3 192..319, D4 320..447, A_P 448..511 (fp16 128 K -> 64 cols),
  // A_Xs reuses 448..511 after D2 has been read out.
  const int row = tid;
  {  // P -> fp16 A operand
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c)
        v[c] = f16m ? pack_f16(a.P[row * 128 + 2 * (c0 + c)], a.P[row * 128 + 2 * (c0 + c) + 1])
                    : pack_bf16(a.P[row * 128 + 2 * (c0 + c)], a.P[row * 128 + 2 * (c0 + c) + 1]);
      tmem_st16(tmem_addr(tb, warp * 32, 448 + c0), v);
    }
    tmem_st_wait();
  }
  {  // S -> bf16, K-major swizzled smem tile(s): row r, K = n
    for (int ch = 0; ch < 16; ++ch) {  // 16-byte chunks of 8 n
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = f16m ? pack_f16(a.S[row * 128 + ch * 8 + 2 * q], a.S[row * 128 + ch * 8 + 2 * q + 1])
                    : pack_bf16(a.S[row * 128 + ch * 8 + 2 * q], a.S[row * 128 + ch * 8 + 2 * q + 1]);
      uint8_t* dst = sS + (ch >> 3) * 16384 + sw128(row, ch & 7);
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(&bar_tma, 0);
    const uint32_t idesc1 = make_idesc(128, 128, fmt, fmt, kMajorK, kMajorK);
    const uint32_t idesc2 = make_idesc(128, 64, fmt, fmt, kMajorK, kMajorMN);
    const uint32_t idesc4 = idesc1;
    if (a.which & 1)
    for (int k = 0; k < 8; ++k) {  // D1 = C * B^T
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      mma_ss(tb + 0, make_sdesc(smem_u32(sC) + off, 16, 1024), make_sdesc(smem_u32(sB) + off, 16, 1024), idesc1, k > 0);
    }
    if (a.which & 2)
    for (int k = 0; k < 8; ++k)  // D2 = P * X   (X MN-major: K step = 16 rows = 2048 B)
      mma_ts(tb + 128, tb + 448 + k * 8, make_sdesc(smem_u32(sX) + k * 2048, 16384, 1024), idesc2, k > 0);
    if (a.which & 4)
    for (int k = 0; k < 8; ++k) {  // D4 = C * S^T
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      mma_ss(tb + 320, make_sdesc(smem_u32(sC) + off, 16, 1024), make_sdesc(smem_u32(sS) + off, 16, 1024), idesc4, k > 0);
    }
    mma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  auto dump = [&](uint32_t col0, int ncols, float* dst) {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tb, warp * 32, col0 + c0), r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) dst[row * ncols + c0 + c] = __uint_as_float(r[c]);
    }
  };
  dump(0, 128, a.D1);
  dump(128, 64, a.D2);
  dump(320, 128, a.D4);
  // second round: D3 = bf16(Xs) * B  with B used MN-major ([j][n]: two 64-wide n blocks, LBO = 16384)
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c)
      v[c] = f16m ? pack_f16(a.Xs[row * 128 + 2 * (c0 + c)], a.Xs[row * 128 + 2 * (c0 + c) + 1])
                  : pack_bf16(a.Xs[row * 128 + 2 * (c0 + c)], a.Xs[row * 128 + 2 * (c0 + c) + 1]);
    tmem_st16(tmem_addr(tb, warp * 32, 448 + c0), v);
  }
  tmem_st_wait();
  if (a.which & 128) {  // Xs as a K-major smem A operand: row r, K = j (two 64-wide halves), for the K-major x MN-major SS form
    for (int ch = 0; ch < 16; ++ch) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = f16m ? pack_f16(a.Xs[row * 128 + ch * 8 + 2 * q], a.Xs[row * 128 + ch * 8 + 2 * q + 1])
                    : pack_bf16(a.Xs[row * 128 + ch * 8 + 2 * q], a.Xs[row * 128 + ch * 8 + 2 * q + 1]);
      *reinterpret_cast<uint4*>(sS + (ch >> 3) * 16384 + sw128(row, ch & 7)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async_smem();
  }
  if (a.which & 64) {  // Xs^T as an fp16 MN-major smem A operand: element (j, r) in row j of the (r >> 6) half
    for (int j = 0; j < 128; ++j) {
      uint8_t* dst = sS + (row >> 6) * 16384 + sw128(j, (row & 63) >> 3) + (row & 7) * 2;
      if (f16m) *reinterpret_cast<__half*>(dst) = __float2half_rn(a.Xs[row * 128 + j]);
      else *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(a.Xs[row * 128 + j]);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc3 = make_idesc(128, 128, fmt, fmt, kMajorK, kMajorMN);
    const uint32_t idesc5 = make_idesc(128, 128, fmt, fmt, kMajorMN, kMajorMN);
    if (a.which & 8)
    for (int k = 0; k < 8; ++k)
      mma_ts(tb + 192, tb + 448 + k * 8, make_sdesc(smem_u32(sB) + k * 2048, 16384, 1024), idesc3, k > 0);
    if (a.which & 64)
    for (int k = 0; k < 8; ++k)
      mma_ss(tb + 192, make_sdesc(smem_u32(sS) + k * 2048, 16384, 1024), make_sdesc(smem_u32(sB) + k * 2048, 16384, 1024),
             idesc5, k > 0);
    if (a.which & 128)
    for (int k = 0; k < 8; ++k)
      mma_ss(tb + 192, make_sdesc(smem_u32(sS) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
             make_sdesc(smem_u32(sB) + k * 2048, 16384, 1024), idesc3, k > 0);
    mma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 1);
  tc_fence_after();
  dump(192, 128, a.D3);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace
}  // namespace omni

using namespace omni;

// Cm, Bm: bf16 [128][128]; X: bf16 [128][64]; P, Xs, S: fp32 [128][128]; D1, D3, D4: fp32 [128][128]; D2: fp32 [128][64]
extern "C" int omni_selftest(const void* Cm, const void* Bm, const void* X, const float* P, const float* Xs, const float* S,
                             float* D1, float* D2, float* D3, float* D4, int which, void* stream) {
  CUtensorMap mC, mB, mX;
  const uint64_t d128[2] = {128, 128}, s128[1] = {256}, d64[2] = {64, 128}, s64[1] = {128};
  const uint32_t box[2] = {64, 128};
  if (int rc = make_tmap_16bit(&mC, Cm, 2, d128, s128, box, true)) return rc;
  if (int rc = make_tmap_16bit(&mB, Bm, 2, d128, s128, box, true)) return rc;
  if (int rc = make_tmap_16bit(&mX, X, 2, d64, s64, box, true)) return rc;
  SelfArgs a{P, Xs, S, D1, D2, D3, D4, which};
  const int smem = 81920 + 32768 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(mC, mB, mX, a);
  OMNI_CUDA_LAUNCH_CHECK("selftest_kernel");
  return OMNI_OK;
}
