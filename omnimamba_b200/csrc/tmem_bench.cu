// omni_debug_tmem_bench: measures tcgen05.ld / tcgen05.st throughput per SM (cycles for a fixed number of bytes) with
// 4, 8 or 12 warps issuing concurrently.  Design input for the SSD kernels (how much TMEM traffic a chunk can afford).
#include "umma.cuh"

namespace omni {
namespace {
using namespace umma;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t* r) {
  uint32_t a[32], b[32];
  tmem_ld32(taddr, a);
  tmem_ld32(taddr + 32, b);
  r[0] ^= a[0] ^ b[31];
}

// mode 0: ld 32x32b.x32 ; mode 1: st 32x32b.x16 ; mode 2: ld .x32 issued in pairs before one wait
__global__ void __launch_bounds__(512) tmem_bench_kernel(long long* out, int mode, int nwarps, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = tmem_addr(tb, (warp & 3) * 32, 0);
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
        uint32_t r[32];
        tmem_ld32(base + ((it * 32) & 511), r);
        tmem_ld_wait();
        acc ^= r[0] ^ r[31];
      }
    } else if (mode == 1) {
      uint32_t s[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = v[i];
      for (int it = 0; it < iters; ++it) {
        tmem_st16(base + ((it * 16) & 511), s);
        tmem_st16(base + ((it * 16 + 16) & 511), s);
      }
      tmem_st_wait();
    } else {
      for (int it = 0; it < iters; it += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld32(base + ((it * 32) & 511), r0);
        tmem_ld32(base + ((it * 32 + 32) & 511), r1);
        tmem_ld_wait();
        acc ^= r0[0] ^ r1[31];
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (lane == 0 && warp < nwarps) out[warp] = t1 - t0;
  if (acc == 0x12345678u) out[100] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace
}  // namespace omni

extern "C" int omni_debug_tmem_bench(long long* out, int mode, int nwarps, int iters, void* stream) {
  omni::tmem_bench_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(out, mode, nwarps, iters);
  OMNI_CUDA_LAUNCH_CHECK("tmem_bench_kernel");
  return OMNI_OK;
}
