// Shared host/device helpers for libomnissm (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/omnissm.h"

namespace omni {

// ---------------------------------------------------------------------------------------------
// host side: error reporting, argument validation, launch bookkeeping
// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);  // returns code
void count_launch(int n = 1);

#define OMNI_CHECK(cond, code, ...)                       \
  do {                                                    \
    if (!(cond)) return ::omni::set_error(code, __VA_ARGS__); \
  } while (0)

#define OMNI_CUDA_LAUNCH_CHECK(name)                                                         \
  do {                                                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                                                 \
    if (e__ != cudaSuccess)                                                                  \
      return ::omni::set_error(OMNI_CUDA_ERROR, "%s launch: %s", name, cudaGetErrorString(e__)); \
    ::omni::count_launch();                                                                  \
  } while (0)

inline bool present(const omni_tensor_t& t) { return t.data != nullptr; }
inline bool is_float_dtype(int dt) { return dt == OMNI_F32 || dt == OMNI_F16 || dt == OMNI_BF16; }
inline int dtype_size(int dt) { return dt == OMNI_F32 || dt == OMNI_I32 ? 4 : (dt == OMNI_I64 ? 8 : (dt == OMNI_U8 ? 1 : 2)); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline bool shape_is(const omni_tensor_t& t, int nd, int64_t a = -1, int64_t b = -1, int64_t c = -1, int64_t d = -1,
                     int64_t e = -1) {
  if (t.ndim != nd) return false;
  const int64_t want[5] = {a, b, c, d, e};
  for (int i = 0; i < nd && i < 5; ++i)
    if (want[i] >= 0 && t.shape[i] != want[i]) return false;
  return true;
}

int sm_count();  // of the current device (cached per device)

// Programmatic dependent launch (PDL) for the chains of small kernels of the decode step (add + norm -> in_proj -> layer core ->
// out_proj, 192 launches per token): a kernel launched through launch_pdl may begin - launch latency, prologue, loads of
// data no earlier kernel writes (weights, last step's states) - while its predecessor in the stream is still running;
// it calls pdl_wait() before it touches anything a predecessor produces or still reads, and pdl_trigger() at its top so
// that ITS successor may be scheduled in turn.  Works the same inside a captured CUDA graph (programmatic edges).
// OMNI_PDL in the environment is a mask of the kernels that may start early (kPdl*); 0 launches everything fully serialised.
enum { kPdlAddNorm = 1, kPdlGemm = 2, kPdlCore = 4, kPdlDefault = 3 };
bool pdl_enabled(int kind);
void set_pdl_mask(int mask);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled(kind)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// dispatch a lambda templated on the element type
#define OMNI_DISPATCH_FLOAT(DT, T, ...)                                             \
  [&]() -> int {                                                                    \
    switch (DT) {                                                                   \
      case OMNI_F32: { using T = float; return __VA_ARGS__(); }                     \
      case OMNI_F16: { using T = __half; return __VA_ARGS__(); }                    \
      case OMNI_BF16: { using T = __nv_bfloat16; return __VA_ARGS__(); }            \
      default: return ::omni::set_error(OMNI_BAD_DTYPE, "unsupported dtype %d", (int)(DT)); \
    }                                                                               \
  }()

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
// PDL (see launch_pdl): both are no-ops in a kernel launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Load/store by runtime dtype (parameters such as A, D, dt_bias, conv weights may be fp32 while
// activations are bf16).
__device__ __forceinline__ float ld_any(const void* p, int dtype, int64_t i) {
  if (dtype == OMNI_F32) return static_cast<const float*>(p)[i];
  if (dtype == OMNI_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(static_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, int64_t i, float v) {
  if (dtype == OMNI_F32) static_cast<float*>(p)[i] = v;
  else if (dtype == OMNI_BF16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else static_cast<__half*>(p)[i] = __float2half_rn(v);
}

// 16-byte vector of 8 (16-bit) or 4 (fp32) elements
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); T v[N]; };

template <typename T, int N> __device__ __forceinline__ void load_vec(const T* p, float (&out)[N]) {
  static_assert((N * sizeof(T)) % 16 == 0, "vector loads are 16B multiples");
  constexpr int CH = N * sizeof(T) / 16;
  constexpr int PER = 16 / sizeof(T);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint4 raw = *reinterpret_cast<const uint4*>(p + c * PER);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < PER; ++i) out[c * PER + i] = to_f<T>(e[i]);
  }
}
template <typename T, int N> __device__ __forceinline__ void store_vec(T* p, const float (&in)[N]) {
  static_assert((N * sizeof(T)) % 16 == 0, "vector stores are 16B multiples");
  constexpr int CH = N * sizeof(T) / 16;
  constexpr int PER = 16 / sizeof(T);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < PER; ++i) e[i] = from_f<T>(in[c * PER + i]);
    *reinterpret_cast<uint4*>(p + c * PER) = raw;
  }
}

// (fast division: one MUFU.RCP + one multiply, <= 2 ulp; the IEEE '/' costs ~12 instructions per element, which made the
// streaming kernels instruction-bound instead of HBM-bound)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ float sigmoid_f(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
// d/dv [v * sigmoid(v)]
__device__ __forceinline__ float dsilu_f(float v) {
  float s = sigmoid_f(v);
  return s * (1.f + v * (1.f - s));
}
// softplus with the upstream cut-over at 20
__device__ __forceinline__ float softplus_f(float v) { return v <= 20.f ? log1pf(expf(v)) : v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

}  // namespace omni
