// Gated (group) RMSNorm / LayerNorm and fused residual-add + norm, forward and backward.
//
// Replaces mamba_ssm/ops/triton/layernorm_gated.py (rmsnorm_fn, RMSNormGated: Mamba2.norm) and
// mamba_ssm/ops/triton/layer_norm.py (layer_norm_fn, RMSNorm: /root/reference/models/stage2/block.py:86-95,
// mixer_seq_simple.py:428-437).  Arithmetic: SURVEY.md Appendix A.4 / A.7, B.
//
// All four kernels are HBM streaming kernels: one CTA owns one (row, group), each thread keeps its
// 8-element chunks of the row in registers between the statistics pass and the output pass, so every
// input/output element crosses HBM once.  Loads/stores are 16-byte vectors when shapes allow.
// The backward kernels are persistent over rows and keep dweight/dbias partial sums in registers; they
// write one fp32 partial row per CTA (deterministic, no atomics) that the caller sums.
#include <algorithm>

#include "common.cuh"

namespace omni {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxChunks = 4;  // max 8-element chunks per thread held in registers (group_size <= 8192)

// ---- runtime-dtype 8-wide chunk access -------------------------------------------------------------
template <bool VEC>
__device__ __forceinline__ void ld8(const void* base, int dtype, int64_t idx, int valid, float (&o)[8]) {
  if constexpr (VEC) {
    if (dtype == OMNI_F32) {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
      float4 a = p[0], b = p[1];
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else if (dtype == OMNI_BF16) {
      load_vec<__nv_bfloat16, 8>(static_cast<const __nv_bfloat16*>(base) + idx, o);
    } else {
      load_vec<__half, 8>(static_cast<const __half*>(base) + idx, o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = i < valid ? ld_any(base, dtype, idx + i) : 0.f;
  }
}
template <bool VEC>
__device__ __forceinline__ void st8(void* base, int dtype, int64_t idx, int valid, const float (&o)[8]) {
  if constexpr (VEC) {
    if (dtype == OMNI_F32) {
      float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + idx);
      p[0] = make_float4(o[0], o[1], o[2], o[3]);
      p[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else if (dtype == OMNI_BF16) {
      store_vec<__nv_bfloat16, 8>(static_cast<__nv_bfloat16*>(base) + idx, o);
    } else {
      store_vec<__half, 8>(static_cast<__half*>(base) + idx, o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < valid) st_any(base, dtype, idx + i, o[i]);
  }
}

struct T2 {  // (rows, cols) view
  void* p; int dtype; int64_t rs;  // row stride (elements); columns are contiguous
};
inline T2 view2(const omni_tensor_t& t) { return T2{t.data, t.dtype, t.ndim == 2 ? t.stride[0] : 0}; }

struct GatedArgs {
  T2 x, z, out, dout, dx, dz, yrec;
  const void* w; const void* b; int w_dtype, b_dtype;
  float* rstd; float* mean; float* dw_part; float* db_part;
  int M, D, gs, ngroups;
  float eps;
  int norm_before_gate, is_rms;
};

// ---- gated norm forward ----------------------------------------------------------------------------
template <bool VEC, int NCH>
__global__ void __launch_bounds__(kThreads) norm_gated_fwd_kernel(GatedArgs a) {
  __shared__ float red[32];
  const int row = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const int64_t c0 = (int64_t)g * a.gs;
  float u[NCH][8], zz[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    const int valid = min(8, a.gs - col);
    if (col < a.gs) {
      ld8<VEC>(a.x.p, a.x.dtype, row * a.x.rs + c0 + col, valid, u[i]);
      if (a.z.p) {
        ld8<VEC>(a.z.p, a.z.dtype, row * a.z.rs + c0 + col, valid, zz[i]);
        if (!a.norm_before_gate) {
#pragma unroll
          for (int k = 0; k < 8; ++k) u[i][k] *= silu_f(zz[i][k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { s1 += u[i][k]; s2 += u[i][k] * u[i][k]; }
    }
  }
  float mu = 0.f, var;
  if (a.is_rms) {
    var = block_sum(s2, red) / a.gs;
  } else {
    mu = block_sum(s1, red) / a.gs;
    float sv = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      if (col < a.gs) {
        const int valid = min(8, a.gs - col);
#pragma unroll
        for (int k = 0; k < 8; ++k) sv += k < valid ? (u[i][k] - mu) * (u[i][k] - mu) : 0.f;
      }
    }
    var = block_sum(sv, red) / a.gs;
  }
  const float rstd = rsqrtf(var + a.eps);
  if (tid == 0) {
    if (a.rstd) a.rstd[(int64_t)row * a.ngroups + g] = rstd;
    if (a.mean && !a.is_rms) a.mean[(int64_t)row * a.ngroups + g] = mu;
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    if (col < a.gs) {
      const int valid = min(8, a.gs - col);
      float w[8], bb[8], y[8];
      ld8<VEC>(a.w, a.w_dtype, c0 + col, valid, w);
      if (a.b) ld8<VEC>(a.b, a.b_dtype, c0 + col, valid, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = (u[i][k] - mu) * rstd * w[k] + (a.b ? bb[k] : 0.f);
        if (a.z.p && a.norm_before_gate) v *= silu_f(zz[i][k]);
        y[k] = v;
      }
      st8<VEC>(a.out.p, a.out.dtype, row * a.out.rs + c0 + col, valid, y);
    }
  }
}

// ---- fast path of the gated RMSNorm forward (Mamba2.norm as OmniMamba configures it) -------------------------------
// y = rmsnorm(x * silu(z)) * w, one group, x / z / out in the same 16-bit type, D = NCH * 2048.  The generic kernel above
// executes ~35 instructions per element (runtime dtype dispatch, mean / bias / gate-order generality) and one CTA per row;
// here a persistent CTA keeps its weight columns in registers, loads the next row while the current one is reduced and
// uses one __syncthreads per row (the cross-warp scratch is double-buffered).
template <typename T> __device__ __forceinline__ void unpack8(const uint4& r, float (&o)[8]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& r, float (&o)[8]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  o[4] = __uint_as_float(r.z << 16); o[5] = __uint_as_float(r.z & 0xffff0000u);
  o[6] = __uint_as_float(r.w << 16); o[7] = __uint_as_float(r.w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& r, float (&o)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    o[2 * i] = f.x; o[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t r);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t r) {
  return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t r) { return __half22float2(*reinterpret_cast<const __half2*>(&r)); }
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kThreads, 4) rms_gated_fast_kernel(GatedArgs a) {
  __shared__ float red[2][kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float w[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) ld8<true>(a.w, a.w_dtype, (i * kThreads + tid) * 8, 8, w[i]);
  const T* xb = static_cast<const T*>(a.x.p);
  const T* zb = static_cast<const T*>(a.z.p);
  T* ob = static_cast<T*>(a.out.p);
  uint4 rx[NCH], rz[NCH];
  auto load = [&](int64_t r) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      rx[i] = __ldg(reinterpret_cast<const uint4*>(xb + r * a.x.rs + (i * kThreads + tid) * 8));
      rz[i] = __ldg(reinterpret_cast<const uint4*>(zb + r * a.z.rs + (i * kThreads + tid) * 8));
    }
  };
  int64_t row = blockIdx.x;
  if (row < a.M) load(row);
  const float inv_d = 1.f / (float)a.D;
#pragma unroll 1
  for (int it = 0; row < a.M; row += gridDim.x, ++it) {
    float u[NCH][8];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float xv[8], zv[8];
      unpack8<T>(rx[i], xv);
      unpack8<T>(rz[i], zv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        u[i][k] = xv[k] * silu_f(zv[k]);
        s2 = fmaf(u[i][k], u[i][k], s2);
      }
    }
    const int64_t nrow = row + gridDim.x;
    if (nrow < a.M) load(nrow);  // next row in flight during the reduction and the stores
    s2 = warp_sum(s2);
    if (lane == 0) red[it & 1][warp] = s2;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) tot += red[it & 1][k];
    const float rstd = rsqrtf(tot * inv_d + a.eps);
    if (tid == 0 && a.rstd) a.rstd[row] = rstd;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      uint4 o;
      o.x = pack2<T>(u[i][0] * rstd * w[i][0], u[i][1] * rstd * w[i][1]);
      o.y = pack2<T>(u[i][2] * rstd * w[i][2], u[i][3] * rstd * w[i][3]);
      o.z = pack2<T>(u[i][4] * rstd * w[i][4], u[i][5] * rstd * w[i][5]);
      o.w = pack2<T>(u[i][6] * rstd * w[i][6], u[i][7] * rstd * w[i][7]);
      *reinterpret_cast<uint4*>(ob + row * a.out.rs + (i * kThreads + tid) * 8) = o;
    }
  }
}

// ---- gated norm backward ---------------------------------------------------------------------------
template <bool VEC, int NCH>
__global__ void __launch_bounds__(kThreads) norm_gated_bwd_kernel(GatedArgs a) {
  __shared__ float red[32];
  const int g = blockIdx.y, tid = threadIdx.x;
  const int64_t c0 = (int64_t)g * a.gs;
  float w[NCH][8], bb[NCH][8], dwa[NCH][8], dba[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    const int valid = min(8, a.gs - col);
#pragma unroll
    for (int k = 0; k < 8; ++k) { dwa[i][k] = 0.f; dba[i][k] = 0.f; w[i][k] = 0.f; bb[i][k] = 0.f; }
    if (col < a.gs) {
      ld8<VEC>(a.w, a.w_dtype, c0 + col, valid, w[i]);
      if (a.b) ld8<VEC>(a.b, a.b_dtype, c0 + col, valid, bb[i]);
    }
  }
  for (int row = blockIdx.x; row < a.M; row += gridDim.x) {
    const float rstd = a.rstd[(int64_t)row * a.ngroups + g];
    const float mu = a.is_rms ? 0.f : a.mean[(int64_t)row * a.ngroups + g];
    float xh[NCH][8], wdy[NCH][8], xr[NCH][8], zz[NCH][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      const int valid = min(8, a.gs - col);
      if (col < a.gs) {
        float dy[8];
        ld8<VEC>(a.x.p, a.x.dtype, row * a.x.rs + c0 + col, valid, xr[i]);
        ld8<VEC>(a.dout.p, a.dout.dtype, row * a.dout.rs + c0 + col, valid, dy);
        if (a.z.p) ld8<VEC>(a.z.p, a.z.dtype, row * a.z.rs + c0 + col, valid, zz[i]);
        float yrec[8], dzv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float u = xr[i][k];
          if (a.z.p && !a.norm_before_gate) u *= silu_f(zz[i][k]);
          const float xhat = k < valid ? (u - mu) * rstd : 0.f;
          xh[i][k] = xhat;
          float yn = xhat * w[i][k] + bb[i][k];  // normalised (pre-gate when norm_before_gate)
          float dyn = dy[k];
          if (a.z.p && a.norm_before_gate) {
            dzv[k] = dy[k] * yn * dsilu_f(zz[i][k]);
            dyn = dy[k] * silu_f(zz[i][k]);
            yn *= silu_f(zz[i][k]);
          }
          yrec[k] = yn;
          dwa[i][k] += dyn * xhat;
          dba[i][k] += dyn;
          wdy[i][k] = k < valid ? dyn * w[i][k] : 0.f;
          c1 += xhat * wdy[i][k];
          c2 += wdy[i][k];
        }
        if (a.yrec.p) st8<VEC>(a.yrec.p, a.yrec.dtype, row * a.yrec.rs + c0 + col, valid, yrec);
        if (a.z.p && a.norm_before_gate) st8<VEC>(a.dz.p, a.dz.dtype, row * a.dz.rs + c0 + col, valid, dzv);
      }
    }
    c1 = block_sum(c1, red) / a.gs;
    c2 = a.is_rms ? 0.f : block_sum(c2, red) / a.gs;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      const int valid = min(8, a.gs - col);
      if (col < a.gs) {
        float dxv[8], dzv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float du = (wdy[i][k] - xh[i][k] * c1 - c2) * rstd;
          if (a.z.p && !a.norm_before_gate) {
            dxv[k] = du * silu_f(zz[i][k]);
            dzv[k] = du * xr[i][k] * dsilu_f(zz[i][k]);
          } else {
            dxv[k] = du;
          }
        }
        st8<VEC>(a.dx.p, a.dx.dtype, row * a.dx.rs + c0 + col, valid, dxv);
        if (a.z.p && !a.norm_before_gate) st8<VEC>(a.dz.p, a.dz.dtype, row * a.dz.rs + c0 + col, valid, dzv);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    const int valid = min(8, a.gs - col);
    if (col < a.gs) {
      st8<VEC>(a.dw_part, OMNI_F32, (int64_t)blockIdx.x * a.D + c0 + col, valid, dwa[i]);
      if (a.db_part) st8<VEC>(a.db_part, OMNI_F32, (int64_t)blockIdx.x * a.D + c0 + col, valid, dba[i]);
    }
  }
}

// ---- fast path of the gated RMSNorm backward (same configuration as rms_gated_fast_kernel) ---------------------------
// One thread owns 8 columns of the row (blockDim = D / 8), so the weight and the dweight partial sums of its columns stay in
// 16 registers for the whole launch; per row it keeps the raw 16-bit x / z / dy vectors, reduces c1 = mean(xhat * w dy)
// across the CTA with one __syncthreads (double-buffered scratch) while the next row's vectors are already in flight, and
// recomputes the cheap per-element terms after the reduction instead of holding them in registers.
template <typename T, bool YREC>
__global__ void __launch_bounds__(1024, 1) rms_gated_bwd_fast_kernel(GatedArgs a) {
  __shared__ float red[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int col = tid * 8;
  float w[8], dwa[8];
  ld8<true>(a.w, a.w_dtype, col, 8, w);
#pragma unroll
  for (int k = 0; k < 8; ++k) dwa[k] = 0.f;
  const T* xb = static_cast<const T*>(a.x.p);
  const T* zb = static_cast<const T*>(a.z.p);
  const T* gb = static_cast<const T*>(a.dout.p);
  uint4 rx, rz, rg;
  float rstd_n = 0.f;
  auto load = [&](int64_t r) {
    rx = __ldg(reinterpret_cast<const uint4*>(xb + r * a.x.rs + col));
    rz = __ldg(reinterpret_cast<const uint4*>(zb + r * a.z.rs + col));
    rg = __ldg(reinterpret_cast<const uint4*>(gb + r * a.dout.rs + col));
    rstd_n = __ldg(a.rstd + r);
  };
  int64_t row = blockIdx.x;
  if (row < a.M) load(row);
  const float inv_d = 1.f / (float)a.D;
#pragma unroll 1
  for (int it = 0; row < a.M; row += gridDim.x, ++it) {
    float gv[8], xh[8], sa[8], sb[8];  // w dy, xhat, silu(z), x dsilu(z)
    const float rstd = rstd_n;
    float c1 = 0.f;
    {
      float xv[8], zv[8];
      unpack8<T>(rx, xv);
      unpack8<T>(rz, zv);
      unpack8<T>(rg, gv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float sg = sigmoid_f(zv[k]);
        sa[k] = zv[k] * sg;
        sb[k] = xv[k] * (sg + sa[k] * (1.f - sg));    // x * d silu / dz,  d silu / dz = sg + silu (1 - sg)
        xh[k] = xv[k] * sa[k] * rstd;                 // xhat = x silu(z) rstd
        dwa[k] = fmaf(gv[k], xh[k], dwa[k]);
        gv[k] *= w[k];                                // w dy
        c1 = fmaf(xh[k], gv[k], c1);
      }
    }
    const int64_t nrow = row + gridDim.x;
    if (nrow < a.M) load(nrow);  // (the raw vectors of this row are fully unpacked above)
    c1 = warp_sum(c1);
    if (lane == 0) red[it & 1][warp] = c1;
    __syncthreads();
    float tot = 0.f;
    for (int k = 0; k < nwarps; ++k) tot += red[it & 1][k];
    c1 = tot * inv_d;
    uint32_t odx[4], odz[4], oy[4];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float du0 = (gv[k] - xh[k] * c1) * rstd, du1 = (gv[k + 1] - xh[k + 1] * c1) * rstd;
      odx[k >> 1] = pack2<T>(du0 * sa[k], du1 * sa[k + 1]);
      odz[k >> 1] = pack2<T>(du0 * sb[k], du1 * sb[k + 1]);
      oy[k >> 1] = pack2<T>(xh[k] * w[k], xh[k + 1] * w[k + 1]);
    }
    *reinterpret_cast<uint4*>(static_cast<T*>(a.dx.p) + row * a.dx.rs + col) = make_uint4(odx[0], odx[1], odx[2], odx[3]);
    *reinterpret_cast<uint4*>(static_cast<T*>(a.dz.p) + row * a.dz.rs + col) = make_uint4(odz[0], odz[1], odz[2], odz[3]);
    if (YREC) *reinterpret_cast<uint4*>(static_cast<T*>(a.yrec.p) + row * a.yrec.rs + col) = make_uint4(oy[0], oy[1], oy[2], oy[3]);
  }
  float4* dst = reinterpret_cast<float4*>(a.dw_part + (int64_t)blockIdx.x * a.D + col);
  dst[0] = make_float4(dwa[0], dwa[1], dwa[2], dwa[3]);
  dst[1] = make_float4(dwa[4], dwa[5], dwa[6], dwa[7]);
}

// ---- fused add + norm ------------------------------------------------------------------------------
struct AddNormArgs {
  T2 x, res, y, res_out, dy, dres_in, dx, dres_out;
  const void* w; const void* b; int w_dtype, b_dtype;
  float* rstd; float* mean; float* dw_part; float* db_part;
  int M, D;
  float eps;
  int is_rms;
};

template <bool VEC, int NCH>
__global__ void __launch_bounds__(kThreads) add_norm_fwd_kernel(AddNormArgs a) {
  __shared__ float red[32];
  const int row = blockIdx.x, tid = threadIdx.x;
  float r[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    const int valid = min(8, a.D - col);
    if (col < a.D) {
      ld8<VEC>(a.x.p, a.x.dtype, row * a.x.rs + col, valid, r[i]);
      if (a.res.p) {
        float q[8];
        ld8<VEC>(a.res.p, a.res.dtype, row * a.res.rs + col, valid, q);
#pragma unroll
        for (int k = 0; k < 8; ++k) r[i][k] += q[k];
      }
      if (a.res_out.p) st8<VEC>(a.res_out.p, a.res_out.dtype, row * a.res_out.rs + col, valid, r[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s1 += r[i][k]; s2 += r[i][k] * r[i][k]; }
    }
  }
  float mu = 0.f, var;
  if (a.is_rms) {
    var = block_sum(s2, red) / a.D;
  } else {
    mu = block_sum(s1, red) / a.D;
    float sv = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      if (col < a.D) {
        const int valid = min(8, a.D - col);
#pragma unroll
        for (int k = 0; k < 8; ++k) sv += k < valid ? (r[i][k] - mu) * (r[i][k] - mu) : 0.f;
      }
    }
    var = block_sum(sv, red) / a.D;
  }
  const float rstd = rsqrtf(var + a.eps);
  if (tid == 0) {
    if (a.rstd) a.rstd[row] = rstd;
    if (a.mean && !a.is_rms) a.mean[row] = mu;
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    if (col < a.D) {
      const int valid = min(8, a.D - col);
      float w[8], bb[8], y[8];
      ld8<VEC>(a.w, a.w_dtype, col, valid, w);
      if (a.b) ld8<VEC>(a.b, a.b_dtype, col, valid, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = (r[i][k] - mu) * rstd * w[k] + (a.b ? bb[k] : 0.f);
      st8<VEC>(a.y.p, a.y.dtype, row * a.y.rs + col, valid, y);
    }
  }
}

// ---- fast path of the fused residual-add + RMSNorm forward (Block.forward, /root/reference/models/stage2/block.py:86-95) -
// x in a 16-bit type, residual / residual_out fp32 (residual_in_fp32=True) or absent, y in x's type, no bias, D = 8 columns
// per thread (1024 / 2048 / 4096).  Persistent CTAs: weight columns in registers, the next row's x and residual vectors in
// flight during the reduction and the stores, one __syncthreads per row.
template <typename T>
__global__ void __launch_bounds__(512) add_rms_fast_kernel(AddNormArgs a) {
  __shared__ float red[2][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int col = tid * 8;
  pdl_trigger();   // (PDL, common.cuh: the next kernel of the decode chain may start its prologue)
  float w[8];
  ld8<true>(a.w, a.w_dtype, col, 8, w);
  pdl_wait();      // x / residual come from the previous kernel
  const T* xb = static_cast<const T*>(a.x.p);
  const float* rb = static_cast<const float*>(a.res.p);
  uint4 rx;
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
  auto load = [&](int64_t r) {
    rx = __ldg(reinterpret_cast<const uint4*>(xb + r * a.x.rs + col));
    if (rb) {
      r0 = __ldg(reinterpret_cast<const float4*>(rb + r * a.res.rs + col));
      r1 = __ldg(reinterpret_cast<const float4*>(rb + r * a.res.rs + col + 4));
    }
  };
  int64_t row = blockIdx.x;
  if (row < a.M) load(row);
  const float inv_d = 1.f / (float)a.D;
#pragma unroll 1
  for (int it = 0; row < a.M; row += gridDim.x, ++it) {
    float v[8];
    unpack8<T>(rx, v);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    const int64_t nrow = row + gridDim.x;
    if (nrow < a.M) load(nrow);
    if (a.res_out.p) {
      float* ro = static_cast<float*>(a.res_out.p) + row * a.res_out.rs + col;
      *reinterpret_cast<float4*>(ro) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(ro + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s2 = fmaf(v[k], v[k], s2);
    s2 = warp_sum(s2);
    if (lane == 0) red[it & 1][warp] = s2;
    __syncthreads();
    float tot = 0.f;
    for (int k = 0; k < nwarps; ++k) tot += red[it & 1][k];
    const float rstd = rsqrtf(tot * inv_d + a.eps);
    if (tid == 0 && a.rstd) a.rstd[row] = rstd;
    uint4 o;
    o.x = pack2<T>(v[0] * rstd * w[0], v[1] * rstd * w[1]);
    o.y = pack2<T>(v[2] * rstd * w[2], v[3] * rstd * w[3]);
    o.z = pack2<T>(v[4] * rstd * w[4], v[5] * rstd * w[5]);
    o.w = pack2<T>(v[6] * rstd * w[6], v[7] * rstd * w[7]);
    *reinterpret_cast<uint4*>(static_cast<T*>(a.y.p) + row * a.y.rs + col) = o;
  }
}

// ---- fast path of the add + RMSNorm backward (the configuration of the 48-layer training loop: fp32 residual stream, 16-bit
// dy / dx, RMSNorm without bias, 1024 / 2048 / 4096 columns).  As in the gated backward a thread owns 8 columns for the whole
// launch (weight and dweight partial sums in registers), the next row's vectors are in flight during the reduction and the
// stores, and there is one __syncthreads per row (double-buffered scratch).  The generic kernel below walks its rows without
// a prefetch and ran at 60 % of the HBM roofline at (29 610, 2048).
template <typename T, int CPT>   // CPT columns per thread: 4 (1024 / 2048 columns: 45 registers, 32 warps per SM) or 8 (4096)
__global__ void __launch_bounds__(512, CPT == 4 ? 2 : 1) add_rms_bwd_fast_kernel(AddNormArgs a) {
  constexpr int NV = CPT / 4;   // float4 vectors per fp32 row segment
  __shared__ float red[2][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int col = tid * CPT;
  float w[CPT], dwa[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) { w[k] = ld_any(a.w, a.w_dtype, col + k); dwa[k] = 0.f; }
  const float* xb = static_cast<const float*>(a.x.p);
  const T* gb = static_cast<const T*>(a.dy.p);
  const float* qb = static_cast<const float*>(a.dres_in.p);
  float4 xr[NV], qr[NV];
  uint32_t rg[CPT / 2];
#pragma unroll
  for (int v = 0; v < NV; ++v) xr[v] = qr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  float rstd_n = 0.f;
  // the next row's vectors are requested as soon as the registers that hold this row's are dead: x / dy / rstd right after the
  // per-element terms (before the reduction), the incoming residual gradient after it has been added
  auto load_xg = [&](int64_t r) {
#pragma unroll
    for (int v = 0; v < NV; ++v) xr[v] = __ldg(reinterpret_cast<const float4*>(xb + r * a.x.rs + col + 4 * v));
    if constexpr (CPT == 8) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(gb + r * a.dy.rs + col));
      rg[0] = t.x; rg[1] = t.y; rg[2] = t.z; rg[3] = t.w;
    } else {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(gb + r * a.dy.rs + col));
      rg[0] = t.x; rg[1] = t.y;
    }
    rstd_n = __ldg(a.rstd + r);
  };
  auto load_q = [&](int64_t r) {
    if (qb) {
#pragma unroll
      for (int v = 0; v < NV; ++v) qr[v] = __ldg(reinterpret_cast<const float4*>(qb + r * a.dres_in.rs + col + 4 * v));
    }
  };
  int64_t row = blockIdx.x;
  if (row < a.M) { load_xg(row); load_q(row); }
  const float inv_d = 1.f / (float)a.D;
#pragma unroll 1
  for (int it = 0; row < a.M; row += gridDim.x, ++it) {
    const float rstd = rstd_n;
    float gv[CPT], xh[CPT];   // w dy, xhat
#pragma unroll
    for (int k = 0; k < CPT / 2; ++k) {
      const float2 g2 = unpack2<T>(rg[k]);
      gv[2 * k] = g2.x; gv[2 * k + 1] = g2.y;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      xh[4 * v] = xr[v].x * rstd; xh[4 * v + 1] = xr[v].y * rstd; xh[4 * v + 2] = xr[v].z * rstd; xh[4 * v + 3] = xr[v].w * rstd;
    }
    float c1 = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      dwa[k] = fmaf(gv[k], xh[k], dwa[k]);
      gv[k] *= w[k];
      c1 = fmaf(xh[k], gv[k], c1);
    }
    const int64_t nrow = row + gridDim.x;
    if (nrow < a.M) load_xg(nrow);
    c1 = warp_sum(c1);
    if (lane == 0) red[it & 1][warp] = c1;
    __syncthreads();
    float tot = 0.f;
    for (int k = 0; k < nwarps; ++k) tot += red[it & 1][k];
    c1 = tot * inv_d;
    float dxv[CPT];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      dxv[4 * v] = fmaf(gv[4 * v] - xh[4 * v] * c1, rstd, qr[v].x);
      dxv[4 * v + 1] = fmaf(gv[4 * v + 1] - xh[4 * v + 1] * c1, rstd, qr[v].y);
      dxv[4 * v + 2] = fmaf(gv[4 * v + 2] - xh[4 * v + 2] * c1, rstd, qr[v].z);
      dxv[4 * v + 3] = fmaf(gv[4 * v + 3] - xh[4 * v + 3] * c1, rstd, qr[v].w);
    }
    if (nrow < a.M) load_q(nrow);
    T* dxp = static_cast<T*>(a.dx.p) + row * a.dx.rs + col;
    if constexpr (CPT == 8) {
      *reinterpret_cast<uint4*>(dxp) = make_uint4(pack2<T>(dxv[0], dxv[1]), pack2<T>(dxv[2], dxv[3]), pack2<T>(dxv[4], dxv[5]), pack2<T>(dxv[6], dxv[7]));
    } else {
      *reinterpret_cast<uint2*>(dxp) = make_uint2(pack2<T>(dxv[0], dxv[1]), pack2<T>(dxv[2], dxv[3]));
    }
    if (a.dres_out.p) {
      float* ro = static_cast<float*>(a.dres_out.p) + row * a.dres_out.rs + col;
#pragma unroll
      for (int v = 0; v < NV; ++v) *reinterpret_cast<float4*>(ro + 4 * v) = make_float4(dxv[4 * v], dxv[4 * v + 1], dxv[4 * v + 2], dxv[4 * v + 3]);
    }
  }
  float4* dst = reinterpret_cast<float4*>(a.dw_part + (int64_t)blockIdx.x * a.D + col);
#pragma unroll
  for (int v = 0; v < NV; ++v) dst[v] = make_float4(dwa[4 * v], dwa[4 * v + 1], dwa[4 * v + 2], dwa[4 * v + 3]);
}

template <bool VEC, int NCH>
__global__ void __launch_bounds__(kThreads) add_norm_bwd_kernel(AddNormArgs a) {
  __shared__ float red[32];
  const int tid = threadIdx.x;
  float w[NCH][8], dwa[NCH][8], dba[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { dwa[i][k] = 0.f; dba[i][k] = 0.f; w[i][k] = 0.f; }
    if (col < a.D) ld8<VEC>(a.w, a.w_dtype, col, min(8, a.D - col), w[i]);
  }
  for (int row = blockIdx.x; row < a.M; row += gridDim.x) {
    const float rstd = a.rstd[row];
    const float mu = a.is_rms ? 0.f : a.mean[row];
    float xh[NCH][8], wdy[NCH][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      const int valid = min(8, a.D - col);
      if (col < a.D) {
        float xr[8], dy[8];
        ld8<VEC>(a.x.p, a.x.dtype, row * a.x.rs + col, valid, xr);
        ld8<VEC>(a.dy.p, a.dy.dtype, row * a.dy.rs + col, valid, dy);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xhat = k < valid ? (xr[k] - mu) * rstd : 0.f;
          xh[i][k] = xhat;
          dwa[i][k] += dy[k] * xhat;
          dba[i][k] += dy[k];
          wdy[i][k] = k < valid ? dy[k] * w[i][k] : 0.f;
          c1 += xhat * wdy[i][k];
          c2 += wdy[i][k];
        }
      }
    }
    c1 = block_sum(c1, red) / a.D;
    c2 = a.is_rms ? 0.f : block_sum(c2, red) / a.D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int col = (i * kThreads + tid) * 8;
      const int valid = min(8, a.D - col);
      if (col < a.D) {
        float dxv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dxv[k] = (wdy[i][k] - xh[i][k] * c1 - c2) * rstd;
        if (a.dres_in.p) {
          float q[8];
          ld8<VEC>(a.dres_in.p, a.dres_in.dtype, row * a.dres_in.rs + col, valid, q);
#pragma unroll
          for (int k = 0; k < 8; ++k) dxv[k] += q[k];
        }
        st8<VEC>(a.dx.p, a.dx.dtype, row * a.dx.rs + col, valid, dxv);
        if (a.dres_out.p) st8<VEC>(a.dres_out.p, a.dres_out.dtype, row * a.dres_out.rs + col, valid, dxv);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int col = (i * kThreads + tid) * 8;
    if (col < a.D) {
      const int valid = min(8, a.D - col);
      st8<VEC>(a.dw_part, OMNI_F32, (int64_t)blockIdx.x * a.D + col, valid, dwa[i]);
      if (a.db_part) st8<VEC>(a.db_part, OMNI_F32, (int64_t)blockIdx.x * a.D + col, valid, dba[i]);
    }
  }
}

// ---- validation helpers ----------------------------------------------------------------------------
int check_rows(const omni_tensor_t& t, int64_t M, int64_t D, const char* name, bool required = true) {
  if (!present(t)) {
    OMNI_CHECK(!required, OMNI_BAD_SHAPE, "%s is required", name);
    return OMNI_OK;
  }
  OMNI_CHECK(t.ndim == 2 && t.shape[0] == M && t.shape[1] == D, OMNI_BAD_SHAPE, "%s must be (%lld, %lld)", name,
             (long long)M, (long long)D);
  OMNI_CHECK(D <= 1 || t.stride[1] == 1, OMNI_BAD_STRIDE, "%s: last dim must be contiguous", name);
  OMNI_CHECK(is_float_dtype(t.dtype), OMNI_BAD_DTYPE, "%s: bad dtype", name);
  return OMNI_OK;
}
int check_vecparam(const omni_tensor_t& t, int64_t D, const char* name, bool required) {
  if (!present(t)) {
    OMNI_CHECK(!required, OMNI_BAD_SHAPE, "%s is required", name);
    return OMNI_OK;
  }
  OMNI_CHECK(t.ndim == 1 && t.shape[0] == D && (D <= 1 || t.stride[0] == 1), OMNI_BAD_SHAPE,
             "%s must be contiguous (%lld)", name, (long long)D);
  OMNI_CHECK(is_float_dtype(t.dtype), OMNI_BAD_DTYPE, "%s: bad dtype", name);
  return OMNI_OK;
}
bool rows_vec_ok(const omni_tensor_t& t, int64_t unit) {
  if (!present(t)) return true;
  const int64_t align = t.dtype == OMNI_F32 ? 16 : 16;
  if (reinterpret_cast<uintptr_t>(t.data) % align) return false;
  if (t.ndim == 2) return t.stride[0] % 8 == 0 && t.shape[1] % 8 == 0 && unit % 8 == 0;
  return t.shape[0] % 8 == 0;
}
bool f32_contig(const omni_tensor_t& t, int64_t n) {
  if (t.dtype != OMNI_F32) return false;
  int64_t tot = 1;
  for (int i = 0; i < t.ndim; ++i) tot *= t.shape[i];
  if (tot != n) return false;
  int64_t exp = 1;
  for (int i = t.ndim - 1; i >= 0; --i) {
    if (t.shape[i] != 1 && t.stride[i] != exp) return false;
    exp *= t.shape[i];
  }
  return true;
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_norm_gated_fwd(const omni_norm_gated_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  OMNI_CHECK(present(p->x) && p->x.ndim == 2, OMNI_BAD_SHAPE, "norm_gated: x must be (rows, dim)");
  const int64_t M = p->x.shape[0], D = p->x.shape[1];
  const int64_t gs = p->group_size > 0 ? p->group_size : D;
  OMNI_CHECK(gs > 0 && D % gs == 0, OMNI_BAD_SHAPE, "norm_gated: dim %lld not divisible by group_size %lld",
             (long long)D, (long long)gs);
  OMNI_CHECK(gs <= kThreads * 8 * kMaxChunks, OMNI_UNSUPPORTED, "norm_gated: group_size > %d", kThreads * 8 * kMaxChunks);
  if (int rc = check_rows(p->x, M, D, "x")) return rc;
  if (int rc = check_rows(p->z, M, D, "z", false)) return rc;
  if (int rc = check_rows(p->out, M, D, "out")) return rc;
  if (int rc = check_vecparam(p->weight, D, "weight", true)) return rc;
  if (int rc = check_vecparam(p->bias, D, "bias", false)) return rc;
  const int64_t ng = D / gs;
  if (present(p->rstd)) OMNI_CHECK(f32_contig(p->rstd, M * ng), OMNI_BAD_SHAPE, "rstd must be contiguous fp32 (rows*ngroups)");
  if (present(p->mean)) OMNI_CHECK(f32_contig(p->mean, M * ng), OMNI_BAD_SHAPE, "mean must be contiguous fp32 (rows*ngroups)");
  if (M == 0 || D == 0) return OMNI_OK;
  GatedArgs a{};
  a.x = view2(p->x); a.z = view2(p->z); a.out = view2(p->out);
  a.w = p->weight.data; a.w_dtype = p->weight.dtype; a.b = p->bias.data; a.b_dtype = p->bias.dtype;
  a.rstd = static_cast<float*>(p->rstd.data); a.mean = static_cast<float*>(p->mean.data);
  a.M = (int)M; a.D = (int)D; a.gs = (int)gs; a.ngroups = (int)ng; a.eps = p->eps;
  a.norm_before_gate = p->norm_before_gate; a.is_rms = p->is_rms_norm;
  const bool vec = rows_vec_ok(p->x, gs) && rows_vec_ok(p->z, gs) && rows_vec_ok(p->out, gs) &&
                   rows_vec_ok(p->weight, gs) && rows_vec_ok(p->bias, gs);
  dim3 grid((unsigned)M, (unsigned)ng);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nch = (int)(((gs) + kThreads * 8 - 1) / (kThreads * 8));
  // fast path: gate-then-RMSNorm, one group of NCH * 2048 columns, 16-bit x / z / out of one type, no bias
  if (vec && a.is_rms && present(p->z) && !a.norm_before_gate && ng == 1 && !present(p->bias) && !present(p->mean) &&
      (D == 2048 || D == 4096 || D == 8192) && p->x.dtype != OMNI_F32 && p->z.dtype == p->x.dtype && p->out.dtype == p->x.dtype) {
    const unsigned gridp = (unsigned)std::min<int64_t>(M, (int64_t)sm_count() * 4);  // 4 resident CTAs per SM (64 registers)
#define OMNI_LAUNCH_FAST(T)                                                               \
    do {                                                                                  \
      if (D == 2048) rms_gated_fast_kernel<T, 1><<<gridp, kThreads, 0, s>>>(a);           \
      else if (D == 4096) rms_gated_fast_kernel<T, 2><<<gridp, kThreads, 0, s>>>(a);      \
      else rms_gated_fast_kernel<T, 4><<<gridp, kThreads, 0, s>>>(a);                     \
    } while (0)
    if (p->x.dtype == OMNI_BF16) OMNI_LAUNCH_FAST(__nv_bfloat16); else OMNI_LAUNCH_FAST(__half);
#undef OMNI_LAUNCH_FAST
    OMNI_CUDA_LAUNCH_CHECK("rms_gated_fast_kernel");
    return OMNI_OK;
  }
#define OMNI_LAUNCH_NCH(V, N) norm_gated_fwd_kernel<V, N><<<grid, kThreads, 0, s>>>(a)
  if (vec) { if (nch <= 1) OMNI_LAUNCH_NCH(true, 1); else if (nch == 2) OMNI_LAUNCH_NCH(true, 2); else OMNI_LAUNCH_NCH(true, 4); }
  else { if (nch <= 1) OMNI_LAUNCH_NCH(false, 1); else if (nch == 2) OMNI_LAUNCH_NCH(false, 2); else OMNI_LAUNCH_NCH(false, 4); }
#undef OMNI_LAUNCH_NCH
  OMNI_CUDA_LAUNCH_CHECK("norm_gated_fwd_kernel");
  return OMNI_OK;
}

extern "C" int omni_norm_gated_bwd(const omni_norm_gated_bwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  OMNI_CHECK(present(p->x) && p->x.ndim == 2, OMNI_BAD_SHAPE, "norm_gated bwd: x must be (rows, dim)");
  const int64_t M = p->x.shape[0], D = p->x.shape[1];
  const int64_t gs = p->group_size > 0 ? p->group_size : D;
  OMNI_CHECK(gs > 0 && D % gs == 0, OMNI_BAD_SHAPE, "norm_gated bwd: bad group_size");
  OMNI_CHECK(gs <= kThreads * 8 * kMaxChunks, OMNI_UNSUPPORTED, "norm_gated bwd: group_size too large");
  if (int rc = check_rows(p->x, M, D, "x")) return rc;
  if (int rc = check_rows(p->z, M, D, "z", false)) return rc;
  if (int rc = check_rows(p->dout, M, D, "dout")) return rc;
  if (int rc = check_rows(p->dx, M, D, "dx")) return rc;
  if (int rc = check_rows(p->dz, M, D, "dz", present(p->z))) return rc;
  if (int rc = check_rows(p->out_recompute, M, D, "out_recompute", false)) return rc;
  if (int rc = check_vecparam(p->weight, D, "weight", true)) return rc;
  if (int rc = check_vecparam(p->bias, D, "bias", false)) return rc;
  const int64_t ng = D / gs;
  OMNI_CHECK(present(p->rstd) && f32_contig(p->rstd, M * ng), OMNI_BAD_SHAPE, "rstd must be contiguous fp32");
  if (!p->is_rms_norm) OMNI_CHECK(present(p->mean) && f32_contig(p->mean, M * ng), OMNI_BAD_SHAPE, "mean required");
  OMNI_CHECK(present(p->dw_part) && p->dw_part.ndim == 2 && p->dw_part.shape[1] == D && p->dw_part.shape[0] >= 1 &&
                 f32_contig(p->dw_part, p->dw_part.shape[0] * D),
             OMNI_BAD_SHAPE, "dw_part must be contiguous fp32 (nparts, dim)");
  const int64_t nparts = p->dw_part.shape[0];
  if (present(p->bias))
    OMNI_CHECK(present(p->db_part) && f32_contig(p->db_part, nparts * D), OMNI_BAD_SHAPE,
               "db_part must be contiguous fp32 (nparts, dim)");
  if (D == 0) return OMNI_OK;
  GatedArgs a{};
  a.x = view2(p->x); a.z = view2(p->z); a.dout = view2(p->dout); a.dx = view2(p->dx); a.dz = view2(p->dz);
  a.yrec = view2(p->out_recompute);
  a.w = p->weight.data; a.w_dtype = p->weight.dtype; a.b = p->bias.data; a.b_dtype = p->bias.dtype;
  a.rstd = static_cast<float*>(p->rstd.data); a.mean = static_cast<float*>(p->mean.data);
  a.dw_part = static_cast<float*>(p->dw_part.data); a.db_part = static_cast<float*>(p->db_part.data);
  a.M = (int)M; a.D = (int)D; a.gs = (int)gs; a.ngroups = (int)ng; a.eps = p->eps;
  a.norm_before_gate = p->norm_before_gate; a.is_rms = p->is_rms_norm;
  const bool vec = rows_vec_ok(p->x, gs) && rows_vec_ok(p->z, gs) && rows_vec_ok(p->dout, gs) && rows_vec_ok(p->dx, gs) &&
                   rows_vec_ok(p->dz, gs) && rows_vec_ok(p->out_recompute, gs) && rows_vec_ok(p->weight, gs) &&
                   rows_vec_ok(p->bias, gs);
  dim3 grid((unsigned)nparts, (unsigned)ng);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nch = (int)(((gs) + kThreads * 8 - 1) / (kThreads * 8));
  // fast path: gate-then-RMSNorm, one group of 2048 / 4096 / 8192 columns, 16-bit tensors of one type, no bias
  if (vec && a.is_rms && present(p->z) && !a.norm_before_gate && ng == 1 && !present(p->bias) &&
      (D == 2048 || D == 4096 || D == 8192) && p->x.dtype != OMNI_F32 && p->z.dtype == p->x.dtype &&
      p->dout.dtype == p->x.dtype && p->dx.dtype == p->x.dtype && p->dz.dtype == p->x.dtype &&
      (!present(p->out_recompute) || p->out_recompute.dtype == p->x.dtype)) {
    const unsigned threads = (unsigned)(D / 8);
    const bool yr = present(p->out_recompute);
#define OMNI_LAUNCH_FAST(T)                                                                       \
    do {                                                                                          \
      if (yr) rms_gated_bwd_fast_kernel<T, true><<<(unsigned)nparts, threads, 0, s>>>(a);         \
      else rms_gated_bwd_fast_kernel<T, false><<<(unsigned)nparts, threads, 0, s>>>(a);           \
    } while (0)
    if (p->x.dtype == OMNI_BF16) OMNI_LAUNCH_FAST(__nv_bfloat16); else OMNI_LAUNCH_FAST(__half);
#undef OMNI_LAUNCH_FAST
    OMNI_CUDA_LAUNCH_CHECK("rms_gated_bwd_fast_kernel");
    return OMNI_OK;
  }
#define OMNI_LAUNCH_NCH(V, N) norm_gated_bwd_kernel<V, N><<<grid, kThreads, 0, s>>>(a)
  if (vec) { if (nch <= 1) OMNI_LAUNCH_NCH(true, 1); else if (nch == 2) OMNI_LAUNCH_NCH(true, 2); else OMNI_LAUNCH_NCH(true, 4); }
  else { if (nch <= 1) OMNI_LAUNCH_NCH(false, 1); else if (nch == 2) OMNI_LAUNCH_NCH(false, 2); else OMNI_LAUNCH_NCH(false, 4); }
#undef OMNI_LAUNCH_NCH
  OMNI_CUDA_LAUNCH_CHECK("norm_gated_bwd_kernel");
  return OMNI_OK;
}

extern "C" int omni_add_norm_fwd(const omni_add_norm_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  OMNI_CHECK(present(p->x) && p->x.ndim == 2, OMNI_BAD_SHAPE, "add_norm: x must be (rows, dim)");
  const int64_t M = p->x.shape[0], D = p->x.shape[1];
  OMNI_CHECK(D <= kThreads * 8 * kMaxChunks, OMNI_UNSUPPORTED, "add_norm: dim > %d", kThreads * 8 * kMaxChunks);
  if (int rc = check_rows(p->x, M, D, "x")) return rc;
  if (int rc = check_rows(p->residual, M, D, "residual", false)) return rc;
  if (int rc = check_rows(p->y, M, D, "y")) return rc;
  if (int rc = check_rows(p->residual_out, M, D, "residual_out", false)) return rc;
  if (int rc = check_vecparam(p->weight, D, "weight", true)) return rc;
  if (int rc = check_vecparam(p->bias, D, "bias", false)) return rc;
  if (present(p->rstd)) OMNI_CHECK(f32_contig(p->rstd, M), OMNI_BAD_SHAPE, "rstd must be contiguous fp32 (rows)");
  if (present(p->mean)) OMNI_CHECK(f32_contig(p->mean, M), OMNI_BAD_SHAPE, "mean must be contiguous fp32 (rows)");
  if (M == 0 || D == 0) return OMNI_OK;
  AddNormArgs a{};
  a.x = view2(p->x); a.res = view2(p->residual); a.y = view2(p->y); a.res_out = view2(p->residual_out);
  a.w = p->weight.data; a.w_dtype = p->weight.dtype; a.b = p->bias.data; a.b_dtype = p->bias.dtype;
  a.rstd = static_cast<float*>(p->rstd.data); a.mean = static_cast<float*>(p->mean.data);
  a.M = (int)M; a.D = (int)D; a.eps = p->eps; a.is_rms = p->is_rms_norm;
  const bool vec = rows_vec_ok(p->x, D) && rows_vec_ok(p->residual, D) && rows_vec_ok(p->y, D) &&
                   rows_vec_ok(p->residual_out, D) && rows_vec_ok(p->weight, D) && rows_vec_ok(p->bias, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nch = (int)(((D) + kThreads * 8 - 1) / (kThreads * 8));
  // fast path: RMSNorm, 16-bit x / y of one type, fp32 (or no) residual in and fp32 (or no) residual out, no bias
  if (vec && a.is_rms && !present(p->bias) && !present(p->mean) && (D == 1024 || D == 2048 || D == 4096) &&
      p->x.dtype != OMNI_F32 && p->y.dtype == p->x.dtype && (!present(p->residual) || p->residual.dtype == OMNI_F32) &&
      (!present(p->residual_out) || p->residual_out.dtype == OMNI_F32)) {
    const unsigned threads = (unsigned)(D / 8);
    const unsigned gridp = (unsigned)std::min<int64_t>(M, (int64_t)sm_count() * (2048 / threads));
    if (p->x.dtype == OMNI_BF16) launch_pdl(kPdlAddNorm, add_rms_fast_kernel<__nv_bfloat16>, dim3(gridp), dim3(threads), 0, s, 1, a);
    else launch_pdl(kPdlAddNorm, add_rms_fast_kernel<__half>, dim3(gridp), dim3(threads), 0, s, 1, a);
    OMNI_CUDA_LAUNCH_CHECK("add_rms_fast_kernel");
    return OMNI_OK;
  }
#define OMNI_LAUNCH_NCH(V, N) add_norm_fwd_kernel<V, N><<<(unsigned)M, kThreads, 0, s>>>(a)
  if (vec) { if (nch <= 1) OMNI_LAUNCH_NCH(true, 1); else if (nch == 2) OMNI_LAUNCH_NCH(true, 2); else OMNI_LAUNCH_NCH(true, 4); }
  else { if (nch <= 1) OMNI_LAUNCH_NCH(false, 1); else if (nch == 2) OMNI_LAUNCH_NCH(false, 2); else OMNI_LAUNCH_NCH(false, 4); }
#undef OMNI_LAUNCH_NCH
  OMNI_CUDA_LAUNCH_CHECK("add_norm_fwd_kernel");
  return OMNI_OK;
}

extern "C" int omni_add_norm_bwd(const omni_add_norm_bwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  OMNI_CHECK(present(p->xres) && p->xres.ndim == 2, OMNI_BAD_SHAPE, "add_norm bwd: xres must be (rows, dim)");
  const int64_t M = p->xres.shape[0], D = p->xres.shape[1];
  OMNI_CHECK(D <= kThreads * 8 * kMaxChunks, OMNI_UNSUPPORTED, "add_norm bwd: dim too large");
  if (int rc = check_rows(p->xres, M, D, "xres")) return rc;
  if (int rc = check_rows(p->dy, M, D, "dy")) return rc;
  if (int rc = check_rows(p->dresidual_in, M, D, "dresidual_in", false)) return rc;
  if (int rc = check_rows(p->dx, M, D, "dx")) return rc;
  if (int rc = check_rows(p->dresidual, M, D, "dresidual", false)) return rc;
  if (int rc = check_vecparam(p->weight, D, "weight", true)) return rc;
  OMNI_CHECK(present(p->rstd) && f32_contig(p->rstd, M), OMNI_BAD_SHAPE, "rstd must be contiguous fp32 (rows)");
  if (!p->is_rms_norm) OMNI_CHECK(present(p->mean) && f32_contig(p->mean, M), OMNI_BAD_SHAPE, "mean required");
  OMNI_CHECK(present(p->dw_part) && p->dw_part.ndim == 2 && p->dw_part.shape[1] == D && p->dw_part.shape[0] >= 1 &&
                 f32_contig(p->dw_part, p->dw_part.shape[0] * D),
             OMNI_BAD_SHAPE, "dw_part must be contiguous fp32 (nparts, dim)");
  const int64_t nparts = p->dw_part.shape[0];
  if (present(p->db_part))
    OMNI_CHECK(f32_contig(p->db_part, nparts * D), OMNI_BAD_SHAPE, "db_part must be contiguous fp32 (nparts, dim)");
  if (D == 0) return OMNI_OK;
  AddNormArgs a{};
  a.x = view2(p->xres); a.dy = view2(p->dy); a.dres_in = view2(p->dresidual_in); a.dx = view2(p->dx);
  a.dres_out = view2(p->dresidual);
  a.w = p->weight.data; a.w_dtype = p->weight.dtype;
  a.rstd = static_cast<float*>(p->rstd.data); a.mean = static_cast<float*>(p->mean.data);
  a.dw_part = static_cast<float*>(p->dw_part.data); a.db_part = static_cast<float*>(p->db_part.data);
  a.M = (int)M; a.D = (int)D; a.eps = p->eps; a.is_rms = p->is_rms_norm;
  const bool vec = rows_vec_ok(p->xres, D) && rows_vec_ok(p->dy, D) && rows_vec_ok(p->dresidual_in, D) &&
                   rows_vec_ok(p->dx, D) && rows_vec_ok(p->dresidual, D) && rows_vec_ok(p->weight, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nch = (int)(((D) + kThreads * 8 - 1) / (kThreads * 8));
  // fast path: RMSNorm without bias, fp32 residual stream, 16-bit dy / dx of one type, fp32 (or no) residual gradients
  if (vec && a.is_rms && !present(p->db_part) && (D == 1024 || D == 2048 || D == 4096) && p->xres.dtype == OMNI_F32 &&
      p->dy.dtype != OMNI_F32 && p->dx.dtype == p->dy.dtype && (!present(p->dresidual_in) || p->dresidual_in.dtype == OMNI_F32) &&
      (!present(p->dresidual) || p->dresidual.dtype == OMNI_F32)) {
    if (D == 4096) {
      if (p->dy.dtype == OMNI_BF16) add_rms_bwd_fast_kernel<__nv_bfloat16, 8><<<(unsigned)nparts, 512, 0, s>>>(a);
      else add_rms_bwd_fast_kernel<__half, 8><<<(unsigned)nparts, 512, 0, s>>>(a);
    } else {
      const unsigned threads = (unsigned)(D / 4);
      if (p->dy.dtype == OMNI_BF16) add_rms_bwd_fast_kernel<__nv_bfloat16, 4><<<(unsigned)nparts, threads, 0, s>>>(a);
      else add_rms_bwd_fast_kernel<__half, 4><<<(unsigned)nparts, threads, 0, s>>>(a);
    }
    OMNI_CUDA_LAUNCH_CHECK("add_rms_bwd_fast_kernel");
    return OMNI_OK;
  }
#define OMNI_LAUNCH_NCH(V, N) add_norm_bwd_kernel<V, N><<<(unsigned)nparts, kThreads, 0, s>>>(a)
  if (vec) { if (nch <= 1) OMNI_LAUNCH_NCH(true, 1); else if (nch == 2) OMNI_LAUNCH_NCH(true, 2); else OMNI_LAUNCH_NCH(true, 4); }
  else { if (nch <= 1) OMNI_LAUNCH_NCH(false, 1); else if (nch == 2) OMNI_LAUNCH_NCH(false, 2); else OMNI_LAUNCH_NCH(false, 4); }
#undef OMNI_LAUNCH_NCH
  OMNI_CUDA_LAUNCH_CHECK("add_norm_bwd_kernel");
  return OMNI_OK;
}
