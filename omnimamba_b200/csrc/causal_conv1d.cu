// Depthwise causal conv1d: forward, backward, decode update.  HBM-bound streaming kernels.
//
// Replaces causal_conv1d_fn / causal_conv1d_update of causal-conv1d==1.4.0
// (/root/reference/requirements.txt:12) as reached from Mamba2.forward / Mamba2.step
// (/root/reference/models/stage2/block.py:117).  Arithmetic: SURVEY.md Appendix A.2.
//
// Layout: the hot call passes xBC as a transposed slice of zxbcdt, i.e. (B, D, L) with
// stride(D) == 1 ("channel-last", row pitch 8512 elements).  Each thread owns VEC consecutive
// channels (one 16-byte vector) and walks TL consecutive tokens with the W-1 previous inputs kept
// in registers, so every x / out element crosses HBM exactly once (+ (W-1)/TL halo re-reads that
// hit L2).  A warp covers 32*VEC contiguous channels = 512 B per token row: fully coalesced.
#include "common.cuh"

namespace omni {
namespace {

constexpr int kMaxW = 4;
constexpr int kTL = 64;       // tokens per thread
constexpr int kSegs = 4;      // token segments (warps) per block

struct ConvArgs {
  const void* x; const void* w; const void* bias; const int* seq_idx; const void* init;
  void* out; void* fin;
  const void* dout; void* dx; float* dw; float* db; void* dinit;
  int64_t xs_b, xs_d, xs_l, os_b, os_d, os_l;     // x / out (fwd) strides
  int64_t gs_b, gs_d, gs_l, ds_b, ds_d, ds_l;     // dout / dx strides (bwd)
  int64_t ws_d, ws_w, bs_d;
  int64_t is_b, is_d, is_k, fs_b, fs_d, fs_k, dis_b, dis_d, dis_k;
  int64_t ss_b, ss_l;
  int B, D, L;
  int w_dtype, b_dtype;
  int silu;
};

template <typename T, int VEC> __device__ __forceinline__ void ldv(const T* p, float (&o)[VEC]) {
  if constexpr (VEC == 1) o[0] = to_f<T>(*p);
  else load_vec<T, VEC>(p, o);
}
template <typename T, int VEC> __device__ __forceinline__ void stv(T* p, const float (&o)[VEC]) {
  if constexpr (VEC == 1) *p = from_f<T>(o[0]);
  else store_vec<T, VEC>(p, o);
}

// x at token t (t may be negative => initial state / zero) for this thread's channels
template <typename T, int VEC, int W>
__device__ __forceinline__ void load_tok(const ConvArgs& a, int b, int d0, int t, float (&o)[VEC]) {
  if (t >= 0) {
    if (t < a.L) ldv<T, VEC>(static_cast<const T*>(a.x) + b * a.xs_b + d0 * a.xs_d + t * a.xs_l, o);
    else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = 0.f;
    }
  } else if (a.init != nullptr) {
    const T* ip = static_cast<const T*>(a.init) + b * a.is_b + (int64_t)(W - 1 + t) * a.is_k;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = to_f<T>(ip[(d0 + v) * a.is_d]);
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = 0.f;
  }
}

template <typename T, int VEC, int W, bool SEQ>
__global__ void __launch_bounds__(32 * kSegs) conv1d_fwd_kernel(ConvArgs a) {
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kSegs + threadIdx.y) * kTL;
  if (d0 >= a.D) return;

  // final states: the last W-1 inputs (done by the threads of the first segment)
  if (a.fin != nullptr && t0 == 0) {
    T* fp = static_cast<T*>(a.fin) + b * a.fs_b;
#pragma unroll
    for (int k = 0; k < W - 1; ++k) {
      float xv[VEC];
      load_tok<T, VEC, W>(a, b, d0, a.L - (W - 1) + k, xv);
#pragma unroll
      for (int v = 0; v < VEC; ++v) fp[(d0 + v) * a.fs_d + k * a.fs_k] = from_f<T>(xv[v]);
    }
  }
  if (t0 >= a.L) return;

  float wgt[W][VEC], bia[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
#pragma unroll
    for (int k = 0; k < W; ++k) wgt[k][v] = ld_any(a.w, a.w_dtype, (d0 + v) * a.ws_d + k * a.ws_w);
    bia[v] = a.bias ? ld_any(a.bias, a.b_dtype, (d0 + v) * a.bs_d) : 0.f;
  }
  float win[W][VEC];  // win[k] = x[t - (W-1) + k]; win[W-1] is the current token
  int sid[W];
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
    load_tok<T, VEC, W>(a, b, d0, t0 - (W - 1) + k, win[k + 1]);
    const int tt = t0 - (W - 1) + k;
    sid[k + 1] = (SEQ && tt >= 0) ? a.seq_idx[b * a.ss_b + tt * a.ss_l] : 0;
  }
  const int tend = min(t0 + kTL, a.L);
  const T* xp = static_cast<const T*>(a.x) + b * a.xs_b + d0 * a.xs_d;
  T* op = static_cast<T*>(a.out) + b * a.os_b + d0 * a.os_d;
#pragma unroll 8
  for (int t = t0; t < tend; ++t) {
#pragma unroll
    for (int k = 0; k < W - 1; ++k) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) win[k][v] = win[k + 1][v];
      sid[k] = sid[k + 1];
    }
    ldv<T, VEC>(xp + t * a.xs_l, win[W - 1]);
    sid[W - 1] = SEQ ? a.seq_idx[b * a.ss_b + t * a.ss_l] : 0;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = bia[v];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      if constexpr (SEQ) {
        const bool same = (sid[k] == sid[W - 1]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += same ? wgt[k][v] * win[k][v] : 0.f;
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wgt[k][v], win[k][v], acc[v]);
      }
    }
    if (a.silu) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = silu_f(acc[v]);
    }
    stv<T, VEC>(op + t * a.os_l, acc);
  }
}

// ---- fast path of the forward: 16-bit channel-last I/O, width 4, no seq_idx / initial / final states ------------------
// (the call Mamba2.forward makes on the training and prefill paths).  The generic kernel above needs 104 registers at 8
// channels per thread, which limits it to 16 warps per SM and leaves it latency-bound (ncu: 23 % occupancy, 10 warps stalled
// on loads per issue).  Here a thread owns 4 channels (8-byte vectors; a warp still covers 256 contiguous bytes per token),
// walks 64 tokens with the three previous inputs in registers and issues the loads of 8 tokens before touching them.
constexpr int kTLF = 64;
template <typename T> __device__ __forceinline__ void unpack4(uint2 r, float (&o)[4]);
template <> __device__ __forceinline__ void unpack4<__nv_bfloat16>(uint2 r, float (&o)[4]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack4<__half>(uint2 r, float (&o)[4]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T> __device__ __forceinline__ uint2 pack4(const float (&o)[4]);
template <> __device__ __forceinline__ uint2 pack4<__nv_bfloat16>(const float (&o)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <> __device__ __forceinline__ uint2 pack4<__half>(const float (&o)[4]) {
  const __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 operations per issue slot, each lane an ordinary IEEE operation).
// The 16-bit fast kernels below are bound by instruction issue before HBM (forward 15 instructions per element at 74 % of the
// roofline, backward 35 at 49 %), so their arithmetic runs on PAIRS of channels, the sigmoid is ex2.approx.ftz + rcp.approx.ftz
// (no denormal fix-up code) and addresses are 32-bit multiples of the row pitch added to one 64-bit base per batch.
__device__ __forceinline__ float2 fma2p(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 mul2p(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 add2p(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float ex2_ftz(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rcp_ftz(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
// d/dv [v sigmoid(v)] = s (1 + v (1 - s)) for a pair of channels
__device__ __forceinline__ float2 dsilu2(float2 v) {
  const float2 one = make_float2(1.f, 1.f);
  const float2 nl = mul2p(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = add2p(make_float2(ex2_ftz(nl.x), ex2_ftz(nl.y)), one);
  const float2 s = make_float2(rcp_ftz(d.x), rcp_ftz(d.y));
  const float2 oms = fma2p(s, make_float2(-1.f, -1.f), one);
  return mul2p(s, fma2p(v, oms, one));
}
template <typename T> __device__ __forceinline__ void unpack22(uint2 r, float2 (&o)[2]);
template <> __device__ __forceinline__ void unpack22<__nv_bfloat16>(uint2 r, float2 (&o)[2]) {
  o[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
  o[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
}
template <> __device__ __forceinline__ void unpack22<__half>(uint2 r, float2 (&o)[2]) {
  o[0] = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  o[1] = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
}
template <typename T> __device__ __forceinline__ uint2 pack22(const float2 (&o)[2]);
template <> __device__ __forceinline__ uint2 pack22<__nv_bfloat16>(const float2 (&o)[2]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(o[0].x, o[0].y), b = __floats2bfloat162_rn(o[1].x, o[1].y);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <> __device__ __forceinline__ uint2 pack22<__half>(const float2 (&o)[2]) {
  const __half2 a = __floats2half2_rn(o[0].x, o[0].y), b = __floats2half2_rn(o[1].x, o[1].y);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// v sigmoid(v) for a pair of channels
__device__ __forceinline__ float2 silu2(float2 v) {
  const float2 nl = mul2p(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = add2p(make_float2(ex2_ftz(nl.x), ex2_ftz(nl.y)), make_float2(1.f, 1.f));
  return mul2p(v, make_float2(rcp_ftz(d.x), rcp_ftz(d.y)));
}

template <typename T, bool SILU, bool FIXED>   // FIXED: kTLF tokens per thread as a compile-time count (3 % faster at L = 4096)
__global__ void __launch_bounds__(32 * kSegs, 6) conv1d_fwd_fast_kernel(ConvArgs a, int tl_) {
  const int tl = FIXED ? kTLF : tl_;   // tl <= kTLF tokens per thread (even segments, see the launcher)
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kSegs + threadIdx.y) * tl;
  if (d0 >= a.D || t0 >= a.L) return;
  float2 w[4][2], bia[2];  // w[k][pair]: tap k (k = 3 multiplies the current token) of channels d0 + 2 pair, + 1
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    const int64_t da = d0 + 2 * pr, db = d0 + 2 * pr + 1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      w[k][pr] = make_float2(ld_any(a.w, a.w_dtype, da * a.ws_d + k * a.ws_w), ld_any(a.w, a.w_dtype, db * a.ws_d + k * a.ws_w));
    bia[pr] = a.bias ? make_float2(ld_any(a.bias, a.b_dtype, da * a.bs_d), ld_any(a.bias, a.b_dtype, db * a.bs_d)) : make_float2(0.f, 0.f);
  }
  const int xsl = (int)a.xs_l, osl = (int)a.os_l;   // (row pitches fit 32 bits: checked by the host)
  const T* xp = static_cast<const T*>(a.x) + b * a.xs_b + d0 + (int64_t)t0 * xsl;
  T* op = static_cast<T*>(a.out) + b * a.os_b + d0 + (int64_t)t0 * osl;
  float2 p1[2], p2[2], p3[2];  // x[t-1], x[t-2], x[t-3]
  {
    const uint2 z2 = make_uint2(0u, 0u);
    unpack22<T>(t0 >= 1 ? __ldg(reinterpret_cast<const uint2*>(xp - xsl)) : z2, p1);
    unpack22<T>(t0 >= 2 ? __ldg(reinterpret_cast<const uint2*>(xp - 2 * xsl)) : z2, p2);
    unpack22<T>(t0 >= 3 ? __ldg(reinterpret_cast<const uint2*>(xp - 3 * xsl)) : z2, p3);
  }
  auto token = [&](uint2 raw, T* dst) {
    float2 c[2], acc[2];
    unpack22<T>(raw, c);
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      acc[pr] = fma2p(w[3][pr], c[pr], fma2p(w[2][pr], p1[pr], fma2p(w[1][pr], p2[pr], fma2p(w[0][pr], p3[pr], bia[pr]))));
      if (SILU) acc[pr] = silu2(acc[pr]);
      p3[pr] = p2[pr]; p2[pr] = p1[pr]; p1[pr] = c[pr];
    }
    *reinterpret_cast<uint2*>(dst) = pack22<T>(acc);
  };
  const int n = min(tl, a.L - t0);
  int i = 0;
#pragma unroll 1
  for (; i + 16 <= n; i += 16) {  // 16 tokens (128 B per thread) in flight: ~100 KB per SM at 24 resident warps
    const T* px = xp + (int64_t)i * xsl;
    T* po = op + (int64_t)i * osl;
    uint2 raw[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) raw[u] = __ldg(reinterpret_cast<const uint2*>(px + u * xsl));
#pragma unroll
    for (int u = 0; u < 16; ++u) token(raw[u], po + u * osl);
  }
#pragma unroll 1
  for (; i + 8 <= n; i += 8) {
    const T* px = xp + (int64_t)i * xsl;
    T* po = op + (int64_t)i * osl;
    uint2 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = __ldg(reinterpret_cast<const uint2*>(px + u * xsl));
#pragma unroll
    for (int u = 0; u < 8; ++u) token(raw[u], po + u * osl);
  }
#pragma unroll 1
  for (; i < n; ++i) token(__ldg(reinterpret_cast<const uint2*>(xp + (int64_t)i * xsl)), op + (int64_t)i * osl);
}

template <typename T>
bool try_launch_fwd_fast(const ConvArgs& a, int W, const omni_tensor_t& x, const omni_tensor_t& o, cudaStream_t s) {
  if constexpr (sizeof(T) != 2) return false;
  else {
    if (W != 4 || a.seq_idx || a.init || a.fin || a.xs_d != 1 || a.os_d != 1 || a.D % 4 != 0 || a.L == 0) return false;
    auto ok8 = [](const omni_tensor_t& t) {  // 8-byte vectors along the channel dim
      return reinterpret_cast<uintptr_t>(t.data) % 8 == 0 && (t.shape[0] <= 1 || t.stride[0] % 4 == 0) &&
             (t.shape[2] <= 1 || t.stride[2] % 4 == 0);
    };
    if (!ok8(x) || !ok8(o)) return false;
    auto pitch32 = [](int64_t st) { return st >= 0 && st * 16 < (int64_t)0x7fffffff; };   // (u * pitch, u < 16, is formed in 32 bits)
    if (!pitch32(a.xs_l) || !pitch32(a.os_l)) return false;
    // even segments, as in the backward: L = 329 (the stage-1 training length) is 8 x 42 tokens instead of 5 x 64 + 9 with two
    // idle warps in the second block; a 72-token prefill 4 x 18 instead of 64 + 8
    const int nblk = (a.L + kSegs * kTLF - 1) / (kSegs * kTLF);
    const int tl = (a.L + kSegs * nblk - 1) / (kSegs * nblk);
    dim3 block(32, kSegs), grid((a.D + 127) / 128, (a.L + kSegs * tl - 1) / (kSegs * tl), a.B);
    if (tl == kTLF) {
      if (a.silu) conv1d_fwd_fast_kernel<T, true, true><<<grid, block, 0, s>>>(a, tl);
      else conv1d_fwd_fast_kernel<T, false, true><<<grid, block, 0, s>>>(a, tl);
    } else {
      if (a.silu) conv1d_fwd_fast_kernel<T, true, false><<<grid, block, 0, s>>>(a, tl);
      else conv1d_fwd_fast_kernel<T, false, false><<<grid, block, 0, s>>>(a, tl);
    }
    return true;
  }
}

// Backward.  dc[t] = dout[t] * act'(c[t]);  dx[s] = sum_w weight[w] dc[s + W-1 - w];
// dweight[w] = sum_t dc[t] x[t - (W-1) + w];  dbias = sum_t dc[t].
template <typename T, int VEC, int W>
__global__ void __launch_bounds__(32 * kSegs) conv1d_bwd_kernel(ConvArgs a) {
  __shared__ float red[kSegs][32][VEC * (W + 1) + 1];
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kSegs + threadIdx.y) * kTL;
  const bool active = d0 < a.D && t0 < a.L;

  float dwa[W][VEC], dba[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    dba[v] = 0.f;
#pragma unroll
    for (int k = 0; k < W; ++k) dwa[k][v] = 0.f;
  }
  if (active) {
    float wgt[W][VEC], bia[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
      for (int k = 0; k < W; ++k) wgt[k][v] = ld_any(a.w, a.w_dtype, (d0 + v) * a.ws_d + k * a.ws_w);
      bia[v] = a.bias ? ld_any(a.bias, a.b_dtype, (d0 + v) * a.bs_d) : 0.f;
    }
    float win[W][VEC];   // x[t-(W-1) .. t]
    float dcw[W][VEC];   // dcw[k] = dc[t - k]
    float dia[W][VEC];   // dinitial_states accumulators (first segment only)
#pragma unroll
    for (int k = 0; k < W; ++k) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) dia[k][v] = 0.f;
    }
    int sid[W];          // seq id of x window
    int dsid[W];         // seq id of dc window (dsid[k] = seq[t-k])
#pragma unroll
    for (int k = 0; k < W; ++k) {
      dsid[k] = 0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) dcw[k][v] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < W - 1; ++k) {
      load_tok<T, VEC, W>(a, b, d0, t0 - (W - 1) + k, win[k + 1]);
      const int tt = t0 - (W - 1) + k;
      sid[k + 1] = (a.seq_idx && tt >= 0) ? a.seq_idx[b * a.ss_b + tt * a.ss_l] : 0;
    }
    const int own_end = min(t0 + kTL, a.L);
    const int tlast = own_end + (W - 1);  // exclusive; dc beyond L is zero
    const T* gp = static_cast<const T*>(a.dout) + b * a.gs_b + d0 * a.gs_d;
    T* dxp = static_cast<T*>(a.dx) + b * a.ds_b + d0 * a.ds_d;
    for (int t = t0; t < tlast; ++t) {
#pragma unroll
      for (int k = 0; k < W - 1; ++k) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) win[k][v] = win[k + 1][v];
        sid[k] = sid[k + 1];
      }
#pragma unroll
      for (int k = W - 1; k > 0; --k) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) dcw[k][v] = dcw[k - 1][v];
        dsid[k] = dsid[k - 1];
      }
      float g[VEC];
      if (t < a.L) {
        load_tok<T, VEC, W>(a, b, d0, t, win[W - 1]);
        sid[W - 1] = a.seq_idx ? a.seq_idx[b * a.ss_b + t * a.ss_l] : 0;
        ldv<T, VEC>(gp + t * a.gs_l, g);
        if (a.silu) {
          float c[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) c[v] = bia[v];
#pragma unroll
          for (int k = 0; k < W; ++k) {
            const bool same = (sid[k] == sid[W - 1]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) c[v] += same ? wgt[k][v] * win[k][v] : 0.f;
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) g[v] *= dsilu_f(c[v]);
        }
        if (t < own_end) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) dba[v] += g[v];
#pragma unroll
          for (int k = 0; k < W; ++k) {
            const bool same = (sid[k] == sid[W - 1]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) dwa[k][v] += same ? g[v] * win[k][v] : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { g[v] = 0.f; win[W - 1][v] = 0.f; }
        sid[W - 1] = -1;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) dcw[0][v] = g[v];
      dsid[0] = sid[W - 1];
      // dinitial_states: "token" s = kk-(W-1) < 0 reached out[t] through tap kk - t
      if (a.dinit != nullptr && t0 == 0 && t < W - 1) {
#pragma unroll
        for (int kk = 0; kk < W - 1; ++kk) {
          const int wtap = kk - t;
          if (wtap >= 0) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) dia[kk][v] += wgt[wtap][v] * g[v];
          }
        }
      }
      const int s = t - (W - 1);
      if (s >= t0 && s < own_end) {
        float r[VEC];
        const int ssid = a.seq_idx ? a.seq_idx[b * a.ss_b + s * a.ss_l] : 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) {
          // x[s] reached out[t-k] (t = s+W-1) through tap k
          const bool same = !a.seq_idx || (dsid[k] == ssid);
#pragma unroll
          for (int v = 0; v < VEC; ++v) r[v] += same ? wgt[k][v] * dcw[k][v] : 0.f;
        }
        stv<T, VEC>(dxp + s * a.ds_l, r);
      }
    }
    if (a.dinit != nullptr && t0 == 0) {
      T* dip = static_cast<T*>(a.dinit) + b * a.dis_b;
#pragma unroll
      for (int kk = 0; kk < W - 1; ++kk) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) dip[(d0 + v) * a.dis_d + kk * a.dis_k] = from_f<T>(dia[kk][v]);
      }
    }
  }
  // reduce dweight / dbias over the block's token segments, then one atomic per (channel, tap)
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
#pragma unroll
    for (int k = 0; k < W; ++k) red[threadIdx.y][threadIdx.x][v * (W + 1) + k] = dwa[k][v];
    red[threadIdx.y][threadIdx.x][v * (W + 1) + W] = dba[v];
  }
  __syncthreads();
  if (threadIdx.y == 0 && d0 < a.D) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
      for (int k = 0; k <= W; ++k) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < kSegs; ++g) s += red[g][threadIdx.x][v * (W + 1) + k];
        if (k < W) atomicAdd(a.dw + (int64_t)(d0 + v) * W + k, s);
        else if (a.db) atomicAdd(a.db + d0 + v, s);
      }
    }
  }
}

// ---- fast path of the backward: same configuration as conv1d_fwd_fast_kernel ------------------------------------------
// A thread owns 4 channels and `tl` consecutive tokens.  It walks t = t0 .. t0 + tl + 2 keeping the last four inputs and the
// last four dc = dout * act'(c) in registers: c[t] is recomputed from the window, dweight / dbias accumulate over the
// thread's own tokens, dx[t-3] = sum_k w[k] dc[t-k] leaves as soon as its four dc are known (the three tokens past the
// thread's range are re-read: L2 hits).  Loads of 8 tokens (x and dout) are issued before any of them is used.
// The kernel is bound by instruction issue, not by HBM (35 instructions per element in the first version: 62 % issue
// utilisation at 49 % of the HBM roofline), so the arithmetic runs on PAIRS of channels with the packed fp32 instructions
// (fma / mul / add .f32x2: two lanes per issue slot, each lane an ordinary IEEE operation), the sigmoid is ex2.approx.ftz +
// rcp.approx.ftz (no denormal fix-up code), addresses are 32-bit multiples of the row pitch added to one 64-bit base per
// batch, and the tokens that need no edge predicate (own, with a dx row to store) run in a predicate-free loop.
// NP channel pairs per thread (2: 8-byte vectors, 1: 4-byte), CB tokens per batch of loads, MINB resident blocks per SM.
// Measured (profiles/r2c_conv_bwd_channels_ab.txt, (16, 4096)): 4 channels x 8 tokens x 4 blocks 0.383 ms (the default); 2 channels
// per thread - half the register state, more warps - 0.405 ms at 6 blocks per SM and worse beyond (spills at 72 / 64 registers,
// 16-token batches 0.47 - 0.55 ms): fewer bytes in flight per warp cost more than the extra warps bring.
template <int NP> struct RawVec;
template <> struct RawVec<2> { using type = uint2; };
template <> struct RawVec<1> { using type = uint32_t; };
template <typename T, int NP> __device__ __forceinline__ void unpack_np(typename RawVec<NP>::type r, float2 (&o)[NP]) {
  if constexpr (NP == 2) unpack22<T>(r, o);
  else { uint2 t = make_uint2(r, 0u); float2 q[2]; unpack22<T>(t, q); o[0] = q[0]; }
}
template <typename T, int NP> __device__ __forceinline__ typename RawVec<NP>::type pack_np(const float2 (&o)[NP]) {
  if constexpr (NP == 2) return pack22<T>(o);
  else { const float2 q[2] = {o[0], make_float2(0.f, 0.f)}; return pack22<T>(q).x; }
}
template <int NP> __device__ __forceinline__ typename RawVec<NP>::type raw_zero() {
  if constexpr (NP == 2) return make_uint2(0u, 0u); else return 0u;
}
template <typename T, bool SILU, int NP, int CB, int MINB>
__global__ void __launch_bounds__(32 * kSegs, MINB) conv1d_bwd_fast_kernel(ConvArgs a, int tl, const T* __restrict__ xbase, const T* __restrict__ gbase, T* __restrict__ dxbase) {
  // (x / dout / dx also arrive as __restrict__ parameters: without the no-alias guarantee every load of the next batch is kept
  // behind the dx stores of the current one, i.e. at the end of the loop body, and its latency is exposed)
  using Raw = typename RawVec<NP>::type;
  constexpr int kCB = CB, CH = 2 * NP;
  __shared__ float red[kSegs][32][5 * CH + 1];
  const int d0 = (blockIdx.x * 32 + threadIdx.x) * CH;
  const int b = blockIdx.z;
  const int t0 = (blockIdx.y * kSegs + threadIdx.y) * tl;
  const bool active = d0 < a.D && t0 < a.L;
  const float2 zero2 = make_float2(0.f, 0.f);
  float2 dwa[4][NP], dba[NP];   // [tap][channel pair]
#pragma unroll
  for (int pr = 0; pr < NP; ++pr) {
    dba[pr] = zero2;
#pragma unroll
    for (int k = 0; k < 4; ++k) dwa[k][pr] = zero2;
  }
  if (active) {
    float2 w[4][NP], bia[NP];
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) {
      const int64_t da = d0 + 2 * pr, db = d0 + 2 * pr + 1;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k][pr] = make_float2(ld_any(a.w, a.w_dtype, da * a.ws_d + k * a.ws_w), ld_any(a.w, a.w_dtype, db * a.ws_d + k * a.ws_w));
      bia[pr] = a.bias ? make_float2(ld_any(a.bias, a.b_dtype, da * a.bs_d), ld_any(a.bias, a.b_dtype, db * a.bs_d)) : zero2;
    }
    const T* __restrict__ xp = xbase + b * a.xs_b + d0;
    const T* __restrict__ gp = gbase + b * a.gs_b + d0;
    T* __restrict__ dxp = dxbase + b * a.ds_b + d0;
    const int xsl = (int)a.xs_l, gsl = (int)a.gs_l, dsl = (int)a.ds_l;   // (row pitches fit 32 bits: checked by the host)
    const Raw z2 = raw_zero<NP>();
    float2 p1[NP], p2[NP], p3[NP];     // x[t-1], x[t-2], x[t-3]
    float2 g1[NP], g2[NP], g3[NP];     // dc[t-1], dc[t-2], dc[t-3]
    unpack_np<T, NP>(t0 >= 1 ? __ldg(reinterpret_cast<const Raw*>(xp + (int64_t)(t0 - 1) * xsl)) : z2, p1);
    unpack_np<T, NP>(t0 >= 2 ? __ldg(reinterpret_cast<const Raw*>(xp + (int64_t)(t0 - 2) * xsl)) : z2, p2);
    unpack_np<T, NP>(t0 >= 3 ? __ldg(reinterpret_cast<const Raw*>(xp + (int64_t)(t0 - 3) * xsl)) : z2, p3);
#pragma unroll
    for (int pr = 0; pr < NP; ++pr) g1[pr] = g2[pr] = g3[pr] = zero2;
    const int own_end = min(t0 + tl, a.L);
    const int tlast = own_end + 3;   // exclusive; dc beyond L is zero
    // one token: `own` = its dc feeds this thread's dweight / dbias sums, `store` = dx[t - 3] is a row of this thread
    auto token = [&](Raw rx, Raw rg, bool own, bool store, T* dst) {
      float2 c[NP], g[NP], r[NP];
      unpack_np<T, NP>(rx, c);
      unpack_np<T, NP>(rg, g);
#pragma unroll
      for (int pr = 0; pr < NP; ++pr) {
        if (SILU) {
          const float2 pre = fma2p(w[3][pr], c[pr], fma2p(w[2][pr], p1[pr], fma2p(w[1][pr], p2[pr], fma2p(w[0][pr], p3[pr], bia[pr]))));
          g[pr] = mul2p(g[pr], dsilu2(pre));
        }
        if (own) {
          dba[pr] = add2p(dba[pr], g[pr]);
          dwa[3][pr] = fma2p(g[pr], c[pr], dwa[3][pr]);
          dwa[2][pr] = fma2p(g[pr], p1[pr], dwa[2][pr]);
          dwa[1][pr] = fma2p(g[pr], p2[pr], dwa[1][pr]);
          dwa[0][pr] = fma2p(g[pr], p3[pr], dwa[0][pr]);
        }
        r[pr] = fma2p(w[0][pr], g[pr], fma2p(w[1][pr], g1[pr], fma2p(w[2][pr], g2[pr], mul2p(w[3][pr], g3[pr]))));
        p3[pr] = p2[pr]; p2[pr] = p1[pr]; p1[pr] = c[pr];
        g3[pr] = g2[pr]; g2[pr] = g1[pr]; g1[pr] = g[pr];
      }
      if (store) *reinterpret_cast<Raw*>(dst) = pack_np<T, NP>(r);
    };
    auto edge_token = [&](int t) {   // any token of [t0, tlast): loads guarded by L, flags from its position
      const bool in = t < a.L;
      token(in ? __ldg(reinterpret_cast<const Raw*>(xp + (int64_t)t * xsl)) : z2,
            in ? __ldg(reinterpret_cast<const Raw*>(gp + (int64_t)t * gsl)) : z2, t < own_end, t - 3 >= t0,
            dxp + (int64_t)(t - 3) * dsl);
    };
    int t = t0;
    const int head_end = min(t0 + 3, tlast);
#pragma unroll 1
    for (; t < head_end; ++t) edge_token(t);   // (the first three tokens have no dx row of this thread behind them)
    // own tokens with a dx row behind them: no predicates.  The 16 loads of the NEXT batch are issued in the same loop body
    // (ptxas places loads that nothing in the body consumes at its end whatever the source order - half-batch rotation with a
    // __syncwarp as a scheduling fence was measured 25 % slower, more registers at 12 warps per SM 5 % slower).
    if (t + kCB <= own_end) {
      Raw rx[kCB], rg[kCB];
      {
        const T* px = xp + (int64_t)t * xsl;
        const T* pg = gp + (int64_t)t * gsl;
#pragma unroll
        for (int u = 0; u < kCB; ++u) {
          rx[u] = __ldg(reinterpret_cast<const Raw*>(px + u * xsl));
          rg[u] = __ldg(reinterpret_cast<const Raw*>(pg + u * gsl));
        }
      }
#pragma unroll 1
      for (; t + 2 * kCB <= own_end; t += kCB) {
        const T* px = xp + (int64_t)(t + kCB) * xsl;
        const T* pg = gp + (int64_t)(t + kCB) * gsl;
        T* pd = dxp + (int64_t)(t - 3) * dsl;
#pragma unroll
        for (int u = 0; u < kCB; ++u) {
          token(rx[u], rg[u], true, true, pd + u * dsl);
          rx[u] = __ldg(reinterpret_cast<const Raw*>(px + u * xsl));
          rg[u] = __ldg(reinterpret_cast<const Raw*>(pg + u * gsl));
        }
      }
      {   // the last full batch
        T* pd = dxp + (int64_t)(t - 3) * dsl;
#pragma unroll
        for (int u = 0; u < kCB; ++u) token(rx[u], rg[u], true, true, pd + u * dsl);
        t += kCB;
      }
    }
#pragma unroll 1
    for (; t < tlast; ++t) edge_token(t);   // (edge tokens in predicated batches of 8: 3 % faster at L = 329, but the extra
                                            //  live state spills inside the main loop: 6 % slower at L = 4096)
  }
  // reduce dweight / dbias over the block's token segments, then one atomic per (channel, tap)
#pragma unroll
  for (int pr = 0; pr < NP; ++pr) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      red[threadIdx.y][threadIdx.x][(2 * pr) * 5 + k] = dwa[k][pr].x;
      red[threadIdx.y][threadIdx.x][(2 * pr + 1) * 5 + k] = dwa[k][pr].y;
    }
    red[threadIdx.y][threadIdx.x][(2 * pr) * 5 + 4] = dba[pr].x;
    red[threadIdx.y][threadIdx.x][(2 * pr + 1) * 5 + 4] = dba[pr].y;
  }
  __syncthreads();
  if (threadIdx.y == 0 && d0 < a.D) {
#pragma unroll
    for (int v = 0; v < CH; ++v) {
#pragma unroll
      for (int k = 0; k <= 4; ++k) {
        float sum = 0.f;
#pragma unroll
        for (int gsg = 0; gsg < kSegs; ++gsg) sum += red[gsg][threadIdx.x][v * 5 + k];
        if (k < 4) atomicAdd(a.dw + (int64_t)(d0 + v) * 4 + k, sum);
        else if (a.db) atomicAdd(a.db + d0 + v, sum);
      }
    }
  }
}

template <typename T>
bool try_launch_bwd_fast(const ConvArgs& a, int W, const omni_tensor_t& x, const omni_tensor_t& g, const omni_tensor_t& dx,
                         cudaStream_t s) {
  if constexpr (sizeof(T) != 2) return false;
  else {
    if (W != 4 || a.seq_idx || a.init || a.dinit || a.xs_d != 1 || a.gs_d != 1 || a.ds_d != 1 || a.D % 4 != 0) return false;
    auto ok8 = [](const omni_tensor_t& t) {
      return reinterpret_cast<uintptr_t>(t.data) % 8 == 0 && (t.shape[0] <= 1 || t.stride[0] % 4 == 0) &&
             (t.shape[2] <= 1 || t.stride[2] % 4 == 0);
    };
    if (!ok8(x) || !ok8(g) || !ok8(dx)) return false;
    auto pitch32 = [](int64_t st) { return st >= 0 && st * 16 < (int64_t)0x7fffffff; };   // (u * pitch, u < 8, is formed in 32 bits)
    if (!pitch32(a.xs_l) || !pitch32(a.gs_l) || !pitch32(a.ds_l)) return false;
    // tokens per thread: 256 when the sequences are long enough to keep every segment of a block busy with it and there is
    // enough work to fill the GPU (4x fewer partial-sum atomics), else 64 (L = 329, the stage-1 training length, with 256:
    // one segment of 256 tokens, one of 73 and idle ones - the kernel ran at 20 % of the HBM roofline there)
    const int64_t cols = (a.D + 127) / 128;
    int tl = (a.L >= kSegs * 256 && cols * a.B * ((a.L + kSegs * 256 - 1) / (kSegs * 256)) >= 4 * (int64_t)sm_count()) ? 256 : 64;
    if (tl == 64) {  // even segments: every warp of every block gets the same share (L = 329: 8 x 42 instead of 5 x 64 + 9)
      const int nblk = (a.L + kSegs * 64 - 1) / (kSegs * 64);
      tl = (a.L + kSegs * nblk - 1) / (kSegs * nblk);
    }
#ifndef OMNI_CONV_BWD_NP
#define OMNI_CONV_BWD_NP 2
#endif
#ifndef OMNI_CONV_BWD_CB
#define OMNI_CONV_BWD_CB 8
#endif
#ifndef OMNI_CONV_BWD_MINB
#define OMNI_CONV_BWD_MINB 4
#endif
    constexpr int NP = OMNI_CONV_BWD_NP, CB = OMNI_CONV_BWD_CB, MINB = OMNI_CONV_BWD_MINB;
    const int64_t colsn = (a.D + 64 * NP - 1) / (64 * NP);
    dim3 block(32, kSegs), grid((unsigned)colsn, (a.L + kSegs * tl - 1) / (kSegs * tl), a.B);
    const T* xb = static_cast<const T*>(a.x);
    const T* gb = static_cast<const T*>(a.dout);
    T* db = static_cast<T*>(a.dx);
    if (a.silu) conv1d_bwd_fast_kernel<T, true, NP, CB, MINB><<<grid, block, 0, s>>>(a, tl, xb, gb, db);
    else conv1d_bwd_fast_kernel<T, false, NP, CB, MINB><<<grid, block, 0, s>>>(a, tl, xb, gb, db);
    return true;
  }
}

struct UpdArgs {
  const void* x; void* st; const void* w; const void* bias; const int* cs; void* out;
  int64_t xs_b, xs_d, xs_t, ss_b, ss_d, ss_s, os_b, os_d, os_t, ws_d, ws_w, bs_d;
  int B, D, T, S, W;
  int w_dtype, b_dtype, st_dtype;
  int silu;
};

template <typename T>
__global__ void conv1d_update_kernel(UpdArgs a) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (d >= a.D) return;
  float wgt[kMaxW];
#pragma unroll
  for (int k = 0; k < kMaxW; ++k) wgt[k] = k < a.W ? ld_any(a.w, a.w_dtype, d * a.ws_d + k * a.ws_w) : 0.f;
  const float bia = a.bias ? ld_any(a.bias, a.b_dtype, d * a.bs_d) : 0.f;
  const T* xp = static_cast<const T*>(a.x) + b * a.xs_b + d * a.xs_d;
  T* op = static_cast<T*>(a.out) + b * a.os_b + d * a.os_d;
  char* sp = static_cast<char*>(a.st);
  const int64_t sbase = b * a.ss_b + d * a.ss_d;
  auto st_ld = [&](int i) { return ld_any(sp, a.st_dtype, sbase + i * a.ss_s); };
  auto st_st = [&](int i, float v) { st_any(sp, a.st_dtype, sbase + i * a.ss_s, v); };
  if (a.cs == nullptr) {
    // window = [state(S) | x(T)], out[t] taps window[S + t - (W-1) + k]
    for (int t = 0; t < a.T; ++t) {
      float acc = bia;
      for (int k = 0; k < a.W; ++k) {
        const int i = a.S + t - (a.W - 1) + k;
        const float v = i < a.S ? st_ld(i) : to_f<T>(xp[(i - a.S) * a.xs_t]);
        acc += wgt[k] * v;
      }
      op[t * a.os_t] = from_f<T>(a.silu ? silu_f(acc) : acc);
    }
    for (int j = 0; j < a.S; ++j) {  // ascending: reads index T+j > j, not yet overwritten
      const int i = a.T + j;
      st_st(j, i < a.S ? st_ld(i) : to_f<T>(xp[(i - a.S) * a.xs_t]));
    }
  } else {
    const int c0 = a.cs[b];
    for (int t = 0; t < a.T; ++t) {
      float acc = bia;
      for (int k = 0; k < a.W; ++k) {
        const int i = t - (a.W - 1) + k;  // relative to the new tokens
        float v;
        if (i >= 0) v = to_f<T>(xp[i * a.xs_t]);
        else {
          int pos = (c0 + i) % a.S;
          if (pos < 0) pos += a.S;
          v = st_ld(pos);
        }
        acc += wgt[k] * v;
      }
      op[t * a.os_t] = from_f<T>(a.silu ? silu_f(acc) : acc);
    }
    for (int t = 0; t < a.T; ++t) st_st((c0 + t) % a.S, to_f<T>(xp[t * a.xs_t]));
  }
}

template <typename T, int VEC>
int launch_fwd_w(const ConvArgs& a, int W, cudaStream_t s) {
  dim3 block(32, kSegs), grid((a.D + 32 * VEC - 1) / (32 * VEC), (a.L + kSegs * kTL - 1) / (kSegs * kTL), a.B);
  if (grid.y == 0) grid.y = 1;
  switch (W) {
    case 2:
      if (a.seq_idx) conv1d_fwd_kernel<T, VEC, 2, true><<<grid, block, 0, s>>>(a);
      else conv1d_fwd_kernel<T, VEC, 2, false><<<grid, block, 0, s>>>(a);
      break;
    case 3:
      if (a.seq_idx) conv1d_fwd_kernel<T, VEC, 3, true><<<grid, block, 0, s>>>(a);
      else conv1d_fwd_kernel<T, VEC, 3, false><<<grid, block, 0, s>>>(a);
      break;
    default:
      if (a.seq_idx) conv1d_fwd_kernel<T, VEC, 4, true><<<grid, block, 0, s>>>(a);
      else conv1d_fwd_kernel<T, VEC, 4, false><<<grid, block, 0, s>>>(a);
      break;
  }
  OMNI_CUDA_LAUNCH_CHECK("conv1d_fwd_kernel");
  return OMNI_OK;
}
template <typename T, int VEC>
int launch_bwd_w(const ConvArgs& a, int W, cudaStream_t s) {
  dim3 block(32, kSegs), grid((a.D + 32 * VEC - 1) / (32 * VEC), (a.L + kSegs * kTL - 1) / (kSegs * kTL), a.B);
  switch (W) {
    case 2: conv1d_bwd_kernel<T, VEC, 2><<<grid, block, 0, s>>>(a); break;
    case 3: conv1d_bwd_kernel<T, VEC, 3><<<grid, block, 0, s>>>(a); break;
    default: conv1d_bwd_kernel<T, VEC, 4><<<grid, block, 0, s>>>(a); break;
  }
  OMNI_CUDA_LAUNCH_CHECK("conv1d_bwd_kernel");
  return OMNI_OK;
}

// can this (B, D, L) view be accessed with VEC-wide vectors along D?
bool vec_ok(const omni_tensor_t& t, int vec) {
  return t.stride[1] == 1 && t.shape[1] % vec == 0 && t.stride[0] % vec == 0 && t.stride[2] % vec == 0 &&
         aligned16(t.data);
}

int check_common(const omni_tensor_t& x, const omni_tensor_t& w, const omni_tensor_t& bias) {
  OMNI_CHECK(x.ndim == 3, OMNI_BAD_SHAPE, "conv1d: x must be (batch, dim, seqlen)");
  OMNI_CHECK(is_float_dtype(x.dtype), OMNI_BAD_DTYPE, "conv1d: x dtype");
  OMNI_CHECK(w.ndim == 2 && w.shape[0] == x.shape[1], OMNI_BAD_SHAPE, "conv1d: weight must be (dim, width)");
  OMNI_CHECK(w.shape[1] >= 2 && w.shape[1] <= kMaxW, OMNI_UNSUPPORTED, "conv1d: width must be 2..4");
  OMNI_CHECK(is_float_dtype(w.dtype), OMNI_BAD_DTYPE, "conv1d: weight dtype");
  if (present(bias)) {
    OMNI_CHECK(bias.ndim == 1 && bias.shape[0] == x.shape[1], OMNI_BAD_SHAPE, "conv1d: bias must be (dim)");
    OMNI_CHECK(is_float_dtype(bias.dtype), OMNI_BAD_DTYPE, "conv1d: bias dtype");
  }
  return OMNI_OK;
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_causal_conv1d_fwd(const omni_conv1d_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  if (int rc = check_common(p->x, p->weight, p->bias)) return rc;
  const omni_tensor_t &x = p->x, &o = p->out;
  OMNI_CHECK(present(o) && o.ndim == 3 && o.dtype == x.dtype, OMNI_BAD_SHAPE, "conv1d: out must match x");
  for (int i = 0; i < 3; ++i) OMNI_CHECK(o.shape[i] == x.shape[i], OMNI_BAD_SHAPE, "conv1d: out shape != x shape");
  const int W = (int)p->weight.shape[1];
  ConvArgs a{};
  a.x = x.data; a.w = p->weight.data; a.bias = p->bias.data; a.out = o.data;
  a.B = (int)x.shape[0]; a.D = (int)x.shape[1]; a.L = (int)x.shape[2];
  a.xs_b = x.stride[0]; a.xs_d = x.stride[1]; a.xs_l = x.stride[2];
  a.os_b = o.stride[0]; a.os_d = o.stride[1]; a.os_l = o.stride[2];
  a.ws_d = p->weight.stride[0]; a.ws_w = p->weight.stride[1]; a.bs_d = present(p->bias) ? p->bias.stride[0] : 0;
  a.w_dtype = p->weight.dtype; a.b_dtype = p->bias.dtype; a.silu = p->activation == OMNI_ACT_SILU;
  if (present(p->seq_idx)) {
    OMNI_CHECK(p->seq_idx.dtype == OMNI_I32 && shape_is(p->seq_idx, 2, a.B, a.L), OMNI_BAD_SHAPE,
               "conv1d: seq_idx must be int32 (batch, seqlen)");
    a.seq_idx = static_cast<const int*>(p->seq_idx.data); a.ss_b = p->seq_idx.stride[0]; a.ss_l = p->seq_idx.stride[1];
  }
  if (present(p->initial_states)) {
    OMNI_CHECK(shape_is(p->initial_states, 3, a.B, a.D, W - 1) && p->initial_states.dtype == x.dtype, OMNI_BAD_SHAPE,
               "conv1d: initial_states must be (batch, dim, width-1) in x's dtype");
    a.init = p->initial_states.data;
    a.is_b = p->initial_states.stride[0]; a.is_d = p->initial_states.stride[1]; a.is_k = p->initial_states.stride[2];
  }
  if (present(p->final_states)) {
    OMNI_CHECK(shape_is(p->final_states, 3, a.B, a.D, W - 1) && p->final_states.dtype == x.dtype, OMNI_BAD_SHAPE,
               "conv1d: final_states must be (batch, dim, width-1) in x's dtype");
    a.fin = p->final_states.data;
    a.fs_b = p->final_states.stride[0]; a.fs_d = p->final_states.stride[1]; a.fs_k = p->final_states.stride[2];
  }
  if (a.B == 0 || a.D == 0) return OMNI_OK;
  if (a.L == 0 && a.fin == nullptr) return OMNI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return OMNI_DISPATCH_FLOAT(x.dtype, T, [&]() -> int {
    constexpr int V = 16 / sizeof(T);
    if (try_launch_fwd_fast<T>(a, W, x, o, s)) {
      OMNI_CUDA_LAUNCH_CHECK("conv1d_fwd_fast_kernel");
      return OMNI_OK;
    }
    if (vec_ok(x, V) && vec_ok(o, V)) return launch_fwd_w<T, V>(a, W, s);
    return launch_fwd_w<T, 1>(a, W, s);
  });
}

extern "C" int omni_causal_conv1d_bwd(const omni_conv1d_bwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  if (int rc = check_common(p->x, p->weight, p->bias)) return rc;
  const omni_tensor_t &x = p->x, &g = p->dout, &dx = p->dx;
  OMNI_CHECK(present(g) && present(dx) && g.ndim == 3 && dx.ndim == 3 && g.dtype == x.dtype && dx.dtype == x.dtype,
             OMNI_BAD_SHAPE, "conv1d bwd: dout/dx must match x");
  for (int i = 0; i < 3; ++i)
    OMNI_CHECK(g.shape[i] == x.shape[i] && dx.shape[i] == x.shape[i], OMNI_BAD_SHAPE, "conv1d bwd: shape mismatch");
  const int W = (int)p->weight.shape[1];
  OMNI_CHECK(present(p->dweight) && p->dweight.dtype == OMNI_F32 && shape_is(p->dweight, 2, x.shape[1], W) &&
                 p->dweight.stride[1] == 1 && p->dweight.stride[0] == W,
             OMNI_BAD_SHAPE, "conv1d bwd: dweight must be contiguous fp32 (dim, width)");
  if (present(p->dbias))
    OMNI_CHECK(p->dbias.dtype == OMNI_F32 && shape_is(p->dbias, 1, x.shape[1]) && p->dbias.stride[0] == 1,
               OMNI_BAD_SHAPE, "conv1d bwd: dbias must be contiguous fp32 (dim)");
  ConvArgs a{};
  a.x = x.data; a.w = p->weight.data; a.bias = p->bias.data; a.dout = g.data; a.dx = dx.data;
  a.dw = static_cast<float*>(p->dweight.data); a.db = static_cast<float*>(p->dbias.data);
  a.B = (int)x.shape[0]; a.D = (int)x.shape[1]; a.L = (int)x.shape[2];
  a.xs_b = x.stride[0]; a.xs_d = x.stride[1]; a.xs_l = x.stride[2];
  a.gs_b = g.stride[0]; a.gs_d = g.stride[1]; a.gs_l = g.stride[2];
  a.ds_b = dx.stride[0]; a.ds_d = dx.stride[1]; a.ds_l = dx.stride[2];
  a.ws_d = p->weight.stride[0]; a.ws_w = p->weight.stride[1]; a.bs_d = present(p->bias) ? p->bias.stride[0] : 0;
  a.w_dtype = p->weight.dtype; a.b_dtype = p->bias.dtype; a.silu = p->activation == OMNI_ACT_SILU;
  if (present(p->seq_idx)) {
    OMNI_CHECK(p->seq_idx.dtype == OMNI_I32 && shape_is(p->seq_idx, 2, a.B, a.L), OMNI_BAD_SHAPE,
               "conv1d bwd: seq_idx must be int32 (batch, seqlen)");
    a.seq_idx = static_cast<const int*>(p->seq_idx.data); a.ss_b = p->seq_idx.stride[0]; a.ss_l = p->seq_idx.stride[1];
  }
  if (present(p->initial_states)) {
    OMNI_CHECK(shape_is(p->initial_states, 3, a.B, a.D, W - 1) && p->initial_states.dtype == x.dtype, OMNI_BAD_SHAPE,
               "conv1d bwd: initial_states must be (batch, dim, width-1)");
    a.init = p->initial_states.data;
    a.is_b = p->initial_states.stride[0]; a.is_d = p->initial_states.stride[1]; a.is_k = p->initial_states.stride[2];
  }
  if (present(p->dinitial_states)) {
    OMNI_CHECK(shape_is(p->dinitial_states, 3, a.B, a.D, W - 1) && p->dinitial_states.dtype == x.dtype, OMNI_BAD_SHAPE,
               "conv1d bwd: dinitial_states must be (batch, dim, width-1)");
    a.dinit = p->dinitial_states.data;
    a.dis_b = p->dinitial_states.stride[0]; a.dis_d = p->dinitial_states.stride[1];
    a.dis_k = p->dinitial_states.stride[2];
  }
  if (a.B == 0 || a.D == 0 || a.L == 0) return OMNI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return OMNI_DISPATCH_FLOAT(x.dtype, T, [&]() -> int {
    constexpr int V = 16 / sizeof(T);
    if (try_launch_bwd_fast<T>(a, W, x, g, dx, s)) {
      OMNI_CUDA_LAUNCH_CHECK("conv1d_bwd_fast_kernel");
      return OMNI_OK;
    }
    if (vec_ok(x, V) && vec_ok(g, V) && vec_ok(dx, V)) return launch_bwd_w<T, V>(a, W, s);
    return launch_bwd_w<T, 1>(a, W, s);
  });
}

extern "C" int omni_causal_conv1d_update(const omni_conv1d_update_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  if (int rc = check_common(p->x, p->weight, p->bias)) return rc;
  const omni_tensor_t &x = p->x, &st = p->conv_state, &o = p->out;
  const int W = (int)p->weight.shape[1];
  OMNI_CHECK(present(st) && st.ndim == 3 && st.shape[0] == x.shape[0] && st.shape[1] == x.shape[1] &&
                 st.shape[2] >= W - 1 && is_float_dtype(st.dtype),
             OMNI_BAD_SHAPE, "conv1d update: conv_state must be (batch, dim, state_len >= width-1)");
  OMNI_CHECK(present(o) && o.ndim == 3 && o.dtype == x.dtype, OMNI_BAD_SHAPE, "conv1d update: out must match x");
  for (int i = 0; i < 3; ++i) OMNI_CHECK(o.shape[i] == x.shape[i], OMNI_BAD_SHAPE, "conv1d update: out shape");
  UpdArgs a{};
  a.x = x.data; a.st = st.data; a.w = p->weight.data; a.bias = p->bias.data; a.out = o.data;
  a.B = (int)x.shape[0]; a.D = (int)x.shape[1]; a.T = (int)x.shape[2]; a.S = (int)st.shape[2]; a.W = W;
  a.xs_b = x.stride[0]; a.xs_d = x.stride[1]; a.xs_t = x.stride[2];
  a.ss_b = st.stride[0]; a.ss_d = st.stride[1]; a.ss_s = st.stride[2];
  a.os_b = o.stride[0]; a.os_d = o.stride[1]; a.os_t = o.stride[2];
  a.ws_d = p->weight.stride[0]; a.ws_w = p->weight.stride[1]; a.bs_d = present(p->bias) ? p->bias.stride[0] : 0;
  a.w_dtype = p->weight.dtype; a.b_dtype = p->bias.dtype; a.st_dtype = st.dtype;
  a.silu = p->activation == OMNI_ACT_SILU;
  if (present(p->cache_seqlens)) {
    OMNI_CHECK(p->cache_seqlens.dtype == OMNI_I32 && shape_is(p->cache_seqlens, 1, a.B) &&
                   p->cache_seqlens.stride[0] == 1,
               OMNI_BAD_SHAPE, "conv1d update: cache_seqlens must be contiguous int32 (batch)");
    a.cs = static_cast<const int*>(p->cache_seqlens.data);
  }
  if (a.B == 0 || a.D == 0 || a.T == 0) return OMNI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 block(128), grid((a.D + 127) / 128, a.B);
  return OMNI_DISPATCH_FLOAT(x.dtype, T, [&]() -> int {
    conv1d_update_kernel<T><<<grid, block, 0, s>>>(a);
    OMNI_CUDA_LAUNCH_CHECK("conv1d_update_kernel");
    return OMNI_OK;
  });
}
