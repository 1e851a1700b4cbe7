// Single-token recurrent state update (decode).  Pure HBM streaming of the SSM state.
//
// Replaces mamba_ssm/ops/triton/selective_state_update.py::selective_state_update, reached from
// Mamba2.step during the 256-token image decode of /root/reference/scripts/inference_t2i.py
// (models/stage2/generation.py:208-211 replays it inside a CUDA graph).  Arithmetic: SURVEY.md A.5.
//
//   state[b,h,p,:] = state * exp(dt A) + dt * x[b,h,p] * B[b,g,:]     (in place, state dtype kept)
//   out[b,h,p]     = <state[b,h,p,:], C[b,g,:]> + D x  (* silu(z))
//
// Layout / mapping: one warp owns kRows consecutive p-rows of one (b, h); its 32 lanes span d_state with
// 4-element vectors (16 B fp32 / 8 B bf16), so a row of N=128 fp32 is one fully coalesced 512 B
// transaction each way.  B and C (shared by the rows) live in registers.  All kRows rows' loads are issued
// before any math to keep >= 4 independent 16 B requests in flight per lane.  Algorithmic bytes per
// (b,h,p) row: 2*N*sizeof(state); everything else is O(1/N) of that.
#include "common.cuh"

namespace omni {
namespace {

// rows (p) per warp: enough 16-byte loads in flight per lane to cover the HBM latency (8 x 512 B fp32 rows, 16 x 256 B
// 16-bit rows = 4 KB per warp)
template <typename TS> constexpr int rows_of() { return 4; }  // (8 / 16 rows per warp measured slower: fewer warps in flight)
constexpr int kWarps = 4;

struct SsuArgs {
  void* state; const void* x; const void* dt; const void* A; const void* Bm; const void* Cm; const void* D;
  const void* z; const void* dt_bias; void* out;
  int64_t st_b, st_h, st_p;
  int64_t x_b, x_h, x_p, dt_b, dt_h, dt_p, z_b, z_h, z_p, o_b, o_h, o_p;
  int64_t A_h, A_p, A_n, B_b, B_g, C_b, C_g, D_h, D_p, db_h, db_p;
  int B, H, P, N, G;
  int st_dtype, x_dtype, dt_dtype, A_dtype, bc_dtype, D_dtype, db_dtype;
  int dt_softplus;
};

template <typename TS> __device__ __forceinline__ void ld4(const TS* p, float (&o)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <> __device__ __forceinline__ void ld4<__half>(const __half* p, float (&o)[4]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
  const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename TS> __device__ __forceinline__ void st4(TS* p, const float (&o)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[4]) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(o[0], o[1]);
  h[1] = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(p) = raw;
}
template <> __device__ __forceinline__ void st4<__half>(__half* p, const float (&o)[4]) {
  uint2 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
  h[0] = __floats2half2_rn(o[0], o[1]);
  h[1] = __floats2half2_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(p) = raw;
}

// NV = number of 4-element vectors per lane (N = 128 * NV, or N <= 128 with idle lanes when NV == 1).
template <typename TS, int NV, bool TIE_A>
__global__ void __launch_bounds__(32 * kWarps) ssu_kernel(SsuArgs a) {
  constexpr int kRows = rows_of<TS>() / NV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pblocks = (a.P + kRows - 1) / kRows;
  const int64_t task = (int64_t)blockIdx.x * kWarps + warp;
  if (task >= (int64_t)a.B * a.H * pblocks) return;
  const int pb = (int)(task % pblocks);
  const int h = (int)((task / pblocks) % a.H);
  const int b = (int)(task / ((int64_t)pblocks * a.H));
  const int g = h / (a.H / a.G);
  const int p0 = pb * kRows;

  float Bv[NV][4], Cv[NV][4];
  bool nvalid[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int n = (j * 32 + lane) * 4;
    nvalid[j] = n < a.N;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      Bv[j][e] = nvalid[j] ? ld_any(a.Bm, a.bc_dtype, b * a.B_b + g * a.B_g + n + e) : 0.f;
      Cv[j][e] = nvalid[j] ? ld_any(a.Cm, a.bc_dtype, b * a.C_b + g * a.C_g + n + e) : 0.f;
    }
  }
  TS* sbase = static_cast<TS*>(a.state) + b * a.st_b + h * a.st_h;
  float S[kRows][NV][4];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int p = p0 + r;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int n = (j * 32 + lane) * 4;
      if (p < a.P && nvalid[j]) ld4<TS>(sbase + p * a.st_p + n, S[r][j]);
    }
  }
  // per-row scalars of all rows first: their (dependent-free) loads are in flight together with the state rows instead of
  // costing one memory latency per row
  float dtr[kRows], xr[kRows], Ar[kRows], Dr[kRows], zr[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int p = min(p0 + r, a.P - 1);
    dtr[r] = ld_any(a.dt, a.dt_dtype, b * a.dt_b + h * a.dt_h + p * a.dt_p);
    if (a.dt_bias) dtr[r] += ld_any(a.dt_bias, a.db_dtype, h * a.db_h + p * a.db_p);
    xr[r] = ld_any(a.x, a.x_dtype, b * a.x_b + h * a.x_h + p * a.x_p);
    Ar[r] = TIE_A ? ld_any(a.A, a.A_dtype, h * a.A_h + p * a.A_p) : 0.f;
    Dr[r] = a.D ? ld_any(a.D, a.D_dtype, h * a.D_h + p * a.D_p) : 0.f;
    zr[r] = a.z ? ld_any(a.z, a.x_dtype, b * a.z_b + h * a.z_h + p * a.z_p) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int p = p0 + r;
    if (p >= a.P) continue;  // warp-uniform
    float dtv = dtr[r];
    if (a.dt_softplus) dtv = softplus_f(dtv);
    const float xv = xr[r];
    const float dtx = dtv * xv;
    float dA_tied = 0.f;
    if (TIE_A) dA_tied = __expf(dtv * Ar[r]);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int n = (j * 32 + lane) * 4;
      if (nvalid[j]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dA = TIE_A ? dA_tied : __expf(dtv * ld_any(a.A, a.A_dtype, h * a.A_h + p * a.A_p + (n + e) * a.A_n));
          const float s = S[r][j][e] * dA + dtx * Bv[j][e];
          S[r][j][e] = s;
          acc += s * Cv[j][e];
        }
        st4<TS>(sbase + p * a.st_p + n, S[r][j]);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (a.D) acc += xv * Dr[r];
      if (a.z) acc *= silu_f(zr[r]);
      st_any(a.out, a.x_dtype, b * a.o_b + h * a.o_h + p * a.o_p, acc);
    }
  }
}

// ---- fast path: A tied over d_state (Mamba-2), d_state = 128, headdim a multiple of 16 --------------------------------
// The kernel above recomputes every per-row scalar (dt transform, exp, D x, silu(z)) in all 32 lanes and reduces each row
// with its own 5-step shuffle tree: ~60 instructions per state element, instruction-bound at 43 % of the HBM roofline (ncu).
// Here a warp owns 16 (fp32 state) or 32 (16-bit state) rows of one (batch, head): lane r computes the scalars of row r once
// and broadcasts them by shuffle, all rows (8 KB) are in flight before the first is used, and the per-lane partial sums of
// <state, C> are reduced together by a transposing butterfly (16 / 31 shuffles instead of 80 / 160).
template <typename TS> struct RawRow;   // one lane's 4 state elements of a row, as loaded
template <> struct RawRow<float> { using type = float4; };
template <> struct RawRow<__nv_bfloat16> { using type = uint2; };
template <> struct RawRow<__half> { using type = uint2; };
#ifndef OMNI_SSU_ROWS16
#define OMNI_SSU_ROWS16 16   // state rows per warp for a 16-bit state (fp32: half of it)
#endif
#ifndef OMNI_SSU_OCC
#define OMNI_SSU_OCC 6       // CTAs (of kWarps warps) per SM the register budget is cut for
#endif
// (small tasks at a small register budget: many resident warps, each with 2 KB of state in flight, hide the latency of the
// per-row scalars and the butterfly better than few warps with 8 KB each - the same finding as in decode_core.cu)
template <typename TS> constexpr int tied_rows() { return sizeof(TS) == 4 ? OMNI_SSU_ROWS16 / 2 : OMNI_SSU_ROWS16; }

template <typename TS>
__global__ void __launch_bounds__(32 * kWarps, OMNI_SSU_OCC) ssu_tied_kernel(SsuArgs a) {
  constexpr int R = tied_rows<TS>();
  using Raw = typename RawRow<TS>::type;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pblocks = a.P / R;
  const int64_t task = (int64_t)blockIdx.x * kWarps + warp;
  if (task >= (int64_t)a.B * a.H * pblocks) return;
  const int pb = (int)(task % pblocks);
  const int h = (int)((task / pblocks) % a.H);
  const int b = (int)(task / ((int64_t)pblocks * a.H));
  const int g = h / (a.H / a.G);
  const int p0 = pb * R;
  const int n = lane * 4;
  // state rows first: the longest latency (kept as loaded, converted when used)
  TS* sbase = static_cast<TS*>(a.state) + b * a.st_b + h * a.st_h + (int64_t)p0 * a.st_p + n;
  Raw raw[R];
#pragma unroll
  for (int r = 0; r < R; ++r) raw[r] = *reinterpret_cast<const Raw*>(sbase + r * a.st_p);
  float Bv[4], Cv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    Bv[e] = ld_any(a.Bm, a.bc_dtype, b * a.B_b + g * a.B_g + n + e);
    Cv[e] = ld_any(a.Cm, a.bc_dtype, b * a.C_b + g * a.C_g + n + e);
  }
  // per-row scalars, one row per lane (R = 16: lanes 16..31 mirror rows 0..15)
  const int p = p0 + (lane & (R - 1));
  float dtv = ld_any(a.dt, a.dt_dtype, b * a.dt_b + h * a.dt_h + p * a.dt_p);
  if (a.dt_bias) dtv += ld_any(a.dt_bias, a.db_dtype, h * a.db_h + p * a.db_p);
  if (a.dt_softplus) dtv = softplus_f(dtv);
  const float xv = ld_any(a.x, a.x_dtype, b * a.x_b + h * a.x_h + p * a.x_p);
  const float dA_l = __expf(dtv * ld_any(a.A, a.A_dtype, h * a.A_h + p * a.A_p));
  const float dtx_l = dtv * xv;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float dA = __shfl_sync(0xffffffffu, dA_l, r), dtx = __shfl_sync(0xffffffffu, dtx_l, r);
    float S[4];
    if constexpr (sizeof(TS) == 4) { S[0] = raw[r].x; S[1] = raw[r].y; S[2] = raw[r].z; S[3] = raw[r].w; }
    else ld4<TS>(reinterpret_cast<const TS*>(&raw[r]), S);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      S[e] = fmaf(S[e], dA, dtx * Bv[e]);
      sum = fmaf(S[e], Cv[e], sum);
    }
    st4<TS>(sbase + r * a.st_p, S);
    acc[r] = sum;
  }
  // transposing butterfly: each xor step halves the rows a lane carries and doubles the lanes summed; with R = 32 the five
  // steps leave lane l with the complete sum of row l, with R = 16 lane l holds row l >> 1 and one more xor-1 step completes it
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
  const float x_r = __shfl_sync(0xffffffffu, xv, row);  // (the scalars of `row` live in lane `row`)
  if (lane % (32 / R) == 0) {
    const int pr = p0 + row;
    if (a.D) y = fmaf(x_r, ld_any(a.D, a.D_dtype, h * a.D_h + pr * a.D_p), y);
    if (a.z) y *= silu_f(ld_any(a.z, a.x_dtype, b * a.z_b + h * a.z_h + pr * a.z_p));
    st_any(a.out, a.x_dtype, b * a.o_b + h * a.o_h + pr * a.o_p, y);
  }
}

template <typename TS>
int launch(const SsuArgs& a, cudaStream_t s) {
  const bool tie = a.A_n == 0;
  if (tie && a.N == 128 && a.P % tied_rows<TS>() == 0 && a.P > 0) {
    const int64_t tasks16 = (int64_t)a.B * a.H * (a.P / tied_rows<TS>());
    ssu_tied_kernel<TS><<<(unsigned)((tasks16 + kWarps - 1) / kWarps), 32 * kWarps, 0, s>>>(a);
    OMNI_CUDA_LAUNCH_CHECK("ssu_tied_kernel");
    return OMNI_OK;
  }
  const int nv = a.N <= 128 ? 1 : 2;
  const int krows = rows_of<TS>() / nv;
  const int pblocks = (a.P + krows - 1) / krows;
  const int64_t tasks = (int64_t)a.B * a.H * pblocks;
  const unsigned grid = (unsigned)((tasks + kWarps - 1) / kWarps);
  if (a.N <= 128) {
    if (tie) ssu_kernel<TS, 1, true><<<grid, 32 * kWarps, 0, s>>>(a);
    else ssu_kernel<TS, 1, false><<<grid, 32 * kWarps, 0, s>>>(a);
  } else {
    if (tie) ssu_kernel<TS, 2, true><<<grid, 32 * kWarps, 0, s>>>(a);
    else ssu_kernel<TS, 2, false><<<grid, 32 * kWarps, 0, s>>>(a);
  }
  OMNI_CUDA_LAUNCH_CHECK("ssu_kernel");
  return OMNI_OK;
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_selective_state_update(const omni_ssu_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t& st = p->state;
  OMNI_CHECK(present(st) && st.ndim == 4 && is_float_dtype(st.dtype), OMNI_BAD_SHAPE,
             "selective_state_update: state must be (B, H, P, N)");
  const int64_t Bsz = st.shape[0], H = st.shape[1], P = st.shape[2], N = st.shape[3];
  OMNI_CHECK(N % 4 == 0 && N <= 256, OMNI_UNSUPPORTED, "selective_state_update: d_state must be a multiple of 4, <= 256");
  OMNI_CHECK(st.stride[3] == 1 && st.stride[2] % 4 == 0 && st.stride[1] % 4 == 0 && st.stride[0] % 4 == 0 &&
                 aligned16(st.data),
             OMNI_BAD_STRIDE, "selective_state_update: state rows must be contiguous and 16B aligned");
  auto chk3 = [&](const omni_tensor_t& t, const char* name, bool req) -> int {
    if (!present(t)) { OMNI_CHECK(!req, OMNI_BAD_SHAPE, "selective_state_update: %s required", name); return OMNI_OK; }
    OMNI_CHECK(shape_is(t, 3, Bsz, H, P) && is_float_dtype(t.dtype), OMNI_BAD_SHAPE,
               "selective_state_update: %s must be (B, H, P)", name);
    return OMNI_OK;
  };
  if (int rc = chk3(p->x, "x", true)) return rc;
  if (int rc = chk3(p->dt, "dt", true)) return rc;
  if (int rc = chk3(p->z, "z", false)) return rc;
  if (int rc = chk3(p->out, "out", true)) return rc;
  OMNI_CHECK(p->out.dtype == p->x.dtype && (!present(p->z) || p->z.dtype == p->x.dtype), OMNI_BAD_DTYPE,
             "selective_state_update: out/z dtype must equal x dtype");
  OMNI_CHECK(present(p->A) && shape_is(p->A, 3, H, P, N) && is_float_dtype(p->A.dtype), OMNI_BAD_SHAPE,
             "selective_state_update: A must be (H, P, N)");
  OMNI_CHECK(present(p->B) && p->B.ndim == 3 && p->B.shape[0] == Bsz && p->B.shape[2] == N && is_float_dtype(p->B.dtype) &&
                 p->B.stride[2] == 1,
             OMNI_BAD_SHAPE, "selective_state_update: B must be (B, G, N) with contiguous N");
  const int64_t G = p->B.shape[1];
  OMNI_CHECK(G > 0 && H % G == 0, OMNI_BAD_SHAPE, "selective_state_update: nheads %% ngroups != 0");
  OMNI_CHECK(shape_is(p->C, 3, Bsz, G, N) && p->C.dtype == p->B.dtype && p->C.stride[2] == 1, OMNI_BAD_SHAPE,
             "selective_state_update: C must match B");
  auto chk2 = [&](const omni_tensor_t& t, const char* name) -> int {
    if (!present(t)) return OMNI_OK;
    OMNI_CHECK(shape_is(t, 2, H, P) && is_float_dtype(t.dtype), OMNI_BAD_SHAPE,
               "selective_state_update: %s must be (H, P)", name);
    return OMNI_OK;
  };
  if (int rc = chk2(p->D, "D")) return rc;
  if (int rc = chk2(p->dt_bias, "dt_bias")) return rc;
  if (Bsz == 0 || H == 0 || P == 0) return OMNI_OK;
  SsuArgs a{};
  a.state = st.data; a.x = p->x.data; a.dt = p->dt.data; a.A = p->A.data; a.Bm = p->B.data; a.Cm = p->C.data;
  a.D = p->D.data; a.z = p->z.data; a.dt_bias = p->dt_bias.data; a.out = p->out.data;
  a.st_b = st.stride[0]; a.st_h = st.stride[1]; a.st_p = st.stride[2];
  a.x_b = p->x.stride[0]; a.x_h = p->x.stride[1]; a.x_p = p->x.stride[2];
  a.dt_b = p->dt.stride[0]; a.dt_h = p->dt.stride[1]; a.dt_p = p->dt.stride[2];
  if (present(p->z)) { a.z_b = p->z.stride[0]; a.z_h = p->z.stride[1]; a.z_p = p->z.stride[2]; }
  a.o_b = p->out.stride[0]; a.o_h = p->out.stride[1]; a.o_p = p->out.stride[2];
  a.A_h = p->A.stride[0]; a.A_p = p->A.stride[1]; a.A_n = p->A.stride[2];
  a.B_b = p->B.stride[0]; a.B_g = p->B.stride[1]; a.C_b = p->C.stride[0]; a.C_g = p->C.stride[1];
  if (present(p->D)) { a.D_h = p->D.stride[0]; a.D_p = p->D.stride[1]; }
  if (present(p->dt_bias)) { a.db_h = p->dt_bias.stride[0]; a.db_p = p->dt_bias.stride[1]; }
  a.B = (int)Bsz; a.H = (int)H; a.P = (int)P; a.N = (int)N; a.G = (int)G;
  a.st_dtype = st.dtype; a.x_dtype = p->x.dtype; a.dt_dtype = p->dt.dtype; a.A_dtype = p->A.dtype;
  a.bc_dtype = p->B.dtype; a.D_dtype = p->D.dtype; a.db_dtype = p->dt_bias.dtype;
  a.dt_softplus = p->dt_softplus;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return OMNI_DISPATCH_FLOAT(st.dtype, TS, [&]() -> int { return launch<TS>(a, s); });
}
