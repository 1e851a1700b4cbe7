// Exact fp32 SIMT evaluation of the SSD recurrence (forward and backward) for ANY dtype / headdim /
// d_state in {16,32,64,128,256}.  It is (a) the fp32-accurate path behind mamba_chunk_scan_combined
// (north_star: <= 1e-5 vs the reference in fp32), (b) the on-device cross-check for the tcgen05
// chunked kernel at sizes where the CPU oracle is too slow, and (c) the fallback for shapes the
// tensor-core kernel does not cover (seq_idx, D with headdim, odd d_state, fp16...).
//
// Recurrence:  S_t = exp(dt_t A) S_{t-1} + dt_t x_t (x) B_t ;  y_t = S_t C_t + D x_t.
// Backward with G_t = dL/dS_t:
//     G_t = dy_t (x) C_t + exp(a_{t+1}) G_{t+1},   dx_t = dt_t G_t B_t + D dy_t,
//     dB_t = dt_t sum_h G_t^T x_t,  dC_t = sum_h S_t^T dy_t,
//     da_t = alpha_t - dt_t gamma_t + da_{t+1},  alpha_t = <dy_t, S_t C_t>,  gamma_t = <x_t, G_t B_t>,
//     ddt_t = gamma_t + A da_t,  dA = sum dt_t da_t,  dD = sum <dy_t, x_t>.
//
// Two kernel templates cover all four sweeps; both keep the (P x N) state of a head slice in
// registers for the whole sequence and stream 32-token tiles of the operands through shared memory:
//   rn  ("reduce over n"): thread = (p, n-slice);  out[p] = sum_n state[p][n] c[n]   -> y (fwd), dx (bwd)
//   rp  ("reduce over p"): thread = (n, 16 p's);   out[n] = sum_p state[p][n] c[p]   -> dC, dB (bwd)
#include "common.cuh"

namespace omni {
namespace {

constexpr int kTT = 32;   // tokens per shared-memory tile
constexpr int kPB = 16;   // head-dim rows per CTA
constexpr int kTTp = 16;  // tile of the reduce-over-p kernels (three [tile][N] fp32 buffers must fit 48 KB)

struct SsdArgs {
  // forward operands
  const void* x; const void* dt; const float* A; const void* Bm; const void* Cm; const void* D; const void* z;
  const void* dt_bias; const void* init; const int* seq_idx;
  void* out; float* fin;
  // backward operands
  const void* dout; const float* dfin;
  void* dx; void* ddt; float* dB; float* dC; float* dinit; float* dA_part; float* ddtb_part; float* dD_part;
  void* dz; int64_t dz_b, dz_l, dz_h;  // forward-direction "dz pass": dz = dout * y * silu'(z) instead of out
  float* ws_alpha; float* ws_gamma; float* ws_e0;
  int64_t x_b, x_l, x_h;        // x (B, L, H, P), P contiguous
  int64_t o_b, o_l, o_h;        // out / dx
  int64_t z_b, z_l, z_h;
  int64_t g_b, g_l, g_h;        // dout
  int64_t dt_b, dt_l, dt_h;
  int64_t ddt_b, ddt_l, ddt_h;
  int64_t B_b, B_l, B_g, C_b, C_l, C_g;
  int64_t D_h, D_p;
  int64_t i_b, i_h, i_p;        // initial_states (B, H, P, N), N contiguous
  int64_t s_b, s_l;
  int B, L, H, P, G, N;
  int x_dtype, dt_dtype, bc_dtype, D_dtype, dtb_dtype, init_dtype;
  int dt_softplus;
  float dt_min, dt_max;
};

__device__ __forceinline__ float dt_xform(const SsdArgs& a, int b, int t, int h) {
  float v = ld_any(a.dt, a.dt_dtype, b * a.dt_b + t * a.dt_l + h * a.dt_h);
  if (a.dt_bias) v += ld_any(a.dt_bias, a.dtb_dtype, h);
  if (a.dt_softplus) v = softplus_f(v);
  return fminf(fmaxf(v, a.dt_min), a.dt_max);
}
// derivative of the transformed dt w.r.t. the raw dt
__device__ __forceinline__ float dt_xform_grad(const SsdArgs& a, int b, int t, int h) {
  float v = ld_any(a.dt, a.dt_dtype, b * a.dt_b + t * a.dt_l + h * a.dt_h);
  if (a.dt_bias) v += ld_any(a.dt_bias, a.dtb_dtype, h);
  float g = 1.f;
  float u = v;
  if (a.dt_softplus) {
    u = softplus_f(v);
    g = v <= 20.f ? sigmoid_f(v) : 1.f;
  }
  if (u < a.dt_min || u > a.dt_max) g = 0.f;
  return g;
}
// decay applied when stepping INTO token t (0 at packed-sequence boundaries)
__device__ __forceinline__ float decay_into(const SsdArgs& a, int b, int t, int h, float dtv) {
  if (a.seq_idx && t > 0 && a.seq_idx[b * a.s_b + t * a.s_l] != a.seq_idx[b * a.s_b + (t - 1) * a.s_l]) return 0.f;
  return __expf(dtv * a.A[h]);
}

// ---- "reduce over n": forward y, or reverse dx ---------------------------------------------------------
template <int NPT, bool REVERSE>
__global__ void __launch_bounds__(128) ssd_rn_kernel(SsdArgs a) {
  constexpr int N = NPT * 8;
  __shared__ float bs[kTT][N];      // fwd: B   rev: C   (the operand of the outer product)
  __shared__ float cs[kTT][N];      // fwd: C   rev: B   (the operand of the contraction)
  __shared__ float as_[kTT][kPB];   // fwd: x   rev: gated dout
  __shared__ float xs2[kTT][kPB];   // rev only: x (for dD)
  __shared__ float ys[kTT][kPB];
  __shared__ float dts[kTT], das[kTT];
  const int tid = threadIdx.x, pl = tid >> 3, ns = tid & 7;
  const int p0 = blockIdx.x * kPB, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (a.H / a.G);
  const int p = p0 + pl;
  const bool pvalid = p < a.P;

  float S[NPT];
  {
    const float* src = nullptr;
    if (!REVERSE && a.init) {
#pragma unroll
      for (int k = 0; k < NPT; ++k)
        S[k] = pvalid ? ld_any(a.init, a.init_dtype, b * a.i_b + h * a.i_h + p * a.i_p + ns + 8 * k) : 0.f;
    } else if (REVERSE && a.dfin) {
      src = a.dfin + ((int64_t)(b * a.H + h) * a.P + p) * N;
#pragma unroll
      for (int k = 0; k < NPT; ++k) S[k] = pvalid ? src[ns + 8 * k] : 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < NPT; ++k) S[k] = 0.f;
    }
  }
  float dD_acc = 0.f;
  const float Dv = a.D ? (pvalid ? ld_any(a.D, a.D_dtype, h * a.D_h + p * a.D_p) : 0.f) : 0.f;
  const int ntiles = (a.L + kTT - 1) / kTT;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int tile = REVERSE ? ntiles - 1 - ti : ti;
    const int t0 = tile * kTT;
    const int tn = min(kTT, a.L - t0);
    __syncthreads();
    for (int i = tid; i < tn * N; i += 128) {
      const int tt = i / N, n = i % N;
      const float bv = ld_any(a.Bm, a.bc_dtype, b * a.B_b + (t0 + tt) * a.B_l + g * a.B_g + n);
      const float cv = ld_any(a.Cm, a.bc_dtype, b * a.C_b + (t0 + tt) * a.C_l + g * a.C_g + n);
      bs[tt][n] = REVERSE ? cv : bv;
      cs[tt][n] = REVERSE ? bv : cv;
    }
    for (int i = tid; i < tn * kPB; i += 128) {
      const int tt = i / kPB, q = i % kPB;
      float xv = 0.f, av = 0.f;
      if (p0 + q < a.P) {
        xv = ld_any(a.x, a.x_dtype, b * a.x_b + (t0 + tt) * a.x_l + h * a.x_h + p0 + q);
        if (REVERSE) {
          av = ld_any(a.dout, a.x_dtype, b * a.g_b + (t0 + tt) * a.g_l + h * a.g_h + p0 + q);
          if (a.z) av *= silu_f(ld_any(a.z, a.x_dtype, b * a.z_b + (t0 + tt) * a.z_l + h * a.z_h + p0 + q));
        }
      }
      as_[tt][q] = REVERSE ? av : xv;
      if (REVERSE) xs2[tt][q] = xv;
    }
    if (tid < tn) {
      const int t = t0 + tid;
      const float dtv = dt_xform(a, b, t, h);
      dts[tid] = dtv;
      if (!REVERSE) das[tid] = decay_into(a, b, t, h, dtv);
      else das[tid] = (t + 1 < a.L) ? decay_into(a, b, t + 1, h, dt_xform(a, b, t + 1, h)) : 1.f;
    }
    __syncthreads();
    for (int j = 0; j < tn; ++j) {
      const int tt = REVERSE ? tn - 1 - j : j;
      const float da = das[tt];
      const float av = REVERSE ? as_[tt][pl] : dts[tt] * as_[tt][pl];
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NPT; ++k) {
        S[k] = da * S[k] + av * bs[tt][ns + 8 * k];
        acc += S[k] * cs[tt][ns + 8 * k];
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (ns == 0) {
        if (!REVERSE) ys[tt][pl] = acc + Dv * as_[tt][pl];
        else {
          ys[tt][pl] = dts[tt] * acc + Dv * as_[tt][pl];
          dD_acc += as_[tt][pl] * xs2[tt][pl];
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < tn * kPB; i += 128) {
      const int tt = i / kPB, q = i % kPB;
      if (p0 + q < a.P) {
        float v = ys[tt][q];
        if (!REVERSE) {
          if (a.dz) {
            const float zv = ld_any(a.z, a.x_dtype, b * a.z_b + (t0 + tt) * a.z_l + h * a.z_h + p0 + q);
            const float gv = ld_any(a.dout, a.x_dtype, b * a.g_b + (t0 + tt) * a.g_l + h * a.g_h + p0 + q);
            st_any(a.dz, a.x_dtype, b * a.dz_b + (t0 + tt) * a.dz_l + h * a.dz_h + p0 + q, gv * v * dsilu_f(zv));
          } else {
            if (a.z) v *= silu_f(ld_any(a.z, a.x_dtype, b * a.z_b + (t0 + tt) * a.z_l + h * a.z_h + p0 + q));
            st_any(a.out, a.x_dtype, b * a.o_b + (t0 + tt) * a.o_l + h * a.o_h + p0 + q, v);
          }
        } else {
          st_any(a.dx, a.x_dtype, b * a.o_b + (t0 + tt) * a.o_l + h * a.o_h + p0 + q, v);
        }
      }
    }
  }
  if (!REVERSE) {
    if (a.fin && pvalid) {
      float* dst = a.fin + ((int64_t)(b * a.H + h) * a.P + p) * N;
#pragma unroll
      for (int k = 0; k < NPT; ++k) dst[ns + 8 * k] = S[k];
    }
  } else {
    if (a.dD_part && ns == 0 && pvalid) a.dD_part[(int64_t)(b * a.H + h) * a.P + p] = dD_acc;
    if (a.dinit && pvalid) {
      // dS_0 = exp(a_0) G_0 (token 0 always starts a sequence only when seq_idx says so; decay_into(t=0) is plain)
      const float d0 = a.L > 0 ? __expf(dt_xform(a, b, 0, h) * a.A[h]) : 1.f;
      float* dst = a.dinit + ((int64_t)(b * a.H + h) * a.P + p) * N;
#pragma unroll
      for (int k = 0; k < NPT; ++k) dst[ns + 8 * k] = d0 * S[k];
    }
  }
}

// ---- "reduce over p": forward-direction dC (+alpha), reverse-direction dB (+gamma) ----------------------
// CTA = N threads x PG groups; each thread owns one n and kPB p's.
template <int N, int PG, bool REVERSE>
__global__ void __launch_bounds__(N * PG) ssd_rp_kernel(SsdArgs a) {
  constexpr int NT = N * PG;
  constexpr int PBT = kPB * PG;  // p rows per CTA
  __shared__ float bs[kTTp][N];               // fwd: B    rev: C
  __shared__ float ms[kTTp][N];               // fwd: C    rev: B   (for the alpha/gamma dot at tile end)
  __shared__ __align__(16) float av[kTTp][PBT];   // fwd: dt*x  rev: gated dout   (outer-product operand)
  __shared__ __align__(16) float cv[kTTp][PBT];   // fwd: gated dout   rev: x     (contraction operand)
  __shared__ float qs[kTTp][N];               // per-token results, reduced over PG at tile end
  __shared__ float dts[kTTp], das[kTTp];
  const int tid = threadIdx.x, n = tid % N, pg = tid / N;
  const int p0 = blockIdx.x * PBT, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (a.H / a.G);

  float S[kPB];
#pragma unroll
  for (int k = 0; k < kPB; ++k) {
    const int p = p0 + pg * kPB + k;
    float v = 0.f;
    if (p < a.P) {
      if (!REVERSE && a.init) v = ld_any(a.init, a.init_dtype, b * a.i_b + h * a.i_h + p * a.i_p + n);
      if (REVERSE && a.dfin) v = a.dfin[((int64_t)(b * a.H + h) * a.P + p) * N + n];
    }
    S[k] = v;
  }
  const int ntiles = (a.L + kTTp - 1) / kTTp;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int tile = REVERSE ? ntiles - 1 - ti : ti;
    const int t0 = tile * kTTp;
    const int tn = min(kTTp, a.L - t0);
    __syncthreads();
    if (tid < tn) {
      const int t = t0 + tid;
      const float dtv = dt_xform(a, b, t, h);
      dts[tid] = dtv;
      if (!REVERSE) das[tid] = decay_into(a, b, t, h, dtv);
      else das[tid] = (t + 1 < a.L) ? decay_into(a, b, t + 1, h, dt_xform(a, b, t + 1, h)) : 1.f;
    }
    for (int i = tid; i < tn * N; i += NT) {
      const int tt = i / N, nn = i % N;
      const float bv = ld_any(a.Bm, a.bc_dtype, b * a.B_b + (t0 + tt) * a.B_l + g * a.B_g + nn);
      const float cvv = ld_any(a.Cm, a.bc_dtype, b * a.C_b + (t0 + tt) * a.C_l + g * a.C_g + nn);
      bs[tt][nn] = REVERSE ? cvv : bv;
      ms[tt][nn] = REVERSE ? bv : cvv;
      qs[tt][nn] = 0.f;
    }
    __syncthreads();  // dts visible
    for (int i = tid; i < tn * PBT; i += NT) {
      const int tt = i / PBT, q = i % PBT;
      float xv = 0.f, gv = 0.f;
      if (p0 + q < a.P) {
        xv = ld_any(a.x, a.x_dtype, b * a.x_b + (t0 + tt) * a.x_l + h * a.x_h + p0 + q);
        gv = ld_any(a.dout, a.x_dtype, b * a.g_b + (t0 + tt) * a.g_l + h * a.g_h + p0 + q);
        if (a.z) gv *= silu_f(ld_any(a.z, a.x_dtype, b * a.z_b + (t0 + tt) * a.z_l + h * a.z_h + p0 + q));
      }
      av[tt][q] = REVERSE ? gv : dts[tt] * xv;
      cv[tt][q] = REVERSE ? xv : gv;
    }
    __syncthreads();
    for (int j = 0; j < tn; ++j) {
      const int tt = REVERSE ? tn - 1 - j : j;
      const float da = das[tt];
      const float bn = bs[tt][n];
      float q = 0.f;
      const float4* a4 = reinterpret_cast<const float4*>(&av[tt][pg * kPB]);
      const float4* c4 = reinterpret_cast<const float4*>(&cv[tt][pg * kPB]);
#pragma unroll
      for (int k4 = 0; k4 < kPB / 4; ++k4) {
        const float4 aa = a4[k4], cc = c4[k4];
        S[4 * k4 + 0] = da * S[4 * k4 + 0] + aa.x * bn; q += S[4 * k4 + 0] * cc.x;
        S[4 * k4 + 1] = da * S[4 * k4 + 1] + aa.y * bn; q += S[4 * k4 + 1] * cc.y;
        S[4 * k4 + 2] = da * S[4 * k4 + 2] + aa.z * bn; q += S[4 * k4 + 2] * cc.z;
        S[4 * k4 + 3] = da * S[4 * k4 + 3] + aa.w * bn; q += S[4 * k4 + 3] * cc.w;
      }
      if (PG == 1) qs[tt][n] = q;
      else atomicAdd(&qs[tt][n], q);
    }
    __syncthreads();
    // tile end: scatter this CTA's contribution (reduced over its p rows) and the per-token scalar
    float* dst = REVERSE ? a.dB : a.dC;
    for (int i = tid; i < tn * N; i += NT) {
      const int tt = i / N, nn = i % N;
      const float v = REVERSE ? dts[tt] * qs[tt][nn] : qs[tt][nn];
      atomicAdd(dst + ((int64_t)(b * a.L + t0 + tt) * a.G + g) * N + nn, v);
    }
    // alpha_t = <C_t, q_t>  /  gamma_t = <B_t, q_t>: one warp per token
    const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    for (int tt = warp; tt < tn; tt += nw) {
      float s = 0.f;
      for (int nn = lane; nn < N; nn += 32) s += ms[tt][nn] * qs[tt][nn];
      s = warp_sum(s);
      if (lane == 0) atomicAdd((REVERSE ? a.ws_gamma : a.ws_alpha) + (int64_t)(b * a.H + h) * a.L + t0 + tt, s);
    }
  }
  if (!REVERSE && a.dfin && a.ws_e0) {
    // e0 = <dS_final, S_final> seeds the reverse cumulative sum of da
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kPB; ++k) {
      const int p = p0 + pg * kPB + k;
      if (p < a.P) s += S[k] * a.dfin[((int64_t)(b * a.H + h) * a.P + p) * N + n];
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) atomicAdd(a.ws_e0 + b * a.H + h, s);
  }
}

// ---- finalize: da (reverse cumulative sum, warp Hillis-Steele scan in fp64), ddt, dA, ddt_bias ----------
__global__ void __launch_bounds__(128) ssd_bwd_finalize_kernel(SsdArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B * a.H) return;
  const int b = warp / a.H, h = warp % a.H;
  const float Ah = a.A[h];
  const float* alpha = a.ws_alpha + (int64_t)warp * a.L;
  const float* gamma = a.ws_gamma + (int64_t)warp * a.L;
  double carry = a.ws_e0 ? (double)a.ws_e0[warp] : 0.0;  // da_{t+1} entering from the right
  double dA_acc = 0.0, dtb_acc = 0.0;
  for (int base = ((a.L + 31) / 32 - 1) * 32; base >= 0; base -= 32) {
    const int t = base + (31 - lane);  // lane 0 holds the right-most token of this block
    const bool valid = t < a.L;
    float dtv = 0.f, gm = 0.f;
    double v = 0.0;
    bool brk = false;  // token t+1 starts a new packed sequence => nothing flows from the right into t
    if (valid) {
      dtv = dt_xform(a, b, t, h);
      gm = gamma[t];
      v = (double)alpha[t] - (double)dtv * (double)gm;
      if (a.seq_idx && t + 1 < a.L)
        brk = a.seq_idx[b * a.s_b + (t + 1) * a.s_l] != a.seq_idx[b * a.s_b + t * a.s_l];
    }
    // segmented inclusive scan over lanes (lane order = right to left in time)
    double sum = valid ? v : 0.0;  // lanes past the end are neutral
    bool flag = brk;               // a flagged lane does not accept carry-in from lower lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, sum, o);
      const int fup = __shfl_up_sync(0xffffffffu, (int)flag, o);
      if (lane >= o) {
        if (!flag) sum += up;
        flag = flag || (fup != 0);
      }
    }
    // carry from the previously processed (later-in-time) block applies to lanes with no break to their right
    if (!flag) sum += carry;
    double da = sum;
    // a token that itself starts a new packed sequence has a constant decay (0): its da is 0 by definition
    if (valid && a.seq_idx && t > 0 && a.seq_idx[b * a.s_b + t * a.s_l] != a.seq_idx[b * a.s_b + (t - 1) * a.s_l]) {
      // the identity already yields ~0 here; force it to kill rounding noise, but keep `sum` flowing left? no:
      // nothing flows left across the boundary either (decay 0), which is what brk of lane+1 encodes.
      da = 0.0;
    }
    if (valid) {
      const float ddt = gm + Ah * (float)da;
      const float graw = ddt * dt_xform_grad(a, b, t, h);
      st_any(a.ddt, a.dt_dtype, b * a.ddt_b + t * a.ddt_l + h * a.ddt_h, graw);
      dA_acc += (double)dtv * da;
      dtb_acc += (double)graw;
    }
    carry = __shfl_sync(0xffffffffu, sum, 31);  // formal da of this block's left-most token
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dA_acc += __shfl_xor_sync(0xffffffffu, dA_acc, o);
    dtb_acc += __shfl_xor_sync(0xffffffffu, dtb_acc, o);
  }
  if (lane == 0) {
    if (a.dA_part) a.dA_part[warp] = (float)dA_acc;
    if (a.ddtb_part) a.ddtb_part[warp] = (float)dtb_acc;
  }
}

// ---- host ------------------------------------------------------------------------------------------------
int fill_common(SsdArgs& a, const omni_tensor_t& x, const omni_tensor_t& dt, const omni_tensor_t& A,
                const omni_tensor_t& Bm, const omni_tensor_t& Cm, const omni_tensor_t& D, const omni_tensor_t& z,
                const omni_tensor_t& dt_bias, const omni_tensor_t& init, const omni_tensor_t& seq_idx, int dt_softplus,
                float dt_min, float dt_max) {
  OMNI_CHECK(present(x) && x.ndim == 4 && is_float_dtype(x.dtype), OMNI_BAD_SHAPE, "ssd: x must be (B, L, H, P)");
  const int64_t Bsz = x.shape[0], L = x.shape[1], H = x.shape[2], P = x.shape[3];
  OMNI_CHECK(P <= 1 || x.stride[3] == 1, OMNI_BAD_STRIDE, "ssd: x headdim must be contiguous");
  OMNI_CHECK(present(dt) && shape_is(dt, 3, Bsz, L, H) && is_float_dtype(dt.dtype), OMNI_BAD_SHAPE,
             "ssd: dt must be (B, L, H)");
  OMNI_CHECK(present(A) && shape_is(A, 1, H) && A.dtype == OMNI_F32 && (H <= 1 || A.stride[0] == 1), OMNI_BAD_SHAPE,
             "ssd: A must be contiguous fp32 (H)");
  OMNI_CHECK(present(Bm) && Bm.ndim == 4 && Bm.shape[0] == Bsz && Bm.shape[1] == L && is_float_dtype(Bm.dtype),
             OMNI_BAD_SHAPE, "ssd: B must be (B, L, G, N)");
  const int64_t G = Bm.shape[2], N = Bm.shape[3];
  OMNI_CHECK(G > 0 && H % G == 0, OMNI_BAD_SHAPE, "ssd: nheads must be divisible by ngroups");
  OMNI_CHECK(shape_is(Cm, 4, Bsz, L, G, N) && Cm.dtype == Bm.dtype, OMNI_BAD_SHAPE, "ssd: C must match B");
  OMNI_CHECK(Bm.stride[3] == 1 && Cm.stride[3] == 1, OMNI_BAD_STRIDE, "ssd: B/C d_state must be contiguous");
  OMNI_CHECK(N == 16 || N == 32 || N == 64 || N == 128 || N == 256, OMNI_UNSUPPORTED,
             "ssd: d_state must be one of 16/32/64/128/256 (got %lld)", (long long)N);
  a.x = x.data; a.dt = dt.data; a.A = static_cast<const float*>(A.data); a.Bm = Bm.data; a.Cm = Cm.data;
  a.B = (int)Bsz; a.L = (int)L; a.H = (int)H; a.P = (int)P; a.G = (int)G; a.N = (int)N;
  a.x_b = x.stride[0]; a.x_l = x.stride[1]; a.x_h = x.stride[2];
  a.dt_b = dt.stride[0]; a.dt_l = dt.stride[1]; a.dt_h = dt.stride[2];
  a.B_b = Bm.stride[0]; a.B_l = Bm.stride[1]; a.B_g = Bm.stride[2];
  a.C_b = Cm.stride[0]; a.C_l = Cm.stride[1]; a.C_g = Cm.stride[2];
  a.x_dtype = x.dtype; a.dt_dtype = dt.dtype; a.bc_dtype = Bm.dtype;
  if (present(D)) {
    OMNI_CHECK(is_float_dtype(D.dtype) && ((D.ndim == 1 && D.shape[0] == H) || shape_is(D, 2, H, P)), OMNI_BAD_SHAPE,
               "ssd: D must be (H) or (H, P)");
    a.D = D.data; a.D_dtype = D.dtype; a.D_h = D.stride[0]; a.D_p = D.ndim == 2 ? D.stride[1] : 0;
  }
  if (present(z)) {
    OMNI_CHECK(shape_is(z, 4, Bsz, L, H, P) && z.dtype == x.dtype && (P <= 1 || z.stride[3] == 1), OMNI_BAD_SHAPE,
               "ssd: z must match x");
    a.z = z.data; a.z_b = z.stride[0]; a.z_l = z.stride[1]; a.z_h = z.stride[2];
  }
  if (present(dt_bias)) {
    OMNI_CHECK(shape_is(dt_bias, 1, H) && is_float_dtype(dt_bias.dtype) && (H <= 1 || dt_bias.stride[0] == 1),
               OMNI_BAD_SHAPE, "ssd: dt_bias must be contiguous (H)");
    a.dt_bias = dt_bias.data; a.dtb_dtype = dt_bias.dtype;
  }
  if (present(init)) {
    OMNI_CHECK(shape_is(init, 4, Bsz, H, P, N) && is_float_dtype(init.dtype) && init.stride[3] == 1, OMNI_BAD_SHAPE,
               "ssd: initial_states must be (B, H, P, N)");
    a.init = init.data; a.init_dtype = init.dtype; a.i_b = init.stride[0]; a.i_h = init.stride[1]; a.i_p = init.stride[2];
  }
  if (present(seq_idx)) {
    OMNI_CHECK(seq_idx.dtype == OMNI_I32 && shape_is(seq_idx, 2, Bsz, L), OMNI_BAD_SHAPE,
               "ssd: seq_idx must be int32 (B, L)");
    a.seq_idx = static_cast<const int*>(seq_idx.data); a.s_b = seq_idx.stride[0]; a.s_l = seq_idx.stride[1];
  }
  a.dt_softplus = dt_softplus; a.dt_min = dt_min; a.dt_max = dt_max;
  return OMNI_OK;
}

bool contig_f32(const omni_tensor_t& t, std::initializer_list<int64_t> shape) {
  if (t.dtype != OMNI_F32 || t.ndim != (int)shape.size()) return false;
  int i = 0;
  for (int64_t s : shape) if (t.shape[i++] != s) return false;
  int64_t exp = 1;
  for (int d = t.ndim - 1; d >= 0; --d) {
    if (t.shape[d] != 1 && t.stride[d] != exp) return false;
    exp *= t.shape[d];
  }
  return true;
}

template <bool REVERSE>
int launch_rn(const SsdArgs& a, cudaStream_t s) {
  dim3 grid((a.P + kPB - 1) / kPB, a.H, a.B);
  switch (a.N) {
    case 16: ssd_rn_kernel<2, REVERSE><<<grid, 128, 0, s>>>(a); break;
    case 32: ssd_rn_kernel<4, REVERSE><<<grid, 128, 0, s>>>(a); break;
    case 64: ssd_rn_kernel<8, REVERSE><<<grid, 128, 0, s>>>(a); break;
    case 128: ssd_rn_kernel<16, REVERSE><<<grid, 128, 0, s>>>(a); break;
    default: return set_error(OMNI_UNSUPPORTED, "ssd recurrent: d_state %d not instantiated", a.N);
  }
  OMNI_CUDA_LAUNCH_CHECK("ssd_rn_kernel");
  return OMNI_OK;
}
template <bool REVERSE>
int launch_rp(const SsdArgs& a, cudaStream_t s) {
  switch (a.N) {
    case 16: { dim3 grid((a.P + 127) / 128, a.H, a.B); ssd_rp_kernel<16, 8, REVERSE><<<grid, 128, 0, s>>>(a); break; }
    case 32: { dim3 grid((a.P + 63) / 64, a.H, a.B); ssd_rp_kernel<32, 4, REVERSE><<<grid, 128, 0, s>>>(a); break; }
    case 64: { dim3 grid((a.P + 31) / 32, a.H, a.B); ssd_rp_kernel<64, 2, REVERSE><<<grid, 128, 0, s>>>(a); break; }
    case 128: { dim3 grid((a.P + 15) / 16, a.H, a.B); ssd_rp_kernel<128, 1, REVERSE><<<grid, 128, 0, s>>>(a); break; }
    default: return set_error(OMNI_UNSUPPORTED, "ssd recurrent: d_state %d not instantiated", a.N);
  }
  OMNI_CUDA_LAUNCH_CHECK("ssd_rp_kernel");
  return OMNI_OK;
}

}  // namespace

// exported to ssd_dispatch.cu
int ssd_recurrent_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s) {
  SsdArgs a{};
  if (int rc = fill_common(a, p->x, p->dt, p->A, p->B, p->C, p->D, p->z, p->dt_bias, p->initial_states, p->seq_idx,
                           p->dt_softplus, p->dt_min, p->dt_max))
    return rc;
  const omni_tensor_t& o = p->out;
  OMNI_CHECK(present(o) && shape_is(o, 4, a.B, a.L, a.H, a.P) && o.dtype == p->x.dtype && (a.P <= 1 || o.stride[3] == 1),
             OMNI_BAD_SHAPE, "ssd: out must match x");
  a.out = o.data; a.o_b = o.stride[0]; a.o_l = o.stride[1]; a.o_h = o.stride[2];
  if (present(p->final_states)) {
    OMNI_CHECK(contig_f32(p->final_states, {a.B, a.H, a.P, a.N}), OMNI_BAD_SHAPE,
               "ssd: final_states must be contiguous fp32 (B, H, P, N)");
    a.fin = static_cast<float*>(p->final_states.data);
  }
  OMNI_CHECK(a.N <= 128, OMNI_UNSUPPORTED, "ssd recurrent: d_state 256 not instantiated");
  if (a.B == 0 || a.H == 0 || a.P == 0) return OMNI_OK;
  return launch_rn<false>(a, s);
}

int ssd_recurrent_bwd(const omni_ssd_bwd_params_t* p, cudaStream_t s) {
  SsdArgs a{};
  if (int rc = fill_common(a, p->x, p->dt, p->A, p->B, p->C, p->D, p->z, p->dt_bias, p->initial_states, p->seq_idx,
                           p->dt_softplus, p->dt_min, p->dt_max))
    return rc;
  OMNI_CHECK(a.N <= 128, OMNI_UNSUPPORTED, "ssd recurrent: d_state 256 not instantiated");
  const omni_tensor_t &g = p->dout, &dx = p->dx;
  OMNI_CHECK(present(g) && shape_is(g, 4, a.B, a.L, a.H, a.P) && g.dtype == p->x.dtype && (a.P <= 1 || g.stride[3] == 1),
             OMNI_BAD_SHAPE, "ssd bwd: dout must match x");
  OMNI_CHECK(present(dx) && shape_is(dx, 4, a.B, a.L, a.H, a.P) && dx.dtype == p->x.dtype && (a.P <= 1 || dx.stride[3] == 1),
             OMNI_BAD_SHAPE, "ssd bwd: dx must match x");
  a.dout = g.data; a.g_b = g.stride[0]; a.g_l = g.stride[1]; a.g_h = g.stride[2];
  a.dx = dx.data; a.o_b = dx.stride[0]; a.o_l = dx.stride[1]; a.o_h = dx.stride[2];
  OMNI_CHECK(present(p->ddt) && shape_is(p->ddt, 3, a.B, a.L, a.H) && p->ddt.dtype == p->dt.dtype, OMNI_BAD_SHAPE,
             "ssd bwd: ddt must match dt");
  a.ddt = p->ddt.data; a.ddt_b = p->ddt.stride[0]; a.ddt_l = p->ddt.stride[1]; a.ddt_h = p->ddt.stride[2];
  OMNI_CHECK(contig_f32(p->dB, {a.B, a.L, a.G, a.N}) && contig_f32(p->dC, {a.B, a.L, a.G, a.N}), OMNI_BAD_SHAPE,
             "ssd bwd: dB/dC must be contiguous fp32 (B, L, G, N)");
  a.dB = static_cast<float*>(p->dB.data); a.dC = static_cast<float*>(p->dC.data);
  if (present(p->dfinal_states)) {
    OMNI_CHECK(contig_f32(p->dfinal_states, {a.B, a.H, a.P, a.N}), OMNI_BAD_SHAPE,
               "ssd bwd: dfinal_states must be contiguous fp32 (B, H, P, N)");
    a.dfin = static_cast<const float*>(p->dfinal_states.data);
  }
  if (present(p->dinitial_states)) {
    OMNI_CHECK(contig_f32(p->dinitial_states, {a.B, a.H, a.P, a.N}), OMNI_BAD_SHAPE,
               "ssd bwd: dinitial_states must be contiguous fp32 (B, H, P, N)");
    a.dinit = static_cast<float*>(p->dinitial_states.data);
  }
  OMNI_CHECK(contig_f32(p->dA_part, {a.B, a.H}) && contig_f32(p->ddt_bias_part, {a.B, a.H}) &&
                 contig_f32(p->dD_part, {a.B, a.H, a.P}),
             OMNI_BAD_SHAPE, "ssd bwd: dA_part/ddt_bias_part must be fp32 (B, H), dD_part fp32 (B, H, P)");
  a.dA_part = static_cast<float*>(p->dA_part.data); a.ddtb_part = static_cast<float*>(p->ddt_bias_part.data);
  a.dD_part = static_cast<float*>(p->dD_part.data);
  const int64_t need = omni_ssd_bwd_workspace_elems(a.B, a.L, a.H, a.P, a.N);
  OMNI_CHECK(present(p->workspace) && p->workspace.dtype == OMNI_F32 && p->workspace.ndim == 1 &&
                 p->workspace.shape[0] >= need && p->workspace.stride[0] == 1,
             OMNI_BAD_SHAPE, "ssd bwd: workspace must be contiguous fp32 with >= %lld elements (zeroed)", (long long)need);
  float* ws = static_cast<float*>(p->workspace.data);
  a.ws_alpha = ws; a.ws_gamma = ws + (int64_t)a.B * a.H * a.L; a.ws_e0 = ws + 2 * (int64_t)a.B * a.H * a.L;
  if (a.B == 0 || a.H == 0 || a.P == 0 || a.L == 0) return OMNI_OK;
  if (present(p->z)) {
    // dz needs the un-gated y: one forward sweep that writes dz = dout * y * silu'(z) instead of out
    const omni_tensor_t& dz = p->dz;
    OMNI_CHECK(present(dz) && shape_is(dz, 4, a.B, a.L, a.H, a.P) && dz.dtype == p->x.dtype && (a.P <= 1 || dz.stride[3] == 1),
               OMNI_BAD_SHAPE, "ssd bwd: dz must match z");
    SsdArgs f = a;
    f.dz = dz.data; f.dz_b = dz.stride[0]; f.dz_l = dz.stride[1]; f.dz_h = dz.stride[2];
    f.fin = nullptr; f.dfin = nullptr;
    if (int rc = launch_rn<false>(f, s)) return rc;
  }
  if (int rc = launch_rn<true>(a, s)) return rc;   // dx, dD, dinitial_states
  if (int rc = launch_rp<false>(a, s)) return rc;  // dC, alpha, e0
  if (int rc = launch_rp<true>(a, s)) return rc;   // dB, gamma
  const int warps = a.B * a.H;
  ssd_bwd_finalize_kernel<<<(warps * 32 + 127) / 128, 128, 0, s>>>(a);
  OMNI_CUDA_LAUNCH_CHECK("ssd_bwd_finalize_kernel");
  return OMNI_OK;
}

}  // namespace omni

extern "C" int64_t omni_ssd_bwd_workspace_elems(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim,
                                                int64_t dstate) {
  (void)headdim; (void)dstate;
  return 2 * batch * nheads * seqlen + batch * nheads;
}
