// Mamba-1 selective scan (selective_scan_fn), forward.  Exact fp32 sequential recurrence per (batch, channel).
//
// Replaces selective_scan_cuda.fwd of mamba_ssm==2.2.2 (csrc/selective_scan/*), reachable in OmniMamba only
// with ssm_cfg.layer == "Mamba1" (/root/reference/models/stage2/mixer_seq_simple.py:197-201).
// Arithmetic: SURVEY.md A.6:  x_t = exp(d_t A) x_{t-1} + d_t B_t u_t ;  y_t = <x_t, C_t> + D u_t ; y *= silu(z).
//
// Layout: u/delta/z/out are (B, D, L) with L contiguous, B/C are (B, G, N, L).  A CTA owns 64 channels of one
// group; a thread owns one channel and its N-vector state in registers.  32-token tiles are staged through
// shared memory with token-contiguous (coalesced) global accesses; B_t/C_t tiles are shared by the CTA.
#include "common.cuh"

namespace omni {
namespace {

constexpr int kCH = 64;
constexpr int kT = 32;

struct ScanArgs {
  const void* u; const void* delta; const float* A; const void* Bm; const void* Cm; const void* D; const void* z;
  const void* delta_bias; void* out; float* last;
  int64_t u_b, u_d, dl_b, dl_d, z_b, z_d, o_b, o_d;
  int64_t A_d, A_n, B_b, B_g, B_n, C_b, C_g, C_n;
  int B, Dm, L, N, G;
  int io_dtype, dl_dtype, bc_dtype, D_dtype, db_dtype;
  int softplus;
};

template <int MAXN>
__global__ void __launch_bounds__(kCH) selscan_fwd_kernel(ScanArgs a) {
  __shared__ float us[kCH][kT + 1], dls[kCH][kT + 1], ys[kCH][kT + 1];
  __shared__ float bs[MAXN][kT], cs[MAXN][kT];
  const int tid = threadIdx.x;
  const int Dg = a.Dm / a.G;
  const int g = blockIdx.y, b = blockIdx.z;
  const int c0 = blockIdx.x * kCH;           // channel offset inside the group
  const int d = g * Dg + c0 + tid;
  const bool dvalid = c0 + tid < Dg;
  const int nch = min(kCH, Dg - c0);

  float An[MAXN], X[MAXN];
#pragma unroll
  for (int n = 0; n < MAXN; ++n) {
    An[n] = (dvalid && n < a.N) ? a.A[d * a.A_d + n * a.A_n] : 0.f;
    X[n] = 0.f;
  }
  const float Dv = (a.D && dvalid) ? ld_any(a.D, a.D_dtype, d) : 0.f;
  const float db = (a.delta_bias && dvalid) ? ld_any(a.delta_bias, a.db_dtype, d) : 0.f;
  for (int t0 = 0; t0 < a.L; t0 += kT) {
    const int tn = min(kT, a.L - t0);
    __syncthreads();
    for (int i = tid; i < nch * kT; i += kCH) {
      const int c = i / kT, tt = i % kT;
      if (tt < tn) {
        const int dd = g * Dg + c0 + c;
        us[c][tt] = ld_any(a.u, a.io_dtype, b * a.u_b + dd * a.u_d + t0 + tt);
        dls[c][tt] = ld_any(a.delta, a.dl_dtype, b * a.dl_b + dd * a.dl_d + t0 + tt);
      }
    }
    for (int i = tid; i < a.N * kT; i += kCH) {
      const int n = i / kT, tt = i % kT;
      if (tt < tn) {
        bs[n][tt] = ld_any(a.Bm, a.bc_dtype, b * a.B_b + g * a.B_g + n * a.B_n + t0 + tt);
        cs[n][tt] = ld_any(a.Cm, a.bc_dtype, b * a.C_b + g * a.C_g + n * a.C_n + t0 + tt);
      }
    }
    __syncthreads();
    if (dvalid) {
      for (int tt = 0; tt < tn; ++tt) {
        float dl = dls[tid][tt] + db;
        if (a.softplus) dl = softplus_f(dl);
        const float uu = us[tid][tt];
        const float du = dl * uu;
        float y = 0.f;
#pragma unroll
        for (int n = 0; n < MAXN; ++n) {
          if (n < a.N) {
            X[n] = __expf(dl * An[n]) * X[n] + du * bs[n][tt];
            y += X[n] * cs[n][tt];
          }
        }
        ys[tid][tt] = y + Dv * uu;
      }
    }
    __syncthreads();
    for (int i = tid; i < nch * kT; i += kCH) {
      const int c = i / kT, tt = i % kT;
      if (tt < tn) {
        const int dd = g * Dg + c0 + c;
        float v = ys[c][tt];
        if (a.z) v *= silu_f(ld_any(a.z, a.io_dtype, b * a.z_b + dd * a.z_d + t0 + tt));
        st_any(a.out, a.io_dtype, b * a.o_b + dd * a.o_d + t0 + tt, v);
      }
    }
  }
  if (a.last && dvalid) {
#pragma unroll
    for (int n = 0; n < MAXN; ++n)
      if (n < a.N) a.last[((int64_t)b * a.Dm + d) * a.N + n] = X[n];
  }
}

}  // namespace
}  // namespace omni

using namespace omni;

extern "C" int omni_selective_scan_fwd(const omni_selscan_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t& u = p->u;
  OMNI_CHECK(present(u) && u.ndim == 3 && is_float_dtype(u.dtype), OMNI_BAD_SHAPE, "selective_scan: u must be (B, D, L)");
  const int64_t Bsz = u.shape[0], Dm = u.shape[1], L = u.shape[2];
  auto chk = [&](const omni_tensor_t& t, const char* name, bool req, int dtype) -> int {
    if (!present(t)) { OMNI_CHECK(!req, OMNI_BAD_SHAPE, "selective_scan: %s required", name); return OMNI_OK; }
    OMNI_CHECK(shape_is(t, 3, Bsz, Dm, L) && is_float_dtype(t.dtype) && (dtype < 0 || t.dtype == dtype) &&
                   (L <= 1 || t.stride[2] == 1),
               OMNI_BAD_SHAPE, "selective_scan: %s must be (B, D, L) with contiguous L", name);
    return OMNI_OK;
  };
  if (int rc = chk(u, "u", true, -1)) return rc;
  if (int rc = chk(p->delta, "delta", true, -1)) return rc;
  if (int rc = chk(p->z, "z", false, u.dtype)) return rc;
  if (int rc = chk(p->out, "out", true, u.dtype)) return rc;
  OMNI_CHECK(present(p->A) && p->A.ndim == 2 && p->A.shape[0] == Dm && p->A.dtype == OMNI_F32, OMNI_BAD_SHAPE,
             "selective_scan: A must be fp32 (D, N)");
  const int64_t N = p->A.shape[1];
  OMNI_CHECK(N >= 1 && N <= 64, OMNI_UNSUPPORTED, "selective_scan: d_state must be <= 64");
  OMNI_CHECK(present(p->B) && p->B.ndim == 4 && p->B.shape[0] == Bsz && p->B.shape[2] == N && p->B.shape[3] == L &&
                 is_float_dtype(p->B.dtype) && (L <= 1 || p->B.stride[3] == 1),
             OMNI_BAD_SHAPE, "selective_scan: B must be (B, G, N, L) with contiguous L");
  const int64_t G = p->B.shape[1];
  OMNI_CHECK(G > 0 && Dm % G == 0, OMNI_BAD_SHAPE, "selective_scan: dim %% ngroups != 0");
  OMNI_CHECK(shape_is(p->C, 4, Bsz, G, N, L) && p->C.dtype == p->B.dtype && (L <= 1 || p->C.stride[3] == 1),
             OMNI_BAD_SHAPE, "selective_scan: C must match B");
  auto chk1 = [&](const omni_tensor_t& t, const char* name) -> int {
    if (!present(t)) return OMNI_OK;
    OMNI_CHECK(shape_is(t, 1, Dm) && is_float_dtype(t.dtype) && (Dm <= 1 || t.stride[0] == 1), OMNI_BAD_SHAPE,
               "selective_scan: %s must be contiguous (D)", name);
    return OMNI_OK;
  };
  if (int rc = chk1(p->D, "D")) return rc;
  if (int rc = chk1(p->delta_bias, "delta_bias")) return rc;
  ScanArgs a{};
  a.u = u.data; a.delta = p->delta.data; a.A = static_cast<const float*>(p->A.data); a.Bm = p->B.data; a.Cm = p->C.data;
  a.D = p->D.data; a.z = p->z.data; a.delta_bias = p->delta_bias.data; a.out = p->out.data;
  if (present(p->last_state)) {
    OMNI_CHECK(p->last_state.dtype == OMNI_F32 && shape_is(p->last_state, 3, Bsz, Dm, N) && p->last_state.stride[2] == 1 &&
                   p->last_state.stride[1] == N && p->last_state.stride[0] == Dm * N,
               OMNI_BAD_SHAPE, "selective_scan: last_state must be contiguous fp32 (B, D, N)");
    a.last = static_cast<float*>(p->last_state.data);
  }
  a.u_b = u.stride[0]; a.u_d = u.stride[1]; a.dl_b = p->delta.stride[0]; a.dl_d = p->delta.stride[1];
  if (present(p->z)) { a.z_b = p->z.stride[0]; a.z_d = p->z.stride[1]; }
  a.o_b = p->out.stride[0]; a.o_d = p->out.stride[1];
  a.A_d = p->A.stride[0]; a.A_n = p->A.stride[1];
  a.B_b = p->B.stride[0]; a.B_g = p->B.stride[1]; a.B_n = p->B.stride[2];
  a.C_b = p->C.stride[0]; a.C_g = p->C.stride[1]; a.C_n = p->C.stride[2];
  a.B = (int)Bsz; a.Dm = (int)Dm; a.L = (int)L; a.N = (int)N; a.G = (int)G;
  a.io_dtype = u.dtype; a.dl_dtype = p->delta.dtype; a.bc_dtype = p->B.dtype; a.D_dtype = p->D.dtype;
  a.db_dtype = p->delta_bias.dtype; a.softplus = p->delta_softplus;
  if (Bsz == 0 || Dm == 0) return OMNI_OK;
  const int Dg = (int)(Dm / G);
  dim3 grid((Dg + kCH - 1) / kCH, (unsigned)G, (unsigned)Bsz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 16) selscan_fwd_kernel<16><<<grid, kCH, 0, s>>>(a);
  else if (N <= 32) selscan_fwd_kernel<32><<<grid, kCH, 0, s>>>(a);
  else selscan_fwd_kernel<64><<<grid, kCH, 0, s>>>(a);
  OMNI_CUDA_LAUNCH_CHECK("selscan_fwd_kernel");
  return OMNI_OK;
}

extern "C" int omni_selective_scan_bwd(const omni_selscan_bwd_params_t* p, void* stream) {
  (void)p; (void)stream;
  return set_error(OMNI_UNSUPPORTED, "selective_scan backward is not implemented yet (Mamba-1 is off the OmniMamba default path)");
}
