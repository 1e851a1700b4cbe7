// Mamba-1 selective scan (selective_scan_fn), forward and backward.  Exact fp32 sequential recurrence per (batch, channel).
//
// Replaces selective_scan_cuda.fwd of mamba_ssm==2.2.2 (csrc/selective_scan/*), reachable in OmniMamba only
// with ssm_cfg.layer == "Mamba1" (/root/reference/models/stage2/mixer_seq_simple.py:197-201).
// Arithmetic: SURVEY.md A.6:  x_t = exp(d_t A) x_{t-1} + d_t B_t u_t ;  y_t = <x_t, C_t> + D u_t ; y *= silu(z).
//
// Layout: u/delta/z/out are (B, D, L) with L contiguous, B/C are (B, G, N, L).  A CTA owns 64 channels of one
// group; a thread owns one channel and its N-vector state in registers.  32-token tiles are staged through
// shared memory with token-contiguous (coalesced) global accesses; B_t/C_t tiles are shared by the CTA.
#include <mutex>

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


// ---- backward ------------------------------------------------------------------------------------------------------
// With dl_t = softplus(delta_t + bias), a_t[n] = exp(dl_t A[n]), y_t = <X_t, C_t> + D u_t, out_t = y_t silu(z_t):
//   dy_t = dout_t silu(z_t)            dz_t = dout_t y_t silu'(z_t)
//   G_t[n] = dy_t C_t[n] + a_{t+1}[n] G_{t+1}[n]                      (gradient w.r.t. X_t, reverse recurrence)
//   du_t = sum_n G_t[n] dl_t B_t[n] + D dy_t        ddl_t = sum_n G_t[n] (u_t B_t[n] + A[n] a_t[n] X_{t-1}[n])
//   dB_t[n] += G_t[n] dl_t u_t   dC_t[n] += dy_t X_t[n]   (over the channels of the group)
//   dA[n] += G_t[n] dl_t a_t[n] X_{t-1}[n]            dD += dy_t u_t            ddelta_t = ddl_t softplus'(delta_t + bias)
// Pass 1 walks the sequence forward and parks the state entering every tile of T tokens in the workspace; pass 2 walks the
// tiles last to first, rebuilds the T states of the tile in shared memory from its checkpoint and runs the reverse recurrence.
struct ScanBwdArgs {
  ScanArgs f;
  const void* dout; void* du; void* ddelta; void* dz; float* dB; float* dC; float* dA_part; float* dD_part; float* ddb_part;
  float* ckpt;
  int64_t do_b, do_d, du_b, du_d, ddl_b, ddl_d, dz_b, dz_d;
};

template <int MAXN, int T>
__global__ void __launch_bounds__(kCH) selscan_bwd_kernel(ScanBwdArgs p) {
  const ScanArgs& a = p.f;
  extern __shared__ float xs[];  // [T][MAXN][kCH] states X_t of the tile
  __shared__ float us[kCH][T + 1], dls[kCH][T + 1], zs[kCH][T + 1], dos[kCH][T + 1];
  __shared__ float bs[MAXN][T], cs[MAXN][T];
  const int tid = threadIdx.x, lane = tid & 31;
  const int Dg = a.Dm / a.G;
  const int g = blockIdx.y, b = blockIdx.z;
  const int c0 = blockIdx.x * kCH;
  const int d = g * Dg + c0 + tid;
  const bool dvalid = c0 + tid < Dg;
  const int nch = min(kCH, Dg - c0);
  const int ntiles = (a.L + T - 1) / T;

  float An[MAXN], X[MAXN];
#pragma unroll
  for (int n = 0; n < MAXN; ++n) {
    An[n] = (dvalid && n < a.N) ? a.A[d * a.A_d + n * a.A_n] : 0.f;
    X[n] = 0.f;
  }
  const float Dv = (a.D && dvalid) ? ld_any(a.D, a.D_dtype, d) : 0.f;
  const float db = (a.delta_bias && dvalid) ? ld_any(a.delta_bias, a.db_dtype, d) : 0.f;
  float* ck = p.ckpt + ((int64_t)b * a.Dm + (dvalid ? d : 0)) * ntiles * a.N;

  auto load_tile = [&](int t0, int tn, bool bwd) {
    for (int i = tid; i < nch * T; i += kCH) {
      const int c = i / T, tt = i % T;
      if (tt < tn) {
        const int dd = g * Dg + c0 + c;
        us[c][tt] = ld_any(a.u, a.io_dtype, b * a.u_b + dd * a.u_d + t0 + tt);
        dls[c][tt] = ld_any(a.delta, a.dl_dtype, b * a.dl_b + dd * a.dl_d + t0 + tt);
        if (bwd) {
          dos[c][tt] = ld_any(p.dout, a.io_dtype, b * p.do_b + dd * p.do_d + t0 + tt);
          zs[c][tt] = a.z ? ld_any(a.z, a.io_dtype, b * a.z_b + dd * a.z_d + t0 + tt) : 0.f;
        }
      }
    }
    for (int i = tid; i < a.N * T; i += kCH) {
      const int n = i / T, tt = i % T;
      if (tt < tn) {
        bs[n][tt] = ld_any(a.Bm, a.bc_dtype, b * a.B_b + g * a.B_g + n * a.B_n + t0 + tt);
        if (bwd) cs[n][tt] = ld_any(a.Cm, a.bc_dtype, b * a.C_b + g * a.C_g + n * a.C_n + t0 + tt);
      }
    }
  };
  auto dl_of = [&](float raw) { return a.softplus ? softplus_f(raw) : raw; };

  // ---- pass 1: checkpoints
  for (int tile = 0; tile < ntiles; ++tile) {
    const int t0 = tile * T, tn = min(T, a.L - t0);
    __syncthreads();
    load_tile(t0, tn, false);
    __syncthreads();
    if (dvalid) {
#pragma unroll
      for (int n = 0; n < MAXN; ++n)
        if (n < a.N) ck[(int64_t)tile * a.N + n] = X[n];
      if (tile + 1 < ntiles) {  // (the state after the last tile is not needed)
        for (int tt = 0; tt < tn; ++tt) {
          const float dl = dl_of(dls[tid][tt] + db), du = dl * us[tid][tt];
#pragma unroll
          for (int n = 0; n < MAXN; ++n)
            if (n < a.N) X[n] = __expf(dl * An[n]) * X[n] + du * bs[n][tt];
        }
      }
    }
  }

  // ---- pass 2: reverse over tiles
  float Gc[MAXN], dAn[MAXN];
#pragma unroll
  for (int n = 0; n < MAXN; ++n) { Gc[n] = 0.f; dAn[n] = 0.f; }
  float dDv = 0.f, ddb = 0.f;
  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int t0 = tile * T, tn = min(T, a.L - t0);
    __syncthreads();
    load_tile(t0, tn, true);
    __syncthreads();
    float X0[MAXN];
#pragma unroll
    for (int n = 0; n < MAXN; ++n) X0[n] = (dvalid && n < a.N) ? ck[(int64_t)tile * a.N + n] : 0.f;
    // rebuild X_t, t in the tile; dy_t and dz_t on the way (dos <- dy, zs <- dz)
#pragma unroll
    for (int n = 0; n < MAXN; ++n) X[n] = X0[n];
    for (int tt = 0; tt < tn; ++tt) {
      const float uu = us[tid][tt];
      const float dl = dl_of(dls[tid][tt] + db), du = dl * uu;
      float y = 0.f;
#pragma unroll
      for (int n = 0; n < MAXN; ++n) {
        if (n < a.N) {
          X[n] = __expf(dl * An[n]) * X[n] + du * bs[n][tt];
          y += X[n] * cs[n][tt];
          xs[(tt * MAXN + n) * kCH + tid] = X[n];
        }
      }
      y += Dv * uu;
      const float dout = dvalid ? dos[tid][tt] : 0.f;
      if (a.z) {
        const float zv = zs[tid][tt];
        dos[tid][tt] = dout * silu_f(zv);
        zs[tid][tt] = dout * y * dsilu_f(zv);
      }
    }
    // reverse recurrence; du -> us, ddelta -> dls
    for (int tt = tn - 1; tt >= 0; --tt) {
      const float uu = us[tid][tt], raw = dls[tid][tt] + db;
      const float dl = dl_of(raw), dy = dvalid ? dos[tid][tt] : 0.f;
      float du = Dv * dy, ddl = 0.f;
      dDv += dy * uu;
#pragma unroll
      for (int n = 0; n < MAXN; ++n) {
        if (n < a.N) {  // (warp-uniform: the shuffles below are executed by all lanes)
          const float Xt = dvalid ? xs[(tt * MAXN + n) * kCH + tid] : 0.f;
          const float Xp = tt > 0 ? xs[((tt - 1) * MAXN + n) * kCH + tid] : X0[n];
          const float Bv = bs[n][tt];
          const float G = dy * cs[n][tt] + Gc[n];
          const float at = __expf(dl * An[n]);
          const float gax = G * at * Xp;
          ddl += G * uu * Bv + An[n] * gax;
          dAn[n] += dl * gax;
          du += G * dl * Bv;
          Gc[n] = at * G;
          const float sB = warp_sum(dvalid ? G * dl * uu : 0.f), sC = warp_sum(dy * Xt);
          if (lane == 0) {
            const int64_t o = ((int64_t)(b * a.G + g) * a.N + n) * a.L + t0 + tt;
            atomicAdd(p.dB + o, sB);
            atomicAdd(p.dC + o, sC);
          }
        }
      }
      float dd = ddl;
      if (a.softplus && raw <= 20.f) dd *= sigmoid_f(raw);
      ddb += dd;
      us[tid][tt] = du;
      dls[tid][tt] = dd;
    }
    __syncthreads();
    for (int i = tid; i < nch * T; i += kCH) {
      const int c = i / T, tt = i % T;
      if (tt < tn) {
        const int dd = g * Dg + c0 + c;
        st_any(p.du, a.io_dtype, b * p.du_b + dd * p.du_d + t0 + tt, us[c][tt]);
        st_any(p.ddelta, a.dl_dtype, b * p.ddl_b + dd * p.ddl_d + t0 + tt, dls[c][tt]);
        if (p.dz) st_any(p.dz, a.io_dtype, b * p.dz_b + dd * p.dz_d + t0 + tt, zs[c][tt]);
      }
    }
  }
  if (dvalid) {
#pragma unroll
    for (int n = 0; n < MAXN; ++n)
      if (n < a.N) p.dA_part[((int64_t)b * a.Dm + d) * a.N + n] = dAn[n];
    if (p.dD_part) p.dD_part[(int64_t)b * a.Dm + d] = dDv;
    if (p.ddb_part) p.ddb_part[(int64_t)b * a.Dm + d] = ddb;
  }
}

constexpr int tile_tokens(int maxn) { return 256 / maxn; }  // T x MAXN = 256 states per thread in shared memory (64 KB)
inline int maxn_of(int64_t N) { return N <= 16 ? 16 : (N <= 32 ? 32 : 64); }

}  // namespace
}  // namespace omni

using namespace omni;

namespace {
// shared by forward and backward: validates (u, delta, A, B, C, D, z, delta_bias) and fills the common launch arguments
template <typename Params>
int fill_scan_args(const Params* p, const omni_tensor_t& out, ScanArgs& a) {
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
  if (int rc = chk(out, "out", true, u.dtype)) return rc;
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
  a.u = u.data; a.delta = p->delta.data; a.A = static_cast<const float*>(p->A.data); a.Bm = p->B.data; a.Cm = p->C.data;
  a.D = p->D.data; a.z = p->z.data; a.delta_bias = p->delta_bias.data; a.out = out.data;
  a.u_b = u.stride[0]; a.u_d = u.stride[1]; a.dl_b = p->delta.stride[0]; a.dl_d = p->delta.stride[1];
  if (present(p->z)) { a.z_b = p->z.stride[0]; a.z_d = p->z.stride[1]; }
  a.o_b = out.stride[0]; a.o_d = out.stride[1];
  a.A_d = p->A.stride[0]; a.A_n = p->A.stride[1];
  a.B_b = p->B.stride[0]; a.B_g = p->B.stride[1]; a.B_n = p->B.stride[2];
  a.C_b = p->C.stride[0]; a.C_g = p->C.stride[1]; a.C_n = p->C.stride[2];
  a.B = (int)Bsz; a.Dm = (int)Dm; a.L = (int)L; a.N = (int)N; a.G = (int)G;
  a.io_dtype = u.dtype; a.dl_dtype = p->delta.dtype; a.bc_dtype = p->B.dtype; a.D_dtype = p->D.dtype;
  a.db_dtype = p->delta_bias.dtype; a.softplus = p->delta_softplus;
  return OMNI_OK;
}
}  // namespace

extern "C" int omni_selective_scan_fwd(const omni_selscan_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  ScanArgs a{};
  if (int rc = fill_scan_args(p, p->out, a)) return rc;
  const int64_t Bsz = a.B, Dm = a.Dm, N = a.N, G = a.G;
  if (present(p->last_state)) {
    OMNI_CHECK(p->last_state.dtype == OMNI_F32 && shape_is(p->last_state, 3, Bsz, Dm, N) && p->last_state.stride[2] == 1 &&
                   p->last_state.stride[1] == N && p->last_state.stride[0] == Dm * N,
               OMNI_BAD_SHAPE, "selective_scan: last_state must be contiguous fp32 (B, D, N)");
    a.last = static_cast<float*>(p->last_state.data);
  }
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

extern "C" int64_t omni_selective_scan_bwd_workspace_elems(int64_t batch, int64_t dim, int64_t seqlen, int64_t dstate) {
  const int T = tile_tokens(maxn_of(dstate));
  return batch * dim * ((seqlen + T - 1) / T) * dstate;
}

extern "C" int omni_selective_scan_bwd(const omni_selscan_bwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  ScanBwdArgs q{};
  if (int rc = fill_scan_args(p, p->du, q.f)) return rc;  // (du is validated like `out`: (B, D, L), dtype of u)
  const ScanArgs& a = q.f;
  const int64_t Bsz = a.B, Dm = a.Dm, L = a.L, N = a.N, G = a.G;
  auto like_u = [&](const omni_tensor_t& t, int dtype) {
    return shape_is(t, 3, Bsz, Dm, L) && t.dtype == dtype && (L <= 1 || t.stride[2] == 1);
  };
  OMNI_CHECK(present(p->dout) && like_u(p->dout, p->u.dtype), OMNI_BAD_SHAPE, "selective_scan bwd: dout must be like u");
  OMNI_CHECK(present(p->ddelta) && like_u(p->ddelta, p->delta.dtype), OMNI_BAD_SHAPE, "selective_scan bwd: ddelta must be like delta");
  OMNI_CHECK(present(p->z) == present(p->dz) && (!present(p->dz) || like_u(p->dz, p->u.dtype)), OMNI_BAD_SHAPE,
             "selective_scan bwd: dz must be like z");
  auto dense = [&](const omni_tensor_t& t, int nd, int64_t s0, int64_t s1, int64_t s2, int64_t s3) {
    if (!present(t) || t.ndim != nd || t.dtype != OMNI_F32) return false;
    const int64_t want[4] = {s0, s1, s2, s3};
    int64_t st = 1;
    for (int i = nd - 1; i >= 0; --i) {
      if (t.shape[i] != want[i] || (t.shape[i] > 1 && t.stride[i] != st)) return false;
      st *= want[i];
    }
    return true;
  };
  OMNI_CHECK(dense(p->dB, 4, Bsz, G, N, L) && dense(p->dC, 4, Bsz, G, N, L), OMNI_BAD_SHAPE,
             "selective_scan bwd: dB, dC must be contiguous fp32 (B, G, N, L), zeroed by the caller");
  OMNI_CHECK(dense(p->dA_part, 3, Bsz, Dm, N, 0), OMNI_BAD_SHAPE, "selective_scan bwd: dA_part must be contiguous fp32 (B, D, N)");
  OMNI_CHECK(!present(p->dD_part) || dense(p->dD_part, 2, Bsz, Dm, 0, 0), OMNI_BAD_SHAPE, "selective_scan bwd: dD_part (B, D) fp32");
  OMNI_CHECK(!present(p->ddelta_bias_part) || dense(p->ddelta_bias_part, 2, Bsz, Dm, 0, 0), OMNI_BAD_SHAPE,
             "selective_scan bwd: ddelta_bias_part (B, D) fp32");
  const int64_t need = omni_selective_scan_bwd_workspace_elems(Bsz, Dm, L, N);
  OMNI_CHECK(present(p->workspace) && p->workspace.dtype == OMNI_F32 && p->workspace.ndim == 1 && p->workspace.shape[0] >= need &&
                 (need <= 1 || p->workspace.stride[0] == 1),
             OMNI_BAD_SHAPE, "selective_scan bwd: workspace must hold omni_selective_scan_bwd_workspace_elems fp32 values");
  q.dout = p->dout.data; q.du = p->du.data; q.ddelta = p->ddelta.data; q.dz = p->dz.data;
  q.dB = static_cast<float*>(p->dB.data); q.dC = static_cast<float*>(p->dC.data);
  q.dA_part = static_cast<float*>(p->dA_part.data); q.dD_part = static_cast<float*>(p->dD_part.data);
  q.ddb_part = static_cast<float*>(p->ddelta_bias_part.data); q.ckpt = static_cast<float*>(p->workspace.data);
  q.do_b = p->dout.stride[0]; q.do_d = p->dout.stride[1]; q.du_b = p->du.stride[0]; q.du_d = p->du.stride[1];
  q.ddl_b = p->ddelta.stride[0]; q.ddl_d = p->ddelta.stride[1];
  if (present(p->dz)) { q.dz_b = p->dz.stride[0]; q.dz_d = p->dz.stride[1]; }
  if (Bsz == 0 || Dm == 0 || L == 0) return OMNI_OK;
  const int Dg = (int)(Dm / G);
  dim3 grid((Dg + kCH - 1) / kCH, (unsigned)G, (unsigned)Bsz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int kSmem = 256 * kCH * (int)sizeof(float);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(selscan_bwd_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(selscan_bwd_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(selscan_bwd_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  });
  if (N <= 16) selscan_bwd_kernel<16, 16><<<grid, kCH, kSmem, s>>>(q);
  else if (N <= 32) selscan_bwd_kernel<32, 8><<<grid, kCH, kSmem, s>>>(q);
  else selscan_bwd_kernel<64, 4><<<grid, kCH, kSmem, s>>>(q);
  OMNI_CUDA_LAUNCH_CHECK("selscan_bwd_kernel");
  return OMNI_OK;
}
