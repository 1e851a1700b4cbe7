// mamba_split_conv1d_scan_combined forward (Mamba2.forward path A) as ONE C-ABI call:
//   zxbcdt = [z | xBC | dt]  ->  causal conv1d + SiLU on xBC  ->  chunked SSD scan  ->  gated RMSNorm (y * silu(z))  ->  out_proj
// Replaces MambaSplitConv1dScanCombinedFn.forward of mamba_ssm==2.2.2 (one Triton/CUDA launch sequence upstream as well:
// causal_conv1d_cuda.causal_conv1d_fwd, _mamba_chunk_scan_combined_fwd, _layer_norm_fwd, F.linear).  The stages are the
// library's own kernels, launched back to back on the caller's stream; every intermediate the backward needs
// (conv output, pre-norm scan output, rstd) lives in caller-owned tensors.
#include "common.cuh"

using namespace omni;

namespace {
// columns [lo, lo + n) of a (B, L, C) tensor as a (B, L, n) view
omni_tensor_t cols(const omni_tensor_t& t, int64_t lo, int64_t n) {
  omni_tensor_t v = t;
  v.data = static_cast<char*>(t.data) + lo * t.stride[2] * dtype_size(t.dtype);
  v.shape[2] = n;
  return v;
}
// (B, L, H * P) -> (B, L, H, P)
omni_tensor_t heads(const omni_tensor_t& t, int64_t H, int64_t P) {
  omni_tensor_t v = t;
  v.ndim = 4;
  v.shape[2] = H; v.shape[3] = P;
  v.stride[2] = P * t.stride[2]; v.stride[3] = t.stride[2];
  return v;
}
// (B, L, C) -> (B, C, L): the channel-last layout causal_conv1d_fn is called with (xBC.transpose(1, 2))
omni_tensor_t transpose12(const omni_tensor_t& t) {
  omni_tensor_t v = t;
  v.shape[1] = t.shape[2]; v.shape[2] = t.shape[1];
  v.stride[1] = t.stride[2]; v.stride[2] = t.stride[1];
  return v;
}
// (B, L, C) with contiguous batch -> (B * L, C)
bool rows2d(const omni_tensor_t& t, omni_tensor_t& v) {
  if (t.shape[0] > 1 && t.stride[0] != t.shape[1] * t.stride[1]) return false;
  v = t;
  v.ndim = 2;
  v.shape[0] = t.shape[0] * t.shape[1]; v.shape[1] = t.shape[2];
  v.stride[0] = t.stride[1]; v.stride[1] = t.stride[2];
  return true;
}
}  // namespace

extern "C" int omni_split_conv1d_scan_fwd(const omni_split_conv1d_scan_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  const omni_tensor_t& zx = p->zxbcdt;
  OMNI_CHECK(present(zx) && zx.ndim == 3 && is_float_dtype(zx.dtype) && zx.stride[2] == 1, OMNI_BAD_SHAPE,
             "split_conv1d_scan: zxbcdt must be (B, L, 2 dim + 2 G N + H) with contiguous rows");
  const int64_t Bsz = zx.shape[0], L = zx.shape[1], H = p->nheads, P = p->headdim, G = p->ngroups, N = p->dstate;
  const int64_t dim = H * P, conv_dim = dim + 2 * G * N;
  OMNI_CHECK(H > 0 && P > 0 && G > 0 && N > 0 && zx.shape[2] == 2 * dim + 2 * G * N + H, OMNI_BAD_SHAPE,
             "split_conv1d_scan: zxbcdt width %lld != 2 * %lld + 2 * %lld * %lld + %lld (d_mlp lanes are not on the OmniMamba path)",
             (long long)zx.shape[2], (long long)dim, (long long)G, (long long)N, (long long)H);
  OMNI_CHECK(present(p->xbc_conv) && shape_is(p->xbc_conv, 3, Bsz, L, conv_dim) && p->xbc_conv.dtype == zx.dtype && p->xbc_conv.stride[2] == 1,
             OMNI_BAD_SHAPE, "split_conv1d_scan: xbc_conv must be (B, L, dim + 2 G N), dtype of zxbcdt");
  OMNI_CHECK(present(p->scan_out) && shape_is(p->scan_out, 3, Bsz, L, dim) && p->scan_out.stride[2] == 1, OMNI_BAD_SHAPE,
             "split_conv1d_scan: scan_out must be (B, L, dim)");
  const bool normed = present(p->rmsnorm_weight);
  // 1. depthwise causal conv + SiLU over the xBC columns, channel-last views on both sides
  {
    omni_conv1d_fwd_params_t c{};
    c.x = transpose12(cols(zx, dim, conv_dim));
    c.weight = p->conv1d_weight; c.bias = p->conv1d_bias; c.seq_idx = p->seq_idx;
    c.out = transpose12(p->xbc_conv);
    c.activation = p->activation;
    if (int rc = omni_causal_conv1d_fwd(&c, stream)) return rc;
  }
  // 2. the scan on (x, B, C) = column blocks of the conv output, dt = the last H columns of zxbcdt
  {
    omni_ssd_fwd_params_t s{};
    s.x = heads(cols(p->xbc_conv, 0, dim), H, P);
    s.B = heads(cols(p->xbc_conv, dim, G * N), G, N);
    s.C = heads(cols(p->xbc_conv, dim + G * N, G * N), G, N);
    s.dt = cols(zx, dim + conv_dim, H);
    s.A = p->A; s.D = p->D; s.dt_bias = p->dt_bias;
    if (!normed) s.z = heads(cols(zx, 0, dim), H, P);   // no norm: the gate y * silu(z) is applied inside the scan
    s.initial_states = p->initial_states; s.seq_idx = p->seq_idx;
    s.out = heads(p->scan_out, H, P);
    s.final_states = p->final_states; s.workspace = p->workspace; s.chunk_states = p->chunk_states;
    s.chunk_size = p->chunk_size; s.dt_softplus = 1; s.dt_min = p->dt_min; s.dt_max = p->dt_max; s.algo = p->algo;
    if (int rc = omni_ssd_chunk_scan_fwd(&s, stream)) return rc;
  }
  // 3. gated RMSNorm: y = rmsnorm(scan_out * silu(z)) * w   (norm_before_gate = False in OmniMamba)
  const omni_tensor_t* gemm_in = &p->scan_out;
  if (normed) {
    OMNI_CHECK(present(p->y) && shape_is(p->y, 3, Bsz, L, dim) && present(p->rstd), OMNI_BAD_SHAPE,
               "split_conv1d_scan: y (B, L, dim) and rstd are required with rmsnorm_weight");
    omni_norm_gated_fwd_params_t n{};
    omni_tensor_t z3 = cols(zx, 0, dim);
    OMNI_CHECK(rows2d(p->scan_out, n.x) && rows2d(z3, n.z) && rows2d(p->y, n.out), OMNI_BAD_STRIDE,
               "split_conv1d_scan: batch strides must be L rows");
    n.weight = p->rmsnorm_weight; n.rstd = p->rstd;
    n.eps = p->rmsnorm_eps; n.group_size = (int32_t)(dim / G); n.norm_before_gate = p->norm_before_gate; n.is_rms_norm = 1;
    if (int rc = omni_norm_gated_fwd(&n, stream)) return rc;
    gemm_in = &p->y;
  }
  // 4. out_proj (bf16 operands: the tcgen05 GEMM).  Optional: without outproj_weight the caller applies its own projection.
  if (present(p->outproj_weight)) {
    omni_gemm_params_t g{};
    OMNI_CHECK(rows2d(*gemm_in, g.a) && present(p->out) && p->out.ndim == 3 && rows2d(p->out, g.out), OMNI_BAD_STRIDE,
               "split_conv1d_scan: out must be (B, L, d_out) with batch stride L rows");
    g.b = p->outproj_weight;
    if (int rc = omni_gemm_bf16(&g, stream)) return rc;
  }
  return OMNI_OK;
}
