/*
 * omnissm.h - C ABI of libomnissm.so: the B200 (sm_100a) kernels behind the
 * mamba_ssm / causal_conv1d operator surface that hustvl/OmniMamba reaches from
 * models/stage2/mixer_seq_simple.py:15-20,30 and models/stage2/block.py:10,86-95,117.
 *
 * Conventions:
 *   - plain C: POD structs, raw device pointers, element strides; no torch types.
 *   - the caller owns ALL memory (inputs, outputs, workspaces); the library never
 *     allocates, frees or retains device memory and never synchronises the host.
 *   - every entry point launches on the caller-supplied stream (a cudaStream_t passed
 *     as void*), is re-entrant, and is CUDA-graph capturable.
 *   - return 0 on success, an omni_status_t otherwise; omni_last_error() gives the
 *     thread-local message.  No exceptions cross the boundary.
 *   - a tensor argument whose .data is NULL is "absent" (the Python None).
 *
 * Each entry point cites the reference-side Python interface it replaces; the
 * upstream wheels pinned by /root/reference/requirements.txt:12-13 bind the same
 * operations through pybind modules (selective_scan_cuda, causal_conv1d_cuda) and
 * Triton JIT kernels.
 */
#ifndef OMNISSM_H_
#define OMNISSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMNI_ABI_VERSION 8
#if defined(__GNUC__)
#define OMNI_API __attribute__((visibility("default")))
#else
#define OMNI_API
#endif
#define OMNI_MAX_DIMS 6

typedef enum {
  OMNI_OK = 0,
  OMNI_BAD_SHAPE = 1,
  OMNI_BAD_DTYPE = 2,
  OMNI_BAD_STRIDE = 3,
  OMNI_UNSUPPORTED = 4,
  OMNI_CUDA_ERROR = 5
} omni_status_t;

typedef enum { OMNI_F32 = 0, OMNI_F16 = 1, OMNI_BF16 = 2, OMNI_I32 = 3, OMNI_I64 = 4, OMNI_U8 = 5 } omni_dtype_t;

typedef enum { OMNI_ACT_NONE = 0, OMNI_ACT_SILU = 1 } omni_activation_t;

/* Strided view of device memory. Strides are in ELEMENTS and may be 0 (broadcast). */
typedef struct omni_tensor {
  void* data;
  int32_t dtype; /* omni_dtype_t */
  int32_t ndim;
  int64_t shape[OMNI_MAX_DIMS];
  int64_t stride[OMNI_MAX_DIMS];
} omni_tensor_t;

/* SSD algorithm selector (omni_ssd_params_t.algo). AUTO picks the tcgen05 chunked kernel when the
 * shape/dtype allows and the exact SIMT recurrence otherwise. */
typedef enum { OMNI_SSD_AUTO = 0, OMNI_SSD_RECURRENT = 1, OMNI_SSD_CHUNKED_TC = 2 } omni_ssd_algo_t;

/* ---- library ---------------------------------------------------------------------------- */
OMNI_API int omni_version(void);
OMNI_API const char* omni_last_error(void);
/* Comma-separated names of the device kernels this library has launched in this process
 * (evidence for "which native code ran"); the count is cumulative. */
OMNI_API int64_t omni_launch_count(void);
OMNI_API void omni_reset_launch_count(void);

/* ---- causal depthwise conv1d ------------------------------------------------------------- */
/* causal_conv1d_fn(x, weight, bias, seq_idx, initial_states, return_final_states, final_states_out,
 * activation)  [causal_conv1d/causal_conv1d_interface.py; Mamba2.forward paths A/B, SURVEY.md 3.2].
 *   x, out: (B, D, L) any strides (channel-last stride(1)==1 is the fast path);
 *   weight: (D, W) 2<=W<=4; bias: (D) or absent; seq_idx: (B, L) int32 or absent;
 *   initial_states / final_states: (B, D, W-1) or absent. */
typedef struct omni_conv1d_fwd_params {
  omni_tensor_t x, weight, bias, seq_idx, initial_states;
  omni_tensor_t out, final_states;
  int32_t activation; /* omni_activation_t */
} omni_conv1d_fwd_params_t;
OMNI_API int omni_causal_conv1d_fwd(const omni_conv1d_fwd_params_t* p, void* stream);

/* Backward of the above (causal_conv1d_cuda.causal_conv1d_bwd upstream).
 *   dweight (D, W) and dbias (D) are FP32 and must be ZEROED by the caller (accumulated with
 *   atomics); dx like x; dinitial_states (B, D, W-1) optional. */
typedef struct omni_conv1d_bwd_params {
  omni_tensor_t x, weight, bias, dout, seq_idx, initial_states;
  omni_tensor_t dx, dweight, dbias, dinitial_states;
  int32_t activation;
} omni_conv1d_bwd_params_t;
OMNI_API int omni_causal_conv1d_bwd(const omni_conv1d_bwd_params_t* p, void* stream);

/* causal_conv1d_update(x, conv_state, weight, bias, activation, cache_seqlens)
 * [Mamba2.step, SURVEY.md 3.2 path C].  x/out: (B, D, T); conv_state: (B, D, S>=W-1) mutated in place;
 * cache_seqlens: (B) int32 or absent (ring-buffer mode). */
typedef struct omni_conv1d_update_params {
  omni_tensor_t x, conv_state, weight, bias, cache_seqlens;
  omni_tensor_t out;
  int32_t activation;
} omni_conv1d_update_params_t;
OMNI_API int omni_causal_conv1d_update(const omni_conv1d_update_params_t* p, void* stream);

/* ---- SSD chunked scan (Mamba-2) ---------------------------------------------------------- */
/* mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D, z, dt_bias, initial_states, seq_idx,
 * dt_softplus, dt_limit, return_final_states)  [mamba_ssm/ops/triton/ssd_combined.py; SURVEY.md 8(a) a4].
 *   x, z, out: (B, L, H, P); dt: (B, L, H); A: (H) fp32; B, C: (B, L, G, N);
 *   D: (H) or (H, P); dt_bias: (H); initial_states / final_states: (B, H, P, N) (final: fp32);
 *   seq_idx: (B, L) int32.  Innermost dims of x/z/out/B/C must be contiguous.  out may be fp32 while x is bf16
 *   (the tensor-core kernel then stores its fp32 result unrounded; used by the parity tests). */
typedef struct omni_ssd_fwd_params {
  omni_tensor_t x, dt, A, B, C, D, z, dt_bias, initial_states, seq_idx;
  omni_tensor_t out, final_states;
  omni_tensor_t workspace; /* 1-D, contiguous, >= omni_ssd_fwd_workspace_bytes() bytes, 16B aligned; needed by the
                            * tensor-core algorithm (AUTO falls back to the recurrence without it).  It holds the fp16
                            * copies of B and C and the fp32 state hand-off slots + flags of the half-item schedule
                            * (the flags are reset by a memset on `stream` inside the call); contents are scratch */
  omni_tensor_t chunk_states; /* optional OUTPUT for a caller that will run the backward (upstream's autograd function
                            * recomputes the chunk states in its backward; with 180 GB of HBM they can be kept instead):
                            * 1-D fp16, >= omni_ssd_chunk_states_bytes() bytes, 16B aligned.  When
                            * omni_ssd_fwd_saves_chunk_states(p) is 1 the forward also stores the state ENTERING every
                            * 128-token chunk, (B, ceil(L / 128), H * P, N) fp16 - the tile it converts for its own use anyway -
                            * and the backward given the same tensor skips its forward state sweep.  Absent: nothing changes */
  int32_t chunk_size; /* accepted for API parity; the result does not depend on it */
  int32_t dt_softplus;
  float dt_min, dt_max; /* dt_limit */
  int32_t algo;         /* omni_ssd_algo_t */
} omni_ssd_fwd_params_t;
/* 1 when the call would run on the tcgen05 kernels for exactly these params (algo AUTO falls back to the exact fp32
 * recurrence otherwise) - the single eligibility test; callers ask instead of re-deriving it. */
OMNI_API int omni_ssd_fwd_tc_supported(const omni_ssd_fwd_params_t* p);
OMNI_API int64_t omni_ssd_fwd_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim,
                                              int64_t ngroups, int64_t dstate);
OMNI_API int omni_ssd_chunk_scan_fwd(const omni_ssd_fwd_params_t* p, void* stream);
/* bytes of `chunk_states` for these sizes (0: geometry without a tensor-core path) and whether omni_ssd_chunk_scan_fwd
 * called with exactly these params fills it (tensor-core path taken, plain - not piece - schedule, tensor large enough) */
OMNI_API int64_t omni_ssd_chunk_states_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim, int64_t dstate);
OMNI_API int omni_ssd_fwd_saves_chunk_states(const omni_ssd_fwd_params_t* p);

/* Backward.  dout like out.  Outputs: dx like x; ddt (B, L, H) in dt's dtype (gradient w.r.t. the RAW dt);
 * dB, dC: (B, L, G, N) FP32, ZEROED by the caller; dz like z (optional); dinitial_states (B,H,P,N) fp32
 * optional; per-(batch, head) FP32 partials the caller sums over batch: dA_part, ddt_bias_part: (B, H),
 * dD_part: (B, H) or (B, H, P).  dfinal_states (B,H,P,N) optional incoming gradient.
 * workspace: 1-D, 256-byte aligned; omni_ssd_bwd_workspace_elems() FP32 elements for the recurrence,
 * omni_ssd_bwd_tc_workspace_bytes() bytes for the tensor-core algorithm (AUTO uses the latter when
 * the shapes/dtypes qualify and the workspace is large enough; dB/dC/dA_part/... contracts are the same). */
typedef struct omni_ssd_bwd_params {
  omni_tensor_t x, dt, A, B, C, D, z, dt_bias, initial_states, seq_idx;
  omni_tensor_t out; /* reserved (the forward's output, as upstream saves it); no algorithm reads it: pass an absent tensor */
  omni_tensor_t dout, dfinal_states;
  omni_tensor_t dx, ddt, dB, dC, dz, dinitial_states, dA_part, ddt_bias_part, dD_part;
  omni_tensor_t workspace;
  omni_tensor_t chunk_states; /* optional INPUT: the tensor a forward with omni_ssd_fwd_saves_chunk_states() == 1 filled for the
                            * same x / dt / A / B / dt_bias / initial_states; the tensor-core backward then skips its forward
                            * state sweep (the recurrence ignores it) */
  int32_t chunk_size;
  int32_t dt_softplus;
  float dt_min, dt_max;
  int32_t algo;
} omni_ssd_bwd_params_t;
OMNI_API int64_t omni_ssd_bwd_workspace_elems(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim,
                                     int64_t dstate);
OMNI_API int64_t omni_ssd_bwd_tc_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim,
                                                 int64_t ngroups, int64_t dstate);
OMNI_API int omni_ssd_chunk_scan_bwd(const omni_ssd_bwd_params_t* p, void* stream);
OMNI_API int omni_ssd_bwd_tc_supported(const omni_ssd_bwd_params_t* p); /* as omni_ssd_fwd_tc_supported, for the backward */

/* ---- gated RMSNorm / LayerNorm ------------------------------------------------------------ */
/* rmsnorm_fn / layernorm_fn(x, weight, bias, z, eps, group_size, norm_before_gate)
 * [mamba_ssm/ops/triton/layernorm_gated.py; SURVEY.md A.4].  x, z, out: (M, D) row-major (row stride
 * free); weight/bias: (D); rstd/mean: (M, D/group_size) fp32 outputs (mean only for LayerNorm). */
typedef struct omni_norm_gated_fwd_params {
  omni_tensor_t x, weight, bias, z;
  omni_tensor_t out, rstd, mean;
  float eps;
  int32_t group_size;
  int32_t norm_before_gate;
  int32_t is_rms_norm;
} omni_norm_gated_fwd_params_t;
OMNI_API int omni_norm_gated_fwd(const omni_norm_gated_fwd_params_t* p, void* stream);

/* Backward: dx, dz like x; dw_part/db_part: (nparts, D) fp32 partial sums (caller sums over dim 0;
 * nparts = dw_part.shape[0] chosen by the caller, typically the SM count); out_recompute optional
 * (re-materialised forward output, used for the out_proj weight gradient). */
typedef struct omni_norm_gated_bwd_params {
  omni_tensor_t x, weight, bias, z, dout, rstd, mean;
  omni_tensor_t dx, dz, dw_part, db_part, out_recompute;
  float eps;
  int32_t group_size;
  int32_t norm_before_gate;
  int32_t is_rms_norm;
} omni_norm_gated_bwd_params_t;
OMNI_API int omni_norm_gated_bwd(const omni_norm_gated_bwd_params_t* p, void* stream);

/* ---- fused residual-add + norm ------------------------------------------------------------ */
/* layer_norm_fn(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm)
 * [mamba_ssm/ops/triton/layer_norm.py; call sites block.py:86-95, mixer_seq_simple.py:428-437].
 *   x: (M, D); residual: (M, D) or absent; y: (M, D) in x's dtype... see omnimamba_b200/interface;
 *   residual_out: (M, D) (absent => not written); rstd/mean: (M) fp32. */
typedef struct omni_add_norm_fwd_params {
  omni_tensor_t x, residual, weight, bias;
  omni_tensor_t y, residual_out, rstd, mean;
  float eps;
  int32_t is_rms_norm;
} omni_add_norm_fwd_params_t;
OMNI_API int omni_add_norm_fwd(const omni_add_norm_fwd_params_t* p, void* stream);

/* Backward.  xres is the saved normalised-input (= residual_out, or x when there was no add);
 * dy: (M, D); dresidual_in: optional incoming gradient of residual_out (prenorm);
 * dx: (M, D) (also the gradient of the incoming residual - same values); dw_part/db_part as above. */
typedef struct omni_add_norm_bwd_params {
  omni_tensor_t xres, weight, bias, dy, dresidual_in, rstd, mean;
  omni_tensor_t dx, dresidual, dw_part, db_part; /* dresidual: optional copy of dx in the residual's dtype */
  float eps;
  int32_t is_rms_norm;
} omni_add_norm_bwd_params_t;
OMNI_API int omni_add_norm_bwd(const omni_add_norm_bwd_params_t* p, void* stream);

/* ---- single-token recurrent update -------------------------------------------------------- */
/* selective_state_update(state, x, dt, A, B, C, D, z, dt_bias, dt_softplus)
 * [mamba_ssm/ops/triton/selective_state_update.py; Mamba2.step; scripts/inference_t2i.py decode loop via
 * models/stage2/generation.py:208-211].  Head form: state (B,H,P,N) mutated in place; x, dt, z, out:
 * (B,H,P); A: (H,P,N); B, C: (B,G,N); D, dt_bias: (H,P).  Stride-0 broadcasts are expected. */
typedef struct omni_ssu_params {
  omni_tensor_t state, x, dt, A, B, C, D, z, dt_bias;
  omni_tensor_t out;
  int32_t dt_softplus;
} omni_ssu_params_t;
OMNI_API int omni_selective_state_update(const omni_ssu_params_t* p, void* stream);

/* ---- single-token layer core (decode) ------------------------------------------------------- */
/* The three kernels Mamba2.step launches between in_proj and out_proj, as one [mamba_ssm/modules/mamba2.py: step;
 * captured per layer by /root/reference/models/stage2/generation.py:383-431]: causal_conv1d_update (+SiLU) on the xBC
 * columns of zxbcdt, selective_state_update (dt_softplus, dt_bias, A tied over d_state, D per head) and the gated RMSNorm
 * rmsnorm(y * silu(z)) * norm_weight (norm_before_gate = False, one group).  zxbcdt (B, 2 dim + 2 N + H), dim = 64 H;
 * conv_state (B, dim + 2 N, W <= 4) and ssm_state (B, H, 64, 128) are updated IN PLACE; A (H) fp32 = -exp(A_log);
 * out (B, dim).  Built for the OmniMamba geometry (nheads 64, headdim 64, d_state 128, ngroups 1); other geometries use the
 * three separate entry points.  Graph-capturable (one cluster launch, no host synchronisation). */
typedef struct omni_mamba2_decode_core_params {
  omni_tensor_t zxbcdt, conv_state, conv_weight, conv_bias, ssm_state, A, D, dt_bias, norm_weight;
  omni_tensor_t out;
  float eps;
} omni_mamba2_decode_core_params_t;
OMNI_API int omni_mamba2_decode_core(const omni_mamba2_decode_core_params_t* p, void* stream);

/* ---- Mamba-1 selective scan ---------------------------------------------------------------- */
/* selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
 * [mamba_ssm/ops/selective_scan_interface.py; reachable via ssm_cfg.layer="Mamba1",
 * mixer_seq_simple.py:197-201].  u, delta, z, out: (B, D, L); A: (D, N) fp32; B, C: (B, G, N, L);
 * D, delta_bias: (D) fp32; last_state: (B, D, N) fp32 optional. */
typedef struct omni_selscan_fwd_params {
  omni_tensor_t u, delta, A, B, C, D, z, delta_bias;
  omni_tensor_t out, last_state;
  int32_t delta_softplus;
} omni_selscan_fwd_params_t;
OMNI_API int omni_selective_scan_fwd(const omni_selscan_fwd_params_t* p, void* stream);

/* Backward [selective_scan_cuda.bwd upstream; SelectiveScanFn.backward]: du like u, ddelta like delta; dB, dC
 * contiguous (B, G, N, L) FP32, zeroed by the caller (accumulated with atomics over the channels of a group);
 * dz present iff z is; dA_part (B, D, N), dD_part, ddelta_bias_part (B, D) contiguous fp32 partials (the caller sums
 * over the batch); workspace: omni_selective_scan_bwd_workspace_elems(...) fp32 values (state checkpoints). */
typedef struct omni_selscan_bwd_params {
  omni_tensor_t u, delta, A, B, C, D, z, delta_bias, dout;
  omni_tensor_t du, ddelta, dB, dC, dz, dA_part, dD_part, ddelta_bias_part, workspace;
  int32_t delta_softplus;
} omni_selscan_bwd_params_t;
OMNI_API int omni_selective_scan_bwd(const omni_selscan_bwd_params_t* p, void* stream);
OMNI_API int64_t omni_selective_scan_bwd_workspace_elems(int64_t batch, int64_t dim, int64_t seqlen, int64_t dstate);

/* ---- projections ---------------------------------------------------------------------------- */
/* out (M, N) = a (M, K) b (N, K)^T [+ a2 (M, K2) b2 (N, K2)^T]: bf16 operands, fp32 accumulation, out bf16 or fp32.
 * Replaces the F.linear calls of Mamba2.forward / mamba_split_conv1d_scan_combined (in_proj, out_proj; cuBLAS upstream)
 * and their backward GEMMs; the optional second pair is the LoRA branch of the reference's in_proj
 * (/root/reference/models/stage2/lora.py:263-279) accumulated into the same tile.  a, b: either dim contiguous (an
 * operand may be passed as a transposed view), the other stride a multiple of 8 elements, 16-byte aligned base; a2 / b2
 * laid out like a / b; out: contiguous 16-byte aligned rows.  M, N, K need not be tile multiples. */
typedef struct omni_gemm_params {
  omni_tensor_t a, b, a2, b2;
  omni_tensor_t out;
} omni_gemm_params_t;
OMNI_API int omni_gemm_bf16(const omni_gemm_params_t* p, void* stream);
OMNI_API int omni_gemm_bf16_supported(void); /* 1 if the driver exports cuTensorMapEncodeTiled */
/* The same projection with fp32 operands and fp32 result, for decode shapes only (M = batch <= 128 rows, N >= 256, K >= 128,
 * K-major a and b): what F.linear does on fp32 weights inside Mamba2.step when the model runs without autocast
 * [/root/reference/scripts/inference_t2i.py:21-27 loads and runs the checkpoint in fp32;
 * /root/reference/models/stage2/generation.py:383-431 is the decode loop].  The weights are streamed once and multiplied on
 * the tensor cores with the error-compensated 3xTF32 split (hi / lo tiles, three MMAs per product): fp32-accurate
 * (<= 2e-6 relative), HBM-bound instead of SGEMM-bound.  Other shapes: OMNI_UNSUPPORTED (the caller keeps its own GEMM). */
OMNI_API int omni_gemm_f32_decode(const omni_gemm_params_t* p, void* stream);
/* debug: tile scheme of omni_gemm_bf16 - 0 automatic, 1 single-CTA 128 x 256 tiles only, 2 CTA pairs (256 x 256) whenever legal,
 * 3 automatic without the weight-streaming kernel for decode shapes (M <= 128; csrc/gemm_skinny.cu); 10 + k (k = 0, 1, 2, 4):
 * K split of that kernel forced to k CTAs per cluster (0 = automatic again) */
OMNI_API void omni_debug_set_gemm_mode(int mode);
/* debug: which kernels of the decode chain [Mamba2.step behind /root/reference/models/stage2/generation.py:383-431: add + norm
 * -> in_proj -> layer core -> out_proj per layer] may start before their predecessor in the stream has finished
 * (programmatic dependent launch): bit 0 add + norm, bit 1 the weight-streaming GEMM, bit 2 the layer core; 0 = every launch
 * fully serialised.  Default 3 (or OMNI_PDL from the environment).  Results are bit-identical for every mask. */
OMNI_API void omni_debug_set_pdl(int mask);

/* ---- fused training forward (path A) --------------------------------------------------------- */
/* mamba_split_conv1d_scan_combined forward [mamba_ssm/ops/triton/ssd_combined.py: MambaSplitConv1dScanCombinedFn.forward;
 * called by Mamba2.forward when use_mem_eff_path and no cache, i.e. by every training step of
 * /root/reference/models/stage2/block.py:117].  One call = causal conv1d + SiLU on the xBC columns of zxbcdt ->
 * chunked SSD scan (dt = last nheads columns, dt_softplus) -> gated RMSNorm with z = first dim columns ->
 * out_proj (bf16: omni_gemm_bf16).  Caller-owned intermediates (the backward reads them): xbc_conv (B, L, dim + 2 G N),
 * scan_out (B, L, dim) pre-norm, rstd (B L) fp32, y (B, L, dim) normed.  rmsnorm_weight absent: the gate is applied inside
 * the scan and y / rstd are unused.  outproj_weight (d_out, dim) bf16 absent: no projection, `out` unused.
 * workspace: as omni_ssd_chunk_scan_fwd.  The backward is the composition omni_norm_gated_bwd -> omni_ssd_chunk_scan_bwd ->
 * omni_causal_conv1d_bwd (+ omni_gemm_bf16 for out_proj), see INTEGRATION.md. */
typedef struct omni_split_conv1d_scan_fwd_params {
  omni_tensor_t zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, initial_states, seq_idx, rmsnorm_weight, outproj_weight;
  omni_tensor_t xbc_conv, scan_out, rstd, y, out, final_states, workspace;
  omni_tensor_t chunk_states; /* optional, handed to the scan: omni_ssd_fwd_params_t.chunk_states */
  int32_t nheads, headdim, ngroups, dstate, chunk_size;
  int32_t activation;       /* omni_activation_t */
  int32_t norm_before_gate;
  int32_t algo;             /* omni_ssd_algo_t */
  float dt_min, dt_max, rmsnorm_eps;
} omni_split_conv1d_scan_fwd_params_t;
OMNI_API int omni_split_conv1d_scan_fwd(const omni_split_conv1d_scan_fwd_params_t* p, void* stream);

/* ---- head loss ------------------------------------------------------------------------------ */
/* Softmax cross-entropy over a block of fp32 logits (the img_head / lm_head GEMM output of omni_gemm_bf16) - the loss
 * nn.CrossEntropyLoss computes at /root/reference/models/omnimamba.py:276-279 on the shifted logits of
 * models/mamba_vlm.py:96-100.  fwd: lse[r] = logsumexp(logits[r]), loss[r] = lse[r] - logits[r, labels[r]] (0 where
 * labels[r] == ignore_index).  bwd: grad[r, v] = (exp(logits[r, v] - lse[r]) - [v == labels[r]]) * scale[0], bf16 (zero rows
 * where ignored); scale is a DEVICE fp32 scalar (dloss / number of valid labels).  logits (M, V) fp32, labels (M) int64. */
typedef struct omni_softmax_ce_params {
  omni_tensor_t logits, labels, lse, loss, scale, grad;
  int64_t ignore_index;
} omni_softmax_ce_params_t;
OMNI_API int omni_softmax_ce_fwd(const omni_softmax_ce_params_t* p, void* stream);
OMNI_API int omni_softmax_ce_bwd(const omni_softmax_ce_params_t* p, void* stream);

/* ---- self test ------------------------------------------------------------------------------ */
/* Runs the tcgen05 GEMM forms (smem/TMEM operands, K-/MN-major) the chunked SSD kernel is built from on one CTA;
 * tests/test_gpu_tc.py checks the results.  Cm, Bm: 16-bit [128][128]; X: 16-bit [128][64] (bf16, or fp16 when bit8 of
 * `which` is set); P, Xs, S: fp32 [128][128], rounded to the same 16-bit format inside;
 * D1 = Cm Bm^T, D3 = r(Xs) Bm, D4 = Cm r(S)^T: fp32 [128][128]; D2 = r(P) X: fp32 [128][64].
 * which: bit0 D1, bit1 D2, bit2 D4, bit3 D3 (A in TMEM), bit6 D3 (A = Xs^T as an MN-major smem operand),
 * bit7 D3 (A = Xs as a K-major smem operand, B MN-major), bit8 fp16. */
OMNI_API int omni_selftest(const void* Cm, const void* Bm, const void* X, const float* P, const float* Xs,
                           const float* S, float* D1, float* D2, float* D3, float* D4, int which, void* stream);

/* debug: CTA 0 of subsequent tensor-core SSD launches writes clock64() per (chunk, event) into buf[n*32]
 * (device int64; NULL disables) through the TRACE instantiation of the kernel.  `chunks` = 1000 * mode + n: mode 0 the
 * forward, 1 / 2 the state sweeps of the backward.  Used by scripts/trace_tc.py and scripts/trace_sweep.py. */
OMNI_API void omni_debug_set_trace(void* buf, int chunks);
/* debug: as omni_debug_set_trace, for the backward gradient kernel: buf[items * 32] clock64 stamps per phase of CTA 0. */
OMNI_API void omni_debug_set_bwd_trace(void* buf, int items);
/* debug: hold the hand-off flag of the forward's half-item schedule back by delay_us (tests the consumer's wait; a wait
 * that times out traps the kernel - the launch fails, nothing stale is consumed); `schedules` (default 3): bit 0 enables
 * the half-item hand-off schedule, bit 1 the piece schedule (few long sequences cut into independent pieces). */
OMNI_API void omni_debug_set_handoff(unsigned delay_us, int schedules);
/* debug: suspend-time hint (ns) used by the mbarrier waits of the tensor-core SSD kernel (tuning experiments). */
OMNI_API void omni_debug_set_mbar_hint(unsigned ns);
/* debug: cycles for `iters` tcgen05.ld (mode 0/2: 4 KB each per warp) or 2x tcgen05.st (mode 1) per warp with nwarps warps
 * issuing concurrently on one SM; out[warp] = cycles (device int64[128]). */
OMNI_API int omni_debug_tmem_bench(long long* out, int mode, int nwarps, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMNISSM_H_ */
