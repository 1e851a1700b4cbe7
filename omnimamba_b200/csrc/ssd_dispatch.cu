// C-ABI entry points of the SSD scan: argument checks common to all algorithms + algorithm choice.
#include "common.cuh"

namespace omni {
int ssd_recurrent_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s);
int ssd_recurrent_bwd(const omni_ssd_bwd_params_t* p, cudaStream_t s);
bool ssd_tc_fwd_supported(const omni_ssd_fwd_params_t* p);
int ssd_tc_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s);
bool ssd_tc_bwd_supported(const omni_ssd_bwd_params_t* p);
int ssd_tc_bwd(const omni_ssd_bwd_params_t* p, cudaStream_t s);
int64_t ssd_tc_bwd_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t ngroups);
int64_t ssd_tc_chunk_states_bytes(int64_t batch, int64_t seqlen, int64_t nheads);
bool ssd_tc_fwd_saves_states(const omni_ssd_fwd_params_t* p);
}  // namespace omni

using namespace omni;

extern "C" int omni_ssd_chunk_scan_fwd(const omni_ssd_fwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->algo) {
    case OMNI_SSD_RECURRENT: return ssd_recurrent_fwd(p, s);
    case OMNI_SSD_CHUNKED_TC:
      OMNI_CHECK(ssd_tc_fwd_supported(p), OMNI_UNSUPPORTED,
                 "ssd: the tcgen05 chunked kernel needs bf16 x/B/C, headdim 64, d_state 128, no seq_idx, D of shape (H)");
      return ssd_tc_fwd(p, s);
    case OMNI_SSD_AUTO: return ssd_tc_fwd_supported(p) ? ssd_tc_fwd(p, s) : ssd_recurrent_fwd(p, s);
    default: return set_error(OMNI_UNSUPPORTED, "ssd: unknown algo %d", p->algo);
  }
}

extern "C" int omni_ssd_chunk_scan_bwd(const omni_ssd_bwd_params_t* p, void* stream) {
  OMNI_CHECK(p != nullptr, OMNI_BAD_SHAPE, "null params");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->algo) {
    case OMNI_SSD_RECURRENT: return ssd_recurrent_bwd(p, s);
    case OMNI_SSD_CHUNKED_TC:
      OMNI_CHECK(ssd_tc_bwd_supported(p), OMNI_UNSUPPORTED,
                 "ssd bwd: the tcgen05 chunked kernels need bf16 x/B/C/dout/dx/out, headdim 64, d_state 128, no seq_idx / z, "
                 "D of shape (H) and omni_ssd_bwd_tc_workspace_bytes() of 256-byte aligned workspace");
      return ssd_tc_bwd(p, s);
    case OMNI_SSD_AUTO: return ssd_tc_bwd_supported(p) ? ssd_tc_bwd(p, s) : ssd_recurrent_bwd(p, s);
    default: return set_error(OMNI_UNSUPPORTED, "ssd bwd: unknown algo %d", p->algo);
  }
}

extern "C" int64_t omni_ssd_bwd_tc_workspace_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim, int64_t ngroups,
                                                   int64_t dstate) {
  if (headdim != 64 || dstate != 128) return 0;
  return ssd_tc_bwd_workspace_bytes(batch, seqlen, nheads, ngroups);
}

// 1 when omni_ssd_chunk_scan_bwd would take the tensor-core path for exactly these params (dtypes, strides, alignment,
// workspace size, driver support): the ONE eligibility test - the Python surface asks instead of re-deriving it.
extern "C" int omni_ssd_bwd_tc_supported(const omni_ssd_bwd_params_t* p) { return p != nullptr && ssd_tc_bwd_supported(p) ? 1 : 0; }
extern "C" int omni_ssd_fwd_tc_supported(const omni_ssd_fwd_params_t* p) { return p != nullptr && ssd_tc_fwd_supported(p) ? 1 : 0; }

// The optional chunk-state tensor a forward can leave for its backward (omnissm.h: omni_ssd_fwd_params_t.chunk_states).
extern "C" int64_t omni_ssd_chunk_states_bytes(int64_t batch, int64_t seqlen, int64_t nheads, int64_t headdim, int64_t dstate) {
  if (headdim != 64 || dstate != 128 || batch <= 0 || seqlen <= 0 || nheads <= 0) return 0;
  return ssd_tc_chunk_states_bytes(batch, seqlen, nheads);
}
extern "C" int omni_ssd_fwd_saves_chunk_states(const omni_ssd_fwd_params_t* p) {
  if (p == nullptr || p->algo == OMNI_SSD_RECURRENT) return 0;
  return ssd_tc_fwd_saves_states(p) ? 1 : 0;
}
