// C-ABI entry points of the SSD scan: argument checks common to all algorithms + algorithm choice.
#include "common.cuh"

namespace omni {
int ssd_recurrent_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s);
int ssd_recurrent_bwd(const omni_ssd_bwd_params_t* p, cudaStream_t s);
bool ssd_tc_fwd_supported(const omni_ssd_fwd_params_t* p);
int ssd_tc_fwd(const omni_ssd_fwd_params_t* p, cudaStream_t s);
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
  return ssd_recurrent_bwd(p, static_cast<cudaStream_t>(stream));
}
