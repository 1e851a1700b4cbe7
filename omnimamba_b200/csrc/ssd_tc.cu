// tcgen05 / TMA chunked SSD forward for sm_100a (placeholder until the kernel lands: reports unsupported so
// OMNI_SSD_AUTO uses the exact recurrence).
#include "common.cuh"

namespace omni {
bool ssd_tc_fwd_supported(const omni_ssd_fwd_params_t*) { return false; }
int ssd_tc_fwd(const omni_ssd_fwd_params_t*, cudaStream_t) {
  return set_error(OMNI_UNSUPPORTED, "ssd: tcgen05 chunked kernel not built");
}
}  // namespace omni
