// Library-level entry points: version, thread-local error message, launch counter.
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace omni {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static std::atomic<int> g_pdl_mask{-1};  // -1: not read from the environment yet
bool pdl_enabled(int kind) {
  int m = g_pdl_mask.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("OMNI_PDL");
    m = e != nullptr ? atoi(e) & 7 : kPdlDefault;
    g_pdl_mask.store(m, std::memory_order_relaxed);
  }
  return (m & kind) != 0;
}
void set_pdl_mask(int mask) { g_pdl_mask.store(mask & 7, std::memory_order_relaxed); }

}  // namespace omni

extern "C" {
int omni_version(void) { return OMNI_ABI_VERSION; }
const char* omni_last_error(void) { return omni::g_err; }
int64_t omni_launch_count(void) { return omni::g_launches.load(std::memory_order_relaxed); }
void omni_reset_launch_count(void) { omni::g_launches.store(0, std::memory_order_relaxed); }
/* debug: which kernels of the decode chain may start early (programmatic dependent launch): 1 add + norm, 2 weight-streaming
 * GEMM, 4 layer core; 0 = every launch fully serialised.  Default 3, or OMNI_PDL from the environment. */
void omni_debug_set_pdl(int mask) { omni::set_pdl_mask(mask); }
}
