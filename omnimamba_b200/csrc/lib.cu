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

bool pdl_enabled(int kind) {
  static const int mask = [] {
    const char* e = getenv("OMNI_PDL");
    return e != nullptr ? atoi(e) : kPdlDefault;
  }();
  return (mask & kind) != 0;
}

}  // namespace omni

extern "C" {
int omni_version(void) { return OMNI_ABI_VERSION; }
const char* omni_last_error(void) { return omni::g_err; }
int64_t omni_launch_count(void) { return omni::g_launches.load(std::memory_order_relaxed); }
void omni_reset_launch_count(void) { omni::g_launches.store(0, std::memory_order_relaxed); }
}
