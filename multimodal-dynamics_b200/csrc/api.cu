// Library-level entry points: error string, version, launch counter, one-time init.
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;
int igemm_init();
int elementwise_init();

static thread_local char g_err[512] = "";

bool pdl_enabled() {
  static const bool on = std::getenv("MMDYN_NO_PDL") == nullptr;
  return on;
}

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mmdyn

extern "C" const char* mmdyn_last_error(void) { return mmdyn::g_err; }
extern "C" int mmdyn_version(void) { return 100; }
extern "C" long long mmdyn_launch_count(void) { return mmdyn::g_launch_count.load(); }

extern "C" int mmdyn_init(int device) {
  MMDYN_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MMDYN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    mmdyn::set_last_error("mmdyn_init: device %d is sm_%d%d; this library contains sm_100a code only",
                          device, prop.major, prop.minor);
    return MMDYN_ERR_UNSUPPORTED;
  }
  const int rc = mmdyn::igemm_init();
  if (rc != MMDYN_OK) return rc;
  return mmdyn::elementwise_init();
}
