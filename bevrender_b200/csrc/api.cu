// Error reporting and version of the C ABI (include/bevr.h).
#include <stdarg.h>

#include "bevr_common.cuh"

namespace bevr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return 0;
  cudaGetLastError();  // clear the (non-sticky) launch error so later calls are not poisoned
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  cudaGetLastError();
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

// The SM count only sizes grids (waves, grid-stride loops); it is a per-device constant, cached per device ordinal.
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;                      // benign race: every thread writes the same value
  }
  return n;
}

}  // namespace bevr

extern "C" int bevr_version(void) { return BEVR_VERSION; }
extern "C" const char* bevr_last_error(void) { return bevr::g_err; }
