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

}  // namespace bevr

extern "C" int bevr_version(void) { return BEVR_VERSION; }
extern "C" const char* bevr_last_error(void) { return bevr::g_err; }
