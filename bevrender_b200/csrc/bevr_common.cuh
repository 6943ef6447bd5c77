// Shared helpers for libbevr.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "bevr.h"

namespace bevr {

void set_error(const char* fmt, ...);          // thread-local message (api.cu)
int check_launch(const char* what);            // cudaPeekAtLastError -> return code
int check_cuda(cudaError_t e, const char* what);   // same for a runtime call's own return value (memset, attribute)
int sm_count();                                // multiprocessors of the current device (queried once per device)

#define BEVR_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      bevr::set_error(__VA_ARGS__);        \
      return -1;                           \
    }                                      \
  } while (0)

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// same, predicated (no branch around the RED: the gather loops are issue-bound)
__device__ __forceinline__ void red_add_v4_if(bool pred, float* addr, float a, float b, float c, float d) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
      ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)pred)
      : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace bevr
