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

#define BEVR_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      bevr::set_error(__VA_ARGS__);        \
      return -1;                           \
    }                                      \
  } while (0)

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// 16-byte read-only global load (LDG.E.128.CONSTANT)
__device__ __forceinline__ uint4 ldg128(const uint4* p) { return __ldg(p); }

// vector fp32 reduction to global memory (sm_90+: RED.E.ADD.F32x4)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

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

template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
  __device__ static float load1(const float* p) { return __ldg(p); }
  __device__ static void store1(float* p, float v) { *p = v; }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& u, float* f) {
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
    f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
  }
  __device__ static uint4 pack(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                      pack_bf16(f[6], f[7]));
  }
  __device__ static float load1(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
  }
  __device__ static void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

}  // namespace bevr
