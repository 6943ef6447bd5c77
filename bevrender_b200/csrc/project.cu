// K4: camera -> BEV reference-point projection, bit-exact with the reference's torch-CPU result.
//
// Replaces BEV2CameraProjector.bev_grid_to_camera + get_in_bound_mask (bev_cmr_proj.py:61-124).
// One thread per 3-D point.  Every floating-point operation is an explicit round-to-nearest
// intrinsic in the order torch's CPU kernels execute it, so nvcc can neither contract nor
// re-associate:
//   cam = inv(T) @ p          4x4 @ 4xN sgemm  = acc = a0*p0; acc = fma(a_k, p_k, acc), k = 1..3
//   uvw = K[:, :3] @ cam[:3]  3x3 @ 3xN        = same chain, k = 1..2
//   uv  = uvw[:2] / uvw[2]    IEEE division
//   pix = int32(uv)           truncation; out-of-range / NaN -> INT_MIN as x86 cvttps2dq does
//   mask = 0 <= pix_x < W-1  &  0 <= pix_y < H-1
//   uv  = mask ? uv : 0 ; u /= (W-1) ; v /= (H-1) ; uv = uv*2 - 1     (each step rounded)
#include <limits.h>

#include "bevr_common.cuh"

namespace bevr {

struct ProjArgs {
  float T[16];
  float K[9];
};

__device__ __forceinline__ float chain4(const float* a, float p0, float p1, float p2, float p3) {
  float acc = __fmul_rn(a[0], p0);
  acc = __fmaf_rn(a[1], p1, acc);
  acc = __fmaf_rn(a[2], p2, acc);
  return __fmaf_rn(a[3], p3, acc);
}
__device__ __forceinline__ float chain3(const float* a, float p0, float p1, float p2) {
  float acc = __fmul_rn(a[0], p0);
  acc = __fmaf_rn(a[1], p1, acc);
  return __fmaf_rn(a[2], p2, acc);
}
__device__ __forceinline__ int trunc_x86(float v) {
  return (v >= -2147483648.f && v < 2147483648.f) ? (int)v : INT_MIN;
}

__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ pts, const __grid_constant__ ProjArgs a,
                                                      float* __restrict__ uv, int32_t* __restrict__ pix,
                                                      uint8_t* __restrict__ mask, int n, int W, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p0 = pts[i], p1 = pts[n + i], p2 = pts[2 * n + i], p3 = pts[3 * n + i];
  const float c0 = chain4(a.T + 0, p0, p1, p2, p3);
  const float c1 = chain4(a.T + 4, p0, p1, p2, p3);
  const float c2 = chain4(a.T + 8, p0, p1, p2, p3);
  const float u3 = chain3(a.K + 6, c0, c1, c2);
  float u = __fdiv_rn(chain3(a.K + 0, c0, c1, c2), u3);
  float v = __fdiv_rn(chain3(a.K + 3, c0, c1, c2), u3);
  const int iu = trunc_x86(u), iv = trunc_x86(v);
  const bool ok = iv >= 0 && iv < H - 1 && iu >= 0 && iu < W - 1;
  if (!ok) u = v = 0.f;
  u = __fdiv_rn(u, (float)(W - 1));
  v = __fdiv_rn(v, (float)(H - 1));
  uv[i] = __fsub_rn(__fmul_rn(u, 2.f), 1.f);
  uv[n + i] = __fsub_rn(__fmul_rn(v, 2.f), 1.f);
  if (pix) { pix[i] = iu; pix[n + i] = iv; }
  if (mask) mask[i] = ok ? 1 : 0;
}

}  // namespace bevr

extern "C" int bevr_project_points(const float* points, const float* cam_T, const float* K, float* uv,
                                   int32_t* pix, uint8_t* mask, int n, int img_w, int img_h, void* stream) {
  using namespace bevr;
  BEVR_REQUIRE(points && cam_T && K && uv, "project_points: points / cam_T / K / uv must not be NULL");
  BEVR_REQUIRE(n >= 0 && img_w > 1 && img_h > 1, "project_points: bad sizes n=%d img=%dx%d", n, img_w, img_h);
  if (n == 0) return 0;
  ProjArgs a;
  for (int i = 0; i < 16; ++i) a.T[i] = cam_T[i];
  for (int i = 0; i < 9; ++i) a.K[i] = K[i];
  project_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, a, uv, pix, mask, n, img_w, img_h);
  return check_launch("project_points");
}
