// History-BEV ego-motion warp ("next" row f-1 of SURVEY.md section 8): batched replacement for the per-sample
// torchvision F.affine calls of EncoderLayer.project_history_bev_feat (encoder.py:413-466).
//
// One pass = torchvision's tensor affine with bilinear interpolation and fill = 0:
//   centred output pixel (xc, yc) = (x - (W-1)/2, y - (H-1)/2)
//   grid = theta . (xc, yc, 1) / (W/2, H/2)                      (_gen_affine_grid)
//   val  = grid_sample(img, grid, bilinear, zeros, align_corners=False)
//   out  = val * mask, mask = the same sample of an all-ones image  (_apply_grid_transform with fill=0:
//          img*mask + (1-mask)*0 -- border pixels are attenuated twice; reproduced on purpose)
// theta (B, 2, 3) lives on the device, so the 2*B host syncs per layer of the reference disappear.
#include "bevr_common.cuh"

namespace bevr {

__global__ void __launch_bounds__(256) affine_warp_kernel(const float* __restrict__ in, const float* __restrict__ theta,
                                                          float* __restrict__ out, int B, int C, int H, int W) {
  const int hw = H * W;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * hw) return;
  const int b = idx / hw, p = idx - b * hw;
  const int y = p / W, x = p - y * W;
  const float* t = theta + b * 6;
  const float xc = (float)x - 0.5f * (float)W + 0.5f, yc = (float)y - 0.5f * (float)H + 0.5f;
  const float gx = (t[0] * xc + t[1] * yc + t[2]) / (0.5f * (float)W);
  const float gy = (t[3] * xc + t[4] * yc + t[5]) / (0.5f * (float)H);
  const float px = ((gx + 1.f) * (float)W - 1.f) / 2.f, py = ((gy + 1.f) * (float)H - 1.f) / 2.f;
  const float x0f = floorf(px), y0f = floorf(py);
  const float wx1 = px - x0f, wx0 = (x0f + 1.f) - px, wy1 = py - y0f, wy0 = (y0f + 1.f) - py;
  const bool any = (px > -1.f) && (px < (float)W) && (py > -1.f) && (py < (float)H);
  const int x0 = any ? (int)x0f : -2, y0 = any ? (int)y0f : -2;
  float w[4];
  int off[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xi = x0 + (k & 1), yi = y0 + (k >> 1);
    const bool ok = xi >= 0 && xi < W && yi >= 0 && yi < H;
    w[k] = ok ? ((k & 1) ? wx1 : wx0) * ((k >> 1) ? wy1 : wy0) : 0.f;
    off[k] = ok ? yi * W + xi : 0;
  }
  const float mask = w[0] + w[1] + w[2] + w[3];
  const float* src = in + (size_t)b * C * hw;
  float* dst = out + (size_t)b * C * hw + p;
  for (int c = 0; c < C; ++c, src += hw, dst += hw) {
    const float v = w[0] * __ldg(src + off[0]) + w[1] * __ldg(src + off[1]) + w[2] * __ldg(src + off[2]) +
                    w[3] * __ldg(src + off[3]);
    *dst = v * mask;
  }
}

}  // namespace bevr

extern "C" int bevr_affine_warp(const float* in, const float* theta, float* out, int B, int C, int H, int W,
                                void* stream) {
  using namespace bevr;
  BEVR_REQUIRE(in && theta && out, "affine_warp: NULL pointer argument");
  BEVR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (long long)B * H * W < (1ll << 31), "affine_warp: bad sizes");
  BEVR_REQUIRE(in != out, "affine_warp: in-place operation is not supported");
  const int n = B * H * W;
  affine_warp_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, theta, out, B, C, H, W);
  return check_launch("affine_warp");
}
