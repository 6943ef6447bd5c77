// Offset head ("next" row f-2): everything between the depthwise convolution of the offset network and the
// sampling positions the gather consumes, as ONE kernel per direction.
//
// Reference: conv_offset[1:] = LayerNormProxy -> GELU -> Conv2d(Cm, O, 1, bias=False) (TSA_deform_attn.py:54-68,
// SCA_deform_attn.py:56-77; LayerNormProxy model_utils.py:51-59), then tanh * offset_range * factor (or clamp) and
// "+ reference" with two rearranges (TSA_deform_attn.py:158-196, SCA_deform_attn.py:219-277).  In PyTorch that is
// ~12 launches forward / ~25 backward and four permute copies of the (B*g, Cm, H, W) activation; here the
// activation is read in place (NCHW), the statistics live in registers and the positions are written once.
//
// Mapping: 4 lanes own one pixel (each a quarter of the channels, interleaved), 8 consecutive pixels per warp, so a
// warp-wide load touches four 32-byte sectors; per-pixel reductions are two shuffles.  Parameter gradients are
// reduced over the 8 pixels of a warp with shuffles into a shared-memory accumulator PRIVATE to the warp (channel c
// is only ever touched by lane 0 of slice c % 4, so plain read-modify-writes suffice -- shared fp32 atomics are CAS
// loops), then over the CTA's warps, then one global atomic per (CTA, parameter).
#include "bevr_common.cuh"

namespace bevr {

constexpr int OH_OMAX = 8;
constexpr int OH_THREADS = 128;
constexpr int OH_PASSES = 2;          // pixel groups per CTA (fewer global atomics in the backward)

struct OffArgs {
  const float* y;        // (Bg, Cm, H, W)
  const float* gamma;    // (Cm)
  const float* beta;     // (Cm)
  const float* w1;       // (O, Cm)
  const float* ref;      // (Bg or 1, Hk, Wk, 2) in (y, x)
  float* pos;            // (Bg, Hk, Wk, 2)
  float* stats;          // (Bg*H*W, 2) mean, rstd
  const float* grad_pos;
  float* grad_y;
  float* grad_gamma;
  float* grad_beta;
  float* grad_w1;
  long long ref_bstride;
  long long total;       // Bg*H*W
  int Cm, H, W, O;
  int interleave;        // 0: pos[b, y, x, o] (O == 2);  1: pos[b, y>>1, x*O + o, y&1]  ("(h n) w -> n h (w d)")
  int scaled;            // 1: tanh(raw) * range[n] * factor;  0: clamp(raw + ref, -1, 1)
  float ry, rx, factor;  // offset range of the y / x component, offset_range_factor
  float eps;
};

__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752440f)) + z * 0.39894228040143267794f * expf(-0.5f * z * z);
}
__device__ __forceinline__ float slice_sum(float v) {   // over the 4 lanes (lane bits 3, 4) that share a pixel
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
// over the 8 pixels (lane bits 0..2) of a warp; `m` names the 8 lanes of the caller's channel slice only, because
// slices run different trip counts when Cm is not a multiple of 4
__device__ __forceinline__ float pixel_sum(unsigned m, float v) {
  v += __shfl_xor_sync(m, v, 1);
  v += __shfl_xor_sync(m, v, 2);
  v += __shfl_xor_sync(m, v, 4);
  return v;
}

// index of output channel o of pixel (b, yq, xq) inside pos / grad_pos / ref (without the batch term), and its
// component n (0 = y, 1 = x)
__device__ __forceinline__ long long pos_index(const OffArgs& a, int yq, int xq, int o, int& n) {
  if (a.interleave) {
    n = yq & 1;
    return ((long long)(yq >> 1) * ((long long)a.W * a.O) + (long long)xq * a.O + o) * 2 + n;
  }
  n = o;
  return ((long long)yq * a.W + xq) * 2 + o;
}

template <bool BWD>
__global__ void __launch_bounds__(OH_THREADS) offset_head_kernel(const __grid_constant__ OffArgs a) {
  extern __shared__ float sm[];
  const int Cm = a.Cm, O = a.O;
  float* s_gamma = sm;
  float* s_beta = sm + Cm;
  float* s_w1 = sm + 2 * Cm;                 // [O][Cm]
  float* s_acc = sm + (2 + O) * Cm;          // backward, one copy per warp: grad_gamma[Cm], grad_beta[Cm], grad_w1[O][Cm]
  for (int i = threadIdx.x; i < Cm; i += blockDim.x) { s_gamma[i] = a.gamma[i]; s_beta[i] = a.beta[i]; }
  for (int i = threadIdx.x; i < O * Cm; i += blockDim.x) s_w1[i] = a.w1[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (BWD)
    for (int i = threadIdx.x; i < nwarp * (2 + O) * Cm; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float* w_acc = s_acc + warp * (2 + O) * Cm;
  const int pl = lane & 7, s = lane >> 3;
  const unsigned smask = 0xffu << (8 * s);
  const long long HW = (long long)a.H * a.W;
  const long long hk_wk2 = a.interleave ? (long long)(a.H >> 1) * a.W * O * 2 : HW * 2;   // pos elements per batch entry

  for (int pass = 0; pass < OH_PASSES; ++pass) {
    const long long pix = (((long long)blockIdx.x * OH_PASSES + pass) * nwarp + warp) * 8 + pl;
    const bool live = pix < a.total;
    const long long b = live ? pix / HW : 0, r = live ? pix - b * HW : 0;
    const int yq = (int)(r / a.W), xq = (int)(r - (long long)yq * a.W);
    const float* yp = a.y + b * Cm * HW + r;

    float mean, rstd;
    if (!BWD) {
      float sum = 0.f;
      for (int c = s; c < Cm; c += 4) sum += live ? __ldg(yp + c * HW) : 0.f;
      mean = slice_sum(sum) / (float)Cm;
      float sq = 0.f;
      for (int c = s; c < Cm; c += 4) {
        const float d = (live ? __ldg(yp + c * HW) : mean) - mean;
        sq = fmaf(d, d, sq);
      }
      rstd = 1.f / sqrtf(slice_sum(sq) / (float)Cm + a.eps);
      if (live && s == 0 && a.stats) reinterpret_cast<float2*>(a.stats)[pix] = make_float2(mean, rstd);
    } else {
      const float2 st = live ? __ldg(reinterpret_cast<const float2*>(a.stats) + pix) : make_float2(0.f, 1.f);
      mean = st.x; rstd = st.y;
    }

    // raw_o = sum_c w1[o][c] * gelu(z_c)
    float acc[OH_OMAX];
#pragma unroll
    for (int o = 0; o < OH_OMAX; ++o) acc[o] = 0.f;
    for (int c = s; c < Cm; c += 4) {
      const float yv = live ? __ldg(yp + c * HW) : mean;
      const float g = gelu_f(fmaf((yv - mean) * rstd, s_gamma[c], s_beta[c]));
#pragma unroll
      for (int o = 0; o < OH_OMAX; ++o)
        if (o < O) acc[o] = fmaf(s_w1[o * Cm + c], g, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < OH_OMAX; ++o)
      if (o < O) acc[o] = slice_sum(acc[o]);

    // epilogue per output channel: position (forward) or d loss / d raw_o (backward)
    float draw[OH_OMAX];
#pragma unroll
    for (int o = 0; o < OH_OMAX; ++o) {
      draw[o] = 0.f;
      if (o < O && live) {
        int n;
        const long long pi = pos_index(a, yq, xq, o, n);
        const float refv = __ldg(a.ref + b * a.ref_bstride + pi);
        if (!BWD) {
          if ((o & 3) == s) {     // the four lanes of a pixel share the stores
            float p;
            if (a.scaled) p = tanhf(acc[o]) * (n == 0 ? a.ry : a.rx) * a.factor + refv;   // reference op order
            else p = fminf(fmaxf(acc[o] + refv, -1.f), 1.f);
            a.pos[b * hk_wk2 + pi] = p;
          }
        } else {
          const float gp = __ldg(a.grad_pos + b * hk_wk2 + pi);
          if (a.scaled) {
            const float t = tanhf(acc[o]);
            draw[o] = gp * a.factor * (n == 0 ? a.ry : a.rx) * (1.f - t * t);
          } else {
            const float u = acc[o] + refv;
            draw[o] = (u >= -1.f && u <= 1.f) ? gp : 0.f;      // torch.clamp backward: inclusive bounds
          }
        }
      }
    }
    if (!BWD) continue;

    // backward pass 1: layer-norm sums and the parameter gradients
    float s1 = 0.f, s2 = 0.f;
    for (int c = s; c < Cm; c += 4) {
      const float yv = live ? __ldg(yp + c * HW) : mean;
      const float yh = (yv - mean) * rstd;
      const float z = fmaf(yh, s_gamma[c], s_beta[c]);
      float dg = 0.f;
#pragma unroll
      for (int o = 0; o < OH_OMAX; ++o)
        if (o < O) dg = fmaf(s_w1[o * Cm + c], draw[o], dg);
      const float dz = dg * gelu_grad_f(z);
      const float dyh = dz * s_gamma[c];
      s1 += dyh;
      s2 = fmaf(dyh, yh, s2);
      // parameter gradients: sum over the warp's 8 pixels, one shared-memory atomic per (warp, parameter)
      const float gg = pixel_sum(smask, dz * yh), gb = pixel_sum(smask, dz);
      if (pl == 0) { w_acc[c] += gg; w_acc[Cm + c] += gb; }
      if (a.grad_w1) {
        const float g = gelu_f(z);
#pragma unroll
        for (int o = 0; o < OH_OMAX; ++o)
          if (o < O) {
            const float gw = pixel_sum(smask, draw[o] * g);
            if (pl == 0) w_acc[(2 + o) * Cm + c] += gw;
          }
      }
    }
    s1 = slice_sum(s1) / (float)Cm;
    s2 = slice_sum(s2) / (float)Cm;
    // backward pass 2: d loss / d y
    if (a.grad_y) {
      float* gyp = a.grad_y + b * Cm * HW + r;
      for (int c = s; c < Cm; c += 4) {
        const float yv = live ? __ldg(yp + c * HW) : mean;
        const float yh = (yv - mean) * rstd;
        const float z = fmaf(yh, s_gamma[c], s_beta[c]);
        float dg = 0.f;
#pragma unroll
        for (int o = 0; o < OH_OMAX; ++o)
          if (o < O) dg = fmaf(s_w1[o * Cm + c], draw[o], dg);
        const float dyh = dg * gelu_grad_f(z) * s_gamma[c];
        if (live) gyp[c * HW] = rstd * (dyh - s1 - yh * s2);
      }
    }
  }
  if (BWD) {
    __syncthreads();
    const int np = (2 + O) * Cm;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
      float v = 0.f;
      for (int w = 0; w < nwarp; ++w) v += s_acc[w * np + i];
      float* dst = i < Cm ? a.grad_gamma : (i < 2 * Cm ? a.grad_beta : a.grad_w1);
      if (dst) atomicAdd(dst + (i < Cm ? i : (i < 2 * Cm ? i - Cm : i - 2 * Cm)), v);
    }
  }
}

static int fill(OffArgs& a, int Bg, int Cm, int H, int W, int O, int interleave, int scaled, float factor, float eps,
                long long ref_bstride) {
  BEVR_REQUIRE(Bg >= 0 && Cm > 0 && H > 0 && W > 0, "offset_head: Cm, H, W must be > 0 and Bg >= 0");
  BEVR_REQUIRE(O >= 1 && O <= OH_OMAX, "offset_head: 1 <= O <= %d output channels (got %d)", OH_OMAX, O);
  BEVR_REQUIRE(interleave || O == 2, "offset_head: the plain layout has O == 2 (y, x) channels, got %d", O);
  BEVR_REQUIRE(!interleave || (H % 2 == 0), "offset_head: the interleaved layout needs an even number of query rows");
  a.Cm = Cm; a.H = H; a.W = W; a.O = O; a.interleave = interleave; a.scaled = scaled; a.eps = eps;
  a.total = (long long)Bg * H * W;
  a.ref_bstride = ref_bstride;
  const int Hk = interleave ? H / 2 : H, Wk = interleave ? W * O : W;
  // the reference forms 1.0 / (Hk - 1.0) in Python doubles and rounds once (TSA_deform_attn.py:170-174)
  a.ry = (float)(1.0 / ((double)Hk - 1.0));
  a.rx = (float)(1.0 / ((double)Wk - 1.0));
  a.factor = factor;
  return 0;
}

template <bool BWD>
static int launch_oh(const OffArgs& a, cudaStream_t st, const char* what) {
  const size_t smem = (size_t)(BWD ? 1 + OH_THREADS / 32 : 1) * (2 + a.O) * a.Cm * sizeof(float);
  BEVR_REQUIRE(smem <= 200 * 1024, "offset_head: Cm=%d needs %zu B of shared memory", a.Cm, smem);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(offset_head_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long per_cta = (long long)OH_PASSES * (OH_THREADS / 32) * 8;
  const long long grid = (a.total + per_cta - 1) / per_cta;
  BEVR_REQUIRE(grid < (1ll << 31), "offset_head: grid too large");
  offset_head_kernel<BWD><<<(unsigned)grid, OH_THREADS, smem, st>>>(a);
  return check_launch(what);
}

}  // namespace bevr

using namespace bevr;

extern "C" int bevr_offset_head_fwd(const float* y, const float* ln_weight, const float* ln_bias, const float* w1,
                                    const float* ref, long long ref_batch_stride, float* pos, float* stats, int Bg,
                                    int Cm, int H, int W, int O, int interleave, int scaled, float factor, float eps,
                                    void* stream) {
  OffArgs a = {};
  if (int rc = fill(a, Bg, Cm, H, W, O, interleave, scaled, factor, eps, ref_batch_stride)) return rc;
  if (a.total == 0) return 0;
  BEVR_REQUIRE(y && ln_weight && ln_bias && w1 && ref && pos, "offset_head_fwd: NULL argument");
  a.y = y; a.gamma = ln_weight; a.beta = ln_bias; a.w1 = w1; a.ref = ref; a.pos = pos; a.stats = stats;
  return launch_oh<false>(a, (cudaStream_t)stream, "offset_head_fwd");
}

extern "C" int bevr_offset_head_bwd(const float* y, const float* ln_weight, const float* ln_bias, const float* w1,
                                    const float* ref, long long ref_batch_stride, const float* stats,
                                    const float* grad_pos, float* grad_y, float* grad_ln_weight, float* grad_ln_bias,
                                    float* grad_w1, int Bg, int Cm, int H, int W, int O, int interleave, int scaled,
                                    float factor, float eps, void* stream) {
  OffArgs a = {};
  if (int rc = fill(a, Bg, Cm, H, W, O, interleave, scaled, factor, eps, ref_batch_stride)) return rc;
  if (a.total == 0) return 0;
  BEVR_REQUIRE(y && ln_weight && ln_bias && w1 && ref && stats && grad_pos, "offset_head_bwd: NULL argument");
  a.y = y; a.gamma = ln_weight; a.beta = ln_bias; a.w1 = w1; a.ref = ref; a.stats = const_cast<float*>(stats);
  a.grad_pos = grad_pos; a.grad_y = grad_y; a.grad_gamma = grad_ln_weight; a.grad_beta = grad_ln_bias;
  a.grad_w1 = grad_w1;
  return launch_oh<true>(a, (cudaStream_t)stream, "offset_head_bwd");
}
