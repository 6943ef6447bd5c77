// K1 / K2: multi-scale deformable-attention gather, forward and backward (sm_100a).
//
// Replaces F.grid_sample (+ its autograd) at TSA_deform_attn.py:210-217 and
// SCA_deform_attn.py:290-301 of the reference and generalises it to the
// (L levels, P points, h heads) operator of the north star.
//
// Design (DESIGN.md section "K1/K2"):
//  * value is channels-last (B,S,h,Dh): one bilinear tap of one head is one contiguous
//    Dh*sizeof(T) vector (128 B for fp32/Dh=32 = one L1 line = one L1 wavefront).
//  * A "group" of LPG lanes owns one (query, head[, channel chunk]) item; each lane holds a
//    16-byte vector of channels, so every tap is fetched by LPG coalesced LDG.128s.
//  * A CTA owns a 2-D tile of the query grid for ONE head: all taps of the tile fall in a small
//    window of every level, so they are served by L1 after the first touch and HBM sees each
//    value line about once (L2 keeps the (b,head) slice while its tiles are in flight).
//  * Tap set-up (unnormalise, floor, 4 offsets, 4 weights, validity) is done ONCE per sampling
//    point by one lane and staged in shared memory; the gather loop is branch-free:
//    LDS.128 x2 -> LDG.128 x4 -> FFMA.  Out-of-range taps get weight 0 and a safe offset.
//  * Backward recomputes the taps, accumulates grad_value with vector RED.ADD.F32x4 into an fp32
//    buffer, and writes grad_loc / grad_weight directly (intra-group shuffle reduction, no atomics).
#include "bevr_common.cuh"

namespace bevr {

struct MsdaArgs {
  const void* value;
  const float* loc;
  const float* ref;
  const void* weight;
  void* out;
  const void* grad_out;
  float* grad_value;
  float* grad_loc;
  void* grad_weight;
  int B, S, h, Dh, Nq, L, P;
  int qw, qh;        // query grid (qh*qw == Nq; flat list: qh = 1)
  int tw, th;        // query tile
  int tiles_x, tiles_y;
  int nchunk;        // channel chunks per head (Dh / (LPG*VEC))
  int flags;
  int lvl_h[BEVR_MAX_LEVELS], lvl_w[BEVR_MAX_LEVELS], lvl_start[BEVR_MAX_LEVELS];
};

// pixel coordinate of a normalised location, mirroring ATen's grid_sampler_unnormalize
__device__ __forceinline__ float unnormalize(float u, int size, int flags) {
  if (flags & BEVR_LOC_PM1) {
    if (flags & BEVR_ALIGN_CORNERS) return ((u + 1.f) / 2.f) * (float)(size - 1);
    return ((u + 1.f) * (float)size - 1.f) / 2.f;
  }
  if (flags & BEVR_ALIGN_CORNERS) return u * (float)(size - 1);
  return u * (float)size - 0.5f;
}
// d pixel / d location
__device__ __forceinline__ float unnormalize_scale(int size, int flags) {
  float s = (flags & BEVR_ALIGN_CORNERS) ? (float)(size - 1) : (float)size;
  return (flags & BEVR_LOC_PM1) ? 0.5f * s : s;
}

struct Tap {
  int off[4];     // vector offsets (units of 16 B) inside the sample's value; -1 if out of range
  float fx, fy;   // fractional parts (weights of the +1 taps)
  float aw;       // attention weight
  bool any;
};

template <typename T>
__device__ __forceinline__ Tap make_tap(const MsdaArgs& a, int b, int q, int head, int p) {
  Tap t;
  const int LP = a.L * a.P;
  const int l = p / a.P;
  const int H = a.lvl_h[l], W = a.lvl_w[l];
  const size_t li = (((size_t)b * a.Nq + q) * a.h + head) * LP + p;
  float2 xy = __ldg(reinterpret_cast<const float2*>(a.loc) + li);
  if (a.ref) {
    float2 r = __ldg(reinterpret_cast<const float2*>(a.ref) + ((size_t)b * a.Nq + q) * a.L + l);
    if (a.flags & BEVR_OFFSET_PIXELS) { xy.x = xy.x / (float)W; xy.y = xy.y / (float)H; }
    xy.x += r.x; xy.y += r.y;
  }
  t.aw = a.weight ? Vec<T>::load1(reinterpret_cast<const T*>(a.weight) + li) : 1.f;
  const float px = unnormalize(xy.x, W, a.flags), py = unnormalize(xy.y, H, a.flags);
  // negated comparisons: NaN / inf coordinates contribute nothing (as on the CPU path)
  t.any = (px > -1.f) && (px < (float)W) && (py > -1.f) && (py < (float)H);
  t.fx = t.fy = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) t.off[i] = -1;
  if (t.any) {
    const float x0f = floorf(px), y0f = floorf(py);
    const int x0 = (int)x0f, y0 = (int)y0f;
    t.fx = px - x0f;
    t.fy = py - y0f;
    const int vstride = a.h * a.Dh / Vec<T>::N;   // 16-byte vectors per pixel
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int xi = x0 + (i & 1), yi = y0 + (i >> 1);
      if (xi >= 0 && xi < W && yi >= 0 && yi < H) t.off[i] = (a.lvl_start[l] + yi * W + xi) * vstride;
    }
  }
  return t;
}

__device__ __forceinline__ bool tile_query(const MsdaArgs& a, int tile, int qi, int& q) {
  const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
  const int ly = qi / a.tw, lx = qi - ly * a.tw;
  const int y = ty * a.th + ly, x = tx * a.tw + lx;
  q = y * a.qw + x;
  return (y < a.qh) && (x < a.qw);
}

// ------------------------------------------------------------------ forward
template <typename T, int LPG, int MAXT>
__global__ void __launch_bounds__(MAXT) msda_fwd_kernel(const __grid_constant__ MsdaArgs a) {
  constexpr int VEC = Vec<T>::N;
  constexpr int GPW = 32 / LPG;                    // groups per warp
  extern __shared__ uint4 smem[];
  const int LP = a.L * a.P;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int nwarp = blockDim.x >> 5;
  const int g = lane / LPG, lig = lane % LPG;
  int4* s_off = reinterpret_cast<int4*>(smem) + warp * (2 * GPW * (LP + 1));
  float4* s_w = reinterpret_cast<float4*>(s_off + GPW * (LP + 1));

  const int n_tiles = a.tiles_x * a.tiles_y;
  int blk = blockIdx.x;
  const int tile = blk % n_tiles; blk /= n_tiles;
  const int vh = blk % (a.h * a.nchunk);
  const int b = blk / (a.h * a.nchunk);
  const int head = vh / a.nchunk, chunk = vh - head * a.nchunk;

  const uint4* vbase = reinterpret_cast<const uint4*>(a.value) +
                       (size_t)b * a.S * (a.h * a.Dh / VEC) + head * (a.Dh / VEC) + chunk * LPG + lig;
  const int tq = a.tw * a.th;

  for (int base = warp * GPW; base < tq; base += nwarp * GPW) {
    __syncwarp();
    // ---- stage the taps of this warp's GPW items
    for (int pt = lane; pt < GPW * LP; pt += 32) {
      const int gg = pt / LP, p = pt - gg * LP;
      int q;
      int4 o = make_int4(0, 0, 0, 0);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile_query(a, tile, base + gg, q)) {
        Tap t = make_tap<T>(a, b, q, head, p);
        if (t.any) {
          const float wx1 = t.fx, wx0 = 1.f - t.fx, wy1 = t.fy, wy0 = 1.f - t.fy;
          const float bw[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
          int oo[4];
          float ww[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            oo[i] = t.off[i] < 0 ? 0 : t.off[i];
            ww[i] = t.off[i] < 0 ? 0.f : t.aw * bw[i];
          }
          o = make_int4(oo[0], oo[1], oo[2], oo[3]);
          w = make_float4(ww[0], ww[1], ww[2], ww[3]);
        }
      }
      s_off[gg * (LP + 1) + p] = o;
      s_w[gg * (LP + 1) + p] = w;
    }
    __syncwarp();
    // ---- branch-free gather
    int q;
    const bool live = tile_query(a, tile, base + g, q);
    float acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
    const int4* so = s_off + g * (LP + 1);
    const float4* sw = s_w + g * (LP + 1);
#pragma unroll 4
    for (int p = 0; p < LP; ++p) {
      const int4 o = so[p];
      const float4 w = sw[p];
      const uint4 u0 = ldg128(vbase + o.x), u1 = ldg128(vbase + o.y);
      const uint4 u2 = ldg128(vbase + o.z), u3 = ldg128(vbase + o.w);
      float f[VEC];
      Vec<T>::unpack(u0, f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w.x, f[c], acc[c]);
      Vec<T>::unpack(u1, f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w.y, f[c], acc[c]);
      Vec<T>::unpack(u2, f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w.z, f[c], acc[c]);
      Vec<T>::unpack(u3, f);
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w.w, f[c], acc[c]);
    }
    if (live) {
      uint4* o = reinterpret_cast<uint4*>(a.out) +
                 (((size_t)b * a.Nq + q) * a.h + head) * (a.Dh / VEC) + chunk * LPG + lig;
      *o = Vec<T>::pack(acc);
    }
  }
}

// ------------------------------------------------------------------ backward
template <typename T, int LPG, int MAXT>
__global__ void __launch_bounds__(MAXT) msda_bwd_kernel(const __grid_constant__ MsdaArgs a) {
  constexpr int VEC = Vec<T>::N;
  constexpr int GPW = 32 / LPG;
  extern __shared__ uint4 smem[];
  const int LP = a.L * a.P;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int nwarp = blockDim.x >> 5;
  const int g = lane / LPG, lig = lane % LPG;
  int4* s_off = reinterpret_cast<int4*>(smem) + warp * (2 * GPW * (LP + 1));
  float4* s_w = reinterpret_cast<float4*>(s_off + GPW * (LP + 1));   // fx, fy, aw, -

  const int n_tiles = a.tiles_x * a.tiles_y;
  int blk = blockIdx.x;
  const int tile = blk % n_tiles; blk /= n_tiles;
  const int vh = blk % (a.h * a.nchunk);
  const int b = blk / (a.h * a.nchunk);
  const int head = vh / a.nchunk, chunk = vh - head * a.nchunk;

  const int vec_per_pix = a.h * a.Dh / VEC;
  const size_t vec0 = (size_t)b * a.S * vec_per_pix + head * (a.Dh / VEC) + chunk * LPG + lig;
  const uint4* vbase = reinterpret_cast<const uint4*>(a.value) + vec0;
  float* gvbase = a.grad_value ? a.grad_value + vec0 * VEC : nullptr;
  const int tq = a.tw * a.th;
  const unsigned gmask = (LPG == 32) ? 0xffffffffu : (((1u << LPG) - 1u) << (g * LPG));

  for (int base = warp * GPW; base < tq; base += nwarp * GPW) {
    __syncwarp();
    for (int pt = lane; pt < GPW * LP; pt += 32) {
      const int gg = pt / LP, p = pt - gg * LP;
      int q;
      int4 o = make_int4(-1, -1, -1, -1);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile_query(a, tile, base + gg, q)) {
        Tap t = make_tap<T>(a, b, q, head, p);
        o = make_int4(t.off[0], t.off[1], t.off[2], t.off[3]);
        w = make_float4(t.fx, t.fy, t.aw, 0.f);
      }
      s_off[gg * (LP + 1) + p] = o;
      s_w[gg * (LP + 1) + p] = w;
    }
    __syncwarp();
    int q;
    const bool live = tile_query(a, tile, base + g, q);
    float go[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) go[c] = 0.f;
    if (live) {
      const uint4 u = ldg128(reinterpret_cast<const uint4*>(a.grad_out) +
                             (((size_t)b * a.Nq + q) * a.h + head) * (a.Dh / VEC) + chunk * LPG + lig);
      Vec<T>::unpack(u, go);
    }
    const int4* so = s_off + g * (LP + 1);
    const float4* sw = s_w + g * (LP + 1);
    const size_t li0 = live ? (((size_t)b * a.Nq + q) * a.h + head) * LP : 0;
#pragma unroll 2
    for (int p = 0; p < LP; ++p) {
      const int4 o = so[p];
      const float4 w = sw[p];
      const int offs[4] = {o.x, o.y, o.z, o.w};
      const float wx1 = w.x, wx0 = 1.f - w.x, wy1 = w.y, wy0 = 1.f - w.y, aw = w.z;
      const float bw[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = offs[i] >= 0 ? ldg128(vbase + offs[i]) : make_uint4(0, 0, 0, 0);
      float d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f[VEC];
        Vec<T>::unpack(u[i], f);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) s = fmaf(go[c], f[c], s);
        d[i] = s;
        if (gvbase && offs[i] >= 0) {
          const float wt = aw * bw[i];
          float* dst = gvbase + (size_t)offs[i] * VEC;
#pragma unroll
          for (int c = 0; c < VEC; c += 4)
            red_add_v4(dst + c, wt * go[c], wt * go[c + 1], wt * go[c + 2], wt * go[c + 3]);
        }
      }
      // per-lane partials of d/d weight, d/d pixel_x, d/d pixel_y; reduce over the group
      float gw = wy0 * (wx0 * d[0] + wx1 * d[1]) + wy1 * (wx0 * d[2] + wx1 * d[3]);
      float gx = aw * (wy0 * (d[1] - d[0]) + wy1 * (d[3] - d[2]));
      float gy = aw * (wx0 * (d[2] - d[0]) + wx1 * (d[3] - d[1]));
#pragma unroll
      for (int s = LPG / 2; s > 0; s >>= 1) {
        gw += __shfl_xor_sync(gmask, gw, s);
        gx += __shfl_xor_sync(gmask, gx, s);
        gy += __shfl_xor_sync(gmask, gy, s);
      }
      if (live && lig == (p % LPG)) {
        const int l = p / a.P;
        if (a.nchunk == 1) {
          if (a.grad_loc) {
            float2 r = make_float2(gx * unnormalize_scale(a.lvl_w[l], a.flags),
                                   gy * unnormalize_scale(a.lvl_h[l], a.flags));
            reinterpret_cast<float2*>(a.grad_loc)[li0 + p] = r;
          }
          if (a.grad_weight) Vec<T>::store1(reinterpret_cast<T*>(a.grad_weight) + li0 + p, gw);
        } else {  // several channel chunks of one head add into the same slot (fp32 only)
          if (a.grad_loc) {
            atomicAdd(a.grad_loc + 2 * (li0 + p), gx * unnormalize_scale(a.lvl_w[l], a.flags));
            atomicAdd(a.grad_loc + 2 * (li0 + p) + 1, gy * unnormalize_scale(a.lvl_h[l], a.flags));
          }
          if (a.grad_weight) atomicAdd(reinterpret_cast<float*>(a.grad_weight) + li0 + p, gw);
        }
      }
    }
  }
}

__global__ void cast_f32_to_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 f = src[i];
    dst[i] = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
  }
}

// ------------------------------------------------------------------ host side
static int pick_lpg(int Dh, int vec) {
  const int nvec = Dh / vec;
  for (int lpg = 32; lpg >= 1; lpg >>= 1)
    if (nvec % lpg == 0) return lpg;
  return 1;
}

static int fill_args(MsdaArgs& a, const int32_t* spatial_hw, const int32_t* level_start, int B, int S, int h,
                     int Dh, int Nq, int L, int P, int query_w, int dtype, int flags) {
  BEVR_REQUIRE(dtype == BEVR_F32 || dtype == BEVR_BF16, "msda: dtype must be 0 (f32) or 1 (bf16), got %d", dtype);
  BEVR_REQUIRE(B > 0 && S > 0 && h > 0 && Dh > 0 && Nq > 0 && L > 0 && P > 0, "msda: all dimensions must be > 0");
  BEVR_REQUIRE(L <= BEVR_MAX_LEVELS, "msda: at most %d levels (got %d)", BEVR_MAX_LEVELS, L);
  const int vec = dtype == BEVR_F32 ? 4 : 8;
  BEVR_REQUIRE(Dh % vec == 0, "msda: Dh*sizeof(dtype) must be a multiple of 16 bytes (Dh=%d)", Dh);
  BEVR_REQUIRE(spatial_hw && level_start, "msda: spatial_hw / level_start are NULL");
  BEVR_REQUIRE((size_t)S * h * Dh / vec < (1u << 31), "msda: one sample's value exceeds 2^31 vectors");
  a.B = B; a.S = S; a.h = h; a.Dh = Dh; a.Nq = Nq; a.L = L; a.P = P; a.flags = flags;
  for (int l = 0; l < L; ++l) {
    a.lvl_h[l] = spatial_hw[2 * l]; a.lvl_w[l] = spatial_hw[2 * l + 1]; a.lvl_start[l] = level_start[l];
    BEVR_REQUIRE(a.lvl_h[l] > 0 && a.lvl_w[l] > 0 && a.lvl_start[l] >= 0 &&
                 a.lvl_start[l] + a.lvl_h[l] * a.lvl_w[l] <= S, "msda: level %d (%dx%d @%d) outside S=%d",
                 l, a.lvl_h[l], a.lvl_w[l], a.lvl_start[l], S);
  }
  if (query_w > 0) {
    BEVR_REQUIRE(Nq % query_w == 0, "msda: query_w=%d does not divide Nq=%d", query_w, Nq);
    a.qw = query_w; a.qh = Nq / query_w; a.tw = 8; a.th = 8;
  } else {
    a.qw = Nq; a.qh = 1; a.tw = 64; a.th = 1;
  }
  a.tiles_x = (a.qw + a.tw - 1) / a.tw;
  a.tiles_y = (a.qh + a.th - 1) / a.th;
  const int lpg = pick_lpg(Dh, vec);
  a.nchunk = Dh / (lpg * vec);
  return 0;
}

// tile-shape override for tuning sweeps (process-wide; 0 = default)
static int g_tile_w = 0, g_tile_h = 0, g_threads = 0;

template <typename T, int LPG>
static int launch(const MsdaArgs& a, bool bwd, cudaStream_t st) {
  const int threads = g_threads > 0 ? g_threads : 256;
  const int LP = a.L * a.P;
  const size_t smem = (size_t)(threads / 32) * 2 * (32 / LPG) * (LP + 1) * 16;
  BEVR_REQUIRE(smem <= 200 * 1024, "msda: L*P=%d needs %zu B of shared memory", LP, smem);
  const long long grid = (long long)a.B * a.h * a.nchunk * a.tiles_x * a.tiles_y;
  BEVR_REQUIRE(grid < (1ll << 31), "msda: grid too large");
  auto kern = threads > 256 ? (bwd ? msda_bwd_kernel<T, LPG, 1024> : msda_fwd_kernel<T, LPG, 1024>)
                            : (bwd ? msda_bwd_kernel<T, LPG, 256> : msda_fwd_kernel<T, LPG, 256>);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)grid, threads, smem, st>>>(a);
  return check_launch(bwd ? "msda_bwd" : "msda_fwd");
}

template <typename T>
static int dispatch(const MsdaArgs& a, bool bwd, cudaStream_t st) {
  const int lpg = pick_lpg(a.Dh, Vec<T>::N);
  switch (lpg) {
    case 32: return launch<T, 32>(a, bwd, st);
    case 16: return launch<T, 16>(a, bwd, st);
    case 8: return launch<T, 8>(a, bwd, st);
    case 4: return launch<T, 4>(a, bwd, st);
    case 2: return launch<T, 2>(a, bwd, st);
    default: return launch<T, 1>(a, bwd, st);
  }
}

}  // namespace bevr

using namespace bevr;

extern "C" void bevr_msda_set_tile(int tile_w, int tile_h, int threads) {
  g_tile_w = tile_w; g_tile_h = tile_h; g_threads = threads;
}

static void apply_tile_override(MsdaArgs& a) {
  if (g_tile_w > 0 && g_tile_h > 0) {
    if (a.qh == 1) { a.tw = g_tile_w * g_tile_h; a.th = 1; }
    else { a.tw = g_tile_w; a.th = g_tile_h; }
    a.tiles_x = (a.qw + a.tw - 1) / a.tw;
    a.tiles_y = (a.qh + a.th - 1) / a.th;
  }
}

extern "C" int bevr_msda_fwd(const void* value, const int32_t* spatial_hw, const int32_t* level_start,
                             const float* sampling_loc, const float* ref_points, const void* attn_weight,
                             void* out, int B, int S, int h, int Dh, int Nq, int L, int P, int query_w,
                             int dtype, int flags, void* stream) {
  MsdaArgs a = {};
  if (int rc = fill_args(a, spatial_hw, level_start, B, S, h, Dh, Nq, L, P, query_w, dtype, flags)) return rc;
  BEVR_REQUIRE(value && sampling_loc && out, "msda_fwd: value / sampling_loc / out must not be NULL");
  apply_tile_override(a);
  a.value = value; a.loc = sampling_loc; a.ref = ref_points; a.weight = attn_weight; a.out = out;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == BEVR_F32 ? dispatch<float>(a, false, st) : dispatch<__nv_bfloat16>(a, false, st);
}

extern "C" size_t bevr_msda_bwd_workspace_bytes(int B, int S, int h, int Dh, int dtype) {
  return dtype == BEVR_BF16 ? (size_t)B * S * h * Dh * sizeof(float) : 0;
}

extern "C" int bevr_msda_bwd(const void* value, const int32_t* spatial_hw, const int32_t* level_start,
                             const float* sampling_loc, const float* ref_points, const void* attn_weight,
                             const void* grad_out, void* grad_value, float* grad_loc, void* grad_weight,
                             void* workspace, int B, int S, int h, int Dh, int Nq, int L, int P, int query_w,
                             int dtype, int flags, void* stream) {
  MsdaArgs a = {};
  if (int rc = fill_args(a, spatial_hw, level_start, B, S, h, Dh, Nq, L, P, query_w, dtype, flags)) return rc;
  BEVR_REQUIRE(value && sampling_loc && grad_out, "msda_bwd: value / sampling_loc / grad_out must not be NULL");
  BEVR_REQUIRE(!(attn_weight == nullptr && grad_weight != nullptr), "msda_bwd: grad_weight without attn_weight");
  BEVR_REQUIRE(!(dtype == BEVR_BF16 && grad_value && !workspace), "msda_bwd: bf16 grad_value needs the fp32 workspace");
  BEVR_REQUIRE(!(a.nchunk > 1 && dtype == BEVR_BF16 && grad_weight),
               "msda_bwd: bf16 grad_weight with Dh > 256 is not supported");
  apply_tile_override(a);
  cudaStream_t st = (cudaStream_t)stream;
  a.value = value; a.loc = sampling_loc; a.ref = ref_points; a.weight = attn_weight; a.grad_out = grad_out;
  a.grad_loc = grad_loc; a.grad_weight = grad_weight;
  const size_t n = (size_t)B * S * h * Dh;
  a.grad_value = nullptr;
  if (grad_value) {
    a.grad_value = dtype == BEVR_F32 ? (float*)grad_value : (float*)workspace;
    if (!(flags & BEVR_ACCUMULATE) || dtype == BEVR_BF16) cudaMemsetAsync(a.grad_value, 0, n * sizeof(float), st);
  }
  if (a.nchunk > 1) {  // chunks accumulate with atomics -> outputs must start at zero
    const size_t nl = (size_t)B * Nq * h * L * P;
    if (grad_loc) cudaMemsetAsync(grad_loc, 0, nl * 2 * sizeof(float), st);
    if (grad_weight) cudaMemsetAsync(grad_weight, 0, nl * sizeof(float), st);
  }
  int rc = dtype == BEVR_F32 ? dispatch<float>(a, true, st) : dispatch<__nv_bfloat16>(a, true, st);
  if (rc) return rc;
  if (grad_value && dtype == BEVR_BF16) {
    cast_f32_to_bf16_kernel<<<148 * 8, 256, 0, st>>>((const float4*)workspace, (uint2*)grad_value, n / 4);
    return check_launch("msda_bwd cast");
  }
  return 0;
}
