// K1 / K2: multi-scale deformable-attention gather, forward and backward (sm_100a).
//
// Replaces F.grid_sample (+ its autograd) at TSA_deform_attn.py:210-217 and
// SCA_deform_attn.py:290-301 of the reference and generalises it to the
// (L levels, P points, h heads) operator of the north star.
//
// Design (DESIGN.md section "K1/K2"):
//  * value is channels-last (B,S,h,Dh): one bilinear tap of one head is one contiguous
//    Dh*sizeof(T) vector (128 B for fp32/Dh=32 = one L1 line = one L1 wavefront).
//  * A "group" of LPG lanes owns one (query, head[, channel chunk]) item; each lane holds NV
//    channels (16 B in forward), so every tap is fetched by LPG coalesced vector loads.
//  * A CTA owns a 2-D tile of the query grid for ONE head: all taps of the tile fall in a small
//    window of every level, so they are served by L1 after the first touch and HBM sees each
//    value line about once (L2 keeps the (b,head) slice while its tiles are in flight).
//  * Tap set-up (unnormalise, floor, 4 offsets, 4 weights, validity) is done ONCE per sampling
//    point by one lane of the owning group and staged in shared memory; the gather loop is
//    branch-free: LDS.128 x2 -> (IMAD.WIDE.U32 + LDG) x4 -> FFMA.  Out-of-range taps get weight 0
//    and a safe offset.
//  * Backward recomputes the taps, accumulates grad_value with vector RED.ADD.F32x4 into an fp32
//    buffer, and writes grad_loc / grad_weight directly (intra-group shuffle reduction, no atomics).
//  * The kernels are instruction-issue / L1-wavefront / L2-RED-sector bound (profiles/), so the
//    inner loops avoid integer division and 64-bit address chains.
#include "bevr_common.cuh"

// tunables (overridable with -D for sweeps; defaults are the measured best, see profiles/r1_results.md)
#ifndef MSDA_FWD_MINB
#define MSDA_FWD_MINB 6
#endif
#ifndef MSDA_BWD_MINB
#define MSDA_BWD_MINB 4
#endif
#ifndef MSDA_FWD_UNROLL
#define MSDA_FWD_UNROLL 4
#endif
#ifndef MSDA_BWD_UNROLL
#define MSDA_BWD_UNROLL 2
#endif
#define MSDA_STR2(x) #x
#define MSDA_STR(x) MSDA_STR2(x)

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
  int tw_shift, th;  // query tile: width 1 << tw_shift, height th
  int tiles_x, tiles_y;
  int nchunk;        // channel chunks per head (Dh / (LPG*NV))
  int p_shift;       // log2(P) if P is a power of two, else -1
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

// NV channels of T starting at a byte address (read-only path); fp32 accumulators out
template <typename T, int NV> struct Chan;
template <> struct Chan<float, 4> {
  static constexpr int BYTES = 16;
  __device__ static void load(const char* p, float* f) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static void store(char* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Chan<__nv_bfloat16, 8> {
  static constexpr int BYTES = 16;
  __device__ static void load(const char* p, float* f) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
    f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
  }
  __device__ static void store(char* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                              pack_bf16(f[6], f[7]));
  }
};
template <> struct Chan<__nv_bfloat16, 4> {   // backward: 4 channels/lane so a lane's RED.F32x4 is contiguous
  static constexpr int BYTES = 8;
  __device__ static void load(const char* p, float* f) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  }
};

template <typename T> __device__ __forceinline__ float load_scalar(const T* p);
template <> __device__ __forceinline__ float load_scalar<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_scalar<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
__device__ __forceinline__ void store_scalar(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_scalar(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct TapRec {
  unsigned off[4];   // vector index (units of NV channels) inside the sample; 0 if out of range
  float fx, fy, aw;
  unsigned mask;     // bit t set: tap t (nw, ne, sw, se) is inside the level
};

// one sampling point: coordinates -> taps.  `vec_per_pix` = h*Dh/NV
template <typename T>
__device__ __forceinline__ TapRec make_tap(const MsdaArgs& a, size_t item, int bq, int p, int vec_per_pix) {
  TapRec t;
  const int LP = a.L * a.P;
  const int l = a.p_shift >= 0 ? (p >> a.p_shift) : (p / a.P);
  const int H = a.lvl_h[l], W = a.lvl_w[l];
  const size_t li = item * LP + p;
  float2 xy = __ldg(reinterpret_cast<const float2*>(a.loc) + li);
  if (a.ref) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(a.ref) + (size_t)bq * a.L + l);
    if (a.flags & BEVR_OFFSET_PIXELS) { xy.x = xy.x / (float)W; xy.y = xy.y / (float)H; }
    xy.x += r.x; xy.y += r.y;
  }
  t.aw = a.weight ? load_scalar(reinterpret_cast<const T*>(a.weight) + li) : 1.f;
  const float px = unnormalize(xy.x, W, a.flags), py = unnormalize(xy.y, H, a.flags);
  // negated comparisons: NaN / inf coordinates contribute nothing (as on the CPU path)
  const bool any = (px > -1.f) && (px < (float)W) && (py > -1.f) && (py < (float)H);
  const float x0f = floorf(px), y0f = floorf(py);
  t.fx = px - x0f;
  t.fy = py - y0f;
  const int x0 = any ? (int)x0f : -2, y0 = any ? (int)y0f : -2;
  const bool vx0 = x0 >= 0, vx1 = x0 + 1 < W && x0 + 1 >= 0, vy0 = y0 >= 0, vy1 = y0 + 1 < H && y0 + 1 >= 0;
  const unsigned base = (unsigned)(a.lvl_start[l] + y0 * W + x0) * (unsigned)vec_per_pix;
  const unsigned dx = (unsigned)vec_per_pix, dy = (unsigned)(W * vec_per_pix);
  t.mask = (vx0 && vy0 ? 1u : 0u) | (vx1 && vy0 ? 2u : 0u) | (vx0 && vy1 ? 4u : 0u) | (vx1 && vy1 ? 8u : 0u);
  t.off[0] = (t.mask & 1u) ? base : 0u;
  t.off[1] = (t.mask & 2u) ? base + dx : 0u;
  t.off[2] = (t.mask & 4u) ? base + dy : 0u;
  t.off[3] = (t.mask & 8u) ? base + dy + dx : 0u;
  return t;
}

struct TileCtx {
  int b, head, chunk, q0x, q0y;
};
__device__ __forceinline__ TileCtx tile_ctx(const MsdaArgs& a) {
  TileCtx c;
  const int n_tiles = a.tiles_x * a.tiles_y;
  int blk = blockIdx.x;
  const int tile = blk % n_tiles; blk /= n_tiles;
  const int vh = blk % (a.h * a.nchunk);
  c.b = blk / (a.h * a.nchunk);
  c.head = vh / a.nchunk; c.chunk = vh - c.head * a.nchunk;
  const int ty = tile / a.tiles_x;
  c.q0y = ty * a.th; c.q0x = (tile - ty * a.tiles_x) << a.tw_shift;
  return c;
}
__device__ __forceinline__ bool tile_query(const MsdaArgs& a, const TileCtx& c, int qi, int& q) {
  const int y = c.q0y + (qi >> a.tw_shift), x = c.q0x + (qi & ((1 << a.tw_shift) - 1));
  q = y * a.qw + x;
  return (y < a.qh) && (x < a.qw);
}

__device__ __forceinline__ const char* addr_u32(const char* base, unsigned idx, unsigned scale) {
  return base + (size_t)idx * scale;       // IMAD.WIDE.U32
}
// Hide a per-thread base pointer's provenance from the optimiser: otherwise ptxas keeps the kernel-parameter
// part in uniform registers and rebuilds every tap address with IMAD.WIDE + IADD3 + IADD3.X (3 issue slots
// instead of 1) -- the gather loops are issue-bound.
__device__ __forceinline__ const char* opaque(const char* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// ------------------------------------------------------------------ forward
template <typename T, int NV, int LPG>
__global__ void __launch_bounds__(256, MSDA_FWD_MINB) msda_fwd_kernel(const __grid_constant__ MsdaArgs a) {
  constexpr int GPW = 32 / LPG;                    // groups per warp
  constexpr unsigned VB = Chan<T, NV>::BYTES;
  extern __shared__ uint4 smem[];
  const int LP = a.L * a.P;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int nwarp = blockDim.x >> 5;
  const int g = lane / LPG, lig = lane % LPG;
  uint4* s_off = smem + warp * (2 * GPW * (LP + 1)) + g * (LP + 1);
  float4* s_w = reinterpret_cast<float4*>(smem + warp * (2 * GPW * (LP + 1)) + GPW * (LP + 1)) + g * (LP + 1);

  const TileCtx c = tile_ctx(a);
  const int vec_per_pix = a.h * a.Dh / NV;
  const int head_vec = c.head * (a.Dh / NV) + c.chunk * LPG + lig;
  const char* vbase =
      opaque(reinterpret_cast<const char*>(a.value) + ((size_t)c.b * a.S * vec_per_pix + head_vec) * VB);
  const int tq = a.th << a.tw_shift;

  for (int base = warp * GPW; base < tq; base += nwarp * GPW) {
    int q;
    const bool live = tile_query(a, c, base + g, q);
    const int bq = c.b * a.Nq + q;
    const size_t item = (size_t)bq * a.h + c.head;
    __syncwarp();
    // ---- stage the taps of this group's item: lane `lig` sets up points lig, lig+LPG, ...
    for (int p = lig; p < LP; p += LPG) {
      uint4 o = make_uint4(0, 0, 0, 0);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const TapRec t = make_tap<T>(a, item, bq, p, vec_per_pix);
        const float wx1 = t.fx * t.aw, wx0 = t.aw - wx1, wy1 = t.fy, wy0 = 1.f - t.fy;
        o = make_uint4(t.off[0], t.off[1], t.off[2], t.off[3]);
        w = make_float4((t.mask & 1u) ? wx0 * wy0 : 0.f, (t.mask & 2u) ? wx1 * wy0 : 0.f,
                        (t.mask & 4u) ? wx0 * wy1 : 0.f, (t.mask & 8u) ? wx1 * wy1 : 0.f);
      }
      s_off[p] = o;
      s_w[p] = w;
    }
    __syncwarp();
    // ---- branch-free gather
    float acc[NV];
#pragma unroll
    for (int ch = 0; ch < NV; ++ch) acc[ch] = 0.f;
_Pragma(MSDA_STR(unroll MSDA_FWD_UNROLL))
    for (int p = 0; p < LP; ++p) {
      const uint4 o = s_off[p];
      const float4 w = s_w[p];
      float f0[NV], f1[NV], f2[NV], f3[NV];
      Chan<T, NV>::load(addr_u32(vbase, o.x, VB), f0);
      Chan<T, NV>::load(addr_u32(vbase, o.y, VB), f1);
      Chan<T, NV>::load(addr_u32(vbase, o.z, VB), f2);
      Chan<T, NV>::load(addr_u32(vbase, o.w, VB), f3);
#pragma unroll
      for (int ch = 0; ch < NV; ++ch) {
        acc[ch] = fmaf(w.x, f0[ch], acc[ch]);
        acc[ch] = fmaf(w.y, f1[ch], acc[ch]);
        acc[ch] = fmaf(w.z, f2[ch], acc[ch]);
        acc[ch] = fmaf(w.w, f3[ch], acc[ch]);
      }
    }
    if (live)
      Chan<T, NV>::store(reinterpret_cast<char*>(a.out) + (item * (a.Dh / NV) + c.chunk * LPG + lig) * VB, acc);
  }
}

// ------------------------------------------------------------------ backward
template <typename T, int LPG>
__global__ void __launch_bounds__(256, MSDA_BWD_MINB) msda_bwd_kernel(const __grid_constant__ MsdaArgs a) {
  constexpr int NV = 4;
  constexpr int GPW = 32 / LPG;
  constexpr unsigned VB = Chan<T, NV>::BYTES;
  extern __shared__ uint4 smem[];
  const int LP = a.L * a.P;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int nwarp = blockDim.x >> 5;
  const int g = lane / LPG, lig = lane % LPG;
  uint4* s_off = smem + warp * (2 * GPW * (LP + 1)) + g * (LP + 1);
  float4* s_w = reinterpret_cast<float4*>(smem + warp * (2 * GPW * (LP + 1)) + GPW * (LP + 1)) + g * (LP + 1);

  const TileCtx c = tile_ctx(a);
  const int vec_per_pix = a.h * a.Dh / NV;
  const int head_vec = c.head * (a.Dh / NV) + c.chunk * LPG + lig;
  const size_t vec0 = (size_t)c.b * a.S * vec_per_pix + head_vec;
  const char* vbase = opaque(reinterpret_cast<const char*>(a.value) + vec0 * VB);
  const char* gvbase = a.grad_value ? opaque(reinterpret_cast<const char*>(a.grad_value) + vec0 * 16) : nullptr;
  const int tq = a.th << a.tw_shift;
  const unsigned gmask = (LPG == 32) ? 0xffffffffu : (((1u << LPG) - 1u) << (g * LPG));

  for (int base = warp * GPW; base < tq; base += nwarp * GPW) {
    int q;
    const bool live = tile_query(a, c, base + g, q);
    const int bq = c.b * a.Nq + q;
    const size_t item = (size_t)bq * a.h + c.head;
    __syncwarp();
    for (int p = lig; p < LP; p += LPG) {
      uint4 o = make_uint4(0, 0, 0, 0);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const TapRec t = make_tap<T>(a, item, bq, p, vec_per_pix);
        o = make_uint4(t.off[0], t.off[1], t.off[2], t.off[3]);
        w = make_float4(t.fx, t.fy, t.aw, __uint_as_float(t.mask));
      }
      s_off[p] = o;
      s_w[p] = w;
    }
    __syncwarp();
    float go[NV];
#pragma unroll
    for (int ch = 0; ch < NV; ++ch) go[ch] = 0.f;
    if (live)
      Chan<T, NV>::load(reinterpret_cast<const char*>(a.grad_out) + (item * (a.Dh / NV) + c.chunk * LPG + lig) * VB, go);
    const size_t li0 = item * LP;
    int l = 0, pl = 0;                                   // level of point p, tracked without division
    float sxl = unnormalize_scale(a.lvl_w[0], a.flags), syl = unnormalize_scale(a.lvl_h[0], a.flags);
_Pragma(MSDA_STR(unroll MSDA_BWD_UNROLL))
    for (int p = 0; p < LP; ++p) {
      const uint4 o = s_off[p];
      const float4 w = s_w[p];
      const unsigned mask = __float_as_uint(w.w);
      const unsigned offs[4] = {o.x, o.y, o.z, o.w};
      const float wx1 = w.x, wx0 = 1.f - w.x, wy1 = w.y, wy0 = 1.f - w.y, aw = w.z;
      float f[4][NV];
#pragma unroll
      for (int t = 0; t < 4; ++t) Chan<T, NV>::load(addr_u32(vbase, offs[t], VB), f[t]);
      float d[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float s = go[0] * f[t][0];
#pragma unroll
        for (int ch = 1; ch < NV; ++ch) s = fmaf(go[ch], f[t][ch], s);
        d[t] = (mask >> t) & 1u ? s : 0.f;               // out-of-range taps read a safe line: drop them
      }
      if (gvbase) {
        const float bw[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float wt = aw * bw[t];
          float* dst = reinterpret_cast<float*>(const_cast<char*>(addr_u32(gvbase, offs[t], 16)));
          red_add_v4_if(((mask >> t) & 1u) && wt != 0.f, dst, wt * go[0], wt * go[1], wt * go[2], wt * go[3]);
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
      if (live && lig == (p & (LPG - 1))) {
        if (a.nchunk == 1) {
          if (a.grad_loc) reinterpret_cast<float2*>(a.grad_loc)[li0 + p] = make_float2(gx * sxl, gy * syl);
          if (a.grad_weight) store_scalar(reinterpret_cast<T*>(a.grad_weight) + li0 + p, gw);
        } else {  // several channel chunks of one head add into the same slot (fp32 only)
          if (a.grad_loc) {
            atomicAdd(a.grad_loc + 2 * (li0 + p), gx * sxl);
            atomicAdd(a.grad_loc + 2 * (li0 + p) + 1, gy * syl);
          }
          if (a.grad_weight) atomicAdd(reinterpret_cast<float*>(a.grad_weight) + li0 + p, gw);
        }
      }
      if (++pl == a.P) {
        pl = 0; ++l;
        if (l < a.L) { sxl = unnormalize_scale(a.lvl_w[l], a.flags); syl = unnormalize_scale(a.lvl_h[l], a.flags); }
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
static int pick_lpg(int Dh, int nv) {
  const int nvec = Dh / nv;
  for (int lpg = 32; lpg >= 1; lpg >>= 1)
    if (nvec % lpg == 0) return lpg;
  return 1;
}

// tile-shape override for tuning sweeps (process-wide; 0 = default)
static int g_tile_w = 0, g_tile_h = 0;

static int fill_args(MsdaArgs& a, const int32_t* spatial_hw, const int32_t* level_start, int B, int S, int h,
                     int Dh, int Nq, int L, int P, int query_w, int dtype, int flags, int nv) {
  BEVR_REQUIRE(dtype == BEVR_F32 || dtype == BEVR_BF16, "msda: dtype must be 0 (f32) or 1 (bf16), got %d", dtype);
  BEVR_REQUIRE(B >= 0 && S > 0 && h > 0 && Dh > 0 && Nq >= 0 && L > 0 && P > 0,
               "msda: S, h, Dh, L, P must be > 0 and B, Nq >= 0");
  BEVR_REQUIRE(L <= BEVR_MAX_LEVELS, "msda: at most %d levels (got %d)", BEVR_MAX_LEVELS, L);
  const int vec = dtype == BEVR_F32 ? 4 : 8;
  BEVR_REQUIRE(Dh % vec == 0, "msda: Dh*sizeof(dtype) must be a multiple of 16 bytes (Dh=%d)", Dh);
  BEVR_REQUIRE(spatial_hw && level_start, "msda: spatial_hw / level_start are NULL");
  BEVR_REQUIRE((size_t)S * h * Dh / 4 < (1ull << 32), "msda: one sample's value exceeds 2^32 vectors");
  BEVR_REQUIRE((size_t)B * Nq < (1ull << 31), "msda: B*Nq exceeds 2^31");
  a.B = B; a.S = S; a.h = h; a.Dh = Dh; a.Nq = Nq; a.L = L; a.P = P; a.flags = flags;
  a.p_shift = -1;
  for (int s = 0; s < 16; ++s)
    if ((1 << s) == P) a.p_shift = s;
  for (int l = 0; l < L; ++l) {
    a.lvl_h[l] = spatial_hw[2 * l]; a.lvl_w[l] = spatial_hw[2 * l + 1]; a.lvl_start[l] = level_start[l];
    BEVR_REQUIRE(a.lvl_h[l] > 0 && a.lvl_w[l] > 0 && a.lvl_start[l] >= 0 &&
                 a.lvl_start[l] + a.lvl_h[l] * a.lvl_w[l] <= S, "msda: level %d (%dx%d @%d) outside S=%d",
                 l, a.lvl_h[l], a.lvl_w[l], a.lvl_start[l], S);
  }
  int tw = 8, th = 8;
  if (g_tile_w > 0 && g_tile_h > 0) { tw = g_tile_w; th = g_tile_h; }
  if (query_w > 0) {
    BEVR_REQUIRE(Nq % query_w == 0, "msda: query_w=%d does not divide Nq=%d", query_w, Nq);
    a.qw = query_w; a.qh = Nq / query_w;
  } else {
    a.qw = Nq > 0 ? Nq : 1; a.qh = 1; tw = tw * th; th = 1;
  }
  a.tw_shift = 0;
  while ((2 << a.tw_shift) <= tw) ++a.tw_shift;       // round the tile width down to a power of two
  a.th = th;
  a.tiles_x = (a.qw + (1 << a.tw_shift) - 1) >> a.tw_shift;
  a.tiles_y = (a.qh + a.th - 1) / a.th;
  const int lpg = pick_lpg(Dh, nv);
  a.nchunk = Dh / (lpg * nv);
  return 0;
}

template <typename K>
static int launch(K kern, const MsdaArgs& a, int lpg, const char* what, cudaStream_t st) {
  const int threads = 256;
  const int LP = a.L * a.P;
  const size_t smem = (size_t)(threads / 32) * 2 * (32 / lpg) * (LP + 1) * 16;
  BEVR_REQUIRE(smem <= 200 * 1024, "msda: L*P=%d needs %zu B of shared memory", LP, smem);
  const long long grid = (long long)a.B * a.h * a.nchunk * a.tiles_x * a.tiles_y;
  BEVR_REQUIRE(grid < (1ll << 31), "msda: grid too large");
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)grid, threads, smem, st>>>(a);
  return check_launch(what);
}

template <typename T, int NV>
static int dispatch_fwd(const MsdaArgs& a, cudaStream_t st) {
  const int lpg = pick_lpg(a.Dh, NV);
  switch (lpg) {
    case 32: return launch(msda_fwd_kernel<T, NV, 32>, a, 32, "msda_fwd", st);
    case 16: return launch(msda_fwd_kernel<T, NV, 16>, a, 16, "msda_fwd", st);
    case 8: return launch(msda_fwd_kernel<T, NV, 8>, a, 8, "msda_fwd", st);
    case 4: return launch(msda_fwd_kernel<T, NV, 4>, a, 4, "msda_fwd", st);
    case 2: return launch(msda_fwd_kernel<T, NV, 2>, a, 2, "msda_fwd", st);
    default: return launch(msda_fwd_kernel<T, NV, 1>, a, 1, "msda_fwd", st);
  }
}
template <typename T>
static int dispatch_bwd(const MsdaArgs& a, cudaStream_t st) {
  const int lpg = pick_lpg(a.Dh, 4);
  switch (lpg) {
    case 32: return launch(msda_bwd_kernel<T, 32>, a, 32, "msda_bwd", st);
    case 16: return launch(msda_bwd_kernel<T, 16>, a, 16, "msda_bwd", st);
    case 8: return launch(msda_bwd_kernel<T, 8>, a, 8, "msda_bwd", st);
    case 4: return launch(msda_bwd_kernel<T, 4>, a, 4, "msda_bwd", st);
    case 2: return launch(msda_bwd_kernel<T, 2>, a, 2, "msda_bwd", st);
    default: return launch(msda_bwd_kernel<T, 1>, a, 1, "msda_bwd", st);
  }
}

}  // namespace bevr

using namespace bevr;

extern "C" void bevr_msda_set_tile(int tile_w, int tile_h, int threads) {
  (void)threads;
  g_tile_w = tile_w; g_tile_h = tile_h;
}

extern "C" int bevr_msda_fwd(const void* value, const int32_t* spatial_hw, const int32_t* level_start,
                             const float* sampling_loc, const float* ref_points, const void* attn_weight,
                             void* out, int B, int S, int h, int Dh, int Nq, int L, int P, int query_w,
                             int dtype, int flags, void* stream) {
  MsdaArgs a = {};
  if (int rc = fill_args(a, spatial_hw, level_start, B, S, h, Dh, Nq, L, P, query_w, dtype, flags,
                         dtype == BEVR_F32 ? 4 : 8))
    return rc;
  if (B == 0 || Nq == 0) return 0;                 // empty batch / empty query list: nothing to do
  BEVR_REQUIRE(value && sampling_loc && out, "msda_fwd: value / sampling_loc / out must not be NULL");
  a.value = value; a.loc = sampling_loc; a.ref = ref_points; a.weight = attn_weight; a.out = out;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == BEVR_F32 ? dispatch_fwd<float, 4>(a, st) : dispatch_fwd<__nv_bfloat16, 8>(a, st);
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
  if (int rc = fill_args(a, spatial_hw, level_start, B, S, h, Dh, Nq, L, P, query_w, dtype, flags, 4)) return rc;
  if (B == 0) return 0;
  if (Nq == 0) {                                   // no queries: grad_value is all zeros, nothing else to write
    if (grad_value && !(flags & BEVR_ACCUMULATE))
      cudaMemsetAsync(grad_value, 0, (size_t)B * S * h * Dh * (dtype == BEVR_F32 ? 4 : 2), (cudaStream_t)stream);
    return 0;
  }
  BEVR_REQUIRE(value && sampling_loc && grad_out, "msda_bwd: value / sampling_loc / grad_out must not be NULL");
  BEVR_REQUIRE(!(attn_weight == nullptr && grad_weight != nullptr), "msda_bwd: grad_weight without attn_weight");
  BEVR_REQUIRE(!(dtype == BEVR_BF16 && grad_value && !workspace), "msda_bwd: bf16 grad_value needs the fp32 workspace");
  BEVR_REQUIRE(!(a.nchunk > 1 && dtype == BEVR_BF16 && grad_weight),
               "msda_bwd: bf16 grad_weight with Dh > 128 is not supported");
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
  int rc = dtype == BEVR_F32 ? dispatch_bwd<float>(a, st) : dispatch_bwd<__nv_bfloat16>(a, st);
  if (rc) return rc;
  if (grad_value && dtype == BEVR_BF16) {
    cast_f32_to_bf16_kernel<<<148 * 8, 256, 0, st>>>((const float4*)workspace, (uint2*)grad_value, n / 4);
    return check_launch("msda_bwd cast");
  }
  return 0;
}
