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
#define MSDA_FWD_MINB 4
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
  int lp_shift;      // log2(L*P) likewise
  int sort_lvl_off;  // sorted backward: byte offset of the per-level table in shared memory
  int sort_go_smem;  // sorted backward: the tile's grad_out rows are staged in shared memory (else read through L1)
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

// ------------------------------------------------------------------ backward, direct (one RED per tap)
// Body of the query-major backward: recompute the taps, gather value for the per-point dot products, scatter grad_value
// with one vector RED per tap.  It is the whole kernel for Dh > 128 (several channel chunks per head) and the per-CTA
// fallback of the sorted backward below (sampling windows too large to bin).
template <typename T, int LPG>
__device__ __forceinline__ void bwd_direct_tile(const MsdaArgs& a, const TileCtx& c, uint4* smem) {
  constexpr int NV = 4;
  constexpr int GPW = 32 / LPG;
  constexpr unsigned VB = Chan<T, NV>::BYTES;
  const int LP = a.L * a.P;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int nwarp = blockDim.x >> 5;
  const int g = lane / LPG, lig = lane % LPG;
  uint4* s_off = smem + warp * (2 * GPW * (LP + 1)) + g * (LP + 1);
  float4* s_w = reinterpret_cast<float4*>(smem + warp * (2 * GPW * (LP + 1)) + GPW * (LP + 1)) + g * (LP + 1);

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

template <typename T, int LPG>
__global__ void __launch_bounds__(256, MSDA_BWD_MINB) msda_bwd_kernel(const __grid_constant__ MsdaArgs a) {
  extern __shared__ uint4 smem[];
  bwd_direct_tile<T, LPG>(a, tile_ctx(a), smem);
}

// ------------------------------------------------------------------ backward, tile-sorted (the default for Dh <= 128)
// The direct backward issues one 128-byte RED per bilinear tap: 655 M RED sectors per launch at BASELINE config 2, which
// is the whole 3 ms (the L2 atomic units retire 216 G sectors/s whichever way the adds leave the SM -- RED.v4, bulk TMA
// reduce, tensor TMA reduce; profiles/r1_msda_ceilings.md).  The only way down is to add fewer, fatter numbers.  A CTA
// that owns an 8x8 tile of queries for one head issues 4096 taps, but they land on only ~600 distinct value pixels:
// this kernel SORTS the CTA's sampling points by the pixel they hit and reduces them in registers before anything
// leaves the SM -- CTA-wide aggregation by address, the "warp-aggregated atomics" of the north star taken one level up.
//
//   1. every thread computes its share of the tile's points: (x0, y0, fx, fy, weight); a shared-memory bounding box of
//      the touched nw-pixels is formed per level (warp REDUX min/max, then one shared atomic per warp and level);
//   2. counting sort by (level, y0, x0): shared-memory histogram over the bounding boxes (ATOMS returns the rank inside
//      the bin), exclusive scan, scatter of 16-byte records {run id | point id, fx, fy, weight};
//   3. segmented reduction: a group of LPG lanes (one 128-byte channel row) walks an equal share of the sorted records.
//      While the nw pixel stays the same ("a run") the 2x2 value rows AND the 2x2 grad_value accumulators live in
//      registers; a record costs one shared-memory read of its query's grad_out row, 32 FMAs per lane, and a 4-shuffle
//      transposed reduction of (d/dweight, d/dx, d/dy).  When the run changes to the x-neighbour the right column slides
//      to the left (2 REDs + 2 row loads instead of 4 + 4); REDs are issued per RUN, not per tap: ~5x fewer;
//   4. grad_loc / grad_weight are staged in shared memory by point id and written coalesced.
// A CTA whose bounding boxes do not fit the histogram (offsets spread over > MAXBINS pixels) runs bwd_direct_tile.
#ifndef MSDA_SORT_UNROLL
#define MSDA_SORT_UNROLL 1         // unroll of the record walk
#endif
#ifndef MSDA_SORT_MINB
#define MSDA_SORT_MINB 3           // resident CTAs per SM the register budget is set for (4 channels per lane)
#endif
#ifndef MSDA_SORT_MINB8
#define MSDA_SORT_MINB8 2          // ... with 8 channels per lane (twice the register-resident rows)
#endif

#ifndef MSDA_SORT_BATCH
#define MSDA_SORT_BATCH 4          // points per thread whose coordinate loads are issued together in the per-point steps
#endif

constexpr int RUNCLS = 32;   // run-length classes of the sorted backward: 1 .. 31 records, and "32 or more"

struct SortLvl {        // per level, in shared memory
  int xmin, ymin, xmax, ymax;   // bounding box of the nw pixels (x0 in [-1, W-1], y0 in [-1, H-1])
  int bw, off;                  // box width, first bin
  int W, H;
  int start_pix;                // lvl_start
  unsigned dy_vec;              // W * vec_per_pix: one row down, in value vectors
  float sx, sy;                 // d pixel / d location
};

// raw operands of one sampling point (location, reference point, weight): loads only, so that a thread can have the
// loads of several points in flight before it touches the first one
struct PointRaw {
  float2 xy, r;
  float aw;
};
template <typename T>
__device__ __forceinline__ PointRaw point_load(const MsdaArgs& a, size_t item, int bq, int p, int l) {
  PointRaw o;
  const size_t li = item * (a.L * a.P) + p;
  o.xy = __ldg(reinterpret_cast<const float2*>(a.loc) + li);
  o.r = make_float2(0.f, 0.f);
  if (a.ref) o.r = __ldg(reinterpret_cast<const float2*>(a.ref) + (size_t)bq * a.L + l);
  o.aw = a.weight ? load_scalar(reinterpret_cast<const T*>(a.weight) + li) : 1.f;
  return o;
}
__device__ __forceinline__ bool point_finish(const MsdaArgs& a, const PointRaw& o, int l, float& px, float& py) {
  const int H = a.lvl_h[l], W = a.lvl_w[l];
  float2 xy = o.xy;
  if (a.ref) {
    if (a.flags & BEVR_OFFSET_PIXELS) { xy.x = xy.x / (float)W; xy.y = xy.y / (float)H; }
    xy.x += o.r.x; xy.y += o.r.y;
  }
  px = unnormalize(xy.x, W, a.flags); py = unnormalize(xy.y, H, a.flags);
  return (px > -1.f) && (px < (float)W) && (py > -1.f) && (py < (float)H);     // false for NaN / inf
}

// NV fp32 channels of a row of T (NV = 4 or 8) as float2 pairs
template <typename T, int NV>
__device__ __forceinline__ void load_row(const char* p, float2* v) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + j);
      v[2 * j] = make_float2(__uint_as_float(u.x), __uint_as_float(u.y));
      v[2 * j + 1] = make_float2(__uint_as_float(u.z), __uint_as_float(u.w));
    }
  } else if constexpr (NV == 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = make_float2(bf16lo(u.x), bf16hi(u.x)); v[1] = make_float2(bf16lo(u.y), bf16hi(u.y));
    v[2] = make_float2(bf16lo(u.z), bf16hi(u.z)); v[3] = make_float2(bf16lo(u.w), bf16hi(u.w));
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = make_float2(bf16lo(u.x), bf16hi(u.x)); v[1] = make_float2(bf16lo(u.y), bf16hi(u.y));
  }
}

// sum of four values over the LPG lanes of a group, TRANSPOSED for LPG = 4 / 8 (each exchange halves the number of live
// values, 3 / 4 shuffles instead of 8 / 12); returns the index j of the value d[j] this lane ends with in `which`
// (-1: this lane holds a duplicate and must not write)
template <int LPG>
__device__ __forceinline__ float group_sum4(float d0, float d1, float d2, float d3, int lig, int& which) {
  constexpr unsigned FULL = 0xffffffffu;
  if constexpr (LPG >= 16) {
#pragma unroll
    for (int s = LPG / 2; s >= 8; s >>= 1) {
      d0 += __shfl_xor_sync(FULL, d0, s); d1 += __shfl_xor_sync(FULL, d1, s);
      d2 += __shfl_xor_sync(FULL, d2, s); d3 += __shfl_xor_sync(FULL, d3, s);
    }
  }
  if constexpr (LPG >= 8) {
    const bool hi = lig & 4;
    float ka = hi ? d2 : d0, kb = hi ? d3 : d1;
    ka += __shfl_xor_sync(FULL, hi ? d0 : d2, 4);
    kb += __shfl_xor_sync(FULL, hi ? d1 : d3, 4);
    float k1 = (lig & 2) ? kb : ka;
    k1 += __shfl_xor_sync(FULL, (lig & 2) ? ka : kb, 2);
    k1 += __shfl_xor_sync(FULL, k1, 1);
    which = ((lig & 1) || lig >= 8) ? -1 : ((lig >> 1) & 3);
    return k1;
  } else if constexpr (LPG == 4) {
    const bool hi = lig & 2;
    float ka = hi ? d2 : d0, kb = hi ? d3 : d1;
    ka += __shfl_xor_sync(FULL, hi ? d0 : d2, 2);
    kb += __shfl_xor_sync(FULL, hi ? d1 : d3, 2);
    float k1 = (lig & 1) ? kb : ka;
    k1 += __shfl_xor_sync(FULL, (lig & 1) ? ka : kb, 1);
    which = lig;
    return k1;
  } else {
    static_assert(LPG == 4 || LPG >= 8, "group_sum4 serves groups of 4, 8, 16 or 32 lanes");
    which = 0;
    return d0;
  }
}

// T, NV channels per lane (4 or 8), LPG = Dh / NV lanes per sampling point, FAST: L*P and P are powers of two and the
// tile's grad_out rows are staged in shared memory (the shape every benchmark and the reference use)
template <typename T, int NV, int LPG, bool FAST>
__global__ void __launch_bounds__(256, (NV == 8 ? MSDA_SORT_MINB8 : MSDA_SORT_MINB)) msda_bwd_sorted_kernel(const __grid_constant__ MsdaArgs a,
                                                                              const int maxbins) {
  constexpr unsigned VB = NV * sizeof(T);                // bytes of one lane's slice of a value row
  constexpr int H2 = NV / 2;                             // float2 pairs per lane
  extern __shared__ uint4 smem[];
  const int LP = a.L * a.P;
  const int tq = a.th << a.tw_shift;                     // queries per tile
  const int NP = tq * LP;                                // sampling points per CTA (<= 8192)
  const int row4 = a.Dh / 4;                             // float4 per grad_out row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool go_smem = FAST || a.sort_go_smem;
  auto point_q = [&](int i) { return (FAST || a.lp_shift >= 0) ? (i >> a.lp_shift) : (i / LP); };
  auto point_l = [&](int p) { return (FAST || a.p_shift >= 0) ? (p >> a.p_shift) : (p / a.P); };

  // shared memory carve-up (16-byte units first)
  float4* s_w = reinterpret_cast<float4*>(smem);         // [NP] sorted: the four tap weights aw * bilinear weight
  uint4* s_tmp = reinterpret_cast<uint4*>(s_w + NP);     // [NP] unsorted records; later float4 dot products by point
  float4* s_go = reinterpret_cast<float4*>(s_tmp + NP);  // [tq][row4] grad_out rows, fp32
  uint2* s_key = reinterpret_cast<uint2*>(s_go + (go_smem ? tq * row4 : 0));   // [NP] sorted: {pixel, point | mask | level}
  unsigned* s_hist = reinterpret_cast<unsigned*>(s_key + NP);                 // [maxbins + 1]
  unsigned* s_rt = s_hist + maxbins + 4;                                      // [NP] run table {first record | length << 16}
  unsigned short* s_rank = reinterpret_cast<unsigned short*>(s_rt + NP);      // [NP]
  SortLvl* s_lvl = reinterpret_cast<SortLvl*>(reinterpret_cast<char*>(smem) + a.sort_lvl_off);     // [L]
  int* s_misc = reinterpret_cast<int*>(s_lvl + BEVR_MAX_LEVELS);   // [0] nbins or -1 (overflow), [1] runs; [8..16) warp sums
  unsigned* s_h2 = reinterpret_cast<unsigned*>(s_misc + 16);       // [RUNCLS + 1] runs per length class, then class offsets

  const TileCtx c = tile_ctx(a);
  const int vec_per_pix = a.h * a.Dh / NV;

  // ---- 0. initialise the boxes / histogram, stage the tile's grad_out rows
  for (int i = tid; i <= maxbins / 4; i += 256) reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid <= RUNCLS) s_h2[tid] = 0u;
  if (tid < a.L) {
    SortLvl& lv = s_lvl[tid];
    lv.xmin = lv.ymin = INT_MAX; lv.xmax = lv.ymax = INT_MIN;
    lv.W = a.lvl_w[tid]; lv.H = a.lvl_h[tid];
    lv.start_pix = a.lvl_start[tid];
    lv.dy_vec = (unsigned)(a.lvl_w[tid] * vec_per_pix);
    lv.sx = unnormalize_scale(a.lvl_w[tid], a.flags); lv.sy = unnormalize_scale(a.lvl_h[tid], a.flags);
  }
  if (go_smem) {
    for (int i = tid; i < tq * row4; i += 256) {
      const int qi = i / row4, part = i - qi * row4;
      int q;
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
      if (tile_query(a, c, qi, q)) {
        const size_t item = ((size_t)c.b * a.Nq + q) * a.h + c.head;
        Chan<T, 4>::load(reinterpret_cast<const char*>(a.grad_out) + (item * row4 + part) * (4 * sizeof(T)), g4);
      }
      s_go[i] = make_float4(g4[0], g4[1], g4[2], g4[3]);
    }
  }
  __syncthreads();

  // ---- 1. points -> unsorted records + per-level bounding boxes.  The box is grown with shared atomics, but a thread
  //         only issues one when its point lies outside the box it reads first: after the first few points of a tile
  //         almost none do.
  for (int i0 = tid; i0 < NP; i0 += 256 * MSDA_SORT_BATCH) {
    PointRaw raw[MSDA_SORT_BATCH];
    int lv_of[MSDA_SORT_BATCH];                          // level, or -1: no such point / query outside the grid
#pragma unroll
    for (int k = 0; k < MSDA_SORT_BATCH; ++k) {          // all loads first: one round trip per batch, not per point
      const int i = i0 + 256 * k;
      lv_of[k] = -1;
      if (i < NP) {
        const int qi = point_q(i);
        const int p = i - qi * LP;
        int q;
        if (tile_query(a, c, qi, q)) {
          const int bq = c.b * a.Nq + q;
          lv_of[k] = point_l(p);
          raw[k] = point_load<T>(a, (size_t)bq * a.h + c.head, bq, p, lv_of[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MSDA_SORT_BATCH; ++k) {
      const int i = i0 + 256 * k;
      if (i >= NP) break;
      uint4 rec = make_uint4(0u, 0u, 0u, 0u);            // invalid point (no record; step 4 writes zero gradients)
      const int l = lv_of[k];
      float px, py;
      if (l >= 0 && point_finish(a, raw[k], l, px, py)) {
        const float x0f = floorf(px), y0f = floorf(py);
        const int x0 = (int)x0f, y0 = (int)y0f;
        rec = make_uint4(0x80000000u | ((unsigned)(y0 + 1) << 15) | (unsigned)(x0 + 1), __float_as_uint(px - x0f),
                         __float_as_uint(py - y0f), __float_as_uint(raw[k].aw));
        SortLvl& lv = s_lvl[l];
        if (x0 < lv.xmin) atomicMin(&lv.xmin, x0);
        if (x0 > lv.xmax) atomicMax(&lv.xmax, x0);
        if (y0 < lv.ymin) atomicMin(&lv.ymin, y0);
        if (y0 > lv.ymax) atomicMax(&lv.ymax, y0);
      }
      s_tmp[i] = rec;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int nb = 0;
    bool ok = true;
    for (int l = 0; l < a.L; ++l) {
      SortLvl& lv = s_lvl[l];
      lv.off = nb;
      if (lv.xmin > lv.xmax) { lv.bw = 0; continue; }   // no valid point on this level
      lv.bw = lv.xmax - lv.xmin + 1;
      const int bh = lv.ymax - lv.ymin + 1;
      ok = ok && lv.bw <= 256 && bh <= 256;
      nb += ok ? lv.bw * bh : 0;
      ok = ok && nb <= maxbins;
    }
    s_misc[0] = ok ? nb : -1;
  }
  __syncthreads();
  const int nbins = s_misc[0];
  if (nbins < 0) {                                       // CTA-uniform: windows too large to bin
    __syncthreads();
    bwd_direct_tile<T, (LPG * NV) / 4>(a, c, smem);
    return;
  }

  // ---- 2a. histogram; the shared atomic returns the point's rank inside its bin
  for (int i = tid; i < NP; i += 256) {
    const unsigned k = s_tmp[i].x;
    if (k & 0x80000000u) {
      const int qi = point_q(i);
      const SortLvl& lv = s_lvl[point_l(i - qi * LP)];
      const int xr = (int)(k & 0x7fffu) - 1 - lv.xmin, yr = (int)((k >> 15) & 0xffffu) - 1 - lv.ymin;
      s_rank[i] = (unsigned short)atomicAdd(&s_hist[lv.off + yr * lv.bw + xr], 1u);
    }
  }
  __syncthreads();
  // ---- 2b. exclusive scan of the histogram (in place)
  {
    const int per = (nbins + 255) >> 8;
    const int b0 = min(tid * per, nbins), b1 = min(b0 + per, nbins);
    unsigned sum = 0;
    for (int b = b0; b < b1; ++b) sum += s_hist[b];
    unsigned inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    if (lane == 31) s_misc[8 + warp] = (int)inc;
    __syncthreads();
    unsigned base = inc - sum;
    for (int w = 0; w < warp; ++w) base += (unsigned)s_misc[8 + w];
    for (int b = b0; b < b1; ++b) { const unsigned cnt = s_hist[b]; s_hist[b] = base; base += cnt; }
    if (tid == 255) s_hist[nbins] = base;                // total number of valid points
  }
  __syncthreads();
  // ---- 2c. scatter; everything scalar about a point is computed HERE, one thread per point, so that the walk of
  //          step 3 (where LPG lanes share a point and the groups of a warp diverge) stays lean
  for (int i = tid; i < NP; i += 256) {
    const uint4 r = s_tmp[i];
    if (r.x & 0x80000000u) {
      const int qi = point_q(i);
      const int l = point_l(i - qi * LP);
      const SortLvl& lv = s_lvl[l];
      const int x = (int)(r.x & 0x7fffu) - 1, y = (int)((r.x >> 15) & 0xffffu) - 1;
      const unsigned pos = s_hist[lv.off + (y - lv.ymin) * lv.bw + (x - lv.xmin)] + s_rank[i];
      const bool vx0 = x >= 0, vx1 = x + 1 < lv.W, vy0 = y >= 0, vy1 = y + 1 < lv.H;
      const unsigned m = (vx0 && vy0 ? 1u : 0u) | (vx1 && vy0 ? 2u : 0u) | (vx0 && vy1 ? 4u : 0u) | (vx1 && vy1 ? 8u : 0u);
      const float wx1 = __uint_as_float(r.y), wy1 = __uint_as_float(r.z), aw = __uint_as_float(r.w);
      const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
      s_key[pos] = make_uint2((unsigned)(lv.start_pix + y * lv.W + x), (unsigned)i | (m << 13) | ((unsigned)l << 17));
      s_w[pos] = make_float4(aw * (wx0 * wy0), aw * (wx1 * wy0), aw * (wx0 * wy1), aw * (wx1 * wy1));
    }
  }
  __syncthreads();

  // ---- 2d. run table.  A "run" = the records of one non-empty bin (same nw pixel).  The walk of step 3 keeps a run's
  //          value rows and accumulators in registers, and a warp walks GPW runs side by side: to keep its control flow
  //          UNIFORM (the first version let every group switch runs on its own: the divergent switch path ran in ~90 %
  //          of the iterations and was half of all instructions) the runs are counting-sorted by LENGTH, longest first,
  //          so that the runs a warp walks together have the same number of records.
  {
    unsigned* scratch = reinterpret_cast<unsigned*>(s_tmp);                    // the unsorted records are dead now
    for (int b = tid; b < nbins; b += 256) {
      const unsigned cnt = s_hist[b + 1] - s_hist[b];
      if (cnt) scratch[b] = atomicAdd(&s_h2[min(cnt, (unsigned)RUNCLS)], 1u);
    }
    __syncthreads();
    if (warp == 0) {                                     // class offsets, longest class first: lane <-> length lane + 1
      const unsigned v = s_h2[lane + 1];
      unsigned inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned u = __shfl_down_sync(0xffffffffu, inc, d);
        if (lane + d < 32) inc += u;
      }
      s_h2[lane + 1] = inc - v;                          // runs longer than this class
      if (lane == 0) s_misc[1] = (int)inc;               // number of runs
    }
    __syncthreads();
    for (int b = tid; b < nbins; b += 256) {
      const unsigned st = s_hist[b], cnt = s_hist[b + 1] - st;
      if (cnt) s_rt[s_h2[min(cnt, (unsigned)RUNCLS)] + scratch[b]] = st | (cnt << 16);
    }
    __syncthreads();
  }

  // ---- 3. segmented reduction: a group of LPG lanes (one channel row) takes one run, the GPW groups of a warp take
  //         GPW runs of (nearly) equal length; everything about a run is uniform across the warp
  {
    constexpr int GPW = 32 / LPG;
    const int nruns = s_misc[1];
    const int gw = lane / LPG, lig = lane % LPG;
    const size_t vec0 = (size_t)c.b * a.S * vec_per_pix + c.head * LPG + lig;
    const char* vbase = opaque(reinterpret_cast<const char*>(a.value) + vec0 * VB);
    const char* gvbase = a.grad_value ? opaque(reinterpret_cast<const char*>(a.grad_value) + vec0 * (NV * 4)) : nullptr;
    float* s_dot = reinterpret_cast<float*>(s_tmp);
    const unsigned dx = (unsigned)vec_per_pix;

    // rounds of 8 batches (one per warp), longest runs first; the warp order alternates from round to round so that no
    // warp collects the longest batch of every round
    for (int round = 0; round * 8 * GPW < nruns; ++round) {            // warp-uniform
      const int base = (round * 8 + ((round & 1) ? 7 - warp : warp)) * GPW;
      if (base >= nruns) continue;
      const int ri = base + gw;
      const unsigned rd = ri < nruns ? s_rt[ri] : 0u;
      const int start = (int)(rd & 0xffffu), cnt = (int)(rd >> 16);
      const int maxc = __reduce_max_sync(0xffffffffu, cnt);
      const uint2 k0 = cnt ? s_key[start] : make_uint2(0u, 0u);
      const unsigned m = cnt ? ((k0.y >> 13) & 15u) : 0u;              // taps of the run that lie inside the level
      const unsigned dy = s_lvl[(k0.y >> 17) & 7u].dy_vec;
      const unsigned o0 = k0.x * dx;                     // wraps for x0 = -1 / y0 = -1; such taps are never dereferenced
      const unsigned off[4] = {o0, o0 + dx, o0 + dy, o0 + dy + dx};
      float2 val[4][H2], acc[4][H2];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int j = 0; j < H2; ++j) val[t][j] = acc[t][j] = make_float2(0.f, 0.f);
        if ((m >> t) & 1u) load_row<T, NV>(addr_u32(vbase, off[t], VB), val[t]);
      }
_Pragma(MSDA_STR(unroll MSDA_SORT_UNROLL))
      for (int it = 0; it < maxc; ++it) {
        const bool live = it < cnt;
        unsigned ky = 0u;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) { ky = s_key[start + it].y; w = s_w[start + it]; }
        const int src = (int)(ky & 0x1fffu);
        float2 g[H2];
        if (go_smem) {
          const float4* gp = s_go + point_q(src) * row4 + lig * (NV / 4);
#pragma unroll
          for (int j = 0; j < NV / 4; ++j) {
            const float4 g4 = gp[j];
            g[2 * j] = make_float2(g4.x, g4.y); g[2 * j + 1] = make_float2(g4.z, g4.w);
          }
        } else {                                         // few points per query (L*P small): each row is used once or twice
          int q;
          tile_query(a, c, point_q(src), q);
          load_row<T, NV>(reinterpret_cast<const char*>(a.grad_out) +
                          ((((size_t)c.b * a.Nq + q) * a.h + c.head) * LPG + lig) * VB, g);
        }
        const float wt[4] = {w.x, w.y, w.z, w.w};
        float d[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {                    // packed FFMA2: 5 issue slots per tap and 4 channels instead of 8
          float2 s2 = __fmul2_rn(g[0], val[t][0]);
#pragma unroll
          for (int j = 1; j < H2; ++j) s2 = __ffma2_rn(g[j], val[t][j], s2);
          d[t] = s2.x + s2.y;
          const float2 w2 = make_float2(wt[t], wt[t]);
#pragma unroll
          for (int j = 0; j < H2; ++j) acc[t][j] = __ffma2_rn(w2, g[j], acc[t][j]);
        }
        // <grad_out, value row> of the four taps, summed over the group's lanes; step 4 turns them into gradients
        if constexpr (LPG == 2) {
          const bool hi = lig & 1;
          float ka = hi ? d[2] : d[0], kb = hi ? d[3] : d[1];
          ka += __shfl_xor_sync(0xffffffffu, hi ? d[0] : d[2], 1);
          kb += __shfl_xor_sync(0xffffffffu, hi ? d[1] : d[3], 1);
          if (live) { s_dot[src * 4 + 2 * hi] = ka; s_dot[src * 4 + 2 * hi + 1] = kb; }
        } else if constexpr (LPG == 1) {
          if (live) reinterpret_cast<float4*>(s_dot)[src] = make_float4(d[0], d[1], d[2], d[3]);
        } else {
          int which;
          const float sum = group_sum4<LPG>(d[0], d[1], d[2], d[3], lig, which);
          if (live && which >= 0) s_dot[src * 4 + which] = sum;
        }
      }
      if (gvbase) {                                      // one vector reduction per value row of the RUN, not per tap
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float* dst = reinterpret_cast<float*>(const_cast<char*>(addr_u32(gvbase, off[t], NV * 4)));
#pragma unroll
          for (int j = 0; j < NV / 4; ++j)
            red_add_v4_if((m >> t) & 1u, dst + 4 * j, acc[t][2 * j].x, acc[t][2 * j].y, acc[t][2 * j + 1].x, acc[t][2 * j + 1].y);
        }
      }
    }
  }
  __syncthreads();

  // ---- 4. one thread per point again: dot products -> d/d location, d/d weight; coalesced writes.  The point's
  //         fractions and weight are recomputed from its (L1-resident) location rather than kept in shared memory.
  for (int i0 = tid; i0 < NP; i0 += 256 * MSDA_SORT_BATCH) {
    PointRaw raw[MSDA_SORT_BATCH];
    int lv_of[MSDA_SORT_BATCH];
    size_t li_of[MSDA_SORT_BATCH];
#pragma unroll
    for (int k = 0; k < MSDA_SORT_BATCH; ++k) {
      const int i = i0 + 256 * k;
      lv_of[k] = -1;
      if (i < NP) {
        const int qi = point_q(i);
        const int p = i - qi * LP;
        int q;
        if (tile_query(a, c, qi, q)) {
          const int bq = c.b * a.Nq + q;
          const size_t item = (size_t)bq * a.h + c.head;
          lv_of[k] = point_l(p);
          li_of[k] = item * LP + p;
          raw[k] = point_load<T>(a, item, bq, p, lv_of[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MSDA_SORT_BATCH; ++k) {
      const int i = i0 + 256 * k;
      const int l = lv_of[k];
      if (l < 0) continue;
      float px, py, gw = 0.f, gx = 0.f, gy = 0.f;
      if (point_finish(a, raw[k], l, px, py)) {          // same predicate as step 1: the point has a record
        const float aw = raw[k].aw;
        const float4 d = reinterpret_cast<const float4*>(s_tmp)[i];
        const float wx1 = px - floorf(px), wy1 = py - floorf(py), wx0 = 1.f - wx1, wy0 = 1.f - wy1;
        gw = wy0 * (wx0 * d.x + wx1 * d.y) + wy1 * (wx0 * d.z + wx1 * d.w);
        gx = aw * (wy0 * (d.y - d.x) + wy1 * (d.w - d.z));
        gy = aw * (wx0 * (d.z - d.x) + wx1 * (d.w - d.y));
      }
      const size_t li = li_of[k];
      if (a.grad_loc) reinterpret_cast<float2*>(a.grad_loc)[li] = make_float2(gx * s_lvl[l].sx, gy * s_lvl[l].sy);
      if (a.grad_weight) store_scalar(reinterpret_cast<T*>(a.grad_weight) + li, gw);
    }
  }
}

// fp32 accumulation workspace -> bf16 grad_value; with `accumulate` the caller's earlier contents are added (in fp32)
__global__ void cast_f32_to_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4, int accumulate) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 f = src[i];
    if (accumulate) {
      const uint2 o = dst[i];
      f.x += bf16lo(o.x); f.y += bf16hi(o.x); f.z += bf16lo(o.y); f.w += bf16hi(o.y);
    }
    dst[i] = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
  }
}

// ------------------------------------------------------------------ host side
static bool pick_lpg_pow2(int nvec) { return nvec >= 1 && nvec <= 32 && (nvec & (nvec - 1)) == 0; }
static int pick_lpg(int Dh, int nv) {
  const int nvec = Dh / nv;
  for (int lpg = 32; lpg >= 1; lpg >>= 1)
    if (nvec % lpg == 0) return lpg;
  return 1;
}

// `tile_q` = queries per CTA (a power of two; split into a near-square 2-D tile when the queries form a grid)
static int fill_args(MsdaArgs& a, const int32_t* spatial_hw, const int32_t* level_start, int B, int S, int h,
                     int Dh, int Nq, int L, int P, int query_w, int dtype, int flags, int nv, int tile_q = 64) {
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
  a.p_shift = a.lp_shift = -1;
  for (int s = 0; s < 16; ++s) {
    if ((1 << s) == P) a.p_shift = s;
    if ((1 << s) == L * P) a.lp_shift = s;
  }
  for (int l = 0; l < L; ++l) {
    a.lvl_h[l] = spatial_hw[2 * l]; a.lvl_w[l] = spatial_hw[2 * l + 1]; a.lvl_start[l] = level_start[l];
    BEVR_REQUIRE(a.lvl_h[l] > 0 && a.lvl_w[l] > 0 && a.lvl_start[l] >= 0 &&
                 a.lvl_start[l] + a.lvl_h[l] * a.lvl_w[l] <= S, "msda: level %d (%dx%d @%d) outside S=%d",
                 l, a.lvl_h[l], a.lvl_w[l], a.lvl_start[l], S);
  }
  int tw = 1, th = tile_q;
  while (tw < th) { tw <<= 1; th >>= 1; }              // 64 -> 8x8, 128 -> 16x8, 1024 -> 32x32
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

// ---- tile-sorted backward: tile size, histogram capacity and launch
#ifndef MSDA_SORT_NP
#define MSDA_SORT_NP 1024          // sampling points per CTA (tile queries = MSDA_SORT_NP / (L*P), a power of two)
#endif
#ifndef MSDA_SORT_MAXBINS
#define MSDA_SORT_MAXBINS 3072     // histogram bins per CTA (sum over levels of the nw-pixel bounding boxes)
#endif

static int sorted_tile_queries(int L, int P) {
  int tq = 1;
  while (tq * 2 * L * P <= MSDA_SORT_NP && tq < 1024) tq <<= 1;
  return tq;
}
static bool sorted_applicable(int Dh, int L, int P) {
  const int tq = sorted_tile_queries(L, P);
  return Dh % 4 == 0 && Dh <= 128 && pick_lpg_pow2(Dh / 4) && tq >= 8 && tq * L * P <= 8192;
}

#ifndef MSDA_SORT_GO_BYTES
#define MSDA_SORT_GO_BYTES (16 * 1024)   // stage the tile's grad_out rows in shared memory up to this size
#endif
#ifndef MSDA_SORT_NV
#define MSDA_SORT_NV 4             // channels per lane of the sorted backward (8 halves the per-point overhead but needs 118 registers:
                                   // 3.47 ms against 3.01 ms at BASELINE config 2, profiles/r2_sweep_msda_sorted.txt)
#endif

template <typename T, int NV, int LPG>
static int launch_sorted(MsdaArgs& a, cudaStream_t st) {
  const int LP = a.L * a.P, tq = a.th << a.tw_shift, NP = tq * LP, row4 = a.Dh / 4;
  const int maxbins = MSDA_SORT_MAXBINS;
  a.sort_go_smem = (size_t)tq * row4 * 16 <= MSDA_SORT_GO_BYTES;
  size_t bytes = (size_t)NP * 44 + (a.sort_go_smem ? (size_t)tq * row4 * 16 : 0) + (size_t)(maxbins + 4) * 4 + (size_t)NP * 2;
  bytes = (bytes + 15) & ~(size_t)15;
  a.sort_lvl_off = (int)bytes;
  bytes += BEVR_MAX_LEVELS * sizeof(SortLvl) + (16 + RUNCLS + 2) * sizeof(int);
  const size_t direct = (size_t)8 * 2 * (32 / ((LPG * NV) / 4)) * (LP + 1) * 16;     // staging of the per-CTA fallback
  BEVR_REQUIRE(direct <= (size_t)a.sort_lvl_off, "msda_bwd: L*P=%d too large for the sorted backward", LP);
  BEVR_REQUIRE(bytes <= 200 * 1024, "msda_bwd: sorted backward needs %zu B of shared memory", bytes);
  const long long grid = (long long)a.B * a.h * a.tiles_x * a.tiles_y;
  BEVR_REQUIRE(grid < (1ll << 31), "msda: grid too large");
  const bool fast = a.sort_go_smem && a.lp_shift >= 0 && a.p_shift >= 0;
  auto launch1 = [&](auto kern) {
    if (bytes > 48 * 1024)
      if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                              "msda_bwd (sorted) shared memory"))
        return rc;
    kern<<<(unsigned)grid, 256, bytes, st>>>(a, maxbins);
    return check_launch("msda_bwd (sorted)");
  };
  return fast ? launch1(msda_bwd_sorted_kernel<T, NV, LPG, true>) : launch1(msda_bwd_sorted_kernel<T, NV, LPG, false>);
}

template <typename T>
static int dispatch_bwd_sorted(MsdaArgs& a, cudaStream_t st) {
  // lanes per point: Dh / NV with NV = 8 channels per lane when that leaves >= 2 lanes, else 4
  const bool nv8 = MSDA_SORT_NV == 8 && a.Dh % 8 == 0 && a.Dh >= 16;
  if (nv8) {
    switch (a.Dh / 8) {
      case 16: return launch_sorted<T, 8, 16>(a, st);
      case 8: return launch_sorted<T, 8, 8>(a, st);
      case 4: return launch_sorted<T, 8, 4>(a, st);
      default: return launch_sorted<T, 8, 2>(a, st);
    }
  }
  switch (a.Dh / 4) {
    case 32: return launch_sorted<T, 4, 32>(a, st);
    case 16: return launch_sorted<T, 4, 16>(a, st);
    case 8: return launch_sorted<T, 4, 8>(a, st);
    case 4: return launch_sorted<T, 4, 4>(a, st);
    case 2: return launch_sorted<T, 4, 2>(a, st);
    default: return launch_sorted<T, 4, 1>(a, st);
  }
}

}  // namespace bevr

using namespace bevr;

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
  const bool direct = (flags & BEVR_BWD_DIRECT) != 0 || L <= 0 || P <= 0 || !sorted_applicable(Dh, L, P);
  if (int rc = fill_args(a, spatial_hw, level_start, B, S, h, Dh, Nq, L, P, query_w, dtype, flags, 4,
                         direct ? 64 : sorted_tile_queries(L, P)))
    return rc;
  if (B == 0) return 0;
  if (Nq == 0) {                                   // no queries: grad_value is all zeros, nothing else to write
    if (grad_value && !(flags & BEVR_ACCUMULATE))
      return check_cuda(cudaMemsetAsync(grad_value, 0, (size_t)B * S * h * Dh *
                                        ((dtype == BEVR_F32 || (flags & BEVR_GRAD_VALUE_F32)) ? 4 : 2),
                                        (cudaStream_t)stream), "msda_bwd zero-fill");
    return 0;
  }
  BEVR_REQUIRE(value && sampling_loc && grad_out, "msda_bwd: value / sampling_loc / grad_out must not be NULL");
  BEVR_REQUIRE(!(attn_weight == nullptr && grad_weight != nullptr), "msda_bwd: grad_weight without attn_weight");
  const bool gv_f32 = dtype == BEVR_F32 || (flags & BEVR_GRAD_VALUE_F32);       // grad_value buffer holds fp32
  BEVR_REQUIRE(!(!gv_f32 && grad_value && !workspace), "msda_bwd: bf16 grad_value needs the fp32 workspace");
  BEVR_REQUIRE(!(a.nchunk > 1 && dtype == BEVR_BF16 && grad_weight),
               "msda_bwd: bf16 grad_weight with Dh > 128 is not supported");
  cudaStream_t st = (cudaStream_t)stream;
  a.value = value; a.loc = sampling_loc; a.ref = ref_points; a.weight = attn_weight; a.grad_out = grad_out;
  a.grad_loc = grad_loc; a.grad_weight = grad_weight;
  const size_t n = (size_t)B * S * h * Dh;
  a.grad_value = nullptr;
  if (grad_value) {
    a.grad_value = gv_f32 ? (float*)grad_value : (float*)workspace;
    // bf16 output: the fp32 workspace always starts at zero; BEVR_ACCUMULATE is honoured by the cast kernel below
    if (!(flags & BEVR_ACCUMULATE) || !gv_f32)
      if (int rc = check_cuda(cudaMemsetAsync(a.grad_value, 0, n * sizeof(float), st), "msda_bwd zero-fill")) return rc;
  }
  if (a.nchunk > 1) {  // chunks accumulate with atomics -> outputs must start at zero
    const size_t nl = (size_t)B * Nq * h * L * P;
    if (grad_loc)
      if (int rc = check_cuda(cudaMemsetAsync(grad_loc, 0, nl * 2 * sizeof(float), st), "msda_bwd zero-fill")) return rc;
    if (grad_weight)
      if (int rc = check_cuda(cudaMemsetAsync(grad_weight, 0, nl * sizeof(float), st), "msda_bwd zero-fill")) return rc;
  }
  int rc;
  if (direct) rc = dtype == BEVR_F32 ? dispatch_bwd<float>(a, st) : dispatch_bwd<__nv_bfloat16>(a, st);
  else rc = dtype == BEVR_F32 ? dispatch_bwd_sorted<float>(a, st) : dispatch_bwd_sorted<__nv_bfloat16>(a, st);
  if (rc) return rc;
  if (grad_value && !gv_f32) {
    cast_f32_to_bf16_kernel<<<sm_count() * 8, 256, 0, st>>>((const float4*)workspace, (uint2*)grad_value, n / 4,
                                                               (flags & BEVR_ACCUMULATE) ? 1 : 0);
    return check_launch("msda_bwd cast");
  }
  return 0;
}
