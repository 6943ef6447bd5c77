// K3: fused attention with sampled relative-position bias, forward and backward (sm_100a).
//
// Replaces, per (batch, head):   attn = (q^T k) * scale                      TSA_deform_attn.py:245-250
//                                bias = grid_sample(rpe_table, (qgrid_i - pos_j)/2)        :264-305
//                                out  = softmax_j(attn + bias) v^T                          :308-330
// (same code at SCA_deform_attn.py:331-411).  The reference materialises attn, bias and the
// displacement tensor at (Nq, Ns); here nothing of size Nq*Ns ever exists: K/V (resp. Q/dO) tiles are
// staged in shared memory, the bias is a bilinear lookup computed on the fly from the query index and
// the staged key position, the softmax is the online (flash) form, and the backward recomputes.
//
// Two kernel families (DESIGN.md section "K3"), picked per call by problem size (use_mma):
//   * mma.sync 3xTF32 -- rpe_attn_fwd_mma_kernel, rpe_attn_bwd_dq_mma_kernel, rpe_attn_bwd_dkv_mma_kernel:
//     one warp owns 16 query rows (or 16 keys); the contractions run on the legacy warp-level tensor-core
//     instruction with a hi/lo operand split that keeps fp32-grade accuracy (a single TF32 product would
//     not hold the 1e-4 budget); bias lookup and softmax stay on the CUDA cores.  Operands are pre-split
//     once per call (rpe_presplit_kernel) and staged with cp.async.
//   * FFMA2 -- rpe_attn_fwd_kernel, rpe_attn_bwd_dq_kernel, rpe_attn_bwd_dkv_kernel: one thread owns one
//     query row or key column, tiles broadcast from shared memory; serves grids below ~40x40.
// Shared by both: the zero-padded table copy (rpe_pad_kernel), the hoisted bias coordinates, the split of the
// streamed dimension over blockIdx.z (rpe_combine_kernel merges the forward's key ranges) and the privatised
// table-gradient copies (rpe_unpad_add_kernel).
#include <math.h>
#include <stdlib.h>

#include "bevr_common.cuh"

namespace bevr {

struct RpeArgs {
  const float *q, *k, *v, *rpe_pad, *pos, *out, *lse, *grad_out, *delta;
  float *o_out, *o_lse, *grad_q, *grad_k, *grad_v, *grad_rpe_pad, *grad_pos, *o_delta;
  int BH, h, hpg, Hq, Wq, Nq, Ns, Rh, Rw, Hp, Wp;
  int gcopies;   // privatised copies of the padded table gradient (CTAs spread their atomics over them)
  // work splitting (blockIdx.z): the reference-native shapes are small (B*h*Nq = 50 k query rows = 11 warps per SM),
  // so every kernel also cuts its STREAMED dimension into `split` ranges of `tiles_per_split` tiles.
  int split, tiles_per_split;
  float *part_o, *part_ml;   // forward partials: (split, BH, DH, Nq) unnormalised outputs, (split, BH, 2, Nq) max / sum
  // mma kernels: operands pre-split into tf32 hi / lo parts, row-major (BH, n rounded up to 64, DH), zero rows beyond n
  const float *kh, *kl, *vh, *vl, *qh, *ql, *gh, *gl;
  int NsP, NqP;
  float scale;
  // attention dropout (TSA_deform_attn.py:313-323, SCA_deform_attn.py:402-410): keep mask from a counter-based hash
  unsigned drop_thresh;      // 24-bit threshold: an element is dropped when hash >> 8 < drop_thresh (0: no dropout)
  unsigned seed_lo, seed_hi;
  float drop_scale;          // 1 / (1 - p)
};

// DH (channels per head) is a template parameter of every kernel: 32 (the reference, encoder.py:23-39) and 16.
// Shared-memory row stride is DH + 4 floats: keeps float4 alignment and spreads the banks.
#define PAD (DH + 4)

// normalised query-grid coordinate, bit-identical to torch's arange -> /(n-1) -> *2 -> -1
__device__ __forceinline__ float grid_coord(int i, int n) {
  return __fsub_rn(__fmul_rn(__fdiv_rn((float)i, (float)(n - 1)), 2.f), 1.f);
}

struct Bilin {
  int off;                    // offset of the nw tap inside the PADDED table (relative to padded (2,2))
  float wx0, wx1, wy0, wy1;   // bilinear weights
};

// The kernels read a zero-padded copy of rpe_table: two zero rows/columns on every side
// (Hp = Rh + 4, Wp = Rw + 4).  With the sample coordinate clamped to [-2, R], all four taps of every
// sample are loadable and out-of-table taps read the zero border -- exactly grid_sample's
// padding_mode="zeros" -- so the inner loops carry no validity masks, no per-tap predicates and no
// clamped indices (the bias lookup was 170 of the 237 instructions per (query, key) pair before).
//
// Coordinates: grid_sample(align_corners=True) on the displacement (query - key)/2 is, per axis,
//   i = (((q - p) * 0.5 + 1) / 2) * (R - 1)
// and it is evaluated with the REFERENCE's own roundings: fl(q - p); fl(. * 0.5 + 1) (the product is exact, so one FMA);
// fl(. * (R-1)/2) (the division by two is exact).  The bias itself is continuous in the coordinate, but d bias / d pos is
// piecewise constant: it jumps where the coordinate crosses a table cell.  Round 1 used the cheaper
// fma(q, c, c2) - p * c (one subtraction per pair, ~1 ulp(R) away from the reference): at the BASELINE sizes one key in
// ~300 then had one query column on the other side of such a jump and its grad_pos was off by up to 30 % (5e-2 of
// max|grad_pos| at N = 50, D = 5; invisible at the N <= 16 golden sizes).  Bit-identical coordinates put every jump
// where the reference has it.  Cost: 2 extra FP32 instructions per axis per pair.
__device__ __forceinline__ float rpe_query_half(float q, int) { return q; }
__device__ __forceinline__ float rpe_key_half(float p, int) { return p; }
__device__ __forceinline__ float rpe_coord(float q, float p, float half_rm1) {
  return __fmul_rn(__fmaf_rn(__fsub_rn(q, p), 0.5f, 1.f), half_rm1);
}

__device__ __forceinline__ Bilin rpe_taps(float ay, float ax, float by, float bx, int Rh, int Rw, int Wp) {
  // clamp to [-2, R]: a clamped sample has ALL four taps in the zero border, so its value and both of its
  // gradients vanish exactly as for an out-of-range grid_sample location (NaN maps to -2 as well)
  const float ix = fminf(fmaxf(rpe_coord(ax, bx, 0.5f * (float)(Rw - 1)), -2.f), (float)Rw);
  const float iy = fminf(fmaxf(rpe_coord(ay, by, 0.5f * (float)(Rh - 1)), -2.f), (float)Rh);
  Bilin b;
  const float x0f = floorf(ix), y0f = floorf(iy);
  b.wx1 = ix - x0f; b.wx0 = 1.f - b.wx1;
  b.wy1 = iy - y0f; b.wy0 = 1.f - b.wy1;
  b.off = (int)y0f * Wp + (int)x0f;
  return b;
}

// bias value; tapv receives the four tap values (zero where the tap is outside the table)
__device__ __forceinline__ float rpe_bias(const float* tbl, const Bilin& b, int Wp, float* tapv) {
  const float* r0 = tbl + b.off;
  const float* r1 = r0 + Wp;
  tapv[0] = __ldg(r0); tapv[1] = __ldg(r0 + 1);
  tapv[2] = __ldg(r1); tapv[3] = __ldg(r1 + 1);
  return b.wy0 * (b.wx0 * tapv[0] + b.wx1 * tapv[1]) + b.wy1 * (b.wx0 * tapv[2] + b.wx1 * tapv[3]);
}

__device__ __forceinline__ const float* opaque_f(const float* p) {
  asm volatile("" : "+l"(p));      // keep the per-thread table base in a register pair (1-instruction addressing)
  return p;
}

// Attention dropout.  The reference applies nn.Dropout to the softmax matrix, i.e. draws from torch's Philox stream over a
// (B*h, Nq, Ns) tensor that is never materialised here; the kernels draw the keep mask from a counter-based hash of
// (seed, b*h, query, key) instead -- the same distribution (Bernoulli(1 - p), kept elements scaled by 1 / (1 - p)), a
// different stream.  The three kernels of a call regenerate the same mask; oracle/deform_attn.py restates the hash.
__device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ unsigned drop_key(const unsigned seed_lo, const unsigned seed_hi, int bh) {
  return mix32(seed_lo + (unsigned)bh * 0x9e3779b9u) ^ seed_hi;
}
// multiplier of probability (query i, key j): 0 or 1 / (1 - p); row_base = i * Ns (mod 2^32)
template <bool DROP>
__device__ __forceinline__ float drop_mul(unsigned thresh, float scale, unsigned key, unsigned row_base, unsigned j) {
  if constexpr (!DROP) return 1.f;
  return (mix32((row_base + j) ^ key) >> 8) >= thresh ? scale : 0.f;
}

// padded (h, Rh+4, Rw+4) copy of the table / accumulation of the padded gradient back into (h, Rh, Rw)
__global__ void rpe_pad_kernel(const float* __restrict__ rpe, float* __restrict__ pad, int h, int Rh, int Rw) {
  const int Hp = Rh + 4, Wp = Rw + 4;
  const int n = h * Hp * Wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % Wp - 2, y = (i / Wp) % Hp - 2, hh = i / (Wp * Hp);
    pad[i] = (x >= 0 && x < Rw && y >= 0 && y < Rh) ? rpe[((size_t)hh * Rh + y) * Rw + x] : 0.f;
  }
}
__global__ void rpe_unpad_add_kernel(const float* __restrict__ gpad, float* __restrict__ grad, int h, int Rh, int Rw,
                                     int copies) {
  const int Hp = Rh + 4, Wp = Rw + 4;
  const int n = h * Rh * Rw;
  const size_t copy_stride = (size_t)h * Hp * Wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % Rw, y = (i / Rw) % Rh, hh = i / (Rw * Rh);
    const float* src = gpad + ((size_t)hh * Hp + y + 2) * Wp + x + 2;
    float s = 0.f;
    for (int c = 0; c < copies; ++c) s += src[c * copy_stride];
    grad[i] += s;
  }
}

// stage rows [j0, j0+T) of a channel-major (DH, n) matrix as smem[T][PAD] (zero beyond n)
template <int T, int DH>
__device__ __forceinline__ void stage_transposed(float* __restrict__ dst, const float* __restrict__ src, int n,
                                                 int j0, float mul) {
  for (int idx = threadIdx.x; idx < T * DH; idx += blockDim.x) {
    const int c = idx / T, jj = idx - c * T;
    const int j = j0 + jj;
    dst[jj * PAD + c] = j < n ? __ldg(src + (size_t)c * n + j) * mul : 0.f;
  }
}

// 32-channel dot product / axpy against a shared-memory row with packed FFMA2 (two fp32 FMAs per issue
// slot on sm_100: the kernels of this file are FMA-issue bound).  Register operands are float2 pairs.
template <int DH>
__device__ __forceinline__ float dot32(const float2* reg, const float* __restrict__ srow) {
  float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);     // two chains: halves the FFMA2 latency path
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 kk = *reinterpret_cast<const float4*>(srow + 4 * c);
    s0 = __ffma2_rn(reg[2 * c], make_float2(kk.x, kk.y), s0);
    s1 = __ffma2_rn(reg[2 * c + 1], make_float2(kk.z, kk.w), s1);
  }
  return (s0.x + s1.x) + (s0.y + s1.y);
}
template <int DH>
__device__ __forceinline__ void axpy32(float2* acc, float a, const float* __restrict__ srow) {
  const float2 a2 = make_float2(a, a);
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 vv = *reinterpret_cast<const float4*>(srow + 4 * c);
    acc[2 * c] = __ffma2_rn(a2, make_float2(vv.x, vv.y), acc[2 * c]);
    acc[2 * c + 1] = __ffma2_rn(a2, make_float2(vv.z, vv.w), acc[2 * c + 1]);
  }
}
__device__ __forceinline__ float& lane_elem(float2* r, int c) { return (c & 1) ? r[c >> 1].y : r[c >> 1].x; }

// ------------------------------------------------------------------ forward (thread = query row)
#ifndef K3_TQ
#define K3_TQ 128
#endif
#ifndef K3_CH
#define K3_CH 8
#endif
#ifndef K3_FWD_MINB
#define K3_FWD_MINB 3
#endif
#ifndef K3_DQ_MINB
#define K3_DQ_MINB 3
#endif
constexpr int TQ = K3_TQ, TK = 64, CH = K3_CH;

// K3_R query rows per thread (rows i and i + TQ/R of the CTA's tile): with R = 2 every broadcast LDS.128 of a K / V row
// feeds two rows, which halves the shared-memory wavefronts per (query, key) pair -- the forward's binding pipe.
#ifndef K3_R
#define K3_R 1
#endif
#ifndef K3_R2_MINB
#define K3_R2_MINB 4
#endif
constexpr int RQ = K3_R;
constexpr int FWD_MINB = RQ == 1 ? K3_FWD_MINB : K3_R2_MINB;

template <int DH, int R>
__device__ __forceinline__ void dotR(const float2 (*reg)[DH / 2], const float* __restrict__ srow, float* out) {
  float2 s0[R], s1[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { s0[r] = make_float2(0.f, 0.f); s1[r] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 kk = *reinterpret_cast<const float4*>(srow + 4 * c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s0[r] = __ffma2_rn(reg[r][2 * c], make_float2(kk.x, kk.y), s0[r]);
      s1[r] = __ffma2_rn(reg[r][2 * c + 1], make_float2(kk.z, kk.w), s1[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = (s0[r].x + s1[r].x) + (s0[r].y + s1[r].y);
}
template <int DH, int R>
__device__ __forceinline__ void axpyR(float2 (*acc)[DH / 2], const float* a, const float* __restrict__ srow) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 vv = *reinterpret_cast<const float4*>(srow + 4 * c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 a2 = make_float2(a[r], a[r]);
      acc[r][2 * c] = __ffma2_rn(a2, make_float2(vv.x, vv.y), acc[r][2 * c]);
      acc[r][2 * c + 1] = __ffma2_rn(a2, make_float2(vv.z, vv.w), acc[r][2 * c + 1]);
    }
  }
}

template <int DH, bool DROP>
__global__ void __launch_bounds__(TQ / RQ, FWD_MINB) rpe_attn_fwd_kernel(const __grid_constant__ RpeArgs a) {
  constexpr int R = RQ, NT = TQ / R;
  __shared__ __align__(16) float sK[TK * PAD];
  __shared__ __align__(16) float sV[TK * PAD];
  __shared__ float2 sPos[TK];
  const int bh = blockIdx.y;
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  const float* kb = a.k + (size_t)bh * DH * a.Ns;
  const float* vb = a.v + (size_t)bh * DH * a.Ns;
  const float2* pb = reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns;

  int iq[R];
  bool live[R];
  float2 qr[R][DH / 2], o[R][DH / 2];
  float qy[R], qx[R], m[R], l[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    iq[r] = blockIdx.x * TQ + r * NT + threadIdx.x;
    live[r] = iq[r] < a.Nq;
#pragma unroll
    for (int c = 0; c < DH; ++c) {
      lane_elem(qr[r], c) = live[r] ? __ldg(a.q + ((size_t)bh * DH + c) * a.Nq + iq[r]) * a.scale : 0.f;
      lane_elem(o[r], c) = 0.f;
    }
    const int iy = iq[r] / a.Wq, ixq = iq[r] - iy * a.Wq;
    qy[r] = rpe_query_half(grid_coord(iy, a.Hq), a.Rh); qx[r] = rpe_query_half(grid_coord(ixq, a.Wq), a.Rw);
    m[r] = -INFINITY; l[r] = 0.f;
  }
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh);
  const int jbeg = blockIdx.z * a.tiles_per_split * TK, jend = min(a.Ns, jbeg + a.tiles_per_split * TK);

  for (int j0 = jbeg; j0 < jend; j0 += TK) {
    __syncthreads();
    stage_transposed<TK, DH>(sK, kb, a.Ns, j0, 1.f);
    stage_transposed<TK, DH>(sV, vb, a.Ns, j0, 1.f);
    for (int jj = threadIdx.x; jj < TK; jj += blockDim.x) {
      const float2 p = (j0 + jj < a.Ns) ? __ldg(pb + j0 + jj) : make_float2(0.f, 0.f);
      sPos[jj] = make_float2(rpe_key_half(p.x, a.Rh), rpe_key_half(p.y, a.Rw));
    }
    __syncthreads();
#pragma unroll 1
    for (int jc = 0; jc < TK; jc += CH) {
      if (j0 + jc >= a.Ns) break;
      float s[R][CH];
      float mx[R];
#pragma unroll
      for (int r = 0; r < R; ++r) mx[r] = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < CH; ++jj) {
        const float2 p = sPos[jc + jj];
        float d[R];
        dotR<DH, R>(qr, sK + (jc + jj) * PAD, d);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const Bilin b = rpe_taps(qy[r], qx[r], p.x, p.y, a.Rh, a.Rw, a.Wp);
          float tv[4];
          const float sc = d[r] + rpe_bias(table, b, a.Wp, tv);
          s[r][jj] = (j0 + jc + jj < a.Ns) ? sc : -INFINITY;
          mx[r] = fmaxf(mx[r], s[r][jj]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float mn = fmaxf(m[r], mx[r]);
        const float corr = __expf(m[r] - mn);
        l[r] *= corr;
#pragma unroll
        for (int c = 0; c < DH / 2; ++c) { o[r][c].x *= corr; o[r][c].y *= corr; }
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
          s[r][jj] = __expf(s[r][jj] - mn);
          l[r] += s[r][jj];                                   // the softmax denominator is taken before the dropout
          s[r][jj] *= drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, (unsigned)iq[r] * (unsigned)a.Ns, (unsigned)(j0 + jc + jj));
        }
        m[r] = mn;
      }
#pragma unroll
      for (int jj = 0; jj < CH; ++jj) {
        float pr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pr[r] = s[r][jj];
        axpyR<DH, R>(o, pr, sV + (jc + jj) * PAD);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!live[r]) continue;
    const int i = iq[r];
    if (a.split == 1) {
      const float inv = 1.f / l[r];
#pragma unroll
      for (int c = 0; c < DH; ++c) a.o_out[((size_t)bh * DH + c) * a.Nq + i] = lane_elem(o[r], c) * inv;
      a.o_lse[(size_t)bh * a.Nq + i] = m[r] + logf(l[r]);
    } else {      // partial (unnormalised) result of this key range; rpe_combine_kernel merges the ranges
      const size_t pz = (size_t)blockIdx.z * a.BH + bh;
#pragma unroll
      for (int c = 0; c < DH; ++c) a.part_o[(pz * DH + c) * a.Nq + i] = lane_elem(o[r], c);
      a.part_ml[(pz * 2 + 0) * a.Nq + i] = m[r];
      a.part_ml[(pz * 2 + 1) * a.Nq + i] = l[r];
    }
  }
}

// ------------------------------------------------------------------ forward, contractions on mma.sync (3xTF32)
// The two contractions of the forward (S = Q^T K and O += P V) as legacy warp-level mma.sync.m16n8k8 TF32 with the
// big/small operand split (x = hi + lo, hi = tf32(x); hi*hi + hi*lo + lo*hi accumulate in fp32): the dropped lo*lo term
// and the truncation of lo are ~2^-22 relative, i.e. fp32-grade, which a single TF32 product (2^-11) is not.  SURVEY.md
// appendix E allows CUDA-core FMA or legacy mma.sync here; tcgen05 stays reserved for the correlation.  The bias lookup,
// the online softmax and the rescaling stay on the CUDA cores and now dominate the instruction count (~1.6 warp
// instructions per (query, key) pair against 3.0 for the FFMA2 kernel, whose fma pipe alone needs 0.73 SM cycles per pair).
//
// One warp owns 16 query rows (accumulator layout of the m16n8 tile: lane = 4*g + t holds rows g and g+8, columns 2t,
// 2t+1).  K, V tiles are staged pre-split (hi / lo) with a 36-float row pitch, which makes both B-operand access
// patterns bank-conflict free.  For O += P V the eight keys of a k-step are fed in the order (0,2,4,6,1,3,5,7): operand
// column t is key 2t and column t+4 is key 2t+1, so the probabilities are consumed in the registers they were computed in.
#ifndef K3_FWD_MMA
#define K3_FWD_MMA 1
#endif
constexpr int TQM = 64;                      // query rows per CTA (4 warps x 16)
#ifndef K3_MMA_CH
#define K3_MMA_CH 32
#endif
#ifndef K3_FWD_Q_SMEM
#define K3_FWD_Q_SMEM 0                      // forward: Q fragments in shared memory instead of 32 registers (5 CTAs per SM)
#endif
#ifndef K3_MMA_MINB
#define K3_MMA_MINB (K3_FWD_Q_SMEM ? 5 : 4)
#endif
constexpr int CHM = K3_MMA_CH;               // keys per online-softmax step (4 n-tiles)

__device__ __forceinline__ uint32_t tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// rows [j0, j0+T) of a channel-major (DH, n) matrix as hi[T][PAD] / lo[T][PAD] (zero beyond n)
template <int T, int DH>
__device__ __forceinline__ void stage_split(float* __restrict__ hi, float* __restrict__ lo, const float* __restrict__ src,
                                            int n, int j0, float mul = 1.f) {
  for (int idx = threadIdx.x; idx < T * DH; idx += blockDim.x) {
    const int c = idx / T, jj = idx - c * T;
    const int j = j0 + jj;
    const float x = j < n ? __ldg(src + (size_t)c * n + j) * mul : 0.f;
    const float h = __uint_as_float(tf32_hi(x));
    hi[jj * PAD + c] = h;
    lo[jj * PAD + c] = x - h;
  }
}

// One pass per operand instead of one per (CTA, tile): x (BH, DH, n) channel-major -> hi / lo (BH, nP, DH) row-major.
// Every CTA used to transpose, round and split its K / V (Q / dO) tiles itself -- a fifth of the mma kernels' instructions.
template <int DH>
__global__ void __launch_bounds__(256) rpe_presplit_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                           float* __restrict__ lo, int n, int nP, float mul) {
  const int bh = blockIdx.y, j = blockIdx.x * 64 + (threadIdx.x & 63), c0 = (threadIdx.x >> 6) * (DH / 4);
  if (j >= nP) return;
  float h[DH / 4], l[DH / 4];
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float v = j < n ? __ldg(x + ((size_t)bh * DH + c0 + c) * n + j) * mul : 0.f;
    h[c] = __uint_as_float(tf32_hi(v));
    l[c] = v - h[c];
  }
  float4* dh = reinterpret_cast<float4*>(hi + ((size_t)bh * nP + j) * DH + c0);
  float4* dl = reinterpret_cast<float4*>(lo + ((size_t)bh * nP + j) * DH + c0);
#pragma unroll
  for (int c = 0; c < DH / 16; ++c) {
    dh[c] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
    dl[c] = make_float4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
  }
}
// 64 rows of a pre-split operand -> shared memory [64][PAD] with 16-byte async copies (LDGSTS)
template <int DH>
__device__ __forceinline__ void stage_async(float* __restrict__ dst, const float* __restrict__ src_tile) {
  constexpr int CPR = DH / 4;                // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < TK * CPR; idx += blockDim.x) {
    const int row = idx / CPR, ch = idx - row * CPR;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + row * PAD + ch * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_tile + row * DH + ch * 4) : "memory");
  }
}
__device__ __forceinline__ void stage_async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

template <int DH, bool DROP>
__global__ void __launch_bounds__(128, K3_MMA_MINB) rpe_attn_fwd_mma_kernel(const __grid_constant__ RpeArgs a) {
  constexpr int KS = DH / 8;                 // k-steps of S = Q^T K (channels) = n-tiles of O (channels)
  constexpr int NT = CHM / 8;                // key n-tiles per softmax step = k-steps of O += P V
  __shared__ __align__(16) float sKh[TK * PAD], sKl[TK * PAD], sVh[TK * PAD], sVl[TK * PAD];
  __shared__ __align__(16) float2 sPos[TK];
#if K3_FWD_Q_SMEM
  extern __shared__ __align__(16) float4 s_qf[];           // the warp's Q fragments (A operand of S), split when used
  float4 (*sQf)[KS][32] = reinterpret_cast<float4 (*)[KS][32]>(s_qf);
#endif
  const int bh = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  const float2* pb = reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns;

  const int i0 = blockIdx.x * TQM + warp * 16;
  int iq[2] = {i0 + g, i0 + g + 8};
  float ay[2], ax[2], m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#if !K3_FWD_Q_SMEM
  uint32_t qh[KS][4], ql[KS][4];
#endif
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int iy = iq[r] / a.Wq, ixq = iq[r] - iy * a.Wq;
    ay[r] = rpe_query_half(grid_coord(iy, a.Hq), a.Rh);
    ax[r] = rpe_query_half(grid_coord(ixq, a.Wq), a.Rw);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {            // a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
      const int r = e & 1, c = ks * 8 + t + ((e >> 1) << 2);
      const float x = iq[r] < a.Nq ? __ldg(a.q + ((size_t)bh * DH + c) * a.Nq + iq[r]) * a.scale : 0.f;
#if K3_FWD_Q_SMEM
      reinterpret_cast<float*>(&sQf[warp][ks][lane])[e] = x;
#else
      qh[ks][e] = tf32_hi(x);
      ql[ks][e] = __float_as_uint(x - __uint_as_float(qh[ks][e]));
#endif
    }
  float o[KS][4];
#pragma unroll
  for (int nv = 0; nv < KS; ++nv)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nv][e] = 0.f;
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh);
  const unsigned rowb[2] = {(unsigned)iq[0] * (unsigned)a.Ns, (unsigned)iq[1] * (unsigned)a.Ns};

  const int jbeg = blockIdx.z * a.tiles_per_split * TK, jend = min(a.Ns, jbeg + a.tiles_per_split * TK);
  for (int j0 = jbeg; j0 < jend; j0 += TK) {
    __syncthreads();
    {
      const size_t tile = ((size_t)bh * a.NsP + j0) * DH;
      stage_async<DH>(sKh, a.kh + tile); stage_async<DH>(sKl, a.kl + tile);
      stage_async<DH>(sVh, a.vh + tile); stage_async<DH>(sVl, a.vl + tile);
    }
    for (int jj = threadIdx.x; jj < TK; jj += blockDim.x) {
      const float2 p = (j0 + jj < a.Ns) ? __ldg(pb + j0 + jj) : make_float2(0.f, 0.f);
      sPos[jj] = make_float2(rpe_key_half(p.x, a.Rh), rpe_key_half(p.y, a.Rw));
    }
    stage_async_wait();
    __syncthreads();
#pragma unroll 1
    for (int jc = 0; jc < TK; jc += CHM) {
      if (j0 + jc >= a.Ns) break;
      // ---- scores of 16 rows x 32 keys
      float sc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { sc[nt][0] = 0.f; sc[nt][1] = 0.f; sc[nt][2] = 0.f; sc[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kh[NT][2], kl[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (jc + nt * 8 + g) * PAD + ks * 8 + t;      // B[k = t (+4)][n = g] = K[key g][channel t (+4)]
          kh[nt][0] = __float_as_uint(sKh[off]); kh[nt][1] = __float_as_uint(sKh[off + 4]);
          kl[nt][0] = __float_as_uint(sKl[off]); kl[nt][1] = __float_as_uint(sKl[off + 4]);
        }
        // the three products of one accumulator are issued NT tiles apart (they depend on each other)
#if K3_FWD_Q_SMEM
        uint32_t ah[4], al[4];
        {
          const float4 f = sQf[warp][ks][lane];
          const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { ah[e] = tf32_hi(fv[e]); al[e] = __float_as_uint(fv[e] - __uint_as_float(ah[e])); }
        }
#else
        const uint32_t (&ah)[4] = qh[ks];
        const uint32_t (&al)[4] = ql[ks];
#endif
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, kh[nt][0], kh[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, kl[nt][0], kl[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], al, kh[nt][0], kh[nt][1]);
      }
      // ---- + sampled bias; element e of tile nt: row (e >> 1), key jc + nt*8 + 2t + (e & 1)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float4 pp = *reinterpret_cast<const float4*>(&sPos[jc + nt * 8 + 2 * t]);   // two keys' halves
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1, kk = e & 1;
          const Bilin b = rpe_taps(ay[r], ax[r], kk ? pp.z : pp.x, kk ? pp.w : pp.y, a.Rh, a.Rw, a.Wp);
          float tv[4];
          const float v = sc[nt][e] + rpe_bias(table, b, a.Wp, tv);
          sc[nt][e] = (j0 + jc + nt * 8 + 2 * t + kk < a.Ns) ? v : -INFINITY;
          mx[r] = fmaxf(mx[r], sc[nt][e]);
        }
      }
      // ---- online softmax (a row lives in the four lanes of a quad)
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float mn = fmaxf(m[r], mx[r]);
        corr[r] = __expf(m[r] - mn);
        l[r] *= corr[r];
        m[r] = mn;
      }
#pragma unroll
      for (int nv = 0; nv < KS; ++nv) { o[nv][0] *= corr[0]; o[nv][1] *= corr[0]; o[nv][2] *= corr[1]; o[nv][3] *= corr[1]; }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sc[nt][e] = __expf(sc[nt][e] - m[e >> 1]);
          l[e >> 1] += sc[nt][e];                  // per-lane partial row sum, reduced over the quad at the end
          sc[nt][e] *= drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, rowb[e >> 1], (unsigned)(j0 + jc + nt * 8 + 2 * t + (e & 1)));
        }
      // ---- O += P V; k-step kt = the probabilities of key tile kt, operand columns (t, t+4) = keys (2t, 2t+1)
#pragma unroll
      for (int kt = 0; kt < NT; ++kt) {
        uint32_t ph[4], pl[4];
        const float pv[4] = {sc[kt][0], sc[kt][2], sc[kt][1], sc[kt][3]};   // a0 (g,2t) a1 (g+8,2t) a2 (g,2t+1) a3 (g+8,2t+1)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ph[e] = tf32_hi(pv[e]);
          pl[e] = __float_as_uint(pv[e] - __uint_as_float(ph[e]));
        }
        uint32_t vh[KS][2], vl[KS][2];
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) {
          const int off = (jc + kt * 8 + 2 * t) * PAD + nv * 8 + g;  // B[k = t][n = g] = V[key 2t][channel g]; k = t+4: key 2t+1
          vh[nv][0] = __float_as_uint(sVh[off]); vh[nv][1] = __float_as_uint(sVh[off + PAD]);
          vl[nv][0] = __float_as_uint(sVl[off]); vl[nv][1] = __float_as_uint(sVl[off + PAD]);
        }
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(o[nv], ph, vh[nv][0], vh[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(o[nv], ph, vl[nv][0], vl[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(o[nv], pl, vh[nv][0], vh[nv][1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  // accumulator element e of channel tile nv: row (e >> 1), channel nv*8 + 2t + (e & 1)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = iq[r];
    if (i >= a.Nq) continue;
    if (a.split == 1) {
      const float inv = 1.f / l[r];
#pragma unroll
      for (int nv = 0; nv < KS; ++nv)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          a.o_out[((size_t)bh * DH + nv * 8 + 2 * t + kk) * a.Nq + i] = o[nv][2 * r + kk] * inv;
      if (t == 0) a.o_lse[(size_t)bh * a.Nq + i] = m[r] + logf(l[r]);
    } else {
      const size_t pz = (size_t)blockIdx.z * a.BH + bh;
#pragma unroll
      for (int nv = 0; nv < KS; ++nv)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          a.part_o[(pz * DH + nv * 8 + 2 * t + kk) * a.Nq + i] = o[nv][2 * r + kk];
      if (t == 0) {
        a.part_ml[(pz * 2 + 0) * a.Nq + i] = m[r];
        a.part_ml[(pz * 2 + 1) * a.Nq + i] = l[r];
      }
    }
  }
}

// merge the key ranges of the split forward: out = sum_z o_z e^{m_z - m} / sum_z l_z e^{m_z - m}
template <int DH>
__global__ void __launch_bounds__(128) rpe_combine_kernel(const __grid_constant__ RpeArgs a) {
  const int bh = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.Nq) return;
  float m = -INFINITY;
  for (int z = 0; z < a.split; ++z) m = fmaxf(m, __ldg(a.part_ml + (((size_t)z * a.BH + bh) * 2) * a.Nq + i));
  float l = 0.f, o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;
  for (int z = 0; z < a.split; ++z) {
    const size_t pz = (size_t)z * a.BH + bh;
    const float w = __expf(__ldg(a.part_ml + (pz * 2) * a.Nq + i) - m);     // empty range: e^{-inf} = 0
    l = fmaf(__ldg(a.part_ml + (pz * 2 + 1) * a.Nq + i), w, l);
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = fmaf(__ldg(a.part_o + (pz * DH + c) * a.Nq + i), w, o[c]);
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < DH; ++c) a.o_out[((size_t)bh * DH + c) * a.Nq + i] = o[c] * inv;
  a.o_lse[(size_t)bh * a.Nq + i] = m + logf(l);
}

// ------------------------------------------------------------------ backward
// delta_i = sum_c dO[c][i] * O[c][i]
template <int DH>
__global__ void __launch_bounds__(256) rpe_delta_kernel(const __grid_constant__ RpeArgs a) {
  const int bh = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.Nq) return;
  float s = 0.f;
#pragma unroll 8
  for (int c = 0; c < DH; ++c) {
    const size_t idx = ((size_t)bh * DH + c) * a.Nq + i;
    s = fmaf(__ldg(a.grad_out + idx), __ldg(a.out + idx), s);
  }
  a.o_delta[(size_t)bh * a.Nq + i] = s;
}

// dQ: thread = query row, streams K/V tiles
template <int DH, bool DROP>
__global__ void __launch_bounds__(TQ, K3_DQ_MINB) rpe_attn_bwd_dq_kernel(const __grid_constant__ RpeArgs a) {
  __shared__ __align__(16) float sK[TK * PAD];
  __shared__ __align__(16) float sV[TK * PAD];
  __shared__ float2 sPos[TK];
  const int bh = blockIdx.y, i = blockIdx.x * TQ + threadIdx.x;
  const bool live = i < a.Nq;
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  const float* kb = a.k + (size_t)bh * DH * a.Ns;
  const float* vb = a.v + (size_t)bh * DH * a.Ns;
  const float2* pb = reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns;

  float2 qr[DH / 2], go[DH / 2], dq[DH / 2];
#pragma unroll
  for (int c = 0; c < DH; ++c) {
    const size_t idx = ((size_t)bh * DH + c) * a.Nq + i;
    lane_elem(qr, c) = live ? __ldg(a.q + idx) * a.scale : 0.f;
    lane_elem(go, c) = live ? __ldg(a.grad_out + idx) : 0.f;
    lane_elem(dq, c) = 0.f;
  }
  const float lse = live ? __ldg(a.lse + (size_t)bh * a.Nq + i) : 0.f;
  const float delta = live ? __ldg(a.delta + (size_t)bh * a.Nq + i) : 0.f;
  const int iy = i / a.Wq, ixq = i - iy * a.Wq;
  const float qy = rpe_query_half(grid_coord(iy, a.Hq), a.Rh), qx = rpe_query_half(grid_coord(ixq, a.Wq), a.Rw);
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh), rowb = (unsigned)i * (unsigned)a.Ns;

  const int jbeg = blockIdx.z * a.tiles_per_split * TK, jend = min(a.Ns, jbeg + a.tiles_per_split * TK);
  for (int j0 = jbeg; j0 < jend; j0 += TK) {
    __syncthreads();
    stage_transposed<TK, DH>(sK, kb, a.Ns, j0, 1.f);
    stage_transposed<TK, DH>(sV, vb, a.Ns, j0, 1.f);
    for (int jj = threadIdx.x; jj < TK; jj += blockDim.x) {
      const float2 p = (j0 + jj < a.Ns) ? __ldg(pb + j0 + jj) : make_float2(0.f, 0.f);
      sPos[jj] = make_float2(rpe_key_half(p.x, a.Rh), rpe_key_half(p.y, a.Rw));
    }
    __syncthreads();
    const int jn = min(TK, a.Ns - j0);
#pragma unroll 4
    for (int jj = 0; jj < jn; ++jj) {
      const float2 p = sPos[jj];
      const Bilin b = rpe_taps(qy, qx, p.x, p.y, a.Rh, a.Rw, a.Wp);
      float tv[4];
      const float s = dot32<DH>(qr, sK + jj * PAD) + rpe_bias(table, b, a.Wp, tv);
      const float pr = __expf(s - lse);
      const float dp = dot32<DH>(go, sV + jj * PAD) * drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, rowb, (unsigned)(j0 + jj));
      axpy32<DH>(dq, pr * (dp - delta), sK + jj * PAD);
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < DH; ++c) {
      float* dst = a.grad_q + ((size_t)bh * DH + c) * a.Nq + i;
      if (a.split == 1) *dst = lane_elem(dq, c) * a.scale;
      else atomicAdd(dst, lane_elem(dq, c) * a.scale);      // key ranges add into the zero-filled gradient
    }
  }
}

// dQ with the three contractions on mma.sync 3xTF32 (same warp tile and staging as rpe_attn_fwd_mma_kernel):
// S = Q^T K and dP = dO^T V use K / V as the "col" operand indexed (key = g, channel = t); dQ += dS K uses K indexed
// (key = 2t / 2t+1, channel = g) so that dS is consumed in the accumulator registers it was computed in.
#ifndef K3_DQ_MMA
#define K3_DQ_MMA 1
#endif
#ifndef K3_DQ_QG_SMEM
#define K3_DQ_QG_SMEM 0                      // dQ kernel: Q / dO fragments in shared memory instead of 64 registers (measured: slower here, 7.24 vs 7.05 ms SCA bwd)
#endif
#ifndef K3_DQ_MMA_MINB
#define K3_DQ_MMA_MINB (K3_DQ_QG_SMEM ? 4 : 3)
#endif
template <int DH, bool DROP>
__global__ void __launch_bounds__(128, K3_DQ_MMA_MINB) rpe_attn_bwd_dq_mma_kernel(const __grid_constant__ RpeArgs a) {
  constexpr int KS = DH / 8, NT = CHM / 8;
  __shared__ __align__(16) float sKh[TK * PAD], sKl[TK * PAD], sVh[TK * PAD], sVl[TK * PAD];
  __shared__ __align__(16) float2 sPos[TK];
#if K3_DQ_QG_SMEM
  // the warp's Q / dO fragments (A operands of S and dP) in shared memory, split into tf32 hi / lo when used (see the dK/dV kernel)
  extern __shared__ __align__(16) float4 s_qgf[];
  float4 (*sQf)[KS][32] = reinterpret_cast<float4 (*)[KS][32]>(s_qgf);
  float4 (*sGf)[KS][32] = reinterpret_cast<float4 (*)[KS][32]>(s_qgf + 4 * KS * 32);
#endif
  const int bh = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  const float2* pb = reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns;

  const int i0 = blockIdx.x * TQM + warp * 16;
  const int iq[2] = {i0 + g, i0 + g + 8};
  float ay[2], ax[2], lse[2], delta[2];
#if !K3_DQ_QG_SMEM
  uint32_t qh[KS][4], ql[KS][4], gh[KS][4], gl[KS][4];
#endif
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int iy = iq[r] / a.Wq, ixq = iq[r] - iy * a.Wq;
    ay[r] = rpe_query_half(grid_coord(iy, a.Hq), a.Rh);
    ax[r] = rpe_query_half(grid_coord(ixq, a.Wq), a.Rw);
    const bool live = iq[r] < a.Nq;
    lse[r] = live ? __ldg(a.lse + (size_t)bh * a.Nq + iq[r]) : 0.f;
    delta[r] = live ? __ldg(a.delta + (size_t)bh * a.Nq + iq[r]) : 0.f;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {            // a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
      const int r = e & 1, c = ks * 8 + t + ((e >> 1) << 2);
      const size_t idx = ((size_t)bh * DH + c) * a.Nq + iq[r];
      const float x = iq[r] < a.Nq ? __ldg(a.q + idx) * a.scale : 0.f;
      const float y = iq[r] < a.Nq ? __ldg(a.grad_out + idx) : 0.f;
#if K3_DQ_QG_SMEM
      reinterpret_cast<float*>(&sQf[warp][ks][lane])[e] = x;
      reinterpret_cast<float*>(&sGf[warp][ks][lane])[e] = y;
#else
      qh[ks][e] = tf32_hi(x); ql[ks][e] = __float_as_uint(x - __uint_as_float(qh[ks][e]));
      gh[ks][e] = tf32_hi(y); gl[ks][e] = __float_as_uint(y - __uint_as_float(gh[ks][e]));
#endif
    }
  float dq[KS][4];
#pragma unroll
  for (int nv = 0; nv < KS; ++nv)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[nv][e] = 0.f;
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh);
  const unsigned rowb[2] = {(unsigned)iq[0] * (unsigned)a.Ns, (unsigned)iq[1] * (unsigned)a.Ns};

  const int jbeg = blockIdx.z * a.tiles_per_split * TK, jend = min(a.Ns, jbeg + a.tiles_per_split * TK);
  for (int j0 = jbeg; j0 < jend; j0 += TK) {
    __syncthreads();
    {
      const size_t tile = ((size_t)bh * a.NsP + j0) * DH;
      stage_async<DH>(sKh, a.kh + tile); stage_async<DH>(sKl, a.kl + tile);
      stage_async<DH>(sVh, a.vh + tile); stage_async<DH>(sVl, a.vl + tile);
    }
    for (int jj = threadIdx.x; jj < TK; jj += blockDim.x) {
      const float2 p = (j0 + jj < a.Ns) ? __ldg(pb + j0 + jj) : make_float2(0.f, 0.f);
      sPos[jj] = make_float2(rpe_key_half(p.x, a.Rh), rpe_key_half(p.y, a.Rw));
    }
    stage_async_wait();
    __syncthreads();
#pragma unroll 1
    for (int jc = 0; jc < TK; jc += CHM) {
      if (j0 + jc >= a.Ns) break;
      float sc[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { sc[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh_[NT][2], bl_[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (jc + nt * 8 + g) * PAD + ks * 8 + t;
          bh_[nt][0] = __float_as_uint(sKh[off]); bh_[nt][1] = __float_as_uint(sKh[off + 4]);
          bl_[nt][0] = __float_as_uint(sKl[off]); bl_[nt][1] = __float_as_uint(sKl[off + 4]);
        }
#if K3_DQ_QG_SMEM
        uint32_t ah[4], al[4];
        {
          const float4 f = sQf[warp][ks][lane];
          const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { ah[e] = tf32_hi(fv[e]); al[e] = __float_as_uint(fv[e] - __uint_as_float(ah[e])); }
        }
#else
        const uint32_t (&ah)[4] = qh[ks];
        const uint32_t (&al)[4] = ql[ks];
#endif
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, bl_[nt][0], bl_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], al, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (jc + nt * 8 + g) * PAD + ks * 8 + t;
          bh_[nt][0] = __float_as_uint(sVh[off]); bh_[nt][1] = __float_as_uint(sVh[off + 4]);
          bl_[nt][0] = __float_as_uint(sVl[off]); bl_[nt][1] = __float_as_uint(sVl[off + 4]);
        }
#if K3_DQ_QG_SMEM
        uint32_t wh[4], wl[4];
        {
          const float4 f = sGf[warp][ks][lane];
          const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { wh[e] = tf32_hi(fv[e]); wl[e] = __float_as_uint(fv[e] - __uint_as_float(wh[e])); }
        }
#else
        const uint32_t (&wh)[4] = gh[ks];
        const uint32_t (&wl)[4] = gl[ks];
#endif
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wh, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wh, bl_[nt][0], bl_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wl, bh_[nt][0], bh_[nt][1]);
      }
      // dS = P (dP - delta), P = exp(S + bias - lse); element e of tile nt: row (e >> 1), key jc + nt*8 + 2t + (e & 1)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float4 pp = *reinterpret_cast<const float4*>(&sPos[jc + nt * 8 + 2 * t]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1, kk = e & 1;
          const Bilin b = rpe_taps(ay[r], ax[r], kk ? pp.z : pp.x, kk ? pp.w : pp.y, a.Rh, a.Rw, a.Wp);
          float tv[4];
          const float pr = __expf(sc[nt][e] + rpe_bias(table, b, a.Wp, tv) - lse[r]);
          const float dm = drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, rowb[r], (unsigned)(j0 + jc + nt * 8 + 2 * t + kk));
          sc[nt][e] = (j0 + jc + nt * 8 + 2 * t + kk < a.Ns) ? pr * (dm * dp[nt][e] - delta[r]) : 0.f;
        }
      }
      // dQ += dS K; operand columns (t, t+4) = keys (2t, 2t+1)
#pragma unroll
      for (int kt = 0; kt < NT; ++kt) {
        uint32_t ph[4], pl[4];
        const float pv[4] = {sc[kt][0], sc[kt][2], sc[kt][1], sc[kt][3]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ph[e] = tf32_hi(pv[e]);
          pl[e] = __float_as_uint(pv[e] - __uint_as_float(ph[e]));
        }
        uint32_t kh[KS][2], kl[KS][2];
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) {
          const int off = (jc + kt * 8 + 2 * t) * PAD + nv * 8 + g;
          kh[nv][0] = __float_as_uint(sKh[off]); kh[nv][1] = __float_as_uint(sKh[off + PAD]);
          kl[nv][0] = __float_as_uint(sKl[off]); kl[nv][1] = __float_as_uint(sKl[off + PAD]);
        }
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dq[nv], ph, kh[nv][0], kh[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dq[nv], ph, kl[nv][0], kl[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dq[nv], pl, kh[nv][0], kh[nv][1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (iq[r] >= a.Nq) continue;
#pragma unroll
    for (int nv = 0; nv < KS; ++nv)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        float* dst = a.grad_q + ((size_t)bh * DH + nv * 8 + 2 * t + kk) * a.Nq + iq[r];
        if (a.split == 1) *dst = dq[nv][2 * r + kk] * a.scale;
        else atomicAdd(dst, dq[nv][2 * r + kk] * a.scale);
      }
  }
}

// dK, dV, dpos, drpe: thread = key column, streams Q/dO tiles
#ifndef K3_TQB
#define K3_TQB 128
#endif
#ifndef K3_DKV_MINB
#define K3_DKV_MINB 1
#endif
constexpr int TKB = 128, TQB = K3_TQB;

// K3_CS lanes share one key, each owning DH / K3_CS channels of k, v, dk, dv (the 128 registers of per-key state are what
// limits this kernel to 8 warps per SM): the two half dot products are combined with one shuffle each, the bias, the
// softmax probability and the bias gradients are evaluated by both lanes (identical values), and lane 0 of the pair
// issues the atomics.
#ifndef K3_CS
#define K3_CS 1
#endif
#ifndef K3_DKV_MINB2
#define K3_DKV_MINB2 2
#endif
constexpr int CS = K3_CS;
constexpr int DKV_MINB = CS == 1 ? K3_DKV_MINB : K3_DKV_MINB2;

template <int DH, bool DROP>
__global__ void __launch_bounds__(TKB * CS, DKV_MINB) rpe_attn_bwd_dkv_kernel(const __grid_constant__ RpeArgs a) {
  constexpr int DL = DH / CS;         // channels per lane
  __shared__ __align__(16) float sQ[TQB * PAD];
  __shared__ __align__(16) float sG[TQB * PAD];
  __shared__ float4 sRow[TQB];    // query halves of the bias coordinates (y, x), lse, delta
  const int bh = blockIdx.y, j = blockIdx.x * TKB + threadIdx.x / CS;
  const int part = threadIdx.x % CS, c0 = part * DL;
  const bool live = j < a.Ns;
  const unsigned lmask = __ballot_sync(0xffffffffu, live);    // lanes that run the pair loop (whole key groups)
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  // the table gradient is a hot spot (every (query, key) pair of a head lands on ~12k cells): spread the atomics of
  // different CTAs over `gcopies` private copies, summed by rpe_unpad_add_kernel (-47 % backward time at N=56)
  float* gtable = a.grad_rpe_pad ? a.grad_rpe_pad + ((size_t)((blockIdx.x + blockIdx.y) % a.gcopies) * a.h + head) * a.Hp * a.Wp +
                                       2 * a.Wp + 2 : nullptr;
  if (part != 0) gtable = nullptr;
  const float* qb = a.q + (size_t)bh * DH * a.Nq;
  const float* gb = a.grad_out + (size_t)bh * DH * a.Nq;

  float2 kr[DL / 2], vr[DL / 2], dk[DL / 2], dv[DL / 2];
#pragma unroll
  for (int c = 0; c < DL; ++c) {
    const size_t idx = ((size_t)bh * DH + c0 + c) * a.Ns + j;
    lane_elem(kr, c) = live ? __ldg(a.k + idx) : 0.f;
    lane_elem(vr, c) = live ? __ldg(a.v + idx) : 0.f;
    lane_elem(dk, c) = 0.f; lane_elem(dv, c) = 0.f;
  }
  float2 pj = make_float2(0.f, 0.f);
  if (live) pj = __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns + j);
  pj = make_float2(rpe_key_half(pj.x, a.Rh), rpe_key_half(pj.y, a.Rw));
  float giy = 0.f, gix = 0.f;
  const bool want_bias_grad = (a.grad_rpe_pad != nullptr) || (a.grad_pos != nullptr);
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh);

  const int ibeg = blockIdx.z * a.tiles_per_split * TQB, iend = min(a.Nq, ibeg + a.tiles_per_split * TQB);
  for (int i0 = ibeg; i0 < iend; i0 += TQB) {
    __syncthreads();
    stage_transposed<TQB, DH>(sQ, qb, a.Nq, i0, a.scale);
    stage_transposed<TQB, DH>(sG, gb, a.Nq, i0, 1.f);
    for (int ii = threadIdx.x; ii < TQB; ii += blockDim.x) {
      const int i = i0 + ii;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < a.Nq) {
        const int iy = i / a.Wq, ixq = i - iy * a.Wq;
        r = make_float4(rpe_query_half(grid_coord(iy, a.Hq), a.Rh), rpe_query_half(grid_coord(ixq, a.Wq), a.Rw),
                        __ldg(a.lse + (size_t)bh * a.Nq + i),
                        __ldg(a.delta + (size_t)bh * a.Nq + i));
      }
      sRow[ii] = r;
    }
    __syncthreads();
    if (!live) continue;
    const int in = min(TQB, a.Nq - i0);
#pragma unroll 2
    for (int ii = 0; ii < in; ++ii) {
      const float4 r = sRow[ii];
      const Bilin b = rpe_taps(r.x, r.y, pj.x, pj.y, a.Rh, a.Rw, a.Wp);
      float tv[4];
      float qk = dot32<DL>(kr, sQ + ii * PAD + c0);
      float dp = dot32<DL>(vr, sG + ii * PAD + c0);
#pragma unroll
      for (int sh = 1; sh < CS; sh <<= 1) {
        qk += __shfl_xor_sync(lmask, qk, sh);
        dp += __shfl_xor_sync(lmask, dp, sh);
      }
      const float s = qk + rpe_bias(table, b, a.Wp, tv);
      const float pr = __expf(s - r.z);
      const float dm = drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, (unsigned)(i0 + ii) * (unsigned)a.Ns, (unsigned)j);
      axpy32<DL>(dv, pr * dm, sG + ii * PAD + c0);
      const float ds = pr * (dm * dp - r.w);
      axpy32<DL>(dk, ds, sQ + ii * PAD + c0);
      if (want_bias_grad) {
        gix = fmaf(ds, b.wy0 * (tv[1] - tv[0]) + b.wy1 * (tv[3] - tv[2]), gix);
        giy = fmaf(ds, b.wx0 * (tv[2] - tv[0]) + b.wx1 * (tv[3] - tv[1]), giy);
        if (gtable) {
          float* g0 = gtable + b.off;
          float* g1 = g0 + a.Wp;
          const float dx0 = ds * b.wx0, dx1 = ds * b.wx1;
          atomicAdd(g0, dx0 * b.wy0); atomicAdd(g0 + 1, dx1 * b.wy0);
          atomicAdd(g1, dx0 * b.wy1); atomicAdd(g1 + 1, dx1 * b.wy1);
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < DL; ++c) {
      const size_t idx = ((size_t)bh * DH + c0 + c) * a.Ns + j;
      if (a.split == 1) {
        if (a.grad_k) a.grad_k[idx] = lane_elem(dk, c);    // sQ already carries `scale`
        if (a.grad_v) a.grad_v[idx] = lane_elem(dv, c);
      } else {                                              // query ranges add into the zero-filled gradients
        if (a.grad_k) atomicAdd(a.grad_k + idx, lane_elem(dk, c));
        if (a.grad_v) atomicAdd(a.grad_v + idx, lane_elem(dv, c));
      }
    }
    if (a.grad_pos && part == 0) {
      // table coord = (Rx-1) * ((q - pos)/4 + 1/2)  ->  d coord / d pos = -(Rx-1)/4
      float* gp = a.grad_pos + ((size_t)bg * a.Ns + j) * 2;
      atomicAdd(gp, giy * (-0.25f * (float)(a.Rh - 1)));
      atomicAdd(gp + 1, gix * (-0.25f * (float)(a.Rw - 1)));
    }
  }
}

// dK / dV / dpos / drpe with the four contractions on mma.sync 3xTF32.  One warp owns 16 KEYS (the M rows of every tile):
//   S^T = K Q^T, dP^T = V dO^T   (A = resident K / V fragments, B = staged Q / dO indexed (query = g, channel = t))
//   dV += P^T dO, dK += dS^T Q   (A = the probabilities / score gradients in their accumulator registers with the
//                                 (0,2,4,6,1,3,5,7) query order, B = staged dO / Q indexed (query = 2t / 2t+1, channel = g))
// Element e of score tile nt: key row (e >> 1), query jc + nt*8 + 2t + (e & 1).
#ifndef K3_DKV_MMA
#define K3_DKV_MMA 1
#endif
#ifndef K3_DKV_KV_SMEM
#define K3_DKV_KV_SMEM 1                     // dK/dV kernel: the resident K / V fragments in shared memory instead of 64 registers
#endif
#ifndef K3_DKV_MMA_MINB
#define K3_DKV_MMA_MINB (K3_DKV_KV_SMEM ? 4 : 3)
#endif
constexpr int TKM = 64;                      // keys per CTA (4 warps x 16)

template <int DH, bool DROP, bool MERGE>
__global__ void __launch_bounds__(128, K3_DKV_MMA_MINB) rpe_attn_bwd_dkv_mma_kernel(const __grid_constant__ RpeArgs a) {
  constexpr int KS = DH / 8, NT = CHM / 8;
  __shared__ __align__(16) float sQh[TK * PAD], sQl[TK * PAD], sGh[TK * PAD], sGl[TK * PAD];
  __shared__ __align__(16) float4 sRow[TK];  // query halves of the bias coordinates (y, x), lse, delta
#if K3_DKV_KV_SMEM
  // the warp's K / V fragments (A operands of S^T and dP^T) live in shared memory, one float4 {a0, a1, a2, a3} per
  // (k-step, lane), and are split into tf32 hi / lo when used: 64 registers less per thread (168 -> 4 CTAs per SM at 128)
  // (dynamic: with the staged Q / dO tiles the kernel needs more than the 48 KB a kernel may declare statically)
  extern __shared__ __align__(16) float4 s_kvf[];
  float4 (*sKf)[KS][32] = reinterpret_cast<float4 (*)[KS][32]>(s_kvf);
  float4 (*sVf)[KS][32] = reinterpret_cast<float4 (*)[KS][32]>(s_kvf + 4 * KS * 32);
#endif
  const int bh = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int head = bh % a.h, bg = bh / a.hpg;
  const float* table = opaque_f(a.rpe_pad + (size_t)head * a.Hp * a.Wp + 2 * a.Wp + 2);
  float* gtable = a.grad_rpe_pad ? a.grad_rpe_pad + ((size_t)((blockIdx.x + blockIdx.y) % a.gcopies) * a.h + head) * a.Hp * a.Wp +
                                       2 * a.Wp + 2 : nullptr;
  const bool want_bias_grad = (a.grad_rpe_pad != nullptr) || (a.grad_pos != nullptr);

  const int j0w = blockIdx.x * TKM + warp * 16;
  const int jk[2] = {j0w + g, j0w + g + 8};
  float by[2], bx[2], giy[2] = {0.f, 0.f}, gix[2] = {0.f, 0.f};
#if !K3_DKV_KV_SMEM
  uint32_t kh[KS][4], kl[KS][4], vh[KS][4], vl[KS][4];
#endif
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float2 p = make_float2(0.f, 0.f);
    if (jk[r] < a.Ns) p = __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)bg * a.Ns + jk[r]);
    by[r] = rpe_key_half(p.x, a.Rh); bx[r] = rpe_key_half(p.y, a.Rw);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {            // a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
      const int r = e & 1, c = ks * 8 + t + ((e >> 1) << 2);
      const size_t idx = ((size_t)bh * DH + c) * a.Ns + jk[r];
      const float x = jk[r] < a.Ns ? __ldg(a.k + idx) : 0.f;
      const float y = jk[r] < a.Ns ? __ldg(a.v + idx) : 0.f;
#if K3_DKV_KV_SMEM
      reinterpret_cast<float*>(&sKf[warp][ks][lane])[e] = x;
      reinterpret_cast<float*>(&sVf[warp][ks][lane])[e] = y;
#else
      kh[ks][e] = tf32_hi(x); kl[ks][e] = __float_as_uint(x - __uint_as_float(kh[ks][e]));
      vh[ks][e] = tf32_hi(y); vl[ks][e] = __float_as_uint(y - __uint_as_float(vh[ks][e]));
#endif
    }
  float dk[KS][4], dv[KS][4];
#pragma unroll
  for (int nv = 0; nv < KS; ++nv)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[nv][e] = 0.f; dv[nv][e] = 0.f; }
  const unsigned dkey = drop_key(a.seed_lo, a.seed_hi, bh);

  const int ibeg = blockIdx.z * a.tiles_per_split * TK, iend = min(a.Nq, ibeg + a.tiles_per_split * TK);
  for (int i0 = ibeg; i0 < iend; i0 += TK) {
    __syncthreads();
    {
      const size_t tile = ((size_t)bh * a.NqP + i0) * DH;
      stage_async<DH>(sQh, a.qh + tile); stage_async<DH>(sQl, a.ql + tile);
      stage_async<DH>(sGh, a.gh + tile); stage_async<DH>(sGl, a.gl + tile);
    }
    for (int ii = threadIdx.x; ii < TK; ii += blockDim.x) {
      const int i = i0 + ii;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < a.Nq) {
        const int iy = i / a.Wq, ixq = i - iy * a.Wq;
        r = make_float4(rpe_query_half(grid_coord(iy, a.Hq), a.Rh), rpe_query_half(grid_coord(ixq, a.Wq), a.Rw),
                        __ldg(a.lse + (size_t)bh * a.Nq + i), __ldg(a.delta + (size_t)bh * a.Nq + i));
      }
      sRow[ii] = r;
    }
    stage_async_wait();
    __syncthreads();
#pragma unroll 1
    for (int jc = 0; jc < TK; jc += CHM) {
      if (i0 + jc >= a.Nq) break;
      float sc[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { sc[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh_[NT][2], bl_[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (jc + nt * 8 + g) * PAD + ks * 8 + t;
          bh_[nt][0] = __float_as_uint(sQh[off]); bh_[nt][1] = __float_as_uint(sQh[off + 4]);
          bl_[nt][0] = __float_as_uint(sQl[off]); bl_[nt][1] = __float_as_uint(sQl[off + 4]);
        }
#if K3_DKV_KV_SMEM
        uint32_t ah[4], al[4];
        {
          const float4 f = sKf[warp][ks][lane];
          const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { ah[e] = tf32_hi(fv[e]); al[e] = __float_as_uint(fv[e] - __uint_as_float(ah[e])); }
        }
#else
        const uint32_t (&ah)[4] = kh[ks];
        const uint32_t (&al)[4] = kl[ks];
#endif
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], ah, bl_[nt][0], bl_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(sc[nt], al, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (jc + nt * 8 + g) * PAD + ks * 8 + t;
          bh_[nt][0] = __float_as_uint(sGh[off]); bh_[nt][1] = __float_as_uint(sGh[off + 4]);
          bl_[nt][0] = __float_as_uint(sGl[off]); bl_[nt][1] = __float_as_uint(sGl[off + 4]);
        }
#if K3_DKV_KV_SMEM
        uint32_t wh[4], wl[4];
        {
          const float4 f = sVf[warp][ks][lane];
          const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { wh[e] = tf32_hi(fv[e]); wl[e] = __float_as_uint(fv[e] - __uint_as_float(wh[e])); }
        }
#else
        const uint32_t (&wh)[4] = vh[ks];
        const uint32_t (&wl)[4] = vl[ks];
#endif
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wh, bh_[nt][0], bh_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wh, bl_[nt][0], bl_[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(dp[nt], wl, bh_[nt][0], bh_[nt][1]);
      }
      // P^T -> sc, dS^T -> dp (in place), bias gradients
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float4 r0 = sRow[jc + nt * 8 + 2 * t], r1 = sRow[jc + nt * 8 + 2 * t + 1];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1, kk = e & 1;
          const float4 rq = kk ? r1 : r0;
          const Bilin b = rpe_taps(rq.x, rq.y, by[r], bx[r], a.Rh, a.Rw, a.Wp);
          float tv[4];
          const bool ok = (i0 + jc + nt * 8 + 2 * t + kk < a.Nq) && (jk[r] < a.Ns);
          const float bias = rpe_bias(table, b, a.Wp, tv);          // always sampled: the taps feed the gradients below
          const float pr = ok ? __expf(sc[nt][e] + bias - rq.z) : 0.f;
          const float dm = drop_mul<DROP>(a.drop_thresh, a.drop_scale, dkey, (unsigned)(i0 + jc + nt * 8 + 2 * t + kk) * (unsigned)a.Ns, (unsigned)jk[r]);
          const float ds = pr * (dm * dp[nt][e] - rq.w);
          sc[nt][e] = pr * dm;
          dp[nt][e] = ds;
          if (want_bias_grad) {
            gix[r] = fmaf(ds, b.wy0 * (tv[1] - tv[0]) + b.wy1 * (tv[3] - tv[2]), gix[r]);
            giy[r] = fmaf(ds, b.wx0 * (tv[2] - tv[0]) + b.wx1 * (tv[3] - tv[1]), giy[r]);
            if (gtable) {
              float* g0 = gtable + b.off;
              float* g1 = g0 + a.Wp;
              const float dx0 = ds * b.wx0, dx1 = ds * b.wx1;
              float l0 = dx0 * b.wy0, l1 = dx0 * b.wy1;            // taps (y0, x0), (y0+1, x0)
              float r0 = dx1 * b.wy0, r1 = dx1 * b.wy1;            // taps (y0, x0+1), (y0+1, x0+1)
              // Neighbouring key rows of the tile (lanes 4 apart) sit one table cell apart, so the right-hand taps of
              // key g+1 are the left-hand taps of key g: hand them over with shuffles and halve the RED sectors.
              const int off_up = __shfl_down_sync(0xffffffffu, b.off, 4), off_dn = __shfl_up_sync(0xffffffffu, b.off, 4);
              // Key rows that address the SAME cell (SCA's masked reference points, once the module has ordered the keys
              // by table row and x: ~60 keys per cell) are not coalesced by the memory system -- every lane of a RED
              // instruction with an equal address is its own L2 reduction.  When any two neighbouring key rows of the warp
              // agree (warp-uniform test, free: off_up is needed anyway), the runs of equal cells of each query column are
              // summed with a segmented suffix scan over the 8 key rows and only the run heads issue atomics.
              // MERGE is a per-call flag (BEVR_K3_RUN_MERGE): the instantiation without it is the plain kernel.
              if (MERGE && __any_sync(0xffffffffu, g < 7 && off_up == b.off)) {
                const bool head = g == 0 || off_dn != b.off;
                const unsigned hx = (__ballot_sync(0xffffffffu, head) >> t) & 0x11111111u;   // heads of this column, bit 4 g'
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                  const float p0 = __shfl_down_sync(0xffffffffu, l0, 4 * d), p1 = __shfl_down_sync(0xffffffffu, l1, 4 * d);
                  const float p2 = __shfl_down_sync(0xffffffffu, r0, 4 * d), p3 = __shfl_down_sync(0xffffffffu, r1, 4 * d);
                  // rows g+1 .. g+d hold no run head <=> row g+d continues the run of row g
                  const unsigned span = (g + 1 < 8 ? (0x11111111u << (4 * (g + 1))) : 0u) &
                                        ~(g + d + 1 < 8 ? (0x11111111u << (4 * (g + d + 1))) : 0u);
                  if (g + d < 8 && !(hx & span)) { l0 += p0; l1 += p1; r0 += p2; r1 += p3; }
                }
                if (head) { atomicAdd(g0, l0); atomicAdd(g1, l1); atomicAdd(g0 + 1, r0); atomicAdd(g1 + 1, r1); }
                continue;
              }
              const float in0 = __shfl_down_sync(0xffffffffu, r0, 4), in1 = __shfl_down_sync(0xffffffffu, r1, 4);
              if (g < 7 && off_up + 1 == b.off) { l0 += in0; l1 += in1; }
              atomicAdd(g0, l0); atomicAdd(g1, l1);
              if (!(g > 0 && b.off + 1 == off_dn)) { atomicAdd(g0 + 1, r0); atomicAdd(g1 + 1, r1); }
            }
          }
        }
      }
      // dV += P^T dO, dK += dS^T Q; operand columns (t, t+4) = queries (2t, 2t+1)
#pragma unroll
      for (int kt = 0; kt < NT; ++kt) {
        uint32_t ph[4], pl[4], sh[4], sl[4];
        const float pv[4] = {sc[kt][0], sc[kt][2], sc[kt][1], sc[kt][3]};
        const float sv[4] = {dp[kt][0], dp[kt][2], dp[kt][1], dp[kt][3]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ph[e] = tf32_hi(pv[e]); pl[e] = __float_as_uint(pv[e] - __uint_as_float(ph[e]));
          sh[e] = tf32_hi(sv[e]); sl[e] = __float_as_uint(sv[e] - __uint_as_float(sh[e]));
        }
        uint32_t xh[KS][2], xl[KS][2];
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) {
          const int off = (jc + kt * 8 + 2 * t) * PAD + nv * 8 + g;
          xh[nv][0] = __float_as_uint(sGh[off]); xh[nv][1] = __float_as_uint(sGh[off + PAD]);
          xl[nv][0] = __float_as_uint(sGl[off]); xl[nv][1] = __float_as_uint(sGl[off + PAD]);
        }
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dv[nv], ph, xh[nv][0], xh[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dv[nv], ph, xl[nv][0], xl[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dv[nv], pl, xh[nv][0], xh[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) {
          const int off = (jc + kt * 8 + 2 * t) * PAD + nv * 8 + g;
          xh[nv][0] = __float_as_uint(sQh[off]); xh[nv][1] = __float_as_uint(sQh[off + PAD]);
          xl[nv][0] = __float_as_uint(sQl[off]); xl[nv][1] = __float_as_uint(sQl[off + PAD]);
        }
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dk[nv], sh, xh[nv][0], xh[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dk[nv], sh, xl[nv][0], xl[nv][1]);
#pragma unroll
        for (int nv = 0; nv < KS; ++nv) mma_tf32(dk[nv], sl, xh[nv][0], xh[nv][1]);
      }
    }
  }
  // bias-coordinate gradients of a key row live in the four lanes of a quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    giy[r] += __shfl_xor_sync(0xffffffffu, giy[r], 1); giy[r] += __shfl_xor_sync(0xffffffffu, giy[r], 2);
    gix[r] += __shfl_xor_sync(0xffffffffu, gix[r], 1); gix[r] += __shfl_xor_sync(0xffffffffu, gix[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = jk[r];
    if (j >= a.Ns) continue;
#pragma unroll
    for (int nv = 0; nv < KS; ++nv)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const size_t idx = ((size_t)bh * DH + nv * 8 + 2 * t + kk) * a.Ns + j;
        if (a.split == 1) {
          if (a.grad_k) a.grad_k[idx] = dk[nv][2 * r + kk];      // the staged Q carries `scale`
          if (a.grad_v) a.grad_v[idx] = dv[nv][2 * r + kk];
        } else {
          if (a.grad_k) atomicAdd(a.grad_k + idx, dk[nv][2 * r + kk]);
          if (a.grad_v) atomicAdd(a.grad_v + idx, dv[nv][2 * r + kk]);
        }
      }
    if (a.grad_pos && t == 0) {
      float* gp = a.grad_pos + ((size_t)bg * a.Ns + j) * 2;
      atomicAdd(gp, giy[r] * (-0.25f * (float)(a.Rh - 1)));
      atomicAdd(gp + 1, gix[r] * (-0.25f * (float)(a.Rw - 1)));
    }
  }
}

static int validate(int BH, int h, int hpg, int Dh, int Hq, int Wq, int Ns, int Rh, int Rw) {
  BEVR_REQUIRE(Dh == 32 || Dh == 16, "rpe_attn: only Dh == 32 and Dh == 16 are built (got %d)", Dh);
  BEVR_REQUIRE(BH > 0 && h > 0 && hpg > 0 && BH % h == 0 && h % hpg == 0, "rpe_attn: bad BH=%d h=%d hpg=%d", BH, h, hpg);
  BEVR_REQUIRE(Hq > 1 && Wq > 1 && Ns > 0 && Rh > 0 && Rw > 0, "rpe_attn: bad sizes Hq=%d Wq=%d Ns=%d Rh=%d Rw=%d",
               Hq, Wq, Ns, Rh, Rw);
  BEVR_REQUIRE(BH <= 65535, "rpe_attn: B*h=%d exceeds gridDim.y", BH);
  return 0;
}

}  // namespace bevr

using namespace bevr;

static size_t pad_floats(int h, int Rh, int Rw) { return (size_t)h * (Rh + 4) * (Rw + 4); }
static int grad_copies(int h, int Rh, int Rw) {     // up to 16 private gradient tables within 64 MB
  const size_t bytes = pad_floats(h, Rh, Rw) * sizeof(float);
  size_t c = ((size_t)64 << 20) / bytes;
  return c < 1 ? 1 : (c > 16 ? 16 : (int)c);
}

// Split of a kernel's streamed dimension (n_tiles tiles) so that about 3.5 waves of CTAs exist: `base` CTAs without
// splitting, `per_sm` CTAs resident per SM (register-limited: 4 forward, 3 dQ, 2 dK/dV).  Returns the number of ranges
// and sets tiles_per_split.  Deterministic in the shapes, so the workspace query and the launch agree.
static int pick_split(long long base, int per_sm, int n_tiles, int* tiles_per_split) {
  long long want = (7ll * sm_count() * per_sm / 2 + base - 1) / base;
  if (want > 8) want = 8;
  if (want > n_tiles) want = n_tiles;
  if (want < 1) want = 1;
  const int tps = (n_tiles + (int)want - 1) / (int)want;
  *tiles_per_split = tps;
  return (n_tiles + tps - 1) / tps;
}
// Which kernel family serves a call: the mma.sync kernels win from the reference-native shape (56x56) upwards; at very
// small grids (28x28: 0.6 M (query, key) pairs per head) their fixed costs (operand pre-split launches, fragment loads of
// 2-tile CTAs) lose to the FFMA2 kernels.  The per-call flags BEVR_K3_FFMA / BEVR_K3_MMA force one family (tests, benchmarks).
static bool use_mma(int Nq, int Ns, int flags) {
  if (flags & BEVR_K3_MMA) return true;
  if (flags & BEVR_K3_FFMA) return false;
  return (long long)Nq * Ns >= (3ll << 20);
}

static int fwd_split(int BH, int Nq, int Ns, int flags, int* tps) {
  if (K3_FWD_MMA && use_mma(Nq, Ns, flags)) return pick_split((long long)((Nq + TQM - 1) / TQM) * BH, K3_MMA_MINB, (Ns + TK - 1) / TK, tps);
  return pick_split((long long)((Nq + TQ - 1) / TQ) * BH, FWD_MINB, (Ns + TK - 1) / TK, tps);
}

static size_t round64(int n) { return ((size_t)n + 63) / 64 * 64; }

// workspace layout (floats): [padded table][backward: gcopies padded gradient tables][forward: key-range partials]
//                            [K hi, K lo, V hi, V lo : BH * NsP * Dh each][backward: Q hi, Q lo, dO hi, dO lo : BH * NqP * Dh each]
// The size covers BOTH kernel families (their forward key splits differ), so it does not depend on the flags of the call.
extern "C" size_t bevr_rpe_attn_workspace_bytes(int BH, int h, int Dh, int Nq, int Ns, int Rh, int Rw, int backward) {
  size_t n = pad_floats(h, Rh, Rw) * (backward ? 1 + grad_copies(h, Rh, Rw) : 1);
  if (!backward) {
    int tps;
    const int k0 = fwd_split(BH, Nq, Ns, BEVR_K3_FFMA, &tps), k1 = fwd_split(BH, Nq, Ns, BEVR_K3_MMA, &tps);
    const int ks = k0 > k1 ? k0 : k1;
    if (ks > 1) n += (size_t)ks * BH * (Dh + 2) * Nq;
  }
  n += 4 * (size_t)BH * round64(Ns) * Dh;
  if (backward) n += 4 * (size_t)BH * round64(Nq) * Dh;
  return n * sizeof(float) + 256;          // slack: the pre-split operands start on a 256-byte boundary
}

template <int DH>
static int presplit(const float* x, float* hi, float* lo, int BH, int n, int nP, float mul, cudaStream_t st) {
  rpe_presplit_kernel<DH><<<dim3(nP / 64, BH), 256, 0, st>>>(x, hi, lo, n, nP, mul);
  return check_launch("rpe_attn presplit");
}
// K and V of a call -> hi / lo row-major copies at `base`; returns the first float after them
static float* presplit_kv(RpeArgs& a, const float* k, const float* v, float* base, int Dh, cudaStream_t st, int* rc) {
  base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(base) + 255) & ~(uintptr_t)255);   // 16-byte async copies
  a.NsP = (int)round64(a.Ns);
  const size_t n = (size_t)a.BH * a.NsP * Dh;
  float *kh = base, *kl = base + n, *vh = base + 2 * n, *vl = base + 3 * n;
  a.kh = kh; a.kl = kl; a.vh = vh; a.vl = vl;
  *rc = Dh == 32 ? presplit<32>(k, kh, kl, a.BH, a.Ns, a.NsP, 1.f, st) : presplit<16>(k, kh, kl, a.BH, a.Ns, a.NsP, 1.f, st);
  if (*rc) return nullptr;
  *rc = Dh == 32 ? presplit<32>(v, vh, vl, a.BH, a.Ns, a.NsP, 1.f, st) : presplit<16>(v, vh, vl, a.BH, a.Ns, a.NsP, 1.f, st);
  return base + 4 * n;
}

// launch KERN<Dh, dropout?> (Dh = 32 or 16)
#define K3_LAUNCH(KERN, GRID, BLOCK)                                   \
  do {                                                                 \
    if (Dh == 32) {                                                    \
      if (a.drop_thresh) KERN<32, true><<<GRID, BLOCK, 0, st>>>(a);    \
      else KERN<32, false><<<GRID, BLOCK, 0, st>>>(a);                 \
    } else {                                                           \
      if (a.drop_thresh) KERN<16, true><<<GRID, BLOCK, 0, st>>>(a);    \
      else KERN<16, false><<<GRID, BLOCK, 0, st>>>(a);                 \
    }                                                                  \
  } while (0)

static int set_dropout(RpeArgs& a, float drop_p, unsigned long long seed) {
  BEVR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "rpe_attn: drop_p must be in [0, 1) (got %g)", (double)drop_p);
  a.drop_thresh = (unsigned)(drop_p * 16777216.0f);       // p < 2^-24 rounds to "no dropout"
  a.drop_scale = 1.f / (1.f - drop_p);
  a.seed_lo = (unsigned)(seed & 0xffffffffull); a.seed_hi = (unsigned)(seed >> 32);
  return 0;
}

static int rpe_fwd_impl(const float* q, const float* k, const float* v, const float* rpe, const float* pos,
                                 float* out, float* lse, void* workspace, int BH, int h, int hpg, int Dh, int Hq,
                                 int Wq, int Ns, int Rh, int Rw, float scale, int flags, float drop_p,
                                 unsigned long long seed, void* stream) {
  if (int rc = validate(BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw)) return rc;
  BEVR_REQUIRE(q && k && v && rpe && pos && out && lse && workspace, "rpe_attn_fwd: NULL pointer argument");
  RpeArgs a = {};
  if (int rc = set_dropout(a, drop_p, seed)) return rc;
  a.q = q; a.k = k; a.v = v; a.rpe_pad = (const float*)workspace; a.pos = pos; a.o_out = out; a.o_lse = lse;
  a.BH = BH; a.h = h; a.hpg = hpg; a.Hq = Hq; a.Wq = Wq; a.Nq = Hq * Wq; a.Ns = Ns; a.Rh = Rh; a.Rw = Rw;
  a.Hp = Rh + 4; a.Wp = Rw + 4;
  a.scale = scale;
  a.split = fwd_split(BH, a.Nq, Ns, flags, &a.tiles_per_split);
  a.part_o = (float*)workspace + pad_floats(h, Rh, Rw);
  a.part_ml = a.part_o + (size_t)a.split * BH * Dh * a.Nq;
  cudaStream_t st = (cudaStream_t)stream;
  rpe_pad_kernel<<<sm_count(), 256, 0, st>>>(rpe, (float*)workspace, h, Rh, Rw);
  if (int rc = check_launch("rpe_attn_fwd pad")) return rc;
  const bool mma = K3_FWD_MMA && use_mma(a.Nq, Ns, flags);
  if (mma) {
    int rc = 0;
    presplit_kv(a, k, v, a.part_o + (a.split > 1 ? (size_t)a.split * BH * (Dh + 2) * a.Nq : 0), Dh, st, &rc);
    if (rc) return rc;
  }
  if (mma) {
    dim3 grid((a.Nq + TQM - 1) / TQM, BH, a.split);
#if K3_FWD_Q_SMEM
    {
      const size_t dyn = 4 * (size_t)(Dh / 8) * 32 * sizeof(float4);
      auto launch_fwd = [&](auto kern) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "rpe_attn_fwd shared memory")) return rc;
        kern<<<grid, 128, dyn, st>>>(a);
        return 0;
      };
      int rc;
      if (Dh == 32) rc = a.drop_thresh ? launch_fwd(rpe_attn_fwd_mma_kernel<32, true>) : launch_fwd(rpe_attn_fwd_mma_kernel<32, false>);
      else rc = a.drop_thresh ? launch_fwd(rpe_attn_fwd_mma_kernel<16, true>) : launch_fwd(rpe_attn_fwd_mma_kernel<16, false>);
      if (rc) return rc;
    }
#else
    K3_LAUNCH(rpe_attn_fwd_mma_kernel, grid, 128);
#endif
  } else {
    dim3 grid((a.Nq + TQ - 1) / TQ, BH, a.split);
    K3_LAUNCH(rpe_attn_fwd_kernel, grid, TQ / RQ);
  }
  if (int rc = check_launch("rpe_attn_fwd")) return rc;
  if (a.split > 1) {
    dim3 cgrid((a.Nq + 127) / 128, BH);
    if (Dh == 32) rpe_combine_kernel<32><<<cgrid, 128, 0, st>>>(a);
    else rpe_combine_kernel<16><<<cgrid, 128, 0, st>>>(a);
    return check_launch("rpe_attn_fwd combine");
  }
  return 0;
}

extern "C" int bevr_rpe_attn_fwd(const float* q, const float* k, const float* v, const float* rpe, const float* pos,
                                 float* out, float* lse, void* workspace, int BH, int h, int hpg, int Dh, int Hq,
                                 int Wq, int Ns, int Rh, int Rw, float scale, int flags, void* stream) {
  return rpe_fwd_impl(q, k, v, rpe, pos, out, lse, workspace, BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, scale, flags, 0.f, 0ull, stream);
}
extern "C" int bevr_rpe_attn_fwd_dropout(const float* q, const float* k, const float* v, const float* rpe,
                                         const float* pos, float* out, float* lse, void* workspace, int BH, int h,
                                         int hpg, int Dh, int Hq, int Wq, int Ns, int Rh, int Rw, float scale, int flags,
                                         float drop_p, unsigned long long seed, void* stream) {
  return rpe_fwd_impl(q, k, v, rpe, pos, out, lse, workspace, BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, scale, flags, drop_p, seed, stream);
}

static int rpe_bwd_impl(const float* q, const float* k, const float* v, const float* rpe, const float* pos,
                                 const float* out, const float* lse, const float* grad_out, float* grad_q,
                                 float* grad_k, float* grad_v, float* grad_rpe, float* grad_pos, float* delta_ws,
                                 void* workspace, int BH, int h, int hpg, int Dh, int Hq, int Wq, int Ns, int Rh,
                                 int Rw, float scale, int flags, float drop_p, unsigned long long seed, void* stream) {
  if (int rc = validate(BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw)) return rc;
  BEVR_REQUIRE(q && k && v && rpe && pos && out && lse && grad_out && delta_ws && workspace,
               "rpe_attn_bwd: NULL pointer argument");
  RpeArgs a = {};
  if (int rc = set_dropout(a, drop_p, seed)) return rc;
  float* pad = (float*)workspace;
  float* gpad = pad + pad_floats(h, Rh, Rw);
  a.q = q; a.k = k; a.v = v; a.rpe_pad = pad; a.pos = pos; a.out = out; a.lse = lse; a.grad_out = grad_out;
  a.delta = delta_ws; a.o_delta = delta_ws;
  a.grad_q = grad_q; a.grad_k = grad_k; a.grad_v = grad_v; a.grad_rpe_pad = grad_rpe ? gpad : nullptr;
  a.gcopies = grad_copies(h, Rh, Rw);
  a.grad_pos = grad_pos;
  a.BH = BH; a.h = h; a.hpg = hpg; a.Hq = Hq; a.Wq = Wq; a.Nq = Hq * Wq; a.Ns = Ns; a.Rh = Rh; a.Rw = Rw;
  a.Hp = Rh + 4; a.Wp = Rw + 4;
  a.scale = scale;
  cudaStream_t st = (cudaStream_t)stream;
  rpe_pad_kernel<<<sm_count(), 256, 0, st>>>(rpe, pad, h, Rh, Rw);
  if (int rc = check_launch("rpe_attn_bwd pad")) return rc;
  if (grad_rpe)
    if (int rc = check_cuda(cudaMemsetAsync(gpad, 0, pad_floats(h, Rh, Rw) * sizeof(float) * a.gcopies, st), "rpe_attn_bwd zero-fill")) return rc;
  const bool mma_dq = K3_DQ_MMA && use_mma(a.Nq, Ns, flags), mma_dkv = K3_DKV_MMA && use_mma(a.Nq, Ns, flags);
  if (mma_dq || mma_dkv) {
    int rc = 0;
    float* nxt = presplit_kv(a, k, v, gpad + pad_floats(h, Rh, Rw) * a.gcopies, Dh, st, &rc);
    if (rc) return rc;
    a.NqP = (int)round64(a.Nq);
    const size_t n = (size_t)BH * a.NqP * Dh;
    float *qh = nxt, *ql = nxt + n, *gh = nxt + 2 * n, *gl = nxt + 3 * n;
    a.qh = qh; a.ql = ql; a.gh = gh; a.gl = gl;
    rc = Dh == 32 ? presplit<32>(q, qh, ql, BH, a.Nq, a.NqP, scale, st) : presplit<16>(q, qh, ql, BH, a.Nq, a.NqP, scale, st);
    if (rc) return rc;
    rc = Dh == 32 ? presplit<32>(grad_out, gh, gl, BH, a.Nq, a.NqP, 1.f, st)
                  : presplit<16>(grad_out, gh, gl, BH, a.Nq, a.NqP, 1.f, st);
    if (rc) return rc;
  }
  if (Dh == 32) rpe_delta_kernel<32><<<dim3((a.Nq + 255) / 256, BH), 256, 0, st>>>(a);
  else rpe_delta_kernel<16><<<dim3((a.Nq + 255) / 256, BH), 256, 0, st>>>(a);
  if (int rc = check_launch("rpe_attn_bwd delta")) return rc;
  if (grad_q) {
    const int tq = mma_dq ? TQM : TQ;
    const int nqt = (a.Nq + tq - 1) / tq;
    a.split = pick_split((long long)nqt * BH, mma_dq ? K3_DQ_MMA_MINB : K3_DQ_MINB, (Ns + TK - 1) / TK, &a.tiles_per_split);
    if (a.split > 1)
      if (int rc = check_cuda(cudaMemsetAsync(grad_q, 0, (size_t)BH * Dh * a.Nq * sizeof(float), st), "rpe_attn_bwd zero-fill")) return rc;
    if (mma_dq) {
#if K3_DQ_QG_SMEM
      const size_t dyn = 2 * 4 * (size_t)(Dh / 8) * 32 * sizeof(float4);      // Q and dO fragments of the CTA's 4 warps
      auto launch_dq = [&](auto kern) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "rpe_attn_bwd dq shared memory")) return rc;
        kern<<<dim3(nqt, BH, a.split), 128, dyn, st>>>(a);
        return 0;
      };
      int rc;
      if (Dh == 32) rc = a.drop_thresh ? launch_dq(rpe_attn_bwd_dq_mma_kernel<32, true>) : launch_dq(rpe_attn_bwd_dq_mma_kernel<32, false>);
      else rc = a.drop_thresh ? launch_dq(rpe_attn_bwd_dq_mma_kernel<16, true>) : launch_dq(rpe_attn_bwd_dq_mma_kernel<16, false>);
      if (rc) return rc;
#else
      K3_LAUNCH(rpe_attn_bwd_dq_mma_kernel, dim3(nqt, BH, a.split), 128);
#endif
    }
    else K3_LAUNCH(rpe_attn_bwd_dq_kernel, dim3(nqt, BH, a.split), TQ);
    if (int rc = check_launch("rpe_attn_bwd dq")) return rc;
  }
  if (grad_k || grad_v || grad_rpe || grad_pos) {
    const int nkt = mma_dkv ? (Ns + TKM - 1) / TKM : (Ns + TKB - 1) / TKB;
    const int nqt = mma_dkv ? (a.Nq + TK - 1) / TK : (a.Nq + TQB - 1) / TQB;
    const int per_sm = mma_dkv ? K3_DKV_MMA_MINB : (DKV_MINB < 2 ? 2 : DKV_MINB);
    a.split = pick_split((long long)nkt * BH, per_sm, nqt, &a.tiles_per_split);
    if (a.split > 1) {
      if (grad_k)
        if (int rc = check_cuda(cudaMemsetAsync(grad_k, 0, (size_t)BH * Dh * Ns * sizeof(float), st), "rpe_attn_bwd zero-fill")) return rc;
      if (grad_v)
        if (int rc = check_cuda(cudaMemsetAsync(grad_v, 0, (size_t)BH * Dh * Ns * sizeof(float), st), "rpe_attn_bwd zero-fill")) return rc;
    }
    if (mma_dkv) {
#if K3_DKV_KV_SMEM
      const size_t dyn = 2 * 4 * (size_t)(Dh / 8) * 32 * sizeof(float4);      // K and V fragments of the CTA's 4 warps
      auto launch_dkv = [&](auto kern) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "rpe_attn_bwd dkv shared memory")) return rc;
        kern<<<dim3(nkt, BH, a.split), 128, dyn, st>>>(a);
        return 0;
      };
      int rc;
      const bool mg = (flags & BEVR_K3_RUN_MERGE) != 0;
      if (Dh == 32) {
        if (mg) rc = a.drop_thresh ? launch_dkv(rpe_attn_bwd_dkv_mma_kernel<32, true, true>) : launch_dkv(rpe_attn_bwd_dkv_mma_kernel<32, false, true>);
        else rc = a.drop_thresh ? launch_dkv(rpe_attn_bwd_dkv_mma_kernel<32, true, false>) : launch_dkv(rpe_attn_bwd_dkv_mma_kernel<32, false, false>);
      } else {
        if (mg) rc = a.drop_thresh ? launch_dkv(rpe_attn_bwd_dkv_mma_kernel<16, true, true>) : launch_dkv(rpe_attn_bwd_dkv_mma_kernel<16, false, true>);
        else rc = a.drop_thresh ? launch_dkv(rpe_attn_bwd_dkv_mma_kernel<16, true, false>) : launch_dkv(rpe_attn_bwd_dkv_mma_kernel<16, false, false>);
      }
      if (rc) return rc;
#else
      if (Dh == 32) { if (a.drop_thresh) rpe_attn_bwd_dkv_mma_kernel<32, true, true><<<dim3(nkt, BH, a.split), 128, 0, st>>>(a); else rpe_attn_bwd_dkv_mma_kernel<32, false, true><<<dim3(nkt, BH, a.split), 128, 0, st>>>(a); }
      else { if (a.drop_thresh) rpe_attn_bwd_dkv_mma_kernel<16, true, true><<<dim3(nkt, BH, a.split), 128, 0, st>>>(a); else rpe_attn_bwd_dkv_mma_kernel<16, false, true><<<dim3(nkt, BH, a.split), 128, 0, st>>>(a); }
#endif
    }
    else K3_LAUNCH(rpe_attn_bwd_dkv_kernel, dim3(nkt, BH, a.split), TKB * CS);
    if (int rc = check_launch("rpe_attn_bwd dkv")) return rc;
  }
  if (grad_rpe) {
    rpe_unpad_add_kernel<<<sm_count(), 256, 0, st>>>(gpad, grad_rpe, h, Rh, Rw, a.gcopies);
    if (int rc = check_launch("rpe_attn_bwd unpad")) return rc;
  }
  return 0;
}

extern "C" int bevr_rpe_attn_bwd(const float* q, const float* k, const float* v, const float* rpe, const float* pos,
                                 const float* out, const float* lse, const float* grad_out, float* grad_q,
                                 float* grad_k, float* grad_v, float* grad_rpe, float* grad_pos, float* delta_ws,
                                 void* workspace, int BH, int h, int hpg, int Dh, int Hq, int Wq, int Ns, int Rh,
                                 int Rw, float scale, int flags, void* stream) {
  return rpe_bwd_impl(q, k, v, rpe, pos, out, lse, grad_out, grad_q, grad_k, grad_v, grad_rpe, grad_pos, delta_ws, workspace,
                      BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, scale, flags, 0.f, 0ull, stream);
}
extern "C" int bevr_rpe_attn_bwd_dropout(const float* q, const float* k, const float* v, const float* rpe,
                                         const float* pos, const float* out, const float* lse, const float* grad_out,
                                         float* grad_q, float* grad_k, float* grad_v, float* grad_rpe, float* grad_pos,
                                         float* delta_ws, void* workspace, int BH, int h, int hpg, int Dh, int Hq, int Wq,
                                         int Ns, int Rh, int Rw, float scale, int flags, float drop_p,
                                         unsigned long long seed, void* stream) {
  return rpe_bwd_impl(q, k, v, rpe, pos, out, lse, grad_out, grad_q, grad_k, grad_v, grad_rpe, grad_pos, delta_ws, workspace,
                      BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, scale, flags, drop_p, seed, stream);
}
