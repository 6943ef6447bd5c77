// K5: rotation-sweep correlation of a BEV template against an aerial-map tile (sm_100a, tcgen05).
//
//   out[r, y, x] = sum_{c,u,v} map[c, y+u, x+v] * rot_r(tmpl)[c, u, v]        (valid mode)
//
// The reference has no registration code: the operator is defined by this repository
// (oracle/correlate.py) and its parity is "unpinned vs reference".
//
// This is the one dense contraction of the path, so it runs on the 5th-generation tensor cores:
// an implicit GEMM  D[m, n] += A[m, k] B[n, k]  with  m = 128 consecutive output columns x of one output row y,
// n = rotation r (padded to a multiple of 16), k = template column v, accumulated over (c, u) -- bf16 operands,
// fp32 accumulators in TMEM, issued by one thread with tcgen05.mma, read back with tcgen05.ld.
//
// The im2col operand is never materialised.  A[m, k] = map[c, y+u, x0+m+k] is a Hankel matrix; a shared-memory
// (UMMA) descriptor only needs 16-byte-aligned core-matrix rows, so the map is pre-expanded once into 8 copies
// shifted by 0..7 elements (S[s][row][x] = map[row][x+s]).  Staging the 16-byte blocks G[blk][s] =
// S[s][row][x0 + 8 blk ..] and giving the A descriptor LBO = SBO = 128 B makes row-group mb / k-chunk j read
// block (mb + j): the whole 128 x 224 operand of one (c, map row) is 44 blocks = 5.6 KB of shared memory, and
// the next 16-wide K step is the same tile 256 B further on.  One CTA owns 4 output rows (4 accumulators of
// 128 x NP fp32 in TMEM): every staged map row feeds up to 4 MMA groups (u = row - y), and every rotated-template
// slice B(c, u) is loaded once per CTA and kept in a 4-deep ring while the 4 rows that need it go by.
#include <math.h>

#include "bevr_common.cuh"

namespace bevr {

constexpr int CORR_ROWS = 4;       // max output rows (accumulators) per CTA
constexpr int CORR_M = 128;        // output columns per CTA = UMMA M
constexpr int CORR_XPAD = 128;     // zero columns appended to every shifted map row (A tiles over-read)

struct CorrArgs {
  const __nv_bfloat16* smap;   // (8, C, Hm, Wmp) shifted copies
  const __nv_bfloat16* btile;  // (C, Ht, Wt/8, NP/8, 8, 8) rotated templates in UMMA K-major core-matrix order
  float* out;                  // (R, Ho, Wo)
  int C, Hm, Wmp, Ht, Wt, R, NP, Ho, Wo;
  int rows;                    // output rows (TMEM accumulators) per CTA, <= CORR_ROWS
};

// ---- operand preparation -------------------------------------------------------------------------
__global__ void corr_shift_map_kernel(const float* __restrict__ map, __nv_bfloat16* __restrict__ smap, int C, int Hm,
                                      int Wm, int Wmp) {
  const size_t n = (size_t)8 * C * Hm * Wmp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wmp);
    const size_t row = (i / Wmp) % ((size_t)C * Hm);
    const int s = (int)(i / ((size_t)Wmp * C * Hm));
    const int xs = x + s;
    smap[i] = __float2bfloat16_rn(xs < Wm ? map[row * Wm + xs] : 0.f);
  }
}

// rot_r(tmpl)[c,u,v] = bilinear(tmpl[c], centre + R(theta_r) (v - cx, u - cy)), zeros outside
struct RotArgs {
  float cs[2 * 128];           // (cos, sin) per rotation
};
__global__ void corr_rotate_kernel(const float* __restrict__ tmpl, const __grid_constant__ RotArgs rot,
                                   __nv_bfloat16* __restrict__ btile, int C, int Ht, int Wt, int R, int NP) {
  const size_t n = (size_t)C * Ht * Wt * NP;
  const float cx = 0.5f * (float)(Wt - 1), cy = 0.5f * (float)(Ht - 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // decode the UMMA order: [c][u][v/8][r/8][r%8][v%8]
    const int v8 = (int)(i % 8), r8 = (int)((i / 8) % 8);
    const int rb = (int)((i / 64) % (NP / 8));
    const int vb = (int)((i / (64 * (size_t)(NP / 8))) % (Wt / 8));
    const size_t cu = i / (64 * (size_t)(NP / 8) * (Wt / 8));
    const int u = (int)(cu % Ht), c = (int)(cu / Ht);
    const int r = rb * 8 + r8, v = vb * 8 + v8;
    float val = 0.f;
    if (r < R) {
      const float co = rot.cs[2 * r], si = rot.cs[2 * r + 1];
      const float dx = (float)v - cx, dy = (float)u - cy;
      const float sx = cx + co * dx - si * dy, sy = cy + si * dx + co * dy;
      const float x0f = floorf(sx), y0f = floorf(sy);
      const float fx = sx - x0f, fy = sy - y0f;
      const int x0 = (int)x0f, y0 = (int)y0f;
      const float* t = tmpl + (size_t)c * Ht * Wt;
      auto at = [&](int yy, int xx) { return (yy >= 0 && yy < Ht && xx >= 0 && xx < Wt) ? t[yy * Wt + xx] : 0.f; };
      val = (1.f - fy) * ((1.f - fx) * at(y0, x0) + fx * at(y0, x0 + 1)) +
            fy * ((1.f - fx) * at(y0 + 1, x0) + fx * at(y0 + 1, x0 + 1));
    }
    btile[i] = __float2bfloat16_rn(val);
  }
}

// ---- tcgen05 / mbarrier primitives (PTX as in the vendored CUTLASS cute/arch/*sm100*.hpp) ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- main kernel ---------------------------------------------------------------------------------
// Warp roles: warp 0 / lane 0 issues the MMAs; warps 1-3 (96 threads) stage operands with cp.async.
// CORR_STAGES G stages + a (rows + CORR_STAGES - 1)-deep, K-chunk-interleaved ring of B slices; full[s] (loaders -> MMA, 96 arrivals) and empty[s]
// (tcgen05.commit -> loaders) mbarriers let the staging of map row i+1 overlap the MMAs of row i.
constexpr int CORR_LOADERS = 96;
constexpr int CORR_STAGES = 2;     // staged map rows in flight (a deeper pipeline measured no gain)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) corr_umma_kernel(const __grid_constant__ CorrArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nblk = CORR_M / 8 + a.Wt / 8;                   // 16-byte blocks of one staged map row, per shift
  const int g_bytes = (nblk * 128 + 127) & ~127;
  const int b_bytes = (a.Wt / 8) * a.NP * 16;               // one rotated-template slice B(c, u)
  const int ring = a.rows + CORR_STAGES - 1;
  uint8_t* sG = smem;                                       // CORR_STAGES stages
  uint8_t* sB = smem + CORR_STAGES * g_bytes;               // `ring` slices
  __shared__ __align__(8) uint64_t s_full[CORR_STAGES], s_empty[CORR_STAGES], s_done;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int x0 = blockIdx.x * CORR_M, y0 = blockIdx.y * a.rows;

  if (tid == 0) {
    for (int i = 0; i < CORR_STAGES; ++i) { mbar_init(smem_u32(&s_full[i]), CORR_LOADERS); mbar_init(smem_u32(&s_empty[i]), 1); }
    mbar_init(smem_u32(&s_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: `rows` accumulators of NP fp32 columns (<= 512 columns)
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int vrows = min(a.rows, a.Ho - y0);                  // valid output rows of this CTA
  const int rows_c = vrows + a.Ht - 1;                       // map rows touched per channel
  const int n_it = a.C * rows_c;

  if (warp > 0) {
    // ================= loaders =================
    const int lt = tid - 32;
    const size_t plane = (size_t)a.C * a.Hm * a.Wmp;        // elements of one shifted copy
    int c = 0, rr = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % CORR_STAGES, k = it / CORR_STAGES;
      if (k >= 1) mbar_wait(smem_u32(&s_empty[s]), (uint32_t)(k - 1) & 1u);   // MMAs of iteration it-CORR_STAGES retired
      const __nv_bfloat16* rowp = a.smap + ((size_t)c * a.Hm + (y0 + rr)) * a.Wmp + x0;
      const uint32_t g0 = smem_u32(sG + s * g_bytes);
      for (int i = lt; i < nblk * 8; i += CORR_LOADERS)     // G[blk][shift], 16 B each
        cp_async16(g0 + i * 16, rowp + (size_t)(i & 7) * plane + 8 * (i >> 3));
      if (rr < a.Ht) {                                       // a new template row u = rr enters the ring
        // slice t goes to ring position q(t) = -t mod ring (positions DEscend with u, so the slices of output rows
        // y, y+1, y+2 -- u, u-1, u-2 -- are consecutive), interleaved by K chunk: [k8][position][NP rows x 16 B]
        const int t = c * a.Ht + rr;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.btile) + (size_t)t * b_bytes;
        const int q = (ring - t % ring) % ring;
        const uint32_t d0 = smem_u32(sB) + q * a.NP * 16;
        if (lt < 8) {
          // eight threads move the slice with 1-D bulk copies (TMA engine, async proxy): Wt/8 runs of NP*16 B,
          // completing on full[s] through its transaction count -- no LDGSTS traffic, no proxy fence for B.
          // (the phase cannot complete early: the 96 loader arrivals are still pending)
          const uint32_t bar = smem_u32(&s_full[s]);
          if (lt == 0)
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"((uint32_t)b_bytes) : "memory");
          const uint32_t run = (uint32_t)a.NP * 16;
          for (int k8 = lt; k8 < a.Wt / 8; k8 += 8)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d0 + k8 * ring * run), "l"(src + (size_t)k8 * run), "r"(run), "r"(bar) : "memory");
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_async_smem();          // generic-proxy writes -> visible to the tensor-core (async) proxy
      mbar_arrive(smem_u32(&s_full[s]));
      if (++rr == rows_c) { rr = 0; ++c; }
    }
  } else if (tid == 0) {
    // ================= MMA issuer =================
    // instruction descriptor: D = f32, A = B = bf16, K-major both, M = 128, N = n
    auto idesc_n = [](int n) {
      return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(CORR_M >> 4) << 24);
    };
    const int ksteps = a.Wt / 16;
    const uint32_t lbo = (uint32_t)(ring * a.NP * 16);       // K-chunk stride of the interleaved B ring
    const uint32_t bring = smem_u32(sB);
    uint32_t started = 0;                                    // bit yy: accumulator yy has been written once
    int c = 0, rr = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % CORR_STAGES, k = it / CORR_STAGES;
      mbar_wait(smem_u32(&s_full[s]), (uint32_t)k & 1u);
      tc_fence_after();
      const uint32_t gbase = smem_u32(sG + s * g_bytes);
      // output rows fed by this map row: yy in [ya, yb] <-> template rows u = rr - yy (descending), whose slices
      // sit at consecutive ring positions.  One MMA per K step covers ALL of them (N = rows * NP): the measured
      // cost of an MMA is ~75 + 0.53 N cycles, so merging the rows amortises the fixed part (A-operand fetch).
      const int ya = max(0, rr - (a.Ht - 1)), yb = min(vrows - 1, rr);
      const int nwin = yb - ya + 1;
      const uint32_t need = ((1u << nwin) - 1u) << ya;
      if ((started & need) != need) {
        // start-up (first rows of channel 0): accumulators not initialised yet -> one MMA per row, overwrite first
        for (int yy = ya; yy <= yb; ++yy) {
          const int q = (ring - (c * a.Ht + rr - yy) % ring) % ring;
          for (int ks = 0; ks < ksteps; ++ks) {
            umma_bf16(tmem + yy * a.NP, umma_desc(gbase + ks * 256, 128, 128),
                      umma_desc(bring + q * a.NP * 16 + ks * 2 * lbo, lbo, 128), idesc_n(a.NP), (started >> yy) & 1u);
            started |= 1u << yy;
          }
        }
      } else {
        const int q = (ring - (c * a.Ht + rr - ya) % ring) % ring;      // position of the slice of row ya
        const int n1 = min(nwin, ring - q), n2 = nwin - n1;             // the window may wrap around the ring once
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = umma_desc(gbase + ks * 256, 128, 128);
          umma_bf16(tmem + ya * a.NP, da, umma_desc(bring + q * a.NP * 16 + ks * 2 * lbo, lbo, 128), idesc_n(n1 * a.NP), 1u);
          if (n2 > 0)
            umma_bf16(tmem + (ya + n1) * a.NP, da, umma_desc(bring + ks * 2 * lbo, lbo, 128), idesc_n(n2 * a.NP), 1u);
        }
      }
      umma_commit(smem_u32(&s_empty[s]));   // arrives when the MMAs issued so far have finished reading smem
      if (++rr == rows_c) { rr = 0; ++c; }
    }
    umma_commit(smem_u32(&s_done));
  }
  __syncwarp();
  mbar_wait(smem_u32(&s_done), 0);

  // ---- epilogue: TMEM -> registers -> out[r][y][x]
  tc_fence_after();
  const int m = tid;                                          // TMEM lane = output column x0 + m
  for (int yy = 0; yy < vrows; ++yy) {
    const int y = y0 + yy;
    for (int n0 = 0; n0 < a.NP; n0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(yy * a.NP + n0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (x0 + m < a.Wo) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + j < a.R) a.out[((size_t)(n0 + j) * a.Ho + y) * a.Wo + x0 + m] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static size_t corr_smem(int Wt, int NP, int rows) {
  const int nblk = CORR_M / 8 + Wt / 8;
  return (size_t)CORR_STAGES * ((nblk * 128 + 127) & ~127) + (size_t)(rows + CORR_STAGES - 1) * (Wt / 8) * NP * 16;
}

static int corr_dims(int C, int Hm, int Wm, int Ht, int Wt, int R, int& NP, int& Wmp, int& rows, size_t& smem) {
  BEVR_REQUIRE(C > 0 && Hm >= Ht && Wm >= Wt && Ht > 0 && Wt > 0 && R > 0, "correlate: bad sizes");
  BEVR_REQUIRE(Wt % 16 == 0, "correlate: template width must be a multiple of 16 (UMMA K step), got %d", Wt);
  NP = (R + 15) / 16 * 16;
  BEVR_REQUIRE(NP * CORR_ROWS <= 512, "correlate: at most %d rotations per call (TMEM columns), got %d", 512 / CORR_ROWS, R);
  const int wo = Wm - Wt + 1, ho = Hm - Ht + 1;
  const int need = ((wo + CORR_M - 1) / CORR_M) * CORR_M + Wt + 8;        // furthest column an A tile reads
  Wmp = ((need > Wm + CORR_XPAD ? need : Wm + CORR_XPAD) + 7) / 8 * 8;
  // rows per CTA: fewer rows = more, cheaper CTAs (better wave balance on 148 SMs) but more B-slice traffic;
  // pick the count with the smallest makespan ceil(CTAs / 148) * rows among those whose staging fits in smem
  const int xt = (wo + CORR_M - 1) / CORR_M;
  rows = 0;
  long best_cost = -1;
  for (int r = CORR_ROWS; r >= 1; --r) {
    if (corr_smem(Wt, NP, r) > 225 * 1024 || (r > 1 && r * NP > 256)) continue;   // merged MMA: N = r * NP <= 256
    const long ctas = (long)xt * ((ho + r - 1) / r);
    const long cost = ((ctas + sm_count() - 1) / sm_count()) * r;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; rows = r; }
  }
  BEVR_REQUIRE(rows > 0, "correlate: template %d wide x %d rotations does not fit in shared memory", Wt, R);
  smem = corr_smem(Wt, NP, rows);
  return 0;
}

}  // namespace bevr

using namespace bevr;

extern "C" size_t bevr_correlate_workspace_bytes(int C, int Hm, int Wm, int Ht, int Wt, int R) {
  int NP, Wmp, rows;
  size_t smem;
  if (corr_dims(C, Hm, Wm, Ht, Wt, R, NP, Wmp, rows, smem)) return 0;
  return (size_t)8 * C * Hm * Wmp * 2 + (size_t)C * Ht * Wt * NP * 2 + 256;
}

extern "C" int bevr_correlate(const float* map, const float* tmpl, const float* angles_rad, float* out, void* workspace,
                              int C, int Hm, int Wm, int Ht, int Wt, int R, void* stream) {
  int NP, Wmp, rows;
  size_t smem;
  if (int rc = corr_dims(C, Hm, Wm, Ht, Wt, R, NP, Wmp, rows, smem)) return rc;
  BEVR_REQUIRE(map && tmpl && angles_rad && out && workspace, "correlate: NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  __nv_bfloat16* smap = (__nv_bfloat16*)ws;
  const size_t smap_bytes = ((size_t)8 * C * Hm * Wmp * 2 + 127) & ~(size_t)127;
  __nv_bfloat16* btile = (__nv_bfloat16*)(ws + smap_bytes);
  RotArgs rot;
  for (int r = 0; r < R; ++r) { rot.cs[2 * r] = cosf(angles_rad[r]); rot.cs[2 * r + 1] = sinf(angles_rad[r]); }
  corr_shift_map_kernel<<<sm_count() * 4, 256, 0, st>>>(map, smap, C, Hm, Wm, Wmp);
  if (int rc = check_launch("correlate shift_map")) return rc;
  corr_rotate_kernel<<<sm_count() * 4, 256, 0, st>>>(tmpl, rot, btile, C, Ht, Wt, R, NP);
  if (int rc = check_launch("correlate rotate")) return rc;
  CorrArgs a;
  a.smap = smap; a.btile = btile; a.out = out;
  a.C = C; a.Hm = Hm; a.Wmp = Wmp; a.Ht = Ht; a.Wt = Wt; a.R = R; a.NP = NP;
  a.Ho = Hm - Ht + 1; a.Wo = Wm - Wt + 1;
  const int xt = (a.Wo + CORR_M - 1) / CORR_M;
  a.rows = rows;
  cudaFuncSetAttribute(corr_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(xt, (a.Ho + a.rows - 1) / a.rows);
  corr_umma_kernel<<<grid, 128, smem, st>>>(a);
  return check_launch("correlate");
}
