// Experiment: does the MSDA forward gather get closer to the all-hit L1 rate (1.14 cycles/tap) when the taps that are
// live on one SM share ONE window?  Product kernel today: 5-6 independent 8x8-query CTAs per SM (unrelated tiles),
// L1 hit rate 77 %, 1.5 cycles/tap in the pure-gather microbenchmark (exp_gather_width).
// Variants: CTA = TW x TH queries of one head with THREADS threads, MINB CTAs/SM; PERSIST = grid of 148*MINB CTAs, each
// walking a contiguous range of tiles in serpentine order (neighbouring tiles overlap in every level's window).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_build/exp_gather_tiles scripts/exp_gather_tiles.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int B = 8, NH = 8, NPT = 16, QW = 200, S = 53125;
__constant__ int c_start[4] = {0, 40000, 50000, 52500};

__device__ __forceinline__ void acc_add(float* a, float w, uint4 u) {
  a[0] = fmaf(w, __uint_as_float(u.x), a[0]); a[1] = fmaf(w, __uint_as_float(u.y), a[1]);
  a[2] = fmaf(w, __uint_as_float(u.z), a[2]); a[3] = fmaf(w, __uint_as_float(u.w), a[3]);
}

template <int TW, int TH, int THREADS>
__device__ __forceinline__ void do_tile(const char* __restrict__ V, float* __restrict__ out, int b, int head, int tx, int ty) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, lig = lane & 7;
  const char* vbase = V + (((size_t)b * S * NH + head) * 32 + lig * 4) * 4;
  for (int base = warp * 4; base < TW * TH; base += (THREADS / 32) * 4) {
    const int item = base + g;
    const int qx = tx * TW + (item % TW), qy = ty * TH + (item / TW);
    if (qx >= QW || qy >= QW) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned r = (unsigned)((qy * QW + qx) * 2654435761u) ^ (unsigned)(b * NH + head) * 40503u;
#pragma unroll 4
    for (int p = 0; p < NPT; ++p) {
      const int l = p >> 2, Wl = QW >> l;
      r = r * 1664525u + 1013904223u;
      const int ox = (int)((r >> 8) & 7) + (int)((r >> 11) & 3) - 5, oy = (int)((r >> 16) & 7) + (int)((r >> 19) & 3) - 5;
      const int x0 = min(max((qx >> l) + ox, 0), Wl - 2), y0 = min(max((qy >> l) + oy, 0), Wl - 2);
      const unsigned o0 = (unsigned)(c_start[l] + y0 * Wl + x0) * 1024u, dy = 1024u * Wl;
      const uint4 u0 = __ldg((const uint4*)(vbase + o0)), u1 = __ldg((const uint4*)(vbase + o0 + 1024u)),
                  u2 = __ldg((const uint4*)(vbase + o0 + dy)), u3 = __ldg((const uint4*)(vbase + o0 + dy + 1024u));
      acc_add(acc, 0.25f, u0); acc_add(acc, 0.35f, u1); acc_add(acc, 0.15f, u2); acc_add(acc, 0.25f, u3);
    }
    *(float4*)(out + ((((size_t)b * QW * QW + qy * QW + qx) * NH + head) * 32 + lig * 4)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

template <int TW, int TH, int THREADS, int MINB, int PERSIST, int SYNC>
__global__ void __launch_bounds__(THREADS, MINB) k_gather(const char* __restrict__ V, float* __restrict__ out) {
  constexpr int NTX = (QW + TW - 1) / TW, NTY = (QW + TH - 1) / TH, NT = NTX * NTY;
  if (!PERSIST) {
    int blk = blockIdx.x;
    const int tile = blk % NT; blk /= NT;
    do_tile<TW, TH, THREADS>(V, out, blk / NH, blk % NH, tile % NTX, tile / NTX);
  } else {
    const int total = B * NH * NT;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * per, t1 = min(t0 + per, total);
    for (int t = t0; t < t1; ++t) {
      const int tile = t % NT, bh = t / NT;
      const int ty = tile / NTX;
      int tx = tile % NTX;
      if (ty & 1) tx = NTX - 1 - tx;   // serpentine
      do_tile<TW, TH, THREADS>(V, out, bh / NH, bh % NH, tx, ty);
      if (SYNC) __syncthreads();
    }
  }
}

template <typename F> static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  const size_t nv = (size_t)B * S * NH * 32, no = (size_t)B * QW * QW * NH * 32;
  float *V, *out;
  CK(cudaMalloc(&V, nv * 4)); CK(cudaMalloc(&out, no * 4));
  CK(cudaMemset(V, 0, nv * 4));
  const double taps = (double)B * NH * QW * QW * NPT * 4;
#define RUN(TW, TH, THREADS, MINB, PERSIST, SYNC)                                                                    \
  {                                                                                                                  \
    const int nt = ((QW + TW - 1) / TW) * ((QW + TH - 1) / TH);                                                      \
    const int grid = PERSIST ? 148 * MINB : B * NH * nt;                                                             \
    const float ms = time_ms([&] { k_gather<TW, TH, THREADS, MINB, PERSIST, SYNC><<<grid, THREADS>>>((const char*)V, out); }, 10); \
    printf("tile %2dx%-2d threads %4d ctas/SM %d persist %d sync %d: %.3f ms  %.2f cyc/tap/SM\n", TW, TH, THREADS, MINB, \
           PERSIST, SYNC, ms, ms * 1e-3 * 1.965e9 * 148 / taps);                                                      \
  }
  RUN(8, 8, 256, 4, 0, 0)
  RUN(8, 8, 256, 8, 0, 0)
  RUN(16, 16, 1024, 1, 0, 0)
  RUN(16, 16, 1024, 2, 0, 0)
  RUN(16, 8, 512, 1, 0, 0)
  RUN(16, 8, 512, 2, 0, 0)
  RUN(8, 8, 256, 2, 0, 0)
  RUN(8, 8, 256, 1, 0, 0)
  RUN(16, 16, 1024, 1, 1, 0)
  RUN(16, 16, 1024, 1, 1, 1)
  RUN(16, 8, 512, 2, 1, 0)
  RUN(16, 8, 512, 1, 1, 1)
  RUN(8, 8, 256, 4, 1, 0)
  RUN(8, 8, 512, 2, 1, 1)
  RUN(8, 8, 1024, 1, 1, 1)
  RUN(32, 8, 1024, 1, 1, 1)
  RUN(16, 16, 512, 2, 1, 1)
  return 0;
}
