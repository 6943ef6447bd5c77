// Experiment: L1 gather throughput on sm_100a as a function of the load width, for the MSDA tap pattern.
// A tap is 128 contiguous bytes (32 fp32 channels).  One warp instruction fetches
//   LDG.128: 4 taps (8 lanes x 16 B each)   LDG.64: 2 taps (16 lanes x 8 B)   LDG.32: 1 tap (32 lanes x 4 B)
// so the number of distinct 128 B lines per instruction is 4 / 2 / 1.  The guide's model says an extra line
// INSIDE one LDG costs ~2 cycles (replay) while lines of different LDGs stream at ~1 cycle, and every LDG costs
// ~1.8 cycles of LSU issue -- this measures where the optimum is.  `hot` = every tap in a 16 KB window (all L1 hits).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_build/exp_gather_width scripts/exp_gather_width.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int B = 8, NH = 8, NPT = 16, QW = 200, S = 53125;
__constant__ int c_start[4] = {0, 40000, 50000, 52500};

template <int LPG> struct Ld;
template <> struct Ld<8> { typedef uint4 V; static constexpr int N = 4; };
template <> struct Ld<16> { typedef uint2 V; static constexpr int N = 2; };
template <> struct Ld<32> { typedef uint32_t V; static constexpr int N = 1; };
__device__ __forceinline__ void acc_add(float* a, float w, uint4 u) {
  a[0] = fmaf(w, __uint_as_float(u.x), a[0]); a[1] = fmaf(w, __uint_as_float(u.y), a[1]);
  a[2] = fmaf(w, __uint_as_float(u.z), a[2]); a[3] = fmaf(w, __uint_as_float(u.w), a[3]);
}
__device__ __forceinline__ void acc_add(float* a, float w, uint2 u) {
  a[0] = fmaf(w, __uint_as_float(u.x), a[0]); a[1] = fmaf(w, __uint_as_float(u.y), a[1]);
}
__device__ __forceinline__ void acc_add(float* a, float w, uint32_t u) { a[0] = fmaf(w, __uint_as_float(u), a[0]); }

template <int LPG, int HOT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather(const char* __restrict__ V, float* __restrict__ out) {
  typedef typename Ld<LPG>::V VT;
  constexpr int N = Ld<LPG>::N, GPW = 32 / LPG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane / LPG, lig = lane % LPG;
  int blk = blockIdx.x;
  const int tile = blk % 625; blk /= 625;
  const int head = blk % NH, b = blk / NH;
  const char* vbase = V + (((size_t)b * S * NH + head) * 32 + lig * N) * 4;
  for (int base = warp * GPW; base < 64; base += 8 * GPW) {
    const int item = base + g;
    const int qx = (tile % 25) * 8 + (item & 7), qy = (tile / 25) * 8 + (item >> 3);
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.f;
    unsigned r = (unsigned)(item * 2654435761u) ^ (unsigned)blockIdx.x * 40503u;
#pragma unroll 4
    for (int p = 0; p < NPT; ++p) {
      const int l = p >> 2, Wl = QW >> l;
      r = r * 1664525u + 1013904223u;
      const int ox = (int)((r >> 8) & 7) + (int)((r >> 11) & 3) - 5, oy = (int)((r >> 16) & 7) + (int)((r >> 19) & 3) - 5;
      int x0, y0;
      if (HOT) { x0 = (ox + 5) & 7; y0 = (oy + 5) & 7; }
      else { x0 = min(max((qx >> l) + ox, 0), Wl - 2); y0 = min(max((qy >> l) + oy, 0), Wl - 2); }
      const unsigned o0 = HOT ? (unsigned)(y0 * 16 + x0) * 1024u : (unsigned)(c_start[l] + y0 * Wl + x0) * 1024u;
      const unsigned dy = HOT ? 16u * 1024u : 1024u * Wl;
      const VT u0 = __ldg((const VT*)(vbase + o0)), u1 = __ldg((const VT*)(vbase + o0 + 1024u)),
               u2 = __ldg((const VT*)(vbase + o0 + dy)), u3 = __ldg((const VT*)(vbase + o0 + dy + 1024u));
      acc_add(acc, 0.25f, u0); acc_add(acc, 0.35f, u1); acc_add(acc, 0.15f, u2); acc_add(acc, 0.25f, u3);
    }
    float* o = out + ((((size_t)b * QW * QW + qy * QW + qx) * NH + head) * 32 + lig * N);
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = acc[i];
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
  const int grid = B * NH * 625;
  const double taps = (double)B * NH * QW * QW * NPT * 4;
#define RUN(LPG, HOT, MINB)                                                                                       \
  {                                                                                                               \
    const float ms = time_ms([&] { k_gather<LPG, HOT, MINB><<<grid, 256>>>((const char*)V, out); }, 10);          \
    printf("LDG.%-3d hot=%d minb=%d: %.3f ms  %.2f cyc/tap/SM\n", 1024 / LPG, HOT, MINB, ms, ms * 1e-3 * 1.965e9 * 148 / taps); \
  }
  RUN(8, 0, 4) RUN(16, 0, 4) RUN(32, 0, 4) RUN(8, 0, 8) RUN(16, 0, 8) RUN(32, 0, 8)
  RUN(8, 1, 4) RUN(16, 1, 4) RUN(32, 1, 4) RUN(8, 1, 8) RUN(16, 1, 8) RUN(32, 1, 8)
  return 0;
}
