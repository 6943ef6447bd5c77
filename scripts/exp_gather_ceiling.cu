// Experiment: the L1 ceiling of the MSDA forward gather on sm_100a.
// Same CTA / lane mapping and the same tap pattern as msda_fwd_kernel at config 2 (B=8, levels 200/100/50/25,
// 8 heads x 32 fp32 channels, 16 points, 4 taps each), but the tap origins come from a hash (no coordinate
// loads, no tap records in shared memory): what is left is LDG.128 x4 + 16 FFMA per point per lane.
//   (g)  gather only
//   (g+) gather + two LDS.128 per point (the tap-record broadcast of the product kernel)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_build/exp_gather_ceiling scripts/exp_gather_ceiling.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int B = 8, NH = 8, DH = 32, NPT = 16, QW = 200;
constexpr int S = 200 * 200 + 100 * 100 + 50 * 50 + 25 * 25;
__constant__ int c_start[4] = {0, 40000, 50000, 52500};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int REC, int MINB, int UNROLL>
__global__ void __launch_bounds__(256, MINB) k_gather(const uint4* __restrict__ V, float4* __restrict__ out) {
  __shared__ uint4 s_rec[8 * 4 * 2 * 17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, lig = lane & 7;
  int blk = blockIdx.x;
  const int tile = blk % 625; blk /= 625;
  const int head = blk % NH, b = blk / NH;
  const uint4* vbase = V + ((size_t)b * S * NH + head) * 8 + lig;
  uint4* rec = s_rec + (warp * 4 + g) * 34;
  if (REC) {
    for (int i = lig; i < 34; i += 8) rec[i] = make_uint4(i, lane, warp, 1);
    __syncwarp();
  }
  for (int base = warp * 4; base < 64; base += 32) {
    const int item = base + g;
    const int qx = (tile % 25) * 8 + (item & 7), qy = (tile / 25) * 8 + (item >> 3);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll UNROLL
    for (int p = 0; p < NPT; ++p) {
      const int l = p >> 2, Wl = QW >> l;
      const uint32_t r = hash32((((b * NH + head) * QW + qy) * QW + qx) * NPT + p);
      const int ox = (int)(r & 7) + (int)((r >> 3) & 3) - 5, oy = (int)((r >> 8) & 7) + (int)((r >> 11) & 3) - 5;
      const int x0 = min(max((qx >> l) + ox, 0), Wl - 2), y0 = min(max((qy >> l) + oy, 0), Wl - 2);
      const unsigned o0 = (unsigned)(c_start[l] + y0 * Wl + x0) * 64u;
      float w0 = 0.25f, w1 = 0.35f, w2 = 0.15f, w3 = 0.25f;
      unsigned extra = 0;
      if (REC) {
        const uint4 a = rec[p], c = rec[17 + p];
        extra = (a.w ^ c.w) & 1u;  // always 0, keeps the loads alive
        w0 += __uint_as_float(a.x & c.y & 0u);
      }
      const uint4 u0 = __ldg(vbase + o0 + extra), u1 = __ldg(vbase + o0 + 64u), u2 = __ldg(vbase + o0 + 64u * Wl),
                  u3 = __ldg(vbase + o0 + 64u * Wl + 64u);
      const uint4 u[4] = {u0, u1, u2, u3};
      const float w[4] = {w0, w1, w2, w3};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[0] = fmaf(w[t], __uint_as_float(u[t].x), acc[0]);
        acc[1] = fmaf(w[t], __uint_as_float(u[t].y), acc[1]);
        acc[2] = fmaf(w[t], __uint_as_float(u[t].z), acc[2]);
        acc[3] = fmaf(w[t], __uint_as_float(u[t].w), acc[3]);
      }
    }
    out[(((size_t)b * QW * QW + qy * QW + qx) * NH + head) * 8 + lig] = make_float4(acc[0], acc[1], acc[2], acc[3]);
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
  const size_t nv = (size_t)B * S * NH * DH, no = (size_t)B * QW * QW * NH * DH;
  float *V, *out;
  CK(cudaMalloc(&V, nv * 4)); CK(cudaMalloc(&out, no * 4));
  CK(cudaMemset(V, 0, nv * 4));
  const int grid = B * NH * 625;
  const double taps = (double)B * NH * QW * QW * NPT * 4;
  printf("taps per launch: %.1f M, gathered %.2f GB\n", taps / 1e6, taps * 128 / 1e9);
#define RUN(REC, MINB, UNROLL)                                                                                   \
  {                                                                                                              \
    const float ms = time_ms([&] { k_gather<REC, MINB, UNROLL><<<grid, 256>>>((const uint4*)V, (float4*)out); }, 10); \
    printf("rec=%d minb=%d unroll=%d: %.3f ms  %.2f cyc/tap/SM\n", REC, MINB, UNROLL, ms, ms * 1e-3 * 1.965e9 * 148 / taps); \
  }
  RUN(0, 4, 4) RUN(0, 6, 4) RUN(0, 8, 4) RUN(0, 4, 8) RUN(0, 6, 8) RUN(0, 8, 2) RUN(0, 8, 1)
  RUN(1, 4, 4) RUN(1, 6, 4) RUN(1, 8, 4) RUN(1, 6, 8)
  return 0;
}
