// Experiment: throughput of legacy mma.sync.m16n8k8 TF32 on sm_100a (per SM), against FFMA2.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_build/exp_mma_sync scripts/exp_mma_sync.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = threadIdx.x + 5, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  float2 c[8];
  for (int i = 0; i < 8; ++i) c[i] = make_float2(0.f, 0.f);
  float2 a = make_float2(threadIdx.x * 1e-3f, 0.5f), b = make_float2(1.0001f, 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __ffma2_rn(a, b, c[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += c[i].x + c[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; CK(cudaMalloc(&out, 148 * 8 * 256 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 20000, threads = 128, blocks = 148 * warps / 4;
    k_mma<<<blocks, threads>>>(out, 10); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_mma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas_per_sm = (double)iters * 8 * warps;
    printf("mma.sync m16n8k8 tf32, %2d warps/SM: %.3f ms, %.2f cycles per MMA per SM (%.0f dense TF32 TFLOP/s chip)\n", warps, ms,
           ms * 1e-3 * 1.965e9 / mmas_per_sm, mmas_per_sm * 148 * 2048 / (ms * 1e-3) / 1e12);
    k_ffma2<<<blocks, threads>>>(out, 10); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2,                  %2d warps/SM: %.3f ms, %.2f cycles per warp-FFMA2 per SM\n", warps, ms, ms * 1e-3 * 1.965e9 / mmas_per_sm);
  }
  return 0;
}
