// Experiment: which path carries the MSDA backward's grad_value scatter fastest on sm_100a?
//   (a) red.global.add.v4.f32 from registers (what msda_bwd_kernel does today)
//   (b) 1-D TMA bulk reduce, one 128 B op per tap            (cp.reduce.async.bulk ... add.f32)
//   (c) tensor TMA reduce, one {32 ch, 2 x, 2 y} box per point (cp.reduce.async.bulk.tensor.4d ... add)
//   (s) plain st.global.v4 with the same addresses (the store path's ceiling, not a valid scatter)
// The traffic pattern is config 2's level 0: B=8, 200x200 pixels, 8 heads x 32 fp32 channels, layout
// (B, S, h, Dh); a CTA = 8x8 query tile of one head, 16 points per query with N(0, 2 px)-like offsets.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_build/exp_red_paths scripts/exp_red_paths.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int B = 8, W = 200, H = 200, NH = 8, DH = 32, NPT = 16;
constexpr int S = W * H;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// integer tap origin of point p of query (qx,qy): reference cell + offset in [-5, 4]^2 (triangular ~ sigma 2)
__device__ __forceinline__ void origin(int b, int head, int qx, int qy, int p, int& x0, int& y0) {
  const uint32_t r = hash32((((b * NH + head) * H + qy) * W + qx) * NPT + p);
  const int ox = (int)(r & 7) + (int)((r >> 3) & 3) - 5, oy = (int)((r >> 8) & 7) + (int)((r >> 11) & 3) - 5;
  x0 = min(max(qx + ox, 0), W - 2);
  y0 = min(max(qy + oy, 0), H - 2);
}

struct Ctx { int b, head, qx, qy; };
__device__ __forceinline__ Ctx ctx(int item) {  // item = group's query inside the 8x8 tile
  Ctx c;
  int blk = blockIdx.x;
  const int tile = blk % 625; blk /= 625;
  c.head = blk % NH; c.b = blk / NH;
  c.qx = (tile % 25) * 8 + (item & 7);
  c.qy = (tile / 25) * 8 + (item >> 3);
  return c;
}

template <int MODE>  // 0 = RED.v4, 1 = plain store
__global__ void __launch_bounds__(256, 4) k_red(float* __restrict__ G, const float4* __restrict__ go) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, lig = lane & 7;
  for (int base = warp * 4; base < 64; base += 32) {
    const Ctx c = ctx(base + g);
    const float4 v = go[(((size_t)c.b * S + c.qy * W + c.qx) * NH + c.head) * 8 + lig];
#pragma unroll 2
    for (int p = 0; p < NPT; ++p) {
      int x0, y0;
      origin(c.b, c.head, c.qx, c.qy, p, x0, y0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int pix = (y0 + (t >> 1)) * W + x0 + (t & 1);
        float* dst = G + ((((size_t)c.b * S + pix) * NH + c.head) * DH + lig * 4);
        const float wt = 0.25f + 0.1f * t;
        if (MODE == 0)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(wt * v.x), "f"(wt * v.y),
                       "f"(wt * v.z), "f"(wt * v.w) : "memory");
        else
          *reinterpret_cast<float4*>(dst) = make_float4(wt * v.x, wt * v.y, wt * v.z, wt * v.w);
      }
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (b): stage 16 taps (4 groups x 4 taps) of one point step in shared memory, lanes 0..15 issue one bulk reduce each
template <int SLOTS>
__global__ void __launch_bounds__(256, 4) k_bulk(float* __restrict__ G, const float4* __restrict__ go) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, lig = lane & 7;
  float4* stage = reinterpret_cast<float4*>(smem) + warp * (SLOTS * 16 * 8);  // [slot][tap16][8 x float4]
  int it = 0;
  for (int base = warp * 4; base < 64; base += 32) {
    const Ctx c = ctx(base + g);
    const float4 v = go[(((size_t)c.b * S + c.qy * W + c.qx) * NH + c.head) * 8 + lig];
    for (int p = 0; p < NPT; ++p, ++it) {
      int x0, y0;
      origin(c.b, c.head, c.qx, c.qy, p, x0, y0);
      float4* slot = stage + (it % SLOTS) * 128;
      if (it >= SLOTS) {   // the bulk ops that read this slot must have finished reading
        if (lane < 16) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SLOTS - 1) : "memory");
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float wt = 0.25f + 0.1f * t;
        slot[(g * 4 + t) * 8 + lig] = make_float4(wt * v.x, wt * v.y, wt * v.z, wt * v.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      // lane j < 16 issues tap j = (group j>>2, tap j&3): it needs that group's origin
      const int gx0 = __shfl_sync(0xffffffffu, x0, (lane >> 2) * 8), gy0 = __shfl_sync(0xffffffffu, y0, (lane >> 2) * 8);
      const int gb = __shfl_sync(0xffffffffu, c.b, (lane >> 2) * 8), gh = __shfl_sync(0xffffffffu, c.head, (lane >> 2) * 8);
      if (lane < 16) {
        const int t = lane & 3;
        const int pix = (gy0 + (t >> 1)) * W + gx0 + (t & 1);
        float* dst = G + ((((size_t)gb * S + pix) * NH + gh) * DH);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(dst),
                     "r"(smem_u32(slot + lane * 8)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lane < 16) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (c): one tensor reduce of the whole 2x2 footprint per (group, point); lane lig==0 of each group issues
template <int SLOTS>
__global__ void __launch_bounds__(256, 4) k_tensor(const __grid_constant__ CUtensorMap tm, const float4* __restrict__ go) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, lig = lane & 7;
  float4* stage = reinterpret_cast<float4*>(smem) + warp * (SLOTS * 16 * 8);
  int it = 0;
  for (int base = warp * 4; base < 64; base += 32) {
    const Ctx c = ctx(base + g);
    const float4 v = go[(((size_t)c.b * S + c.qy * W + c.qx) * NH + c.head) * 8 + lig];
    for (int p = 0; p < NPT; ++p, ++it) {
      int x0, y0;
      origin(c.b, c.head, c.qx, c.qy, p, x0, y0);
      float4* slot = stage + (it % SLOTS) * 128;
      if (it >= SLOTS) {
        if (lig == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SLOTS - 1) : "memory");
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float wt = 0.25f + 0.1f * t;
        slot[(g * 4 + t) * 8 + lig] = make_float4(wt * v.x, wt * v.y, wt * v.z, wt * v.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lig == 0) {
        asm volatile(
            "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
            ::"l"(&tm), "r"(c.head * DH), "r"(x0), "r"(y0), "r"(c.b), "r"(smem_u32(slot + g * 32)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lig == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

static double checksum(const float* d, size_t n) {
  float* h = (float*)malloc(n * sizeof(float));
  CK(cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost));
  double s = 0; for (size_t i = 0; i < n; ++i) s += h[i];
  free(h);
  return s;
}

int main() {
  const size_t n = (size_t)B * S * NH * DH;
  float *G, *go;
  CK(cudaMalloc(&G, n * 4)); CK(cudaMalloc(&go, n * 4));
  {
    float* h = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u >> 20) & 1023) / 1024.f - 0.5f;
    CK(cudaMemcpy(go, h, n * 4, cudaMemcpyHostToDevice));
    free(h);
  }
  const int grid = B * NH * 625;
  const double taps = (double)B * NH * S * NPT * 4;
  printf("taps per launch: %.1f M, payload %.2f GB\n", taps / 1e6, taps * 128 / 1e9);

  CK(cudaMemset(G, 0, n * 4));
  k_red<0><<<grid, 256>>>(G, (const float4*)go); CK(cudaDeviceSynchronize());
  const double ref = checksum(G, n);
  float ms = time_ms([&] { k_red<0><<<grid, 256>>>(G, (const float4*)go); }, 10);
  printf("(a) red.v4.f32          : %.3f ms  %.2f cyc/tap/SM  checksum %.6e\n", ms, ms * 1e-3 * 1.965e9 * 148 / taps, ref);
  ms = time_ms([&] { k_red<1><<<grid, 256>>>(G, (const float4*)go); }, 10);
  printf("(s) st.global.v4 ceiling: %.3f ms  %.2f cyc/tap/SM\n", ms, ms * 1e-3 * 1.965e9 * 148 / taps);

  {
    constexpr int SLOTS = 2;
    const size_t sm = 8 * SLOTS * 16 * 128;
    CK(cudaFuncSetAttribute(k_bulk<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaMemset(G, 0, n * 4));
    k_bulk<SLOTS><<<grid, 256, sm>>>(G, (const float4*)go); CK(cudaDeviceSynchronize());
    const double cs = checksum(G, n);
    ms = time_ms([&] { k_bulk<SLOTS><<<grid, 256, sm>>>(G, (const float4*)go); }, 10);
    printf("(b) bulk reduce 128 B   : %.3f ms  %.2f cyc/tap/SM  checksum %.6e (%s)\n", ms,
           ms * 1e-3 * 1.965e9 * 148 / taps, cs, fabs(cs - ref) <= 1e-4 * fabs(ref) + 1.0 ? "ok" : "MISMATCH");
  }
  {
    constexpr int SLOTS = 4;
    const size_t sm = 8 * SLOTS * 16 * 128;
    CK(cudaFuncSetAttribute(k_bulk<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    ms = time_ms([&] { k_bulk<SLOTS><<<grid, 256, sm>>>(G, (const float4*)go); }, 10);
    printf("(b4) bulk reduce, 4 slots: %.3f ms  %.2f cyc/tap/SM\n", ms, ms * 1e-3 * 1.965e9 * 148 / taps);
  }
  {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)NH * DH, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)NH * DH * 4, (cuuint64_t)W * NH * DH * 4, (cuuint64_t)S * NH * DH * 4};
    const cuuint32_t box[4] = {DH, 2, 2, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, G, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    constexpr int SLOTS = 2;
    const size_t sm = 8 * SLOTS * 16 * 128;
    CK(cudaFuncSetAttribute(k_tensor<SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaMemset(G, 0, n * 4));
    k_tensor<SLOTS><<<grid, 256, sm>>>(tm, (const float4*)go); CK(cudaDeviceSynchronize());
    const double cs = checksum(G, n);
    ms = time_ms([&] { k_tensor<SLOTS><<<grid, 256, sm>>>(tm, (const float4*)go); }, 10);
    printf("(c) tensor reduce 2x2x32: %.3f ms  %.2f cyc/tap/SM  checksum %.6e (%s)\n", ms,
           ms * 1e-3 * 1.965e9 * 148 / taps, cs, fabs(cs - ref) <= 1e-4 * fabs(ref) + 1.0 ? "ok" : "MISMATCH");
    constexpr int SLOTS4 = 4;
    const size_t sm4 = 8 * SLOTS4 * 16 * 128;
    CK(cudaFuncSetAttribute(k_tensor<SLOTS4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
    ms = time_ms([&] { k_tensor<SLOTS4><<<grid, 256, sm4>>>(tm, (const float4*)go); }, 10);
    printf("(c4) tensor reduce, 4 slots: %.3f ms  %.2f cyc/tap/SM\n", ms, ms * 1e-3 * 1.965e9 * 148 / taps);
  }
  return 0;
}
