// Retrieval-recall ranks ("next" row f-4 of SURVEY.md section 8; Trainer.get_recall, train.py:551-572).
//
// The reference forms dist = 2 - 2 X Y^T with numpy and then, for each of 11 thresholds, loops over the columns in
// Python counting the rows with dist[:, k] < dist[k, k] (O(11 N^2) interpreter work per validation pass).  Here the
// GEMM stays a library call (cuBLAS, fp32) and this kernel does the counting in one pass over dist:
//   rank[k] = #{ r : dist[r, k] < dist[k, k] }          (strict <, as the reference)
// One warp lane per column (coalesced rows), the CTA's warps split the rows; HBM-bound: n^2 * 4 bytes read once.
#include "bevr_common.cuh"

namespace bevr {

__global__ void __launch_bounds__(256) recall_rank_kernel(const float* __restrict__ dist, int32_t* __restrict__ rank, int n) {
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  int cnt = 0;
  if (k < n) {
    const float gt = __ldg(dist + (size_t)k * n + k);
    for (int r = blockIdx.y * 8 + warp; r < n; r += 8 * gridDim.y) cnt += __ldg(dist + (size_t)r * n + k) < gt ? 1 : 0;
  }
  s_cnt[warp][lane] = cnt;
  __syncthreads();
  if (warp == 0 && k < n) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_cnt[w][lane];
    atomicAdd(rank + k, tot);
  }
}

}  // namespace bevr

// dist dev f32 (n, n) row-major; rank dev i32 (n), OVERWRITTEN.
extern "C" int bevr_recall_ranks(const float* dist, int32_t* rank, int n, void* stream) {
  BEVR_REQUIRE(n >= 0, "recall_ranks: n must be >= 0");
  if (n == 0) return 0;
  BEVR_REQUIRE(dist && rank, "recall_ranks: NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = bevr::check_cuda(cudaMemsetAsync(rank, 0, (size_t)n * sizeof(int32_t), st), "recall_ranks zero-fill")) return rc;
  const int col_blocks = (n + 31) / 32;
  int row_split = (4 * bevr::sm_count() + col_blocks - 1) / col_blocks;      // ~4 CTAs per SM
  row_split = row_split < 1 ? 1 : (row_split > (n + 7) / 8 ? (n + 7) / 8 : row_split);
  bevr::recall_rank_kernel<<<dim3(col_blocks, row_split), 256, 0, st>>>(dist, rank, n);
  return bevr::check_launch("recall_ranks");
}
