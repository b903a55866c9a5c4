// Spatial neighbourhood graphs over spot coordinates (d <= 4).
//   radius graph — StagateGraph(model_name="radius"): sklearn NearestNeighbors(radius=r).radius_neighbors_graph(X)
//   (reference transforms/graph/spatial_graph.py:143-151): A_ij = 1 iff ||x_i - x_j|| <= r, self included; the membership
//   test is done on squared distances in fp64 exactly as sklearn's trees do (rdist <= r²), so the structure is bit-exact.
// The coordinate table is tiny (N·d doubles) and the test is a handful of flops, so this is a brute-force tiled sweep:
// one thread per query spot, reference spots staged through shared memory; two passes (count, then fill in ascending
// column order — the canonical sorted-CSR form scipy produces with sort_indices()).
#include "common.cuh"

#include <cub/device/device_scan.cuh>

namespace b2 {

constexpr int RG_TILE = 1024;
constexpr int RG_MAXD = 4;

template <bool FILL>
__global__ void __launch_bounds__(256)
radius_graph_kernel(const double* __restrict__ X, int64_t ldx, int32_t n, int32_t d, double r2, int32_t* __restrict__ counts,
                    const int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx) {
  __shared__ double tile[RG_TILE * RG_MAXD];
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double xi[RG_MAXD] = {0.0, 0.0, 0.0, 0.0};
  if (i < n)
    for (int c = 0; c < d; ++c) xi[c] = X[(int64_t)i * ldx + c];
  int32_t cnt = 0;
  int32_t wp = (FILL && i < n) ? rowptr[i] : 0;
  for (int32_t j0 = 0; j0 < n; j0 += RG_TILE) {
    const int32_t m = (n - j0 < RG_TILE) ? n - j0 : RG_TILE;
    __syncthreads();
    for (int t = threadIdx.x; t < m * d; t += blockDim.x) tile[t] = X[(int64_t)(j0 + t / d) * ldx + t % d];
    __syncthreads();
    if (i < n) {
      for (int32_t j = 0; j < m; ++j) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) { const double df = xi[c] - tile[j * d + c]; s = __dadd_rn(s, __dmul_rn(df, df)); }   // no FMA contraction, as the host code
        if (s <= r2) {
          if (FILL) colidx[wp++] = j0 + j;
          else ++cnt;
        }
      }
    }
  }
  if (!FILL && i < n) counts[i] = cnt;
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_radius_graph_workspace_bytes(int32_t n) {
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, n + 1);
  return align_up(temp, 256) + align_up(sizeof(int32_t) * ((size_t)n + 1), 256);
}

extern "C" int b2_radius_graph_count(const double* X, int64_t ldx, int32_t n, int32_t d, double radius, int32_t* rowptr,
                                     int64_t* nnz_host, void* workspace, size_t workspace_bytes, void* stream) {
  B2_REQUIRE(X && rowptr && nnz_host && n >= 0 && d >= 1 && d <= RG_MAXD && ldx >= d && radius >= 0.0,
             "b2_radius_graph_count: bad arguments (1 <= d <= 4)");
  B2_REQUIRE(workspace && workspace_bytes >= b2_radius_graph_workspace_bytes(n), "b2_radius_graph_count: workspace too small");
  cudaStream_t st = as_stream(stream);
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, n + 1);
  char* base = reinterpret_cast<char*>(workspace);
  int32_t* counts = reinterpret_cast<int32_t*>(base + align_up(temp, 256));
  B2_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * ((size_t)n + 1), st));
  if (n > 0) {
    radius_graph_kernel<false><<<ceil_div(n, 256), 256, 0, st>>>(X, ldx, n, d, radius * radius, counts, nullptr, nullptr);
    B2_CHECK_LAUNCH("radius_graph_kernel<count>");
  }
  B2_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(base, temp, counts, rowptr, n + 1, st));
  int32_t total = 0;
  B2_CHECK_CUDA(cudaMemcpyAsync(&total, rowptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  B2_CHECK_CUDA(cudaStreamSynchronize(st));
  *nnz_host = total;
  return B2_OK;
}

extern "C" int b2_radius_graph_fill(const double* X, int64_t ldx, int32_t n, int32_t d, double radius, const int32_t* rowptr,
                                    int32_t* colidx, void* stream) {
  B2_REQUIRE(X && rowptr && colidx && n >= 0 && d >= 1 && d <= RG_MAXD && ldx >= d && radius >= 0.0,
             "b2_radius_graph_fill: bad arguments");
  if (n == 0) return B2_OK;
  radius_graph_kernel<true><<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(X, ldx, n, d, radius * radius, nullptr, rowptr, colidx);
  B2_CHECK_LAUNCH("radius_graph_kernel<fill>");
  return B2_OK;
}
