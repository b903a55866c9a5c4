// placeholder until the tcgen05 kernel lands (replaced below in this round)
#include "common.cuh"
namespace b2 {
int gemm_tc(const float*, int64_t, int, const float*, int64_t, int, float*, int64_t, int, int, int, const float*, int,
            const float*, int64_t, float, int, void*, size_t, cudaStream_t) { return B2_ERR_UNSUPPORTED; }
size_t gemm_tc_workspace_bytes(int, int, int, int, int, int) { return 0; }
}  // namespace b2
