// tcgen05 GEMM placeholder: replaced by the TMA + tcgen05/TMEM kernel; until then every call
// reports "not eligible" (-2) and gt_gemm falls through to the CUDA-core kernel.
#include "common.cuh"
namespace gt {
int gemm_tc_launch(int, const void*, int, int64_t, const void*, int, int64_t, void*, int64_t, int64_t, int64_t,
                   int64_t, int64_t, const float*, const void*, int64_t, int, cudaStream_t) {
    set_error("tcgen05 GEMM not built");
    return -2;
}
}  // namespace gt
