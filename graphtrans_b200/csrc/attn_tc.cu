// tcgen05 attention placeholder (see gemm_tc.cu note).
#include "common.cuh"
namespace gt {
int mha_tc_fwd_launch(int, const void*, const int32_t*, const int32_t*, const int32_t*, int64_t, int64_t, int32_t, int32_t, float,
                      void*, float*, float, const uint64_t*, uint64_t, cudaStream_t) {
    set_error("tcgen05 attention not built");
    return -2;
}
int mha_tc_bwd_launch(int, const void*, const void*, const void*, const float*, const int32_t*, const int32_t*,
                      const int32_t*, int64_t, int64_t, int32_t, int32_t, float, void*, float*, float, const uint64_t*,
                      uint64_t, cudaStream_t) {
    set_error("tcgen05 attention not built");
    return -2;
}
}  // namespace gt
