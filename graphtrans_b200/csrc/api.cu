// Error plumbing + version for the C ABI (include/graphtrans_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace gt {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
}  // namespace gt

extern "C" int gt_version(void) { return 100; }
extern "C" const char* gt_last_error(void) { return gt::g_err; }

// profiling hook (not part of the product ABI): one-thread kernel that writes %globaltimer (ns) to buf[idx] on
// `stream` - captured between the launches of a CUDA graph it gives the in-situ (L2-warm, overlapped) timeline
namespace gt {
__global__ void k_stamp(unsigned long long* buf, int idx) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[idx] = t;
}
}  // namespace gt
extern "C" int gtdbg_stamp(unsigned long long* buf, int idx, void* stream) {
    gt::k_stamp<<<1, 1, 0, (cudaStream_t)stream>>>(buf, idx);
    return (int)cudaGetLastError();
}
