// Fused losses of the reference's dataset adapters (caller side of the hot path):
//   * mean over labelled (non-NaN) entries of BCE-with-logits   (reference dataset/mol.py:24-31)
//   * mean cross-entropy over rows                               (reference dataset/code.py:39-45, dataset/tud.py:25-27)
// Each is one forward launch (+ a finalise step done by the last block) and one backward launch instead of the
// 10-30 elementwise / reduction kernels of the eager formulation; everything stays on the device (no host sync).
#include "common.cuh"

namespace gt {

// acc[0] += sum of per-element losses, acc[1] += number of labelled entries; the last block to finish writes
// loss = acc[0] / max(acc[1], 1).  acc (fp32 [3]: sum, count, block ticket) must be zeroed by the caller.
__global__ void __launch_bounds__(256)
k_bce_masked_fwd(const float* __restrict__ x, const float* __restrict__ y, int64_t rows, int cols, int64_t ldx, int64_t ldy,
                 float* __restrict__ acc, float* __restrict__ loss) {
    float s = 0.f, c = 0.f;
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int k = (int)(i - r * cols);
        const float t = y[r * ldy + k];
        if (t == t) {   // labelled (not NaN)
            const float v = x[r * ldx + k];
            s += fmaxf(v, 0.f) - v * t + log1pf(__expf(-fabsf(v)));
            c += 1.f;
        }
    }
    s = warp_sum(s);
    c = warp_sum(c);
    __shared__ float sh[2][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[0][w] = s, sh[1][w] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts = 0.f, tc = 0.f;
        for (int k = 0; k < 8; ++k) ts += sh[0][k], tc += sh[1][k];
        atomicAdd(acc, ts);
        atomicAdd(acc + 1, tc);
        __threadfence();
        const float ticket = atomicAdd(acc + 2, 1.f);
        if (ticket == (float)(gridDim.x - 1)) {
            __threadfence();
            const float S = atomicAdd(acc, 0.f), C = atomicAdd(acc + 1, 0.f);
            *loss = S / fmaxf(C, 1.f);
        }
    }
}

// dx = g * (sigmoid(x) - y) / count on labelled entries, 0 elsewhere
__global__ void k_bce_masked_bwd(const float* __restrict__ x, const float* __restrict__ y, int64_t rows, int cols,
                                 int64_t ldx, int64_t ldy, const float* __restrict__ acc, const float* __restrict__ g,
                                 float* __restrict__ dx, int64_t lddx, int dcols) {
    const float scale = g[0] / fmaxf(acc[1], 1.f);
    const int64_t total = rows * dcols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dcols;
        const int k = (int)(i - r * dcols);
        float o = 0.f;
        if (k < cols) {
            const float t = y[r * ldy + k];
            if (t == t) {
                const float v = x[r * ldx + k];
                o = scale * (1.f / (1.f + __expf(-v)) - t);
            }
        }
        dx[r * lddx + k] = o;
    }
}

// one warp per row: lse[r] = logsumexp(x[r, :cols]); acc[0] += lse - x[r, target]; last block writes the mean
__global__ void __launch_bounds__(256)
k_ce_fwd(const float* __restrict__ x, const int64_t* __restrict__ target, int64_t tstride, int64_t rows, int cols, int64_t ldx,
         float* __restrict__ lse, float* __restrict__ acc, float* __restrict__ loss) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t r = blockIdx.x * 8 + w;
    float li = 0.f;
    if (r < rows) {
        const float* xr = x + r * ldx;
        float m = -INFINITY;
        for (int k = lane; k < cols; k += 32) m = fmaxf(m, xr[k]);
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < cols; k += 32) s += __expf(xr[k] - m);
        s = warp_sum(s);
        const float L = m + logf(s);
        if (lane == 0) {
            lse[r] = L;
            li = L - xr[target[r * tstride]];
        }
    }
    __shared__ float sh[8];
    if (lane == 0) sh[w] = li;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sh[k];
        atomicAdd(acc, t);
        __threadfence();
        const float ticket = atomicAdd(acc + 2, 1.f);
        if (ticket == (float)(gridDim.x - 1)) {
            __threadfence();
            *loss = atomicAdd(acc, 0.f) / (float)rows;
        }
    }
}

// dx[r, k] = g/rows * (exp(x - lse) - [k == target]) for k < cols, 0 for the pad columns up to dcols
__global__ void k_ce_bwd(const float* __restrict__ x, const int64_t* __restrict__ target, int64_t tstride, int64_t rows, int cols,
                         int64_t ldx, const float* __restrict__ lse, const float* __restrict__ g, float* __restrict__ dx,
                         int64_t lddx, int dcols) {
    const float scale = g[0] / (float)rows;
    const int64_t total = rows * dcols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dcols;
        const int k = (int)(i - r * dcols);
        float o = 0.f;
        if (k < cols) {
            o = __expf(x[r * ldx + k] - lse[r]);
            if (k == (int)target[r * tstride]) o -= 1.f;
            o *= scale;
        }
        dx[r * lddx + k] = o;
    }
}

}  // namespace gt

using namespace gt;

extern "C" int gt_bce_masked_fwd(const float* x, const float* y, int64_t rows, int32_t cols, int64_t ldx, int64_t ldy,
                                 float* acc, float* loss, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && ldx >= cols && ldy >= cols, "gt_bce_masked_fwd: bad shape");
    k_bce_masked_fwd<<<blocks_for(rows * cols, 256, kNumSMs), 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols, ldx, ldy, acc, loss);
    GT_LAUNCH_CHECK("gt_bce_masked_fwd");
    return 0;
}

extern "C" int gt_bce_masked_bwd(const float* x, const float* y, int64_t rows, int32_t cols, int64_t ldx, int64_t ldy,
                                 const float* acc, const float* g, float* dx, int64_t lddx, int32_t dcols, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && dcols >= cols && lddx >= dcols, "gt_bce_masked_bwd: bad shape");
    k_bce_masked_bwd<<<blocks_for(rows * dcols, 256), 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols, ldx, ldy, acc, g, dx, lddx, dcols);
    GT_LAUNCH_CHECK("gt_bce_masked_bwd");
    return 0;
}

extern "C" int gt_ce_fwd(const float* x, const int64_t* target, int64_t tstride, int64_t rows, int32_t cols, int64_t ldx,
                         float* lse, float* acc, float* loss, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && ldx >= cols, "gt_ce_fwd: bad shape");
    k_ce_fwd<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, target, tstride, rows, cols, ldx, lse, acc, loss);
    GT_LAUNCH_CHECK("gt_ce_fwd");
    return 0;
}

extern "C" int gt_ce_bwd(const float* x, const int64_t* target, int64_t tstride, int64_t rows, int32_t cols, int64_t ldx,
                         const float* lse, const float* g, float* dx, int64_t lddx, int32_t dcols, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && dcols >= cols && lddx >= dcols, "gt_ce_bwd: bad shape");
    k_ce_bwd<<<blocks_for(rows * dcols, 256), 256, 0, (cudaStream_t)stream>>>(x, target, tstride, rows, cols, ldx, lse, g, dx, lddx, dcols);
    GT_LAUNCH_CHECK("gt_ce_bwd");
    return 0;
}
