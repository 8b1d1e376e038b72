// Fused AdamW over the flat gradient arena (replaces torch.optim.AdamW.step at reference main.py:178 /
// trainers/base_trainer.py:36).  One launch updates every parameter: a descriptor table maps blocks of 2048 elements to
// (parameter tensor, offset into the arena); gradients, first and second moments share the arena layout of
// graphtrans_b200.ddp.GradBuckets.  Learning rate / betas / eps / weight decay and the step counter live in device
// memory, so a captured CUDA graph replays the update with whatever the host (or a scheduler) last wrote there.
#include "common.cuh"

namespace gt {

constexpr int ADAM_PER_BLOCK = 2048;

__global__ void k_adamw_advance(int64_t* step) { step[0] += 1; }

// desc[i] = {param pointer, arena offset (elements), numel, first block}
__global__ void __launch_bounds__(256)
k_adamw_multi(const int64_t* __restrict__ desc, int n, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
              const float* __restrict__ hyper, const int64_t* __restrict__ step, const float* __restrict__ clip) {
    int lo = 0, hi = n;   // largest i with desc[i].first_block <= blockIdx.x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (desc[mid * 4 + 3] <= (int64_t)blockIdx.x) lo = mid; else hi = mid;
    }
    const int64_t* dsc = desc + lo * 4;
    float* p = reinterpret_cast<float*>(dsc[0]);
    const int64_t off = dsc[1], numel = dsc[2];
    const int64_t base = ((int64_t)blockIdx.x - dsc[3]) * ADAM_PER_BLOCK;
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float t = (float)step[0];
    const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
    const float step_size = lr / bc1, decay = 1.f - lr * wd;
    // clip_grad_norm_(parameters, max_norm) of reference trainers/base_trainer.py:34-35 folded into the gradient read:
    // clip = {max_norm, sum of squares of the whole arena (gt_sumsq)}; coefficient = min(1, max_norm / (norm + 1e-6))
    float gscale = 1.f;
    if (clip && clip[0] > 0.f) gscale = fminf(1.f, clip[0] / (sqrtf(clip[1]) + 1e-6f));
#pragma unroll
    for (int k = 0; k < ADAM_PER_BLOCK / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < numel) {
            const float g = grad[off + i] * gscale;
            const float mi = b1 * m[off + i] + (1.f - b1) * g;
            const float vi = b2 * v[off + i] + (1.f - b2) * g * g;
            m[off + i] = mi;
            v[off + i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] * decay - step_size * (mi / denom);
        }
    }
}

}  // namespace gt

using namespace gt;

extern "C" int gt_adamw_multi(const int64_t* desc_dev, int32_t n, int64_t total_blocks, const float* grad_flat, float* m_flat,
                              float* v_flat, const float* hyper_dev, int64_t* step_dev, const float* clip_dev, void* stream) {
    GT_CHECK_ARG(desc_dev && n > 0 && total_blocks > 0 && total_blocks < (1ll << 31) && grad_flat && m_flat && v_flat && hyper_dev && step_dev,
                 "gt_adamw_multi: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    k_adamw_advance<<<1, 1, 0, st>>>(step_dev);
    k_adamw_multi<<<(unsigned)total_blocks, 256, 0, st>>>(desc_dev, n, grad_flat, m_flat, v_flat, hyper_dev, step_dev, clip_dev);
    GT_LAUNCH_CHECK("gt_adamw_multi");
    return 0;
}
