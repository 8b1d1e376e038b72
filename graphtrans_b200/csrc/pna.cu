// PNA multi-aggregator segmented reduce (mean / max / min / std in ONE pass over the in-edges),
// forward and backward.  Replaces the 6 torch_scatter passes + [E, towers, F] message tensor of
// PNAConv.aggregate (reference modules/pna_layer.py:161-167, modules/pna/aggregators.py:11-34).
// Uses the identity W_pre [x_i || x_j] = W_i x_i + W_j x_j: the per-edge
// message is m_e = pi[dst] + pj[src] with pi/pj projected per NODE on tensor cores, so this
// kernel only gathers pj rows.  var(m) = var(pj) (shift invariant), computed two-pass.
#include "common.cuh"

namespace gt {

constexpr int PNA_WARPS = 8;
constexpr float PNA_STD_EPS = 1e-5f;  // aggregators.py:34

template <typename T>
__global__ void __launch_bounds__(PNA_WARPS * 32)
k_pna_fwd(const T* __restrict__ pj, const T* __restrict__ pi, int N, int ld, const int32_t* __restrict__ rp,
          const int32_t* __restrict__ src, T* __restrict__ out, int32_t* __restrict__ amax, int32_t* __restrict__ amin) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * PNA_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * PNA_WARPS;
    for (int i = warp; i < N; i += nwarps) {
        const int b = rp[i], e = rp[i + 1];
        for (int c0 = lane * 4; c0 < ld; c0 += 128) {
            float mean[4] = {0.f, 0.f, 0.f, 0.f}, mx[4], mn[4], sd[4];
            int ix[4] = {-1, -1, -1, -1}, in[4] = {-1, -1, -1, -1};
            float self[4];
            ld4(pi + (int64_t)i * ld + c0, self);
            if (e > b) {
#pragma unroll
                for (int q = 0; q < 4; ++q) mx[q] = -INFINITY, mn[q] = INFINITY;
                for (int p = b; p < e; ++p) {
                    float v[4];
                    ld4(pj + (int64_t)src[p] * ld + c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        mean[q] += v[q];
                        if (v[q] > mx[q]) mx[q] = v[q], ix[q] = p;
                        if (v[q] < mn[q]) mn[q] = v[q], in[q] = p;
                    }
                }
                const float inv = 1.f / (float)(e - b);
                float var[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 4; ++q) mean[q] *= inv;
                for (int p = b; p < e; ++p) {
                    float v[4];
                    ld4(pj + (int64_t)src[p] * ld + c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) var[q] = fmaf(v[q] - mean[q], v[q] - mean[q], var[q]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    sd[q] = sqrtf(var[q] * inv + PNA_STD_EPS);
                    mean[q] += self[q];
                    mx[q] += self[q];
                    mn[q] += self[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) mx[q] = mn[q] = 0.f, sd[q] = sqrtf(PNA_STD_EPS);
            }
            T* o = out + (int64_t)i * 4 * ld + c0;
            st4(o, mean);
            st4(o + ld, mx);
            st4(o + 2 * ld, mn);
            st4(o + 3 * ld, sd);
            *reinterpret_cast<int4*>(amax + (int64_t)i * ld + c0) = make_int4(ix[0], ix[1], ix[2], ix[3]);
            *reinterpret_cast<int4*>(amin + (int64_t)i * ld + c0) = make_int4(in[0], in[1], in[2], in[3]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(PNA_WARPS * 32)
k_pna_bwd(const T* __restrict__ pj, const T* __restrict__ pi, const T* __restrict__ out, const T* __restrict__ dout,
          int N, int ld, const int32_t* __restrict__ rp, const int32_t* __restrict__ src,
          const int32_t* __restrict__ amax, const int32_t* __restrict__ amin, float* __restrict__ dpj,
          T* __restrict__ dpi) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * PNA_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * PNA_WARPS;
    for (int i = warp; i < N; i += nwarps) {
        const int b = rp[i], e = rp[i + 1];
        for (int c0 = lane * 4; c0 < ld; c0 += 128) {
            float gi[4] = {0.f, 0.f, 0.f, 0.f};
            if (e > b) {
                float gmean[4], gmax[4], gmin[4], gstd[4], mean[4], sd[4], self[4];
                const T* go = dout + (int64_t)i * 4 * ld + c0;
                const T* oo = out + (int64_t)i * 4 * ld + c0;
                ld4(go, gmean);
                ld4(go + ld, gmax);
                ld4(go + 2 * ld, gmin);
                ld4(go + 3 * ld, gstd);
                ld4(oo, mean);
                ld4(oo + 3 * ld, sd);
                ld4(pi + (int64_t)i * ld + c0, self);
                const int4 ix = *reinterpret_cast<const int4*>(amax + (int64_t)i * ld + c0);
                const int4 in = *reinterpret_cast<const int4*>(amin + (int64_t)i * ld + c0);
                const int ixa[4] = {ix.x, ix.y, ix.z, ix.w}, ina[4] = {in.x, in.y, in.z, in.w};
                const float inv = 1.f / (float)(e - b);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    gi[q] = gmean[q] + gmax[q] + gmin[q];
                    mean[q] -= self[q];                       // mean of pj
                    // d std / d v = (v - mean) / (n * std) while var > 0 (relu gate, aggregators.py:34)
                    gstd[q] = (sd[q] * sd[q] - PNA_STD_EPS > 0.f) ? gstd[q] * inv / sd[q] : 0.f;
                }
                for (int p = b; p < e; ++p) {
                    const int64_t s = src[p];
                    float v[4];
                    ld4(pj + s * ld + c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float g = gmean[q] * inv + gstd[q] * (v[q] - mean[q]);
                        if (p == ixa[q]) g += gmax[q];
                        if (p == ina[q]) g += gmin[q];
                        atomicAdd(dpj + s * ld + c0 + q, g);
                    }
                }
            }
            st4(dpi + (int64_t)i * ld + c0, gi);
        }
    }
}

}  // namespace gt

using namespace gt;

extern "C" int gt_pna_reduce_fwd(int dt, const void* pj, const void* pi, int64_t N, int32_t d, int32_t ld,
                                 const int32_t* rowptr_dst, const int32_t* src_by_dst, void* out, int32_t* argmax,
                                 int32_t* argmin, void* stream) {
    GT_CHECK_ARG(N > 0 && N < (1ll << 31) && d > 0 && ld >= d && ld % 4 == 0, "gt_pna_reduce_fwd: bad shape");
    const int grid = blocks_for(N, PNA_WARPS, kNumSMs * 8);
    GT_DISPATCH_DT(dt, (k_pna_fwd<T><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)pj, (const T*)pi, (int)N, ld, rowptr_dst, src_by_dst, (T*)out, argmax, argmin)));
    GT_LAUNCH_CHECK("gt_pna_reduce_fwd");
    return 0;
}

extern "C" int gt_pna_reduce_bwd(int dt, const void* pj, const void* pi, const void* out, const void* dout, int64_t N,
                                 int32_t d, int32_t ld, const int32_t* rowptr_dst, const int32_t* src_by_dst,
                                 const int32_t* argmax, const int32_t* argmin, float* dpj, void* dpi, void* stream) {
    GT_CHECK_ARG(N > 0 && N < (1ll << 31) && d > 0 && ld >= d && ld % 4 == 0, "gt_pna_reduce_bwd: bad shape");
    const int grid = blocks_for(N, PNA_WARPS, kNumSMs * 8);
    GT_DISPATCH_DT(dt, (k_pna_bwd<T><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)pj, (const T*)pi, (const T*)out, (const T*)dout, (int)N, ld, rowptr_dst, src_by_dst, argmax, argmin, dpj, (T*)dpi)));
    GT_LAUNCH_CHECK("gt_pna_reduce_bwd");
    return 0;
}
