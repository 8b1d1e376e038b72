// PNA multi-aggregator segmented reduce (mean / max / min / std in ONE pass over the in-edges)
// with the three degree scalers folded in, forward and backward.  Replaces the 6 torch_scatter
// passes, the [E, towers, F] message tensor and the two torch.cat of PNAConv.aggregate/forward
// (reference modules/pna_layer.py:131-167, modules/pna/aggregators.py:11-34,
// modules/pna/scalers.py:10-31; upstream torch_geometric PNAConv as built at
// modules/pna/pna_module.py:43-51: towers=4, divide_input=True, no edge features).
// Uses the identity W_pre [x_i || x_j] = W_i x_i + W_j x_j: the per-edge
// message is m_e = pi[dst] + pj[src] with pi/pj projected per NODE by the dense kernels, so this
// kernel only gathers pj rows.  var(m) = var(pj) (shift invariant), computed two-pass.
// The output row is exactly the operand of the post-MLP, per tower t:
//   [ x_t (F) | id*(mean,max,min,std) (4F) | amp*(...) (4F) | att*(...) (4F) ]   = 13F wide.
#include "common.cuh"

namespace gt {

constexpr int PNA_WARPS = 8;
constexpr float PNA_STD_EPS = 1e-5f;  // aggregators.py:34

struct Scalers { float s[3]; };
__device__ __forceinline__ Scalers pna_scalers(int deg, float delta) {
    Scalers r;
    const float lg = logf((float)deg + 1.f);
    r.s[0] = 1.f;                         // identity
    r.s[1] = lg / delta;                  // amplification (scalers.py:14-15)
    r.s[2] = deg == 0 ? 1.f : delta / lg; // attenuation   (scalers.py:18-21)
    return r;
}

template <typename T, int V>
__global__ void __launch_bounds__(PNA_WARPS * 32)
k_pna_fwd(const T* __restrict__ x, const T* __restrict__ pj, const T* __restrict__ pi, int N, int F, int d, int ld,
          const int32_t* __restrict__ rp, const int32_t* __restrict__ src, float delta, T* __restrict__ out,
          int ld_out, int32_t* __restrict__ amax, int32_t* __restrict__ amin) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * PNA_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * PNA_WARPS;
    for (int i = warp; i < N; i += nwarps) {
        const int b = rp[i], e = rp[i + 1];
        const Scalers sc = pna_scalers(e - b, delta);
        for (int c0 = lane * V; c0 < d; c0 += 32 * V) {
            float mean[V] = {}, mx[V], mn[V], sd[V];
            int ix[V], in[V];
#pragma unroll
            for (int q = 0; q < V; ++q) ix[q] = in[q] = -1;
            if (e > b) {
                float self[V];
                ldv<V>(pi + (int64_t)i * ld + c0, self);
#pragma unroll
                for (int q = 0; q < V; ++q) mx[q] = -INFINITY, mn[q] = INFINITY;
                for (int p = b; p < e; ++p) {
                    float v[V];
                    ldv<V>(pj + (int64_t)src[p] * ld + c0, v);
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        mean[q] += v[q];
                        if (v[q] > mx[q]) mx[q] = v[q], ix[q] = p;
                        if (v[q] < mn[q]) mn[q] = v[q], in[q] = p;
                    }
                }
                const float inv = 1.f / (float)(e - b);
                float var[V] = {};
#pragma unroll
                for (int q = 0; q < V; ++q) mean[q] *= inv;
                for (int p = b; p < e; ++p) {
                    float v[V];
                    ldv<V>(pj + (int64_t)src[p] * ld + c0, v);
#pragma unroll
                    for (int q = 0; q < V; ++q) var[q] = fmaf(v[q] - mean[q], v[q] - mean[q], var[q]);
                }
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    sd[q] = sqrtf(var[q] * inv + PNA_STD_EPS);
                    mean[q] += self[q];
                    mx[q] += self[q];
                    mn[q] += self[q];
                }
            } else {  // empty segment: torch_scatter yields 0 for mean/max/min (Appendix A.3)
#pragma unroll
                for (int q = 0; q < V; ++q) mx[q] = mn[q] = 0.f, sd[q] = sqrtf(PNA_STD_EPS);
            }
            const int t = c0 / F, f = c0 - t * F;
            T* o = out + (int64_t)i * ld_out + t * 13 * F + f;
            float xv[V];
            ldv<V>(x + (int64_t)i * ld + c0, xv);
            stv<V>(o, xv);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                float a[V];
                T* os = o + F + s * 4 * F;
#pragma unroll
                for (int q = 0; q < V; ++q) a[q] = sc.s[s] * mean[q];
                stv<V>(os, a);
#pragma unroll
                for (int q = 0; q < V; ++q) a[q] = sc.s[s] * mx[q];
                stv<V>(os + F, a);
#pragma unroll
                for (int q = 0; q < V; ++q) a[q] = sc.s[s] * mn[q];
                stv<V>(os + 2 * F, a);
#pragma unroll
                for (int q = 0; q < V; ++q) a[q] = sc.s[s] * sd[q];
                stv<V>(os + 3 * F, a);
            }
#pragma unroll
            for (int q = 0; q < V; ++q) {
                amax[(int64_t)i * ld + c0 + q] = ix[q];
                amin[(int64_t)i * ld + c0 + q] = in[q];
            }
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(PNA_WARPS * 32)
k_pna_bwd(const T* __restrict__ pj, const T* __restrict__ out, const T* __restrict__ dout, int N, int F, int d,
          int ld, int ld_out, const int32_t* __restrict__ rp, const int32_t* __restrict__ src, float delta,
          const int32_t* __restrict__ amax, const int32_t* __restrict__ amin, float* __restrict__ dpj,
          T* __restrict__ dpi, T* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * PNA_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * PNA_WARPS;
    for (int i = warp; i < N; i += nwarps) {
        const int b = rp[i], e = rp[i + 1];
        const Scalers sc = pna_scalers(e - b, delta);
        for (int c0 = lane * V; c0 < ld; c0 += 32 * V) {
            float gi[V] = {}, gx[V] = {};
            if (c0 < d) {
                const int t = c0 / F, f = c0 - t * F;
                const T* go = dout + (int64_t)i * ld_out + t * 13 * F + f;
                ldv<V>(go, gx);
                if (e > b) {
                    float gmean[V] = {}, gmax[V] = {};
                    float gmin[V] = {}, gstd[V] = {};
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const T* gs = go + F + s * 4 * F;
                        float a[V];
                        ldv<V>(gs, a);
#pragma unroll
                        for (int q = 0; q < V; ++q) gmean[q] = fmaf(sc.s[s], a[q], gmean[q]);
                        ldv<V>(gs + F, a);
#pragma unroll
                        for (int q = 0; q < V; ++q) gmax[q] = fmaf(sc.s[s], a[q], gmax[q]);
                        ldv<V>(gs + 2 * F, a);
#pragma unroll
                        for (int q = 0; q < V; ++q) gmin[q] = fmaf(sc.s[s], a[q], gmin[q]);
                        ldv<V>(gs + 3 * F, a);
#pragma unroll
                        for (int q = 0; q < V; ++q) gstd[q] = fmaf(sc.s[s], a[q], gstd[q]);
                    }
                    // identity-scaled std (slot s = 0) is the std itself; recompute mean(pj) on the fly
                    float sd[V];
                    ldv<V>(out + (int64_t)i * ld_out + t * 13 * F + f + F + 3 * F, sd);
                    const float inv = 1.f / (float)(e - b);
                    float mean[V] = {};
                    for (int p = b; p < e; ++p) {
                        float v[V];
                        ldv<V>(pj + (int64_t)src[p] * ld + c0, v);
#pragma unroll
                        for (int q = 0; q < V; ++q) mean[q] += v[q];
                    }
                    int ixa[V], ina[V];
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        ixa[q] = amax[(int64_t)i * ld + c0 + q];
                        ina[q] = amin[(int64_t)i * ld + c0 + q];
                    }
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        mean[q] *= inv;
                        gi[q] = gmean[q] + gmax[q] + gmin[q];
                        // d std / d v = (v - mean) / (n * std) while var > 0 (relu gate, aggregators.py:34)
                        gstd[q] = (sd[q] * sd[q] - PNA_STD_EPS > 0.f) ? gstd[q] * inv / sd[q] : 0.f;
                    }
                    for (int p = b; p < e; ++p) {
                        const int64_t s = src[p];
                        float v[V];
                        ldv<V>(pj + s * ld + c0, v);
#pragma unroll
                        for (int q = 0; q < V; ++q) {
                            float g = gmean[q] * inv + gstd[q] * (v[q] - mean[q]);
                            if (p == ixa[q]) g += gmax[q];
                            if (p == ina[q]) g += gmin[q];
                            atomicAdd(dpj + s * ld + c0 + q, g);
                        }
                    }
                }
            }
            stv<V>(dpi + (int64_t)i * ld + c0, gi);
            stv<V>(dx + (int64_t)i * ld + c0, gx);
        }
    }
}

static int check_pna(const char* fn, int64_t N, int towers, int F, int ld, int ld_out, float delta) {
    GT_CHECK_ARG(N > 0 && N < (1ll << 31) && towers > 0 && F > 0, "%s: bad shape", fn);
    GT_CHECK_ARG(ld >= towers * F && ld % 4 == 0 && ld_out >= towers * 13 * F, "%s: bad leading dims", fn);
    GT_CHECK_ARG(delta > 0.f, "%s: avg_deg['log'] must be positive", fn);
    return 0;
}

}  // namespace gt

using namespace gt;

extern "C" int gt_pna_reduce_fwd(int dt, const void* x, const void* pj, const void* pi, int64_t N, int32_t towers,
                                 int32_t F, int32_t ld, const int32_t* rowptr_dst, const int32_t* src_by_dst,
                                 float delta, void* out, int32_t ld_out, int32_t* argmax, int32_t* argmin,
                                 void* stream) {
    if (int r = check_pna("gt_pna_reduce_fwd", N, towers, F, ld, ld_out, delta)) return r;
    const int grid = blocks_for(N, PNA_WARPS, kNumSMs * 8);
    // vector width 4 needs every tower block (and hence every 13F-wide output block) 4-aligned
    const bool v4 = F % 4 == 0 && ld_out % 4 == 0;
    GT_DISPATCH_DT(dt, {
        if (v4) k_pna_fwd<T, 4><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)pj, (const T*)pi, (int)N, F, towers * F, ld, rowptr_dst, src_by_dst, delta, (T*)out, ld_out, argmax, argmin);
        else k_pna_fwd<T, 1><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)pj, (const T*)pi, (int)N, F, towers * F, ld, rowptr_dst, src_by_dst, delta, (T*)out, ld_out, argmax, argmin);
    });
    GT_LAUNCH_CHECK("gt_pna_reduce_fwd");
    return 0;
}

extern "C" int gt_pna_reduce_bwd(int dt, const void* pj, const void* out, const void* dout, int64_t N,
                                 int32_t towers, int32_t F, int32_t ld, int32_t ld_out, const int32_t* rowptr_dst,
                                 const int32_t* src_by_dst, float delta, const int32_t* argmax,
                                 const int32_t* argmin, float* dpj, void* dpi, void* dx, void* stream) {
    if (int r = check_pna("gt_pna_reduce_bwd", N, towers, F, ld, ld_out, delta)) return r;
    const int grid = blocks_for(N, PNA_WARPS, kNumSMs * 8);
    const bool v4 = F % 4 == 0 && ld_out % 4 == 0;
    GT_DISPATCH_DT(dt, {
        if (v4) k_pna_bwd<T, 4><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)pj, (const T*)out, (const T*)dout, (int)N, F, towers * F, ld, ld_out, rowptr_dst, src_by_dst, delta, argmax, argmin, dpj, (T*)dpi, (T*)dx);
        else k_pna_bwd<T, 1><<<grid, PNA_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)pj, (const T*)out, (const T*)dout, (int)N, F, towers * F, ld, ld_out, rowptr_dst, src_by_dst, delta, argmax, argmin, dpj, (T*)dpi, (T*)dx);
    });
    GT_LAUNCH_CHECK("gt_pna_reduce_bwd");
    return 0;
}
