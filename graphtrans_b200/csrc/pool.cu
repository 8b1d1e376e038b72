// Graph-level read-outs and small reductions around the hot path:
//   * global mean / max pooling over the nodes of each graph (PyG global_mean_pool / global_max_pool as used by the
//     reference baselines models/gnn.py:64-69, models/pna.py:74-79, models/transformer.py:49-54; global_add_pool is
//     gt_segment_sum_sorted) and their backward
//   * row-wise argmax of logits (eval: reference dataset/code.py:64 `torch.argmax(pred, dim=1)`)
//   * sum of squares of the flat gradient arena (clip_grad_norm_, reference trainers/base_trainer.py:34-35)
#include "common.cuh"

namespace gt {

// one block per (graph, 128-channel chunk); 8 warps take interleaved rows of the graph, partial results meet in shared
// memory, warp 0 is the single writer (deterministic, no atomics).  mode 1 = mean, 2 = max (+ first-occurrence argmax).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_segment_pool(const T* __restrict__ x, const int32_t* __restrict__ node_off, int ld, int nch, float* __restrict__ out,
               int32_t* __restrict__ arg) {
    __shared__ float4 part[8][32];
    __shared__ int4 parti[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x / nch, c0 = (blockIdx.x - g * nch) * 128 + lane * 4;
    const bool col_ok = c0 < ld;
    const int r0 = node_off[g], r1 = node_off[g + 1];
    float acc[4];
    int idx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = MODE == 2 ? -INFINITY : 0.f, idx[q] = -1;
    if (col_ok) {
        for (int r = r0 + warp; r < r1; r += 8) {
            float v[4];
            ld4(x + (int64_t)r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (MODE == 2) {
                    if (v[q] > acc[q]) acc[q] = v[q], idx[q] = r;      // rows ascend inside a warp: first occurrence kept
                } else {
                    acc[q] += v[q];
                }
            }
        }
    }
    part[warp][lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (MODE == 2) parti[warp][lane] = make_int4(idx[0], idx[1], idx[2], idx[3]);
    __syncthreads();
    if (warp == 0 && col_ok) {
        float o[4];
        int oi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = MODE == 2 ? -INFINITY : 0.f, oi[q] = -1;
        for (int w = 0; w < 8; ++w) {
            const float4 p = part[w][lane];
            const float pv[4] = {p.x, p.y, p.z, p.w};
            if (MODE == 2) {
                const int4 pi4 = parti[w][lane];
                const int pi[4] = {pi4.x, pi4.y, pi4.z, pi4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (pi[q] >= 0 && (pv[q] > o[q] || (pv[q] == o[q] && pi[q] < oi[q]) || oi[q] < 0)) o[q] = pv[q], oi[q] = pi[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] += pv[q];
            }
        }
        const int n = r1 - r0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (MODE == 1) o[q] = n > 0 ? o[q] / (float)n : 0.f;
            if (MODE == 2 && oi[q] < 0) o[q] = 0.f;                     // empty graph -> 0 (torch_scatter semantics)
        }
        *reinterpret_cast<float4*>(out + (int64_t)g * ld + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (MODE == 2) *reinterpret_cast<int4*>(arg + (int64_t)g * ld + c0) = make_int4(oi[0], oi[1], oi[2], oi[3]);
    }
}

// dx[i, :] = dout[g(i), :] / n_g (mean) or dout[g(i), c] * [arg[g(i), c] == i] (max); slack nodes (g < 0) get zeros
template <typename T, int MODE>
__global__ void k_segment_pool_bwd(const float* __restrict__ dout, const int32_t* __restrict__ node_off,
                                   const int32_t* __restrict__ node_graph, const int32_t* __restrict__ arg, int64_t N,
                                   int ld, T* __restrict__ dx) {
    const int vpr = ld / 4;
    const int64_t total = N * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        const int g = node_graph[r];
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (g >= 0) {
            const float4 d4 = *reinterpret_cast<const float4*>(dout + (int64_t)g * ld + c0);
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            if (MODE == 1) {
                const float inv = 1.f / (float)max(node_off[g + 1] - node_off[g], 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = dv[q] * inv;
            } else {
                const int4 a4 = *reinterpret_cast<const int4*>(arg + (int64_t)g * ld + c0);
                const int av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = av[q] == (int)r ? dv[q] : 0.f;
            }
        }
        st4(dx + r * ld + c0, o);
    }
}

// one warp per row: index of the first maximum of x[r, :cols]
__global__ void __launch_bounds__(256)
k_argmax_rows(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* xr = x + r * ldx;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < cols; k += 32) {
        const float v = xr[k];
        if (v > best || (v == best && k < bi)) best = v, bi = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (lane == 0) out[r] = bi == 0x7fffffff ? 0 : bi;
}

// deterministic: every block writes its partial sum to scratch[blockIdx.x]; the last block to finish (ticket in
// scratch[gridDim.x]) adds the partials in index order - the same bits on every rank of a data-parallel job (the clip
// coefficient derived from it must not differ between ranks) and on every replay
__global__ void __launch_bounds__(256)
k_sumsq(const float* __restrict__ x, int64_t n, float* __restrict__ out, float* __restrict__ scratch) {
    float s = 0.f;
    const int64_t n4 = n / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n4 * 4; i < n; ++i) s = fmaf(x[i], x[i], s);
    s = warp_sum(s);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sh[k];
        scratch[blockIdx.x] = t;
        __threadfence();
        unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + gridDim.x);
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        float t = 0.f;                              // fixed order: lane l takes partials l, l + 32, ...; then a fixed shuffle tree
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) t += __ldcg(scratch + i);
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            out[0] = t;
            *reinterpret_cast<unsigned int*>(scratch + gridDim.x) = 0u;     // re-armed for the next launch / replay
        }
    }
}

}  // namespace gt

using namespace gt;

extern "C" int gt_segment_pool_fwd(int dt, int mode, const void* x, const int32_t* node_off, int64_t B, int32_t ld,
                                   float* out, int32_t* arg, void* stream) {
    GT_CHECK_ARG(B > 0 && ld > 0 && ld % 4 == 0 && (mode == 1 || mode == 2), "gt_segment_pool_fwd: bad shape / mode %d", mode);
    GT_CHECK_ARG(mode != 2 || arg, "gt_segment_pool_fwd: max pooling needs the argmax buffer");
    const int nch = (ld + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        GT_DISPATCH_DT(dt, (k_segment_pool<T, 1><<<(unsigned)(B * nch), 256, 0, st>>>((const T*)x, node_off, ld, nch, out, arg)));
    } else {
        GT_DISPATCH_DT(dt, (k_segment_pool<T, 2><<<(unsigned)(B * nch), 256, 0, st>>>((const T*)x, node_off, ld, nch, out, arg)));
    }
    GT_LAUNCH_CHECK("gt_segment_pool_fwd");
    return 0;
}

extern "C" int gt_segment_pool_bwd(int dt, int mode, const float* dout, const int32_t* node_off, const int32_t* node_graph,
                                   const int32_t* arg, int64_t N, int32_t ld, void* dx, void* stream) {
    GT_CHECK_ARG(N > 0 && ld > 0 && ld % 4 == 0 && (mode == 1 || mode == 2), "gt_segment_pool_bwd: bad shape / mode %d", mode);
    GT_CHECK_ARG(mode != 2 || arg, "gt_segment_pool_bwd: max pooling needs the argmax buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = blocks_for(N * (ld / 4), 256);
    if (mode == 1) {
        GT_DISPATCH_DT(dt, (k_segment_pool_bwd<T, 1><<<blocks, 256, 0, st>>>(dout, node_off, node_graph, arg, N, ld, (T*)dx)));
    } else {
        GT_DISPATCH_DT(dt, (k_segment_pool_bwd<T, 2><<<blocks, 256, 0, st>>>(dout, node_off, node_graph, arg, N, ld, (T*)dx)));
    }
    GT_LAUNCH_CHECK("gt_segment_pool_bwd");
    return 0;
}

extern "C" int gt_argmax_rows(const float* x, int64_t rows, int32_t cols, int64_t ldx, int64_t* out, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && ldx >= cols, "gt_argmax_rows: bad shape");
    k_argmax_rows<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, out);
    GT_LAUNCH_CHECK("gt_argmax_rows");
    return 0;
}

extern "C" int gt_sumsq(const float* x, int64_t n, float* out, float* scratch, int32_t n_scratch, void* stream) {
    GT_CHECK_ARG(n > 0 && ((uintptr_t)x % 16) == 0, "gt_sumsq: needs a 16-byte aligned buffer");
    GT_CHECK_ARG(scratch && n_scratch >= 2, "gt_sumsq: needs a zero-initialised scratch of >= 2 floats");
    int blocks = blocks_for(n / 4 + 1, 256, kNumSMs * 4);
    if (blocks > n_scratch - 1) blocks = n_scratch - 1;
    k_sumsq<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out, scratch);
    GT_LAUNCH_CHECK("gt_sumsq");
    return 0;
}
