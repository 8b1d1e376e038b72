// Integer graph preparation: stable counting-sort CSR build, edge-type ids, batch plan.
// Replaces the index handling that PyG MessagePassing / degree / pad_batch do on the fly
// (reference modules/conv.py:28,57,63; modules/utils.py:5-29).  All results are bit-exact
// functions of the inputs (the atomics only pick slots; rows are re-sorted by edge id).
#include "common.cuh"

namespace gt {

__global__ void k_hist2(const int64_t* __restrict__ ei, int64_t E, int32_t* cnt_dst, int32_t* cnt_src) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        atomicAdd(&cnt_src[ei[e]], 1);
        atomicAdd(&cnt_dst[ei[E + e]], 1);
    }
}

// exclusive scan of cnt[0..n) into out[0..n], out[n] = total. One block per array (blockIdx.x).
__global__ void k_scan2(const int32_t* __restrict__ cnt0, int32_t* __restrict__ out0,
                        const int32_t* __restrict__ cnt1, int32_t* __restrict__ out1, int64_t n) {
    const int32_t* cnt = blockIdx.x == 0 ? cnt0 : cnt1;
    int32_t* out = blockIdx.x == 0 ? out0 : out1;
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    constexpr int ITEMS = 8;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)blockDim.x * ITEMS) {
        int32_t v[ITEMS];
        int32_t sum = 0;
        const int64_t i0 = base + (int64_t)tid * ITEMS;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            v[j] = (i0 + j < n) ? cnt[i0 + j] : 0;
            sum += v[j];
        }
        int32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0;
            int32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        int32_t run = carry_s + warp_tot[wid] + (incl - sum);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (i0 + j < n) out[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = run;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

__global__ void k_fill2(const int64_t* __restrict__ ei, int64_t E, const int32_t* __restrict__ rp_dst,
                        const int32_t* __restrict__ rp_src, int32_t* cur_dst, int32_t* cur_src,
                        int32_t* src_by_dst, int32_t* eid_by_dst, int32_t* dst_by_src, int32_t* eid_by_src) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = (int32_t)ei[e], t = (int32_t)ei[E + e];
        int32_t p = rp_dst[t] + atomicAdd(&cur_dst[t], 1);
        src_by_dst[p] = s;
        eid_by_dst[p] = (int32_t)e;
        p = rp_src[s] + atomicAdd(&cur_src[s], 1);
        dst_by_src[p] = t;
        eid_by_src[p] = (int32_t)e;
    }
}

// restore edge-id order inside every row (rows are short): insertion sort, one thread per row
__global__ void k_sort_rows(const int32_t* __restrict__ rp0, int32_t* nb0, int32_t* eid0,
                            const int32_t* __restrict__ rp1, int32_t* nb1, int32_t* eid1, int64_t N) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < 2 * N; r += (int64_t)gridDim.x * blockDim.x) {
        const bool second = r >= N;
        const int64_t row = second ? r - N : r;
        const int32_t* rp = second ? rp1 : rp0;
        int32_t* nb = second ? nb1 : nb0;
        int32_t* eid = second ? eid1 : eid0;
        const int32_t b = rp[row], e = rp[row + 1];
        for (int32_t i = b + 1; i < e; ++i) {
            const int32_t ke = eid[i], kn = nb[i];
            int32_t j = i - 1;
            while (j >= b && eid[j] > ke) {
                eid[j + 1] = eid[j];
                nb[j + 1] = nb[j];
                --j;
            }
            eid[j + 1] = ke;
            nb[j + 1] = kn;
        }
    }
}

__global__ void k_edge_type(const int64_t* __restrict__ attr, int64_t E, int ncol, int m0, int m1, int m2,
                            int m3, int32_t* etype) {
    const int m[4] = {m0, m1, m2, m3};
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int32_t t = 0;
        for (int c = 0; c < ncol; ++c) t += (int32_t)attr[e * ncol + c] * m[c];
        etype[e] = t;
    }
}

// ---------------------------------------------------------------- edges sorted by type (counting sort)
// ntypes is small (<= 1024): block-private shared-memory histograms, one global atomic per (block, type).
constexpr int EBT_MAX_TYPES = 1024;
__global__ void __launch_bounds__(256)
k_ebt_hist(const int32_t* __restrict__ etype, int64_t E, int ntypes, int64_t chunk, int32_t* __restrict__ cnt) {
    __shared__ int32_t sh[EBT_MAX_TYPES];
    for (int i = threadIdx.x; i < ntypes; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int64_t e0 = blockIdx.x * chunk, e1 = min(e0 + chunk, E);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) atomicAdd(&sh[etype[e]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < ntypes; i += blockDim.x)
        if (sh[i]) atomicAdd(&cnt[i], sh[i]);
}
// single block: type_ptr = exclusive scan of cnt, cursor = copy of type_ptr
__global__ void k_ebt_scan(const int32_t* __restrict__ cnt, int ntypes, int32_t* __restrict__ type_ptr,
                           int32_t* __restrict__ cursor) {
    if (threadIdx.x == 0) {
        int32_t run = 0;
        for (int i = 0; i < ntypes; ++i) {
            type_ptr[i] = run;
            cursor[i] = run;
            run += cnt[i];
        }
        type_ptr[ntypes] = run;
    }
}
__global__ void __launch_bounds__(256)
k_ebt_fill(const int64_t* __restrict__ ei, const int32_t* __restrict__ etype, int64_t E, int ntypes, int64_t chunk,
           int32_t* __restrict__ cursor, int32_t* __restrict__ src_t, int32_t* __restrict__ dst_t,
           int32_t* __restrict__ type_t) {
    __shared__ int32_t sh_cnt[EBT_MAX_TYPES];
    __shared__ int32_t sh_base[EBT_MAX_TYPES];
    for (int i = threadIdx.x; i < ntypes; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const int64_t e0 = blockIdx.x * chunk, e1 = min(e0 + chunk, E);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) atomicAdd(&sh_cnt[etype[e]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < ntypes; i += blockDim.x) {
        sh_base[i] = sh_cnt[i] ? atomicAdd(&cursor[i], sh_cnt[i]) : 0;   // this block's slot range of type i
        sh_cnt[i] = 0;
    }
    __syncthreads();
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int ty = etype[e];
        const int32_t p = sh_base[ty] + atomicAdd(&sh_cnt[ty], 1);
        src_t[p] = (int32_t)ei[e];
        dst_t[p] = (int32_t)ei[E + e];
        type_t[p] = ty;
    }
}

// ---------------------------------------------------------------- batch plan
__global__ void k_plan_bounds(const int64_t* __restrict__ batch, int64_t N, int64_t B, int32_t* node_off,
                              int32_t* node_graph) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        // ids >= B mark trailing shape-bucket slack nodes (graphtrans_b200.loader.pad_to_bucket): they belong to no
        // graph (node_graph = -1, no tokens); node_off[B] = first slack node = number of real nodes
        const int64_t g = batch[i] < B ? batch[i] : B;
        node_graph[i] = g < B ? (int32_t)g : -1;
        if (i == 0 || batch[i - 1] != batch[i]) {
            if (g < B || i == 0 || batch[i - 1] < B) node_off[g] = (int32_t)i;
        }
    }
}

// single block: fix empty graphs, kept = min(n, L), tok_off = exclusive scan of kept+1, scalars
__global__ void k_plan_scan(int64_t N, int64_t B, int64_t L, int cls, int32_t* node_off, int32_t* kept,
                            int32_t* tok_off, int32_t* scalars) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s, max_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        if (node_off[B] < 0) node_off[B] = (int32_t)N;
        for (int64_t g = B - 1; g >= 0; --g)
            if (node_off[g] < 0) node_off[g] = node_off[g + 1];
        carry_s = 0;
        max_s = 0;
    }
    __syncthreads();
    int32_t local_max = 0;
    for (int64_t base = 0; base < B; base += blockDim.x) {
        const int64_t g = base + tid;
        int32_t n = 0, k = 0, v = 0;
        if (g < B) {
            n = node_off[g + 1] - node_off[g];
            k = n < L ? n : (int32_t)L;
            kept[g] = k;
            v = k + cls;
            local_max = max(local_max, n);
        }
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0, wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const int32_t excl = carry_s + warp_tot[wid] + incl - v;
        if (g < B) tok_off[g] = excl;
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = excl + v;
        __syncthreads();
    }
    atomicMax(&max_s, local_max);
    __syncthreads();
    if (tid == 0) {
        tok_off[B] = carry_s;
        scalars[0] = max_s < L ? max_s : (int32_t)L;  // S
        scalars[1] = carry_s;                         // n_tok
        scalars[2] = max_s;
        scalars[3] = 0;
    }
}

__global__ void k_plan_maps(int64_t N, int64_t B, int cls, const int32_t* __restrict__ node_off,
                            const int32_t* __restrict__ kept, const int32_t* __restrict__ tok_off,
                            const int32_t* __restrict__ node_graph, int32_t* tok2node, int32_t* tok_graph,
                            int32_t* node2tok, int32_t* cls_rows) {
    const int64_t R = N + B;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < R; t += (int64_t)gridDim.x * blockDim.x) {
        // token row -> graph by binary search over tok_off
        const int32_t ntok = tok_off[B];
        if (t >= ntok) {
            tok2node[t] = -2;
            tok_graph[t] = -1;
        } else {
            int64_t lo = 0, hi = B;  // largest g with tok_off[g] <= t
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (tok_off[mid] <= t) lo = mid; else hi = mid;
            }
            const int32_t g = (int32_t)lo, local = (int32_t)t - tok_off[g], k = kept[g];
            tok_graph[t] = g;
            if (local == k) {  // only reachable when cls != 0
                tok2node[t] = -1;
                cls_rows[g] = (int32_t)t;
            } else {
                const int32_t n = node_off[g + 1] - node_off[g];
                tok2node[t] = node_off[g] + (n - k) + local;
                if (!cls && local == k - 1) cls_rows[g] = (int32_t)t;  // pooling == "last": the last node row
            }
        }
        if (t < N) {
            const int32_t g = node_graph[t];
            if (g < 0) {
                node2tok[t] = -1;
            } else {
                const int32_t n = node_off[g + 1] - node_off[g], k = kept[g];
                const int32_t local = (int32_t)t - node_off[g], skip = n - k;
                node2tok[t] = local >= skip ? tok_off[g] + local - skip : -1;
            }
        }
    }
}

}  // namespace gt

using namespace gt;

extern "C" int gt_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr_dst,
                            int32_t* src_by_dst, int32_t* eid_by_dst, int32_t* rowptr_src,
                            int32_t* dst_by_src, int32_t* eid_by_src, int32_t* work, void* stream) {
    GT_CHECK_ARG(N > 0 && E >= 0 && N < (1ll << 31) && E < (1ll << 31), "gt_csr_build: bad sizes N=%lld E=%lld",
                 (long long)N, (long long)E);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* cnt_dst = work;
    int32_t* cnt_src = work + (N + 1);
    cudaError_t e = cudaMemsetAsync(work, 0, sizeof(int32_t) * 2 * (N + 1), st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_csr_build memset");
    if (E > 0) k_hist2<<<blocks_for(E, 256), 256, 0, st>>>(edge_index, E, cnt_dst, cnt_src);
    k_scan2<<<2, 1024, 0, st>>>(cnt_dst, rowptr_dst, cnt_src, rowptr_src, N);
    e = cudaMemsetAsync(work, 0, sizeof(int32_t) * 2 * (N + 1), st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_csr_build memset2");
    if (E > 0) {
        k_fill2<<<blocks_for(E, 256), 256, 0, st>>>(edge_index, E, rowptr_dst, rowptr_src, cnt_dst, cnt_src,
                                                    src_by_dst, eid_by_dst, dst_by_src, eid_by_src);
        k_sort_rows<<<blocks_for(2 * N, 128), 128, 0, st>>>(rowptr_dst, src_by_dst, eid_by_dst, rowptr_src,
                                                            dst_by_src, eid_by_src, N);
    }
    GT_LAUNCH_CHECK("gt_csr_build");
    return 0;
}

extern "C" int gt_edge_type(const int64_t* edge_attr, int64_t E, int32_t ncol, const int32_t* mult_host,
                            int32_t* etype, void* stream) {
    GT_CHECK_ARG(ncol >= 1 && ncol <= 4, "gt_edge_type: ncol=%d not in 1..4", ncol);
    int m[4] = {0, 0, 0, 0};
    for (int c = 0; c < ncol; ++c) m[c] = mult_host[c];
    if (E > 0)
        k_edge_type<<<blocks_for(E, 256), 256, 0, (cudaStream_t)stream>>>(edge_attr, E, ncol, m[0], m[1], m[2],
                                                                          m[3], etype);
    GT_LAUNCH_CHECK("gt_edge_type");
    return 0;
}

extern "C" int gt_edges_by_type(const int64_t* edge_index, const int32_t* etype, int64_t E, int32_t ntypes,
                                int32_t* type_ptr, int32_t* src_t, int32_t* dst_t, int32_t* type_t, int32_t* work,
                                void* stream) {
    GT_CHECK_ARG(E >= 0 && E < (1ll << 31) && ntypes > 0 && ntypes <= EBT_MAX_TYPES, "gt_edges_by_type: ntypes=%d not in 1..%d", ntypes, EBT_MAX_TYPES);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(work, 0, sizeof(int32_t) * 2 * (size_t)ntypes, st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_edges_by_type");
    int32_t* cnt = work;
    int32_t* cursor = work + ntypes;
    int64_t chunk = (E + 4 * kNumSMs - 1) / (4 * kNumSMs);
    if (chunk < 1024) chunk = 1024;
    const unsigned grid = (unsigned)((E + chunk - 1) / chunk);
    if (E > 0) k_ebt_hist<<<grid, 256, 0, st>>>(etype, E, ntypes, chunk, cnt);
    k_ebt_scan<<<1, 32, 0, st>>>(cnt, ntypes, type_ptr, cursor);
    if (E > 0) k_ebt_fill<<<grid, 256, 0, st>>>(edge_index, etype, E, ntypes, chunk, cursor, src_t, dst_t, type_t);
    GT_LAUNCH_CHECK("gt_edges_by_type");
    return 0;
}

extern "C" int gt_batch_plan(const int64_t* batch, int64_t N, int64_t B, int64_t L, int32_t cls, int32_t* node_off,
                             int32_t* kept, int32_t* tok_off, int32_t* tok2node, int32_t* tok_graph,
                             int32_t* node_graph, int32_t* node2tok, int32_t* cls_rows, int32_t* scalars,
                             void* stream) {
    GT_CHECK_ARG(N > 0 && B > 0 && L > 0, "gt_batch_plan: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(node_off, 0xff, sizeof(int32_t) * (B + 1), st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_batch_plan memset");
    k_plan_bounds<<<blocks_for(N, 256), 256, 0, st>>>(batch, N, B, node_off, node_graph);
    k_plan_scan<<<1, 1024, 0, st>>>(N, B, L, cls ? 1 : 0, node_off, kept, tok_off, scalars);
    k_plan_maps<<<blocks_for(N + B, 256), 256, 0, st>>>(N, B, cls ? 1 : 0, node_off, kept, tok_off, node_graph, tok2node,
                                                        tok_graph, node2tok, cls_rows);
    GT_LAUNCH_CHECK("gt_batch_plan");
    return 0;
}
