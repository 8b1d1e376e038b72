// Stage 1: GCN / GIN message-passing aggregation, forward and adjoint.
// Replaces PyG MessagePassing.propagate -> index_select + message() + torch_scatter atomics
// (reference modules/conv.py:26-33 GIN, :50-68 GCN).  HBM/L2-bound gather + segmented reduce:
// one warp owns a destination (forward) or source (adjoint) node, lanes own 4 consecutive
// channels (8/16-byte vector loads), the per-edge message relu(x_j + ee_e) is recomputed from a
// tiny edge-encoder table instead of materialising [E, d] tensors, and the sum is formed in
// registers in CSR order (deterministic, no atomics on the feature path).
#include "common.cuh"

namespace gt {

constexpr int AGG_WARPS = 8;
constexpr int MAX_KDIM = 4;
constexpr int AGG_TAB_ROWS = 64;   // ogb BondEncoder: 5*6*2 = 60 combined edge types

struct EdgeEnc {
    const float* attr;   // [E, kdim] fp32 (LINEAR)
    const float* w;      // [d, kdim]       (LINEAR; nn.Linear weight)
    const float* b;      // [d]             (LINEAR)
    const int32_t* etype;  // [E]           (TABLE)
    const float* table;  // [ntypes, ld]    (TABLE)
    int kdim;
    int ntypes;          // rows of `table` (adjoint only: <= AGG_TAB_ROWS rows are reduced in shared memory)
};

template <int EK>
struct EdgeRegs {  // per-lane edge-encoder parameters of the current 4-channel chunk
    float w[MAX_KDIM][4];
    float b[4];
    __device__ __forceinline__ void load(const EdgeEnc& en, int c0, int d) {
        if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = c0 + q < d;
                b[q] = ok ? en.b[c0 + q] : 0.f;
#pragma unroll
                for (int k = 0; k < MAX_KDIM; ++k) w[k][q] = (ok && k < en.kdim) ? en.w[(c0 + q) * en.kdim + k] : 0.f;
            }
        }
    }
    __device__ __forceinline__ void embed(const EdgeEnc& en, int eid, int c0, int ld, float (&ee)[4], float (&a)[MAX_KDIM]) const {
        if (EK == GT_EDGE_NONE) {
            ee[0] = ee[1] = ee[2] = ee[3] = 0.f;
        } else if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int k = 0; k < MAX_KDIM; ++k) a[k] = k < en.kdim ? en.attr[(int64_t)eid * en.kdim + k] : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = b[q];
#pragma unroll
                for (int k = 0; k < MAX_KDIM; ++k) v = fmaf(a[k], w[k][q], v);
                ee[q] = v;
            }
        } else {
            ld4(en.table + (int64_t)en.etype[eid] * ld + c0, ee);
        }
    }
};

template <typename T, int CONV, int EK>
__global__ void __launch_bounds__(AGG_WARPS * 32)
k_agg_fwd(const T* __restrict__ x, T* __restrict__ out, int N, int d, int ld,
          const int32_t* __restrict__ rp_dst, const int32_t* __restrict__ src_by_dst,
          const int32_t* __restrict__ eid_by_dst, const int32_t* __restrict__ rp_src, EdgeEnc en,
          const float* __restrict__ self_param) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * AGG_WARPS;
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    for (int c0 = lane * 4; c0 < ld; c0 += 128) {
        EdgeRegs<EK> er;
        er.load(en, c0, d);
        float root[4] = {0.f, 0.f, 0.f, 0.f};
        if (CONV == GT_CONV_GCN) {
#pragma unroll
            for (int q = 0; q < 4; ++q) root[q] = c0 + q < d ? self_param[c0 + q] : 0.f;
        }
        for (int i = warp; i < N; i += nwarps) {
            const int b = rp_dst[i], e = rp_dst[i + 1];
            float dis_i = 1.f, inv_deg_i = 1.f;
            if (CONV == GT_CONV_GCN) {
                const float deg_i = (float)(rp_src[i + 1] - rp_src[i] + 1);
                dis_i = rsqrtf(deg_i);
                inv_deg_i = 1.f / deg_i;
            }
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
            for (int p = b; p < e; ++p) {
                const int j = src_by_dst[p];
                float xv[4], ee[4], a[MAX_KDIM];
                ld4(x + (int64_t)j * ld + c0, xv);
                er.embed(en, EK == GT_EDGE_NONE ? 0 : eid_by_dst[p], c0, ld, ee, a);
                float nrm = 1.f;
                if (CONV == GT_CONV_GCN) nrm = dis_i * rsqrtf((float)(rp_src[j + 1] - rp_src[j] + 1));
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(nrm, fmaxf(xv[q] + ee[q], 0.f), acc[q]);
            }
            float xi[4];
            ld4(x + (int64_t)i * ld + c0, xi);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (CONV == GT_CONV_GCN) acc[q] = fmaf(fmaxf(xi[q] + root[q], 0.f), inv_deg_i, acc[q]);
                else acc[q] = fmaf(eps1, xi[q], acc[q]);
            }
            st4(out + (int64_t)i * ld + c0, acc);
        }
    }
}

// adjoint: one warp per SOURCE node j; dx[j] = sum over out-edges of mask * norm * dout[dst] + self term
template <typename T, int CONV, int EK>
__global__ void __launch_bounds__(AGG_WARPS * 32)
k_agg_bwd(const T* __restrict__ x, const T* __restrict__ dout, T* __restrict__ dx, int N, int d, int ld,
          const int32_t* __restrict__ rp_src, const int32_t* __restrict__ dst_by_src,
          const int32_t* __restrict__ eid_by_src, EdgeEnc en, const float* __restrict__ self_param,
          float* __restrict__ d_edge_w, float* __restrict__ d_edge_b, float* __restrict__ d_table,
          float* __restrict__ d_self) {
    __shared__ float sh_self[128];
    __shared__ float sh_b[128];
    __shared__ float sh_w[MAX_KDIM][128];
    // table-edge gradients: a handful of rows shared by every edge would serialise global atomics, so they are
    // reduced per block in shared memory (lanes own distinct channels: no intra-warp conflicts) and flushed once
    __shared__ float sh_tab[EK == GT_EDGE_TABLE ? AGG_TAB_ROWS * 128 : 1];
    const bool tab_smem = EK == GT_EDGE_TABLE && en.ntypes <= AGG_TAB_ROWS;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * AGG_WARPS;
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    float deps = 0.f;
    const int nchunks = (ld + 127) / 128;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * 128 + lane * 4;
        const bool active = c0 < ld;
        if (threadIdx.x < 128) {
            sh_self[threadIdx.x] = 0.f;
            sh_b[threadIdx.x] = 0.f;
#pragma unroll
            for (int k = 0; k < MAX_KDIM; ++k) sh_w[k][threadIdx.x] = 0.f;
        }
        if (tab_smem)
            for (int i = threadIdx.x; i < en.ntypes * 128; i += blockDim.x) sh_tab[i] = 0.f;
        __syncthreads();
        if (active) {
            EdgeRegs<EK> er;
            er.load(en, c0, d);
            float root[4] = {0.f, 0.f, 0.f, 0.f};
            if (CONV == GT_CONV_GCN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) root[q] = c0 + q < d ? self_param[c0 + q] : 0.f;
            }
            float a_self[4] = {0.f, 0.f, 0.f, 0.f}, a_b[4] = {0.f, 0.f, 0.f, 0.f};
            float a_w[MAX_KDIM][4];
#pragma unroll
            for (int k = 0; k < MAX_KDIM; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) a_w[k][q] = 0.f;
            for (int j = warp; j < N; j += nwarps) {
                const int b = rp_src[j], e = rp_src[j + 1];
                float dis_j = 1.f, inv_deg_j = 1.f;
                if (CONV == GT_CONV_GCN) {
                    const float deg_j = (float)(e - b + 1);
                    dis_j = rsqrtf(deg_j);
                    inv_deg_j = 1.f / deg_j;
                }
                float xj[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
                ld4(x + (int64_t)j * ld + c0, xj);
#pragma unroll 2
                for (int p = b; p < e; ++p) {
                    const int i = dst_by_src[p];
                    const int eid = EK == GT_EDGE_NONE ? 0 : eid_by_src[p];
                    float g[4], ee[4], a[MAX_KDIM];
                    ld4(dout + (int64_t)i * ld + c0, g);
                    er.embed(en, eid, c0, ld, ee, a);
                    float nrm = 1.f;
                    if (CONV == GT_CONV_GCN) nrm = dis_j * rsqrtf((float)(rp_src[i + 1] - rp_src[i] + 1));
                    float gm[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        gm[q] = (xj[q] + ee[q] > 0.f) ? nrm * g[q] : 0.f;
                        acc[q] += gm[q];
                    }
                    if (EK == GT_EDGE_LINEAR) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            a_b[q] += gm[q];
#pragma unroll
                            for (int k = 0; k < MAX_KDIM; ++k) a_w[k][q] = fmaf(a[k], gm[q], a_w[k][q]);
                        }
                    } else if (EK == GT_EDGE_TABLE) {
                        const int ty = en.etype[eid];
                        float* row = tab_smem ? sh_tab + ty * 128 + lane * 4 : d_table + (int64_t)ty * ld + c0;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (gm[q] != 0.f) atomicAdd(row + q, gm[q]);
                    }
                }
                float gj[4];
                ld4(dout + (int64_t)j * ld + c0, gj);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (CONV == GT_CONV_GCN) {
                        const float gs = (xj[q] + root[q] > 0.f) ? gj[q] * inv_deg_j : 0.f;
                        acc[q] += gs;
                        a_self[q] += gs;
                    } else {
                        acc[q] = fmaf(eps1, gj[q], acc[q]);
                        deps = fmaf(xj[q], gj[q], deps);
                    }
                }
                st4(dx + (int64_t)j * ld + c0, acc);
            }
            // block-level pre-reduction of the parameter gradients of this chunk
            if (CONV == GT_CONV_GCN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicAdd(&sh_self[lane * 4 + q], a_self[q]);
            }
            if (EK == GT_EDGE_LINEAR) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    atomicAdd(&sh_b[lane * 4 + q], a_b[q]);
#pragma unroll
                    for (int k = 0; k < MAX_KDIM; ++k)
                        if (k < en.kdim) atomicAdd(&sh_w[k][lane * 4 + q], a_w[k][q]);
                }
            }
        }
        __syncthreads();
        if (tab_smem) {
            for (int i = threadIdx.x; i < en.ntypes * 128; i += blockDim.x) {
                const int c = ch * 128 + (i & 127);
                const float v = sh_tab[i];
                if (c < d && v != 0.f) atomicAdd(d_table + (int64_t)(i >> 7) * ld + c, v);
            }
        }
        if (threadIdx.x < 128) {
            const int c = ch * 128 + threadIdx.x;
            if (c < d) {
                if (CONV == GT_CONV_GCN) atomicAdd(&d_self[c], sh_self[threadIdx.x]);
                if (EK == GT_EDGE_LINEAR) {
                    atomicAdd(&d_edge_b[c], sh_b[threadIdx.x]);
                    for (int k = 0; k < en.kdim; ++k) atomicAdd(&d_edge_w[c * en.kdim + k], sh_w[k][threadIdx.x]);
                }
            }
        }
        __syncthreads();
    }
    if (CONV == GT_CONV_GIN) {
        deps = warp_sum(deps);
        if (lane == 0 && deps != 0.f) atomicAdd(d_self, deps);
    }
}

template <typename T, int CONV>
static int launch_fwd(int ek, const T* x, T* out, int N, int d, int ld, const int32_t* rp_dst,
                      const int32_t* src_by_dst, const int32_t* eid_by_dst, const int32_t* rp_src,
                      EdgeEnc en, const float* self_param, cudaStream_t st) {
    const int grid = blocks_for(N, AGG_WARPS, kNumSMs * 8);
#define L(EK) k_agg_fwd<T, CONV, EK><<<grid, AGG_WARPS * 32, 0, st>>>(x, out, N, d, ld, rp_dst, src_by_dst, eid_by_dst, rp_src, en, self_param)
    if (ek == GT_EDGE_NONE) L(GT_EDGE_NONE);
    else if (ek == GT_EDGE_LINEAR) L(GT_EDGE_LINEAR);
    else L(GT_EDGE_TABLE);
#undef L
    return 0;
}

template <typename T, int CONV>
static int launch_bwd(int ek, const T* x, const T* dout, T* dx, int N, int d, int ld, const int32_t* rp_src,
                      const int32_t* dst_by_src, const int32_t* eid_by_src, EdgeEnc en,
                      const float* self_param, float* dw, float* db, float* dtab, float* dself,
                      cudaStream_t st) {
    const int grid = blocks_for(N, AGG_WARPS, kNumSMs * 8);
#define L(EK) k_agg_bwd<T, CONV, EK><<<grid, AGG_WARPS * 32, 0, st>>>(x, dout, dx, N, d, ld, rp_src, dst_by_src, eid_by_src, en, self_param, dw, db, dtab, dself)
    if (ek == GT_EDGE_NONE) L(GT_EDGE_NONE);
    else if (ek == GT_EDGE_LINEAR) L(GT_EDGE_LINEAR);
    else L(GT_EDGE_TABLE);
#undef L
    return 0;
}

static int check_common(const char* fn, int conv, int64_t N, int32_t d, int32_t ld, int ek, int kdim) {
    GT_CHECK_ARG(conv == GT_CONV_GCN || conv == GT_CONV_GIN, "%s: bad conv kind %d", fn, conv);
    GT_CHECK_ARG(N > 0 && N < (1ll << 31) && d > 0 && ld >= d && ld % 4 == 0, "%s: bad shape N=%lld d=%d ld=%d", fn,
                 (long long)N, d, ld);
    GT_CHECK_ARG(ek >= 0 && ek <= 2, "%s: bad edge kind %d", fn, ek);
    GT_CHECK_ARG(ek != GT_EDGE_LINEAR || (kdim >= 1 && kdim <= MAX_KDIM), "%s: kdim=%d not in 1..%d", fn, kdim, MAX_KDIM);
    return 0;
}

}  // namespace gt

using namespace gt;

extern "C" int gt_aggregate_fwd(int dt, int conv, const void* x, void* out, int64_t N, int32_t d, int32_t ld,
                                const int32_t* rowptr_dst, const int32_t* src_by_dst, const int32_t* eid_by_dst,
                                const int32_t* rowptr_src, int edge_kind, const float* edge_attr, int32_t kdim,
                                const float* edge_w, const float* edge_b, const int32_t* etype,
                                const float* table, const float* self_param, void* stream) {
    if (int r = check_common("gt_aggregate_fwd", conv, N, d, ld, edge_kind, kdim)) return r;
    EdgeEnc en{edge_attr, edge_w, edge_b, etype, table, kdim, 0};
    cudaStream_t st = (cudaStream_t)stream;
    GT_DISPATCH_DT(dt, {
        if (conv == GT_CONV_GCN)
            launch_fwd<T, GT_CONV_GCN>(edge_kind, (const T*)x, (T*)out, (int)N, d, ld, rowptr_dst, src_by_dst, eid_by_dst, rowptr_src, en, self_param, st);
        else
            launch_fwd<T, GT_CONV_GIN>(edge_kind, (const T*)x, (T*)out, (int)N, d, ld, rowptr_dst, src_by_dst, eid_by_dst, rowptr_src, en, self_param, st);
    });
    GT_LAUNCH_CHECK("gt_aggregate_fwd");
    return 0;
}

extern "C" int gt_aggregate_bwd(int dt, int conv, const void* x, const void* dout, void* dx, int64_t N, int32_t d,
                                int32_t ld, const int32_t* rowptr_dst, const int32_t* rowptr_src,
                                const int32_t* dst_by_src, const int32_t* eid_by_src, int edge_kind,
                                const float* edge_attr, int32_t kdim, const float* edge_w, const float* edge_b,
                                const int32_t* etype, const float* table, int32_t ntypes, const float* self_param,
                                float* d_edge_w, float* d_edge_b, float* d_table, float* d_self, void* stream) {
    (void)rowptr_dst;
    if (int r = check_common("gt_aggregate_bwd", conv, N, d, ld, edge_kind, kdim)) return r;
    GT_CHECK_ARG(edge_kind != GT_EDGE_TABLE || ntypes > 0, "gt_aggregate_bwd: ntypes must be the row count of the edge table");
    EdgeEnc en{edge_attr, edge_w, edge_b, etype, table, kdim, ntypes};
    cudaStream_t st = (cudaStream_t)stream;
    GT_DISPATCH_DT(dt, {
        if (conv == GT_CONV_GCN)
            launch_bwd<T, GT_CONV_GCN>(edge_kind, (const T*)x, (const T*)dout, (T*)dx, (int)N, d, ld, rowptr_src, dst_by_src, eid_by_src, en, self_param, d_edge_w, d_edge_b, d_table, d_self, st);
        else
            launch_bwd<T, GT_CONV_GIN>(edge_kind, (const T*)x, (const T*)dout, (T*)dx, (int)N, d, ld, rowptr_src, dst_by_src, eid_by_src, en, self_param, d_edge_w, d_edge_b, d_table, d_self, st);
    });
    GT_LAUNCH_CHECK("gt_aggregate_bwd");
    return 0;
}
