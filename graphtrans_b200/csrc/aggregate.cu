// Stage 1: GCN / GIN message-passing aggregation, forward and adjoint.
// Replaces PyG MessagePassing.propagate -> index_select + message() + torch_scatter atomics
// (reference modules/conv.py:26-33 GIN, :50-68 GCN).  HBM/L2-bound gather + segmented reduce:
// one warp owns a destination (forward) or source (adjoint) node, lanes own 4 consecutive
// channels (8/16-byte vector loads), the per-edge message relu(x_j + ee_e) is recomputed from a
// tiny edge-encoder table instead of materialising [E, d] tensors, and the sum is formed in
// registers in CSR order (deterministic, no atomics on the feature path).
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

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
    // optional per-CSR-slot copies (gt_edge_slots, once per batch, reused by every layer): they remove the dependent
    // "slot -> edge id -> attribute" / "slot -> neighbour -> degree" loads from the per-edge chain
    const float* norm_slot;     // GCN: deg(src)^-1/2 deg(dst)^-1/2 of the edge in this slot
    const int32_t* etype_slot;  // TABLE: combined edge type of the edge in this slot
    const float* attr_slot;     // LINEAR: [E, kdim] attributes in slot order
};

template <int EK, int KD = MAX_KDIM>
struct EdgeRegs {  // per-lane edge-encoder parameters of the current 4-channel chunk (KD = compiled attribute width)
    float w[KD][4];
    float b[4];
    __device__ __forceinline__ void load(const EdgeEnc& en, int c0, int d) {
        if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = c0 + q < d;
                b[q] = ok ? en.b[c0 + q] : 0.f;
#pragma unroll
                for (int k = 0; k < KD; ++k) w[k][q] = (ok && k < en.kdim) ? en.w[(c0 + q) * en.kdim + k] : 0.f;
            }
        }
    }
    // same with the edge attributes / type already fetched (edge-batched kernels)
    __device__ __forceinline__ void embed_pre(const EdgeEnc& en, const float (&a)[KD], int ty, int c0, int ld, float (&ee)[4]) const {
        if (EK == GT_EDGE_NONE) {
            ee[0] = ee[1] = ee[2] = ee[3] = 0.f;
        } else if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = b[q];
#pragma unroll
                for (int k = 0; k < KD; ++k) v = fmaf(a[k], w[k][q], v);
                ee[q] = v;
            }
        } else {
            ld4(en.table + (int64_t)ty * ld + c0, ee);
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
    const bool tab_grad = EK == GT_EDGE_TABLE && d_table != nullptr;   // null: gt_aggregate_table_grad computes it
    const bool tab_smem = tab_grad && en.ntypes <= AGG_TAB_ROWS;
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
                    } else if (EK == GT_EDGE_TABLE && tab_grad) {
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

// ---------------------------------------------------------------------------------------------------------------
// Edge-batched variants (ld <= 128 * NCH, NCH <= 4): the per-edge chain "edge slot -> neighbour id -> degree / edge
// attribute -> feature row" is a sequence of dependent L2 round trips when one warp walks its edges one by one.  Here
// the 32 lanes first fetch up to 32 edges of the node IN PARALLEL (neighbour ids, GCN norms, edge attributes / types),
// then the warp walks them with register shuffles, and every edge issues the loads of ALL channel chunks of the
// neighbour row back to back (the whole 600-byte row per edge, several edges in flight through the unrolled loop).
template <typename T, int CONV, int EK, int NCH, int KD>
__global__ void __launch_bounds__(AGG_WARPS * 32)
k_agg_fwd2(const T* __restrict__ x, T* __restrict__ out, int N, int d, int ld,
           const int32_t* __restrict__ rp_dst, const int32_t* __restrict__ src_by_dst,
           const int32_t* __restrict__ eid_by_dst, const int32_t* __restrict__ rp_src, EdgeEnc en,
           const float* __restrict__ self_param, int nch) {
    // NCH = 1: a warp owns ONE 128-channel chunk (warp % nch) of its nodes, neighbouring warps take the other chunks of
    // the same node (keeps the register footprint small enough for 2-3 resident blocks per SM)
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int cbase = (gwarp % nch) * 128;
    const int warp = gwarp / nch;
    const int nwarps = gridDim.x * AGG_WARPS / nch;
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    EdgeRegs<EK, KD> er[NCH];
    float root[NCH][4];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c0 = cbase + k * 128 + lane * 4;
        if (c0 < ld) er[k].load(en, c0, d);
#pragma unroll
        for (int q = 0; q < 4; ++q) root[k][q] = (CONV == GT_CONV_GCN && c0 + q < d) ? self_param[c0 + q] : 0.f;
    }
    for (int i = warp; i < N; i += nwarps) {
        const int b = rp_dst[i], e = rp_dst[i + 1];
        float dis_i = 1.f, inv_deg_i = 1.f;
        if (CONV == GT_CONV_GCN) {
            const float deg_i = (float)(rp_src[i + 1] - rp_src[i] + 1);
            dis_i = rsqrtf(deg_i);
            inv_deg_i = 1.f / deg_i;
        }
        float acc[NCH][4];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[k][q] = 0.f;
        for (int base = b; base < e; base += 32) {
            const int p = base + lane;
            const bool have = p < e;
            const int j_l = have ? src_by_dst[p] : 0;
            float nrm_l = 1.f, a_l[KD] = {};
            int ty_l = 0;
            if (have) {
                if (CONV == GT_CONV_GCN)
                    nrm_l = en.norm_slot ? en.norm_slot[p] : dis_i * rsqrtf((float)(rp_src[j_l + 1] - rp_src[j_l] + 1));
                if (EK == GT_EDGE_LINEAR) {
                    const float* ap = en.attr_slot ? en.attr_slot + (int64_t)p * en.kdim : en.attr + (int64_t)eid_by_dst[p] * en.kdim;
#pragma unroll
                    for (int k = 0; k < KD; ++k)
                        if (k < en.kdim) a_l[k] = ap[k];
                } else if (EK == GT_EDGE_TABLE) {
                    ty_l = en.etype_slot ? en.etype_slot[p] : en.etype[eid_by_dst[p]];
                }
            }
            const int cnt = min(32, e - base);
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                const int j = __shfl_sync(0xffffffffu, j_l, t);
                const float nrm = CONV == GT_CONV_GCN ? __shfl_sync(0xffffffffu, nrm_l, t) : 1.f;
                float a[KD] = {};
                int ty = 0;
                if (EK == GT_EDGE_LINEAR) {
#pragma unroll
                    for (int k = 0; k < KD; ++k) a[k] = __shfl_sync(0xffffffffu, a_l[k], t);
                } else if (EK == GT_EDGE_TABLE) {
                    ty = __shfl_sync(0xffffffffu, ty_l, t);
                }
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int c0 = cbase + k * 128 + lane * 4;
                    if (c0 < ld) {
                        float xv[4], ee[4];
                        ld4(x + (int64_t)j * ld + c0, xv);
                        er[k].embed_pre(en, a, ty, c0, ld, ee);
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[k][q] = fmaf(nrm, fmaxf(xv[q] + ee[q], 0.f), acc[k][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c0 = cbase + k * 128 + lane * 4;
            if (c0 < ld) {
                float xi[4];
                ld4(x + (int64_t)i * ld + c0, xi);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (CONV == GT_CONV_GCN) acc[k][q] = fmaf(fmaxf(xi[q] + root[k][q], 0.f), inv_deg_i, acc[k][q]);
                    else acc[k][q] = fmaf(eps1, xi[q], acc[k][q]);
                }
                st4(out + (int64_t)i * ld + c0, acc[k]);
            }
        }
    }
}

// adjoint, edge-batched: one warp per SOURCE node j, its out-edges fetched 32 at a time.  Edge-table gradients are
// reduced per block in a dynamic shared-memory table [ntypes][ld] (lanes own distinct channels: conflict-free inside a
// warp); Linear edge-encoder gradients in registers, then shared memory, then one global atomic per (block, element).
template <typename T, int CONV, int EK, int NCH, int KD>
__global__ void __launch_bounds__(AGG_WARPS * 32)
k_agg_bwd2(const T* __restrict__ x, const T* __restrict__ dout, T* __restrict__ dx, int N, int d, int ld,
           const int32_t* __restrict__ rp_src, const int32_t* __restrict__ dst_by_src,
           const int32_t* __restrict__ eid_by_src, EdgeEnc en, const float* __restrict__ self_param,
           float* __restrict__ d_edge_w, float* __restrict__ d_edge_b, float* __restrict__ d_table,
           float* __restrict__ d_self, int nch) {
    extern __shared__ float sh_dyn[];   // TABLE: [ntypes][ld];  LINEAR / GCN: [(2 + KD)][128 * nch]
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int cbase = (gwarp % nch) * 128;   // this warp's 128-channel chunk (see k_agg_fwd2)
    const int warp = gwarp / nch;
    const int nwarps = gridDim.x * AGG_WARPS / nch;
    const int W = 128 * nch;
    const bool tab_grad = EK == GT_EDGE_TABLE && d_table != nullptr;   // null: gt_aggregate_table_grad computes it
    float* sh_tab = sh_dyn;
    float* sh_par = sh_dyn + (tab_grad ? en.ntypes * ld : 0);   // [self | b | w0..w3][W]
    const int n_par = (2 + 4) * W;   // host-side layout: [self | b | w0..w3][W]
    for (int i = threadIdx.x; i < (tab_grad ? en.ntypes * ld : 0) + n_par; i += blockDim.x) sh_dyn[i] = 0.f;
    __syncthreads();
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    float deps = 0.f;
    EdgeRegs<EK, KD> er[NCH];
    float root[NCH][4], a_self[NCH][4], a_b[NCH][4], a_w[NCH][KD][4];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c0 = cbase + k * 128 + lane * 4;
        if (c0 < ld) er[k].load(en, c0, d);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            root[k][q] = (CONV == GT_CONV_GCN && c0 + q < d) ? self_param[c0 + q] : 0.f;
            a_self[k][q] = a_b[k][q] = 0.f;
#pragma unroll
            for (int kk = 0; kk < KD; ++kk) a_w[k][kk][q] = 0.f;
        }
    }
    for (int j = warp; j < N; j += nwarps) {
        const int b = rp_src[j], e = rp_src[j + 1];
        float dis_j = 1.f, inv_deg_j = 1.f;
        if (CONV == GT_CONV_GCN) {
            const float deg_j = (float)(e - b + 1);
            dis_j = rsqrtf(deg_j);
            inv_deg_j = 1.f / deg_j;
        }
        float xj[NCH][4], acc[NCH][4];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c0 = cbase + k * 128 + lane * 4;
            if (c0 < ld) ld4(x + (int64_t)j * ld + c0, xj[k]);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[k][q] = 0.f;
        }
        for (int base = b; base < e; base += 32) {
            const int p = base + lane;
            const bool have = p < e;
            const int i_l = have ? dst_by_src[p] : 0;
            float nrm_l = 1.f, a_l[KD] = {};
            int ty_l = 0;
            if (have) {
                if (CONV == GT_CONV_GCN)
                    nrm_l = en.norm_slot ? en.norm_slot[p] : dis_j * rsqrtf((float)(rp_src[i_l + 1] - rp_src[i_l] + 1));
                if (EK == GT_EDGE_LINEAR) {
                    const float* ap = en.attr_slot ? en.attr_slot + (int64_t)p * en.kdim : en.attr + (int64_t)eid_by_src[p] * en.kdim;
#pragma unroll
                    for (int k = 0; k < KD; ++k)
                        if (k < en.kdim) a_l[k] = ap[k];
                } else if (EK == GT_EDGE_TABLE) {
                    ty_l = en.etype_slot ? en.etype_slot[p] : en.etype[eid_by_src[p]];
                }
            }
            const int cnt = min(32, e - base);
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                const int i = __shfl_sync(0xffffffffu, i_l, t);
                const float nrm = CONV == GT_CONV_GCN ? __shfl_sync(0xffffffffu, nrm_l, t) : 1.f;
                float a[KD] = {};
                int ty = 0;
                if (EK == GT_EDGE_LINEAR) {
#pragma unroll
                    for (int k = 0; k < KD; ++k) a[k] = __shfl_sync(0xffffffffu, a_l[k], t);
                } else if (EK == GT_EDGE_TABLE) {
                    ty = __shfl_sync(0xffffffffu, ty_l, t);
                }
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int c0 = cbase + k * 128 + lane * 4;
                    if (c0 < ld) {
                        float g[4], ee[4], gm[4];
                        ld4(dout + (int64_t)i * ld + c0, g);
                        er[k].embed_pre(en, a, ty, c0, ld, ee);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            gm[q] = (xj[k][q] + ee[q] > 0.f) ? nrm * g[q] : 0.f;
                            acc[k][q] += gm[q];
                        }
                        if (EK == GT_EDGE_LINEAR) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                a_b[k][q] += gm[q];
#pragma unroll
                                for (int kk = 0; kk < KD; ++kk) a_w[k][kk][q] = fmaf(a[kk], gm[q], a_w[k][kk][q]);
                            }
                        } else if (EK == GT_EDGE_TABLE && tab_grad) {
                            float* row = sh_tab + ty * ld + c0;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (gm[q] != 0.f) atomicAdd(row + q, gm[q]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c0 = cbase + k * 128 + lane * 4;
            if (c0 < ld) {
                float gj[4];
                ld4(dout + (int64_t)j * ld + c0, gj);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (CONV == GT_CONV_GCN) {
                        const float gs = (xj[k][q] + root[k][q] > 0.f) ? gj[q] * inv_deg_j : 0.f;
                        acc[k][q] += gs;
                        a_self[k][q] += gs;
                    } else {
                        acc[k][q] = fmaf(eps1, gj[q], acc[k][q]);
                        deps = fmaf(xj[k][q], gj[q], deps);
                    }
                }
                st4(dx + (int64_t)j * ld + c0, acc[k]);
            }
        }
    }
    // block-level reduction of the parameter gradients, then one global atomic per (block, element)
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int cw = cbase + k * 128 + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (CONV == GT_CONV_GCN) atomicAdd(&sh_par[cw + q], a_self[k][q]);
            if (EK == GT_EDGE_LINEAR) {
                atomicAdd(&sh_par[W + cw + q], a_b[k][q]);
#pragma unroll
                for (int kk = 0; kk < KD; ++kk)
                    if (kk < en.kdim) atomicAdd(&sh_par[(2 + kk) * W + cw + q], a_w[k][kk][q]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        if (CONV == GT_CONV_GCN) atomicAdd(&d_self[c], sh_par[c]);
        if (EK == GT_EDGE_LINEAR) {
            atomicAdd(&d_edge_b[c], sh_par[W + c]);
            for (int kk = 0; kk < en.kdim; ++kk) atomicAdd(&d_edge_w[c * en.kdim + kk], sh_par[(2 + kk) * W + c]);
        }
    }
    if (tab_grad) {
        for (int i = threadIdx.x; i < en.ntypes * ld; i += blockDim.x) {
            const float v = sh_tab[i];
            if (v != 0.f && (i % ld) < d) atomicAdd(d_table + i, v);
        }
    }
    if (CONV == GT_CONV_GIN) {
        // d eps: ONE global atomic per block (every warp hitting the single address serialises ~10^4 L2 atomics,
        // which made the adjoint twice as slow as the forward)
        __shared__ float sh_eps[AGG_WARPS];
        deps = warp_sum(deps);
        if (lane == 0) sh_eps[threadIdx.x >> 5] = deps;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < AGG_WARPS; ++w) t += sh_eps[w];
            if (t != 0.f) atomicAdd(d_self, t);
        }
    }
}

// Edge-table gradient, split off the adjoint (gt_aggregate_table_grad): d_table[ty] = sum over the edges of type ty
// of norm_e * dout[dst_e] * 1[x[src_e] + table[ty] > 0].  Inside the adjoint it costs one shared-memory atomic per
// (edge, channel) plus a block-private [ntypes][ld] table (3.5x the forward on the molpcba batch); it is a LEAF
// gradient, so it runs here, off the critical path, over the edges SORTED BY TYPE (gt_edges_by_type, once per batch):
// a warp owns (32 consecutive sorted slots, one 128-channel chunk), keeps the table row of the current type and the
// running sum in registers, gathers TG_U edges' rows at a time (all loads issued before the first use) and issues
// one global atomic per (type run, channel).
constexpr int TG_U = 8;
template <typename T, int CONV>
__global__ void __launch_bounds__(256)
k_agg_table_grad(const T* __restrict__ x, const T* __restrict__ dout, int d, int ld, int nch,
                 const int32_t* __restrict__ rp_src, int64_t E, const int32_t* __restrict__ src_t,
                 const int32_t* __restrict__ dst_t, const int32_t* __restrict__ type_t,
                 const float* __restrict__ table, float* __restrict__ d_table) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t p0 = (item / nch) * 32;
    const int c0 = (int)(item % nch) * 128 + lane * 4;
    if (p0 >= E) return;
    const bool col_ok = c0 < ld;
    const int cnt = (int)min((int64_t)32, E - p0);
    const bool have = lane < cnt;
    const int j_l = have ? src_t[p0 + lane] : 0, i_l = have ? dst_t[p0 + lane] : 0;
    const int ty_l = have ? type_t[p0 + lane] : -1;
    float nrm_l = 1.f;
    if (CONV == GT_CONV_GCN && have)
        nrm_l = rsqrtf((float)(rp_src[j_l + 1] - rp_src[j_l] + 1)) * rsqrtf((float)(rp_src[i_l + 1] - rp_src[i_l] + 1));
    int cur = -1;
    float tab[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = 0; t0 < cnt; t0 += TG_U) {
        int ty[TG_U];
        float nrm[TG_U], xv[TG_U][4], g[TG_U][4];
#pragma unroll
        for (int u = 0; u < TG_U; ++u) {
            const int t = min(t0 + u, cnt - 1);               // tail slots repeat the last edge with weight 0
            const int j = __shfl_sync(0xffffffffu, j_l, t), i = __shfl_sync(0xffffffffu, i_l, t);
            ty[u] = __shfl_sync(0xffffffffu, ty_l, t);
            nrm[u] = CONV == GT_CONV_GCN ? __shfl_sync(0xffffffffu, nrm_l, t) : 1.f;
            if (t0 + u >= cnt) nrm[u] = 0.f;
            if (col_ok) {
                ld4(x + (int64_t)j * ld + c0, xv[u]);
                ld4(dout + (int64_t)i * ld + c0, g[u]);
            }
        }
        if (!col_ok) continue;
#pragma unroll
        for (int u = 0; u < TG_U; ++u) {
            if (ty[u] != cur) {                                // warp-uniform, rare: the slots are sorted by type
                if (cur >= 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c0 + q < d && acc[q] != 0.f) atomicAdd(d_table + (int64_t)cur * ld + c0 + q, acc[q]);
                }
                cur = ty[u];
                ld4(table + (int64_t)cur * ld + c0, tab);
                acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (xv[u][q] + tab[q] > 0.f) acc[q] = fmaf(nrm[u], g[u][q], acc[q]);
        }
    }
    if (col_ok && cur >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c0 + q < d && acc[q] != 0.f) atomicAdd(d_table + (int64_t)cur * ld + c0 + q, acc[q]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// v3: THREAD per (node, 8-channel vector).  blockDim = (ld / 8, nodes per block): the threads of a node read one
// feature row with consecutive 16-byte loads (a fully coalesced 512-608 byte row), every thread walks the CSR slots of
// its node itself (the neighbour ids / norms / edge types are the same address for all threads of the node: one
// broadcast transaction), gathers AGG_U neighbour rows per batch with all loads issued before the first use, and keeps
// the segmented sum of its 8 channels in registers.  Against the warp-per-(node, 128-channel chunk) kernels above this
// removes the shuffle traffic and the per-chunk repetition of the index work (4x fewer instructions per channel on
// the 2-edge-per-node molecule batches, where the per-node overhead dominates) and keeps every lane busy at
// ld = 304 (three 128-channel chunks left 62 % of the last chunk's lanes idle).
constexpr int AGG_U = 4;

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
    t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(p) = t;
}

template <int EK, int KD>
struct EdgeVec {   // edge-encoder parameters of this thread's 8 channels
    float w[KD][8], b[8];
    __device__ __forceinline__ void load(const EdgeEnc& en, int c0, int d) {
        if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const bool ok = c0 + q < d;
                b[q] = ok ? en.b[c0 + q] : 0.f;
#pragma unroll
                for (int k = 0; k < KD; ++k) w[k][q] = (ok && k < en.kdim) ? en.w[(c0 + q) * en.kdim + k] : 0.f;
            }
        }
    }
    __device__ __forceinline__ void embed(const EdgeEnc& en, const float (&a)[KD], int ty, int c0, int ld, float (&ee)[8]) const {
        if (EK == GT_EDGE_NONE) {
#pragma unroll
            for (int q = 0; q < 8; ++q) ee[q] = 0.f;
        } else if (EK == GT_EDGE_LINEAR) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float v = b[q];
#pragma unroll
                for (int k = 0; k < KD; ++k) v = fmaf(a[k], w[k][q], v);
                ee[q] = v;
            }
        } else {
            ld8(en.table + (int64_t)ty * ld + c0, ee);
        }
    }
};

// per-slot edge data of slot p (slot order = the CSR being walked)
template <int CONV, int EK, int KD>
__device__ __forceinline__ void slot_data(const EdgeEnc& en, const int32_t* __restrict__ nbr, const int32_t* __restrict__ eid_slot,
                                          const int32_t* __restrict__ rp_src, int p, int owner, float dis_owner, int& other,
                                          float& nrm, float (&a)[KD], int& ty) {
    other = nbr[p];
    nrm = 1.f;
    ty = 0;
    if (CONV == GT_CONV_GCN)
        nrm = en.norm_slot ? en.norm_slot[p] : dis_owner * rsqrtf((float)(rp_src[other + 1] - rp_src[other] + 1));
    if (EK == GT_EDGE_LINEAR) {
        const float* ap = en.attr_slot ? en.attr_slot + (int64_t)p * en.kdim : en.attr + (int64_t)eid_slot[p] * en.kdim;
#pragma unroll
        for (int k = 0; k < KD; ++k) a[k] = k < en.kdim ? ap[k] : 0.f;
    } else if (EK == GT_EDGE_TABLE) {
        ty = en.etype_slot ? en.etype_slot[p] : en.etype[eid_slot[p]];
    }
    (void)owner;
}

// first AGG_U CSR slots of a node (slots [b, e); tail slots repeat a valid slot with weight 0)
template <int KD>
struct SlotBatch {
    int other[AGG_U], ty[AGG_U];
    float nrm[AGG_U], a[AGG_U][KD];
};
template <int CONV, int EK, int KD>
__device__ __forceinline__ void load_slots(SlotBatch<KD>& sb, const EdgeEnc& en, const int32_t* __restrict__ nbr,
                                           const int32_t* __restrict__ eid_slot, const int32_t* __restrict__ rp_src, int p0, int e,
                                           int owner, float dis_owner) {
#pragma unroll
    for (int u = 0; u < AGG_U; ++u) {
        const int p = max(min(p0 + u, e - 1), 0);
        slot_data<CONV, EK, KD>(en, nbr, eid_slot, rp_src, p, owner, dis_owner, sb.other[u], sb.nrm[u], sb.a[u], sb.ty[u]);
        if (p0 + u >= e) sb.nrm[u] = 0.f;
    }
}

template <typename T, int CONV, int EK, int KD>
__global__ void __launch_bounds__(256)
k_agg_fwd3(const T* __restrict__ x, T* __restrict__ out, int N, int d, int ld,
           const int32_t* __restrict__ rp_dst, const int32_t* __restrict__ src_by_dst,
           const int32_t* __restrict__ eid_by_dst, const int32_t* __restrict__ rp_src, EdgeEnc en,
           const float* __restrict__ self_param) {
    const int c0 = threadIdx.x * 8;
    EdgeVec<EK, KD> ev;
    ev.load(en, c0, d);
    float root[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) root[q] = (CONV == GT_CONV_GCN && c0 + q < d) ? self_param[c0 + q] : 0.f;
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    // Software pipeline over the nodes of a thread (grid-stride): the row pointers are fetched two nodes ahead and the
    // first AGG_U slots (neighbour id, norm, edge type / attributes) one node ahead, so the only exposed latency per
    // node is the feature-row gather itself instead of the chain rowptr -> slot -> row.
    const int stride = gridDim.x * blockDim.y;
    int i = blockIdx.x * blockDim.y + threadIdx.y;
    int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    if (i < N) b0 = rp_dst[i], e0 = rp_dst[i + 1];
    if (i + stride < N) b1 = rp_dst[i + stride], e1 = rp_dst[i + stride + 1];
    SlotBatch<KD> cur;
    if (i < N) load_slots<CONV, EK, KD>(cur, en, src_by_dst, eid_by_dst, rp_src, b0, e0, i, 1.f);
    for (; i < N; i += stride) {
        float xi[8], xv[AGG_U][8];
        ld8(x + (int64_t)i * ld + c0, xi);
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) ld8(x + (int64_t)cur.other[u] * ld + c0, xv[u]);
        // prefetch for the next two nodes while the gathers are in flight
        SlotBatch<KD> nxt;
        const int inext = i + stride;
        if (inext < N) load_slots<CONV, EK, KD>(nxt, en, src_by_dst, eid_by_dst, rp_src, b1, e1, inext, 1.f);
        int b2 = 0, e2 = 0;
        if (inext + stride < N) b2 = rp_dst[inext + stride], e2 = rp_dst[inext + stride + 1];
        float inv_deg_i = 1.f;
        if (CONV == GT_CONV_GCN) inv_deg_i = 1.f / (float)(rp_src[i + 1] - rp_src[i] + 1);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) {
            float ee[8];
            ev.embed(en, cur.a[u], cur.ty[u], c0, ld, ee);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(cur.nrm[u], fmaxf(xv[u][q] + ee[q], 0.f), acc[q]);
        }
        for (int p0 = b0 + AGG_U; p0 < e0; p0 += AGG_U) {         // nodes with more than AGG_U in-edges
            SlotBatch<KD> sb;
            load_slots<CONV, EK, KD>(sb, en, src_by_dst, eid_by_dst, rp_src, p0, e0, i, 1.f);
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) ld8(x + (int64_t)sb.other[u] * ld + c0, xv[u]);
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) {
                float ee[8];
                ev.embed(en, sb.a[u], sb.ty[u], c0, ld, ee);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaf(sb.nrm[u], fmaxf(xv[u][q] + ee[q], 0.f), acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (CONV == GT_CONV_GCN) acc[q] = fmaf(fmaxf(xi[q] + root[q], 0.f), inv_deg_i, acc[q]);
            else acc[q] = fmaf(eps1, xi[q], acc[q]);
        }
        st8(out + (int64_t)i * ld + c0, acc);
        cur = nxt;
        b0 = b1, e0 = e1, b1 = b2, e1 = e2;
    }
}

// adjoint, thread per (source node j, 8-channel vector): dx[j] = sum over out-edges of mask * norm * dout[dst] + self
// term.  Parameter gradients (GCN root, Linear edge encoder, GIN eps) accumulate in registers over the nodes a thread
// visits (its channel vector is fixed), meet in shared memory once per block and leave as one global atomic per
// (block, element).  The edge-TABLE gradient is the separate type-sorted kernel (gt_aggregate_table_grad).
template <typename T, int CONV, int EK, int KD>
__global__ void __launch_bounds__(256, EK == GT_EDGE_LINEAR ? 1 : 3)     // table / no encoder: three blocks per SM
k_agg_bwd3(const T* __restrict__ x, const T* __restrict__ dout, T* __restrict__ dx, int N, int d, int ld,
           const int32_t* __restrict__ rp_src, const int32_t* __restrict__ dst_by_src,
           const int32_t* __restrict__ eid_by_src, EdgeEnc en, const float* __restrict__ self_param,
           float* __restrict__ d_edge_w, float* __restrict__ d_edge_b, float* __restrict__ d_self, T* __restrict__ gm_out) {
    // gm_out (optional, [E, ld] in source-sorted slot order): the per-edge masked message gradient norm * dout[dst] *
    // 1[x[src] + e > 0], written for the edge-table gradient as a contraction (OneHot(type)^T . gm on the tensor cores)
    extern __shared__ float sh_par[];    // [self | b | w0..w(KD-1)][ld], then one eps slot per warp
    const int c0 = threadIdx.x * 8;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    constexpr bool PAR = CONV == GT_CONV_GCN || EK == GT_EDGE_LINEAR;
    if (PAR) {
        for (int i = tid; i < (2 + KD) * ld; i += nthr) sh_par[i] = 0.f;
        __syncthreads();
    }
    EdgeVec<EK, KD> ev;
    ev.load(en, c0, d);
    float root[8], a_self[8], a_b[8], a_w[KD][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        root[q] = (CONV == GT_CONV_GCN && c0 + q < d) ? self_param[c0 + q] : 0.f;
        a_self[q] = a_b[q] = 0.f;
#pragma unroll
        for (int k = 0; k < KD; ++k) a_w[k][q] = 0.f;
    }
    const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
    float deps = 0.f;
    // software pipeline as in the forward: row pointers two nodes ahead, the first AGG_U out-edge slots one node ahead
    const int stride = gridDim.x * blockDim.y;
    int j = blockIdx.x * blockDim.y + threadIdx.y;
    int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    if (j < N) b0 = rp_src[j], e0 = rp_src[j + 1];
    if (j + stride < N) b1 = rp_src[j + stride], e1 = rp_src[j + stride + 1];
    SlotBatch<KD> cur;
    if (j < N) load_slots<CONV, EK, KD>(cur, en, dst_by_src, eid_by_src, rp_src, b0, e0, j, 1.f);
    for (; j < N; j += stride) {
        float xj[8], gj[8], g[AGG_U][8];
        ld8(x + (int64_t)j * ld + c0, xj);
        ld8(dout + (int64_t)j * ld + c0, gj);
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) ld8(dout + (int64_t)cur.other[u] * ld + c0, g[u]);
        SlotBatch<KD> nxt;
        const int jnext = j + stride;
        if (jnext < N) load_slots<CONV, EK, KD>(nxt, en, dst_by_src, eid_by_src, rp_src, b1, e1, jnext, 1.f);
        int b2 = 0, e2 = 0;
        if (jnext + stride < N) b2 = rp_src[jnext + stride], e2 = rp_src[jnext + stride + 1];
        const float inv_deg_j = CONV == GT_CONV_GCN ? 1.f / (float)(e0 - b0 + 1) : 1.f;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto consume = [&](const SlotBatch<KD>& sb, int pbase) {
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) {
                float ee[8], gmv[8];
                ev.embed(en, sb.a[u], sb.ty[u], c0, ld, ee);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float gm = (xj[q] + ee[q] > 0.f) ? sb.nrm[u] * g[u][q] : 0.f;
                    gmv[q] = gm;
                    acc[q] += gm;
                    if (EK == GT_EDGE_LINEAR) {
                        a_b[q] += gm;
#pragma unroll
                        for (int k = 0; k < KD; ++k) a_w[k][q] = fmaf(sb.a[u][k], gm, a_w[k][q]);
                    }
                }
                if (gm_out && pbase + u < e0) st8(gm_out + (int64_t)(pbase + u) * ld + c0, gmv);
            }
        };
        consume(cur, b0);
        for (int p0 = b0 + AGG_U; p0 < e0; p0 += AGG_U) {         // nodes with more than AGG_U out-edges
            SlotBatch<KD> sb;
            load_slots<CONV, EK, KD>(sb, en, dst_by_src, eid_by_src, rp_src, p0, e0, j, 1.f);
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) ld8(dout + (int64_t)sb.other[u] * ld + c0, g[u]);
            consume(sb, p0);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (CONV == GT_CONV_GCN) {
                const float gs = (xj[q] + root[q] > 0.f) ? gj[q] * inv_deg_j : 0.f;
                acc[q] += gs;
                a_self[q] += gs;
            } else {
                acc[q] = fmaf(eps1, gj[q], acc[q]);
                deps = fmaf(xj[q], gj[q], deps);
            }
        }
        st8(dx + (int64_t)j * ld + c0, acc);
        cur = nxt;
        b0 = b1, e0 = e1, b1 = b2, e1 = e2;
    }
    if (PAR) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (CONV == GT_CONV_GCN) atomicAdd(&sh_par[c0 + q], a_self[q]);
            if (EK == GT_EDGE_LINEAR) {
                atomicAdd(&sh_par[ld + c0 + q], a_b[q]);
#pragma unroll
                for (int k = 0; k < KD; ++k)
                    if (k < en.kdim) atomicAdd(&sh_par[(2 + k) * ld + c0 + q], a_w[k][q]);
            }
        }
        __syncthreads();
        for (int c = tid; c < d; c += nthr) {
            if (CONV == GT_CONV_GCN) atomicAdd(&d_self[c], sh_par[c]);
            if (EK == GT_EDGE_LINEAR) {
                atomicAdd(&d_edge_b[c], sh_par[ld + c]);
                for (int k = 0; k < en.kdim; ++k) atomicAdd(&d_edge_w[c * en.kdim + k], sh_par[(2 + k) * ld + c]);
            }
        }
    }
    if (CONV == GT_CONV_GIN) {   // d eps: one global atomic per block
        float* sh_eps = sh_par + (PAR ? (2 + KD) * ld : 0);
        deps = warp_sum(deps);
        const int nw = (nthr + 31) >> 5;
        if ((tid & 31) == 0) sh_eps[tid >> 5] = deps;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < nw; ++w) t += sh_eps[w];
            if (t != 0.f) atomicAdd(d_self, t);
        }
    }
}

// GIN adjoint, bf16, edge-table encoder, PACKED mask arithmetic.  k_agg_bwd3 is issue-bound (~125 warp instructions per
// edge at ~2.1 IPC per SM), and most of them form the ReLU mask and the masked gradient one channel at a time in fp32:
//   gm = (x_j + e_t > 0) ? g : 0.
// The mask is exact in bf16: x_j is a bf16 value, so  x_j + e_t > 0  <=>  x_j > -e_t  <=>  x_j > rd(-e_t)  with rd = round
// DOWN to bf16 (no bf16 value lies in (rd(v), v]).  `th` [ntypes + 1, ld] holds rd(-table) (gt_edge_table_thresholds, once
// per layer; row ntypes = +inf for the tail slots of a 4-slot batch), so a slot costs per channel PAIR one packed compare
// (mask 0xFFFF / 0) and one AND with the packed gradient - which IS the bf16 per-edge gradient that gets stored - plus the
// fp32 accumulation of dx.  One 16-byte table load per slot instead of two.
__global__ void __launch_bounds__(256, 3)
k_agg_bwd3p(const bf16* __restrict__ x, const bf16* __restrict__ dout, bf16* __restrict__ dx, int N, int d, int ld,
            const int32_t* __restrict__ rp_src, const int32_t* __restrict__ dst_by_src,
            const int32_t* __restrict__ eid_by_src, EdgeEnc en, const bf16* __restrict__ th,
            const float* __restrict__ self_param, float* __restrict__ d_self, bf16* __restrict__ gm_out) {
    constexpr int CONV = GT_CONV_GIN, EK = GT_EDGE_TABLE, KD = 1;
    extern __shared__ float sh_par[];    // one eps slot per warp
    const int c0 = threadIdx.x * 8;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const float eps1 = 1.f + self_param[0];
    float deps = 0.f;
    const int stride = gridDim.x * blockDim.y;
    int j = blockIdx.x * blockDim.y + threadIdx.y;
    int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    if (j < N) b0 = rp_src[j], e0 = rp_src[j + 1];
    if (j + stride < N) b1 = rp_src[j + stride], e1 = rp_src[j + stride + 1];
    SlotBatch<KD> cur;
    if (j < N) load_slots<CONV, EK, KD>(cur, en, dst_by_src, eid_by_src, rp_src, b0, e0, j, 1.f);
    auto unpack_add = [](const uint32_t (&w)[4], float (&acc)[8]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[2 * k] += __uint_as_float(w[k] << 16);
            acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
        }
    };
    for (; j < N; j += stride) {
        const uint4 xp = *reinterpret_cast<const uint4*>(x + (int64_t)j * ld + c0);
        const uint4 gjp = *reinterpret_cast<const uint4*>(dout + (int64_t)j * ld + c0);
        uint4 g[AGG_U];
#pragma unroll
        for (int u = 0; u < AGG_U; ++u) g[u] = *reinterpret_cast<const uint4*>(dout + (int64_t)cur.other[u] * ld + c0);
        SlotBatch<KD> nxt;
        const int jnext = j + stride;
        if (jnext < N) load_slots<CONV, EK, KD>(nxt, en, dst_by_src, eid_by_src, rp_src, b1, e1, jnext, 1.f);
        int b2 = 0, e2 = 0;
        if (jnext + stride < N) b2 = rp_src[jnext + stride], e2 = rp_src[jnext + stride + 1];
        const uint32_t xw[4] = {xp.x, xp.y, xp.z, xp.w};
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto consume = [&](const SlotBatch<KD>& sb, int pbase) {
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) {
                const int tyu = sb.nrm[u] != 0.f ? sb.ty[u] : en.ntypes;      // tail slot: the +inf row masks everything
                const uint4 tp = *reinterpret_cast<const uint4*>(th + (int64_t)tyu * ld + c0);
                const uint32_t tw[4] = {tp.x, tp.y, tp.z, tp.w}, gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
                uint32_t m[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    m[k] = gw[k] & __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&xw[k]), *reinterpret_cast<const __nv_bfloat162*>(&tw[k]));
                unpack_add(m, acc);
                if (gm_out && pbase + u < e0)
                    *reinterpret_cast<uint4*>(gm_out + (int64_t)(pbase + u) * ld + c0) = make_uint4(m[0], m[1], m[2], m[3]);
            }
        };
        consume(cur, b0);
        for (int p0 = b0 + AGG_U; p0 < e0; p0 += AGG_U) {         // nodes with more than AGG_U out-edges
            SlotBatch<KD> sb;
            load_slots<CONV, EK, KD>(sb, en, dst_by_src, eid_by_src, rp_src, p0, e0, j, 1.f);
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) g[u] = *reinterpret_cast<const uint4*>(dout + (int64_t)sb.other[u] * ld + c0);
            consume(sb, p0);
        }
        const uint32_t gjw[4] = {gjp.x, gjp.y, gjp.z, gjp.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float g0 = __uint_as_float(gjw[k] << 16), g1 = __uint_as_float(gjw[k] & 0xffff0000u);
            const float x0 = __uint_as_float(xw[k] << 16), x1 = __uint_as_float(xw[k] & 0xffff0000u);
            acc[2 * k] = fmaf(eps1, g0, acc[2 * k]);
            acc[2 * k + 1] = fmaf(eps1, g1, acc[2 * k + 1]);
            deps = fmaf(x0, g0, deps);
            deps = fmaf(x1, g1, deps);
        }
        st8(dx + (int64_t)j * ld + c0, acc);
        cur = nxt;
        b0 = b1, e0 = e1, b1 = b2, e1 = e2;
    }
    deps = warp_sum(deps);                  // d eps: one global atomic per block
    const int nw = (nthr + 31) >> 5;
    if ((tid & 31) == 0) sh_par[tid >> 5] = deps;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += sh_par[w];
        if (t != 0.f) atomicAdd(d_self, t);
    }
}

// th[t, c] = round-down-to-bf16(-table[t, c]) for t < ntypes, +inf for t == ntypes (see k_agg_bwd3p)
__global__ void k_edge_table_thresholds(const float* __restrict__ table, int ntypes, int ld, bf16* __restrict__ th) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (ntypes + 1) * ld) return;
    uint16_t b = 0x7F80;                    // +inf
    if (i < ntypes * ld) {
        const uint32_t v = __float_as_uint(-table[i]);
        b = (uint16_t)(v >> 16);
        if ((v & 0xFFFFu) && (v >> 31)) b += 1;      // negative and inexact: truncation rounds towards zero, i.e. up
        if ((v & 0x7FFFFFFFu) > 0x7F800000u) b = 0x7F80;   // NaN table entry: x + NaN > 0 is false
    }
    reinterpret_cast<uint16_t*>(th)[i] = b;
}

// Adjoint with the edge-TABLE gradient fused in (bf16, table edge encoder, <= 64 edge types, ld = 128 or 256).
// The table gradient is the contraction d_table[t, c] = sum over edges e of type t of gm[e, c] with gm[e, :] =
// norm * dout[dst(e), :] * 1[x[src(e), :] + table[t, :] > 0] - exactly the per-edge vector the adjoint already holds in
// registers.  k_agg_bwd3 wrote gm ([E, ld] bf16, 1.08 GB per layer at config 4) for a separate OneHot^T . gm GEMM that read
// it back; here gm never leaves the SM:
//   worker warps 0..6 : thread = (source node j, 8-channel vector) as in k_agg_bwd3, over a CONTIGUOUS node range per
//                       block, so the block's out-edges are the contiguous CSR slots [P0, P1): slot p is row (p-P0) % 32
//                       of staging tile (p-P0) / 32.  A worker stores its 8 gm values (bf16) into the tile as the
//                       MN-major A operand (k = row, m = channel; 128B swizzle), eight lanes store the row's one-hot type
//                       vector into the MN-major B operand (k = row, n = type), then the node's lane 0 arrives on the
//                       tile's mbarrier once per row (32 arrivals = 32 rows complete).
//   warp 7            : one lane issues tcgen05.mma (M = 128 channels, N = 64 types, K = 16 rows; ld / 128 channel groups
//                       x 2 k-steps per tile) accumulating D[channel, type] in TMEM over all tiles of the block;
//                       the completion of a tile's MMAs (tcgen05.commit) hands the tile (ring of three) back to the workers.
//   epilogue          : TMEM -> registers -> one fp32 global atomic per (block, type, channel).
// Rows past the block's last edge are written as zero rows / zero one-hot columns, so a tile is always complete.
constexpr int BW4_WORKERS = 7 * 32, BW4_THREADS = 8 * 32, BW4_ROWS = 32, BW4_TYPES = 64, BW4_NBUF = 3;

__device__ __forceinline__ bool bw4_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(tc::smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// mbarrier wait that turns a protocol error into a launch failure (message + trap) instead of a hung GPU
__device__ __forceinline__ void bw4_wait(uint64_t* bar, uint32_t parity, int tag, int t, int ntiles) {
    const uint32_t addr = tc::smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok) __nanosleep(64);
        if (!ok && spin > (1u << 23)) {
            printf("k_agg_bwd4: stuck wait tag=%d tile=%d/%d block=%d thread=%d parity=%u\n", tag, t, ntiles, (int)blockIdx.x,
                   (int)threadIdx.x, parity);
            __trap();
        }
    }
}

template <int CONV>
__global__ void __launch_bounds__(BW4_THREADS, 3)
k_agg_bwd4(const bf16* __restrict__ x, const bf16* __restrict__ dout, bf16* __restrict__ dx, int N, int d, int ld,
           const int32_t* __restrict__ rp_src, const int32_t* __restrict__ dst_by_src,
           const int32_t* __restrict__ eid_by_src, EdgeEnc en, const float* __restrict__ self_param,
           float* __restrict__ d_self, float* __restrict__ d_table, int nodes_per_block) {
    using namespace tc;
    constexpr int EK = GT_EDGE_TABLE, KD = 1;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[BW4_NBUF], mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int tiles_done;            // tiles whose MMAs have completed (monotonic; published by the MMA issuer)
    __shared__ float sh_eps[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = ld >> 3;                    // lanes per node: 16 or 32
    const int nb = BW4_WORKERS / L;           // nodes in flight per block iteration
    const int groups = ld >> 7;               // 128-channel accumulator groups (1 or 2)
    const uint32_t a_bytes = (uint32_t)(ld >> 6) * 4096u;           // A tile: ld/64 chunks of [32 rows x 128 B]
    const uint32_t tile_bytes = a_bytes + 4096u;                    // + B tile: [32 rows x 64 types] one-hot, MN-major
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int n_begin = min(blockIdx.x * nodes_per_block, N), n_end = min(n_begin + nodes_per_block, N);
    const int P0 = rp_src[n_begin], P1 = rp_src[n_end];
    const int ntiles = (P1 - P0 + BW4_ROWS - 1) / BW4_ROWS;
    const uint32_t tmem_cols = groups == 1 ? 64u : 128u;

    if (tid == 0) {
        for (int b = 0; b < BW4_NBUF; ++b) mbar_init(&full_bar[b], BW4_ROWS);
        mbar_init(&mma_bar, 1);
        tiles_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 7) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    float deps = 0.f;

    if (warp == 7) {
        if (lane == 0 && ntiles > 0) {   // ===== MMA issuer =====
            // D[m = channel, n = type] += A[m, k] B[n, k]: both operands MN-major (rows = k), M = 128, N = 64
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BW4_TYPES >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // Progress is published as a monotonic tile count, not as an mbarrier phase: a worker whose consecutive rows
            // lie many tiles apart (a node with hundreds of out-edges in between, e.g. the self loops of a shape-bucket
            // slack node) would alias a parity wait on a barrier that has moved on by an even number of phases.  At most
            // one commit is outstanding, so this thread itself never lags mma_bar by more than one phase.
            int pub = 0;                  // tiles published (= MMAs complete, buffer reusable)
            auto publish = [&]() {
                ++pub;
                asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(&tiles_done)), "r"(pub) : "memory");
            };
            for (int t = 0; t < ntiles; ++t) {
                const int b = t % BW4_NBUF;
                const uint32_t a_addr = base + (uint32_t)b * tile_bytes, b_addr = a_addr + a_bytes;
                // wait for tile t; the previous tile is handed back as soon as its MMAs are seen complete
                for (uint32_t spin = 0;; ++spin) {
                    if (pub < t && bw4_try(&mma_bar, (uint32_t)pub & 1u)) publish();
                    if (bw4_try(&full_bar[b], (uint32_t)(t / BW4_NBUF) & 1u)) break;
                    __nanosleep(128);     // a spinning issuer takes issue slots from the workers of three resident blocks
                    if (spin > (1u << 23)) {
                        printf("k_agg_bwd4: tile %d/%d of block %d never filled\n", t, ntiles, (int)blockIdx.x);
                        __trap();
                    }
                }
                if (pub < t) {
                    bw4_wait(&mma_bar, (uint32_t)pub & 1u, 3, t, ntiles);
                    publish();
                }
                tc_fence_after();
                for (int g = 0; g < groups; ++g) {
#pragma unroll
                    for (int ks = 0; ks < BW4_ROWS / 16; ++ks)
                        umma_f16(tmem + (uint32_t)g * BW4_TYPES, desc_mnmajor(a_addr + (uint32_t)g * 8192u + (uint32_t)ks * 2048u, 4096u),
                                 desc_mnmajor(b_addr + (uint32_t)ks * 2048u, 4096u), idesc, (t > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&mma_bar);
            }
            bw4_wait(&mma_bar, (uint32_t)pub & 1u, 3, ntiles, ntiles);
            publish();
        }
    } else {                             // ===== workers =====
        const int tx = tid % L, ty = tid / L;
        const int c0 = tx * 8;
        const uint32_t gmask = L == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
        EdgeVec<EK, KD> ev;
        ev.load(en, c0, d);
        float root[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) root[q] = (CONV == GT_CONV_GCN && c0 + q < d) ? self_param[c0 + q] : 0.f;
        float a_self[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float eps1 = (CONV == GT_CONV_GIN) ? 1.f + self_param[0] : 0.f;
        int t_ok = BW4_NBUF - 1;          // tiles <= t_ok are known to be writable (the first use of a buffer is free)
        // one staged row: 8 gm values of this thread's channels into the A tile; lanes 0..7 of the node write the row's
        // one-hot type vector (64 types = 128 B) into the B tile
        auto write_row = [&](int q, const float (&gmv)[8], int ty_e) {
            const int t = q >> 5;
            const uint32_t r = (uint32_t)q & 31u;
            if (t > t_ok) {               // MMAs of tile t - NBUF must have read the buffer
                int done = 0;
                for (uint32_t spin = 0;; ++spin) {
                    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&tiles_done)) : "memory");
                    if (done >= t - (BW4_NBUF - 1)) break;
                    __nanosleep(64);
                    if (spin > (1u << 24)) __trap();
                }
                t_ok = done + BW4_NBUF - 1;
            }
            const uint32_t a_addr = base + (uint32_t)(t % BW4_NBUF) * tile_bytes, b_addr = a_addr + a_bytes;
            const uint32_t sw = (r & 7u);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(gmv[0], gmv[1]), h1 = __floats2bfloat162_rn(gmv[2], gmv[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(gmv[4], gmv[5]), h3 = __floats2bfloat162_rn(gmv[6], gmv[7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(a_addr + (uint32_t)(tx >> 3) * 4096u + r * 128u + ((((uint32_t)tx & 7u) ^ sw) << 4)),
                           "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                           "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
            if (tx < 8) {                 // types 8 tx .. 8 tx + 7 of the row
                const int idx = ty_e - tx * 8;
                const uint32_t one = (idx & 1) ? 0x3F800000u : 0x00003F80u;
                const uint32_t w0 = (idx >> 1) == 0 ? one : 0u, w1 = (idx >> 1) == 1 ? one : 0u;
                const uint32_t w2 = (idx >> 1) == 2 ? one : 0u, w3 = (idx >> 1) == 3 ? one : 0u;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(b_addr + r * 128u + ((((uint32_t)tx) ^ sw) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            }
        };
        // rows [q0, q0 + n) of this node are in shared memory: make them visible to the tensor core, one arrival per row
        auto publish_rows = [&](int q0, int n) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp(gmask);
            if (tx == 0) {
                const int ta = q0 >> 5, tb = (q0 + n - 1) >> 5;
                const int na = ta == tb ? n : BW4_ROWS - (q0 & 31);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[ta % BW4_NBUF])), "r"(na) : "memory");
                if (tb != ta)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[tb % BW4_NBUF])), "r"(n - na) : "memory");
            }
        };
        const int stride = nb;
        int j = n_begin + ty;
        int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        const bool active = ty < nb;      // L = 16: 14 node slots, no idle threads; kept for safety
        if (active && j < n_end) b0 = rp_src[j], e0 = rp_src[j + 1];
        if (active && j + stride < n_end) b1 = rp_src[j + stride], e1 = rp_src[j + stride + 1];
        SlotBatch<KD> cur;
        if (active && j < n_end) load_slots<CONV, EK, KD>(cur, en, dst_by_src, eid_by_src, rp_src, b0, e0, j, 1.f);
        for (; active && j < n_end; j += stride) {
            float xj[8], gj[8], g[AGG_U][8];
            ld8(x + (int64_t)j * ld + c0, xj);
            ld8(dout + (int64_t)j * ld + c0, gj);
#pragma unroll
            for (int u = 0; u < AGG_U; ++u) ld8(dout + (int64_t)cur.other[u] * ld + c0, g[u]);
            SlotBatch<KD> nxt;
            const int jnext = j + stride;
            if (jnext < n_end) load_slots<CONV, EK, KD>(nxt, en, dst_by_src, eid_by_src, rp_src, b1, e1, jnext, 1.f);
            int b2 = 0, e2 = 0;
            if (jnext + stride < n_end) b2 = rp_src[jnext + stride], e2 = rp_src[jnext + stride + 1];
            const float inv_deg_j = CONV == GT_CONV_GCN ? 1.f / (float)(e0 - b0 + 1) : 1.f;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            auto consume = [&](const SlotBatch<KD>& sb, int pbase) {
#pragma unroll
                for (int u = 0; u < AGG_U; ++u) {
                    float ee[8], gmv[8];
                    ev.embed(en, sb.a[u], sb.ty[u], c0, ld, ee);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float gm = (xj[q] + ee[q] > 0.f) ? sb.nrm[u] * g[u][q] : 0.f;
                        gmv[q] = gm;
                        acc[q] += gm;
                    }
                    if (pbase + u < e0) write_row(pbase + u - P0, gmv, sb.ty[u]);     // node-uniform condition
                }
                if (pbase < e0) publish_rows(pbase - P0, min(AGG_U, e0 - pbase));
            };
            consume(cur, b0);
            for (int p0 = b0 + AGG_U; p0 < e0; p0 += AGG_U) {
                SlotBatch<KD> sb;
                load_slots<CONV, EK, KD>(sb, en, dst_by_src, eid_by_src, rp_src, p0, e0, j, 1.f);
#pragma unroll
                for (int u = 0; u < AGG_U; ++u) ld8(dout + (int64_t)sb.other[u] * ld + c0, g[u]);
                consume(sb, p0);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (CONV == GT_CONV_GCN) {
                    const float gs = (xj[q] + root[q] > 0.f) ? gj[q] * inv_deg_j : 0.f;
                    acc[q] += gs;
                    a_self[q] += gs;
                } else {
                    acc[q] = fmaf(eps1, gj[q], acc[q]);
                    deps = fmaf(xj[q], gj[q], deps);
                }
            }
            st8(dx + (int64_t)j * ld + c0, acc);
            cur = nxt;
            b0 = b1, e0 = e1, b1 = b2, e1 = e2;
        }
        // rows between the block's last edge and the end of its last tile: zero rows, zero one-hot columns
        if (active) {
            const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int q = P1 - P0 + ty; q < ntiles * BW4_ROWS; q += nb) {
                write_row(q, zero, -1);
                publish_rows(q, 1);
            }
        }
        if (CONV == GT_CONV_GCN) {          // d root: one global atomic per (thread, channel)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (c0 + q < d && a_self[q] != 0.f) atomicAdd(&d_self[c0 + q], a_self[q]);
        }
        // ---- table-gradient epilogue: warps 0..3 own TMEM lanes 32 w .. 32 w + 31 = channels of every group
        if (ntiles > 0 && warp < 4) {
            for (uint32_t spin = 0;; ++spin) {
                int done;
                asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&tiles_done)) : "memory");
                if (done >= ntiles) break;
                __nanosleep(128);
                if (spin > (1u << 24)) __trap();
            }
            tc_fence_after();
            for (int g = 0; g < groups; ++g) {
                const int ch = g * 128 + warp * 32 + lane;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t rr[32];
                    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * BW4_TYPES + hf * 32), rr);
                    if (ch < d) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int ty_e = hf * 32 + i;
                            const float v = __uint_as_float(rr[i]);
                            if (ty_e < en.ntypes && v != 0.f) atomicAdd(d_table + (int64_t)ty_e * ld + ch, v);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    if (CONV == GT_CONV_GIN) {   // d eps: one global atomic per block
        deps = warp_sum(deps);
        if (lane == 0) sh_eps[warp] = deps;
    }
    tc_fence_before();
    __syncthreads();
    if (CONV == GT_CONV_GIN && tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 7; ++w) t += sh_eps[w];
        if (t != 0.f) atomicAdd(d_self, t);
    }
    if (warp == 7) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

template <typename T, int CONV>
static int launch_fwd(int ek, const T* x, T* out, int N, int d, int ld, const int32_t* rp_dst,
                      const int32_t* src_by_dst, const int32_t* eid_by_dst, const int32_t* rp_src,
                      EdgeEnc en, const float* self_param, cudaStream_t st) {
    const int grid = blocks_for(N, AGG_WARPS, kNumSMs * 8);
    const int nch = (ld + 127) / 128;
#define L2K(EK, NCH, KD) k_agg_fwd2<T, CONV, EK, NCH, KD><<<grid2, AGG_WARPS * 32, 0, st>>>(x, out, N, d, ld, rp_dst, src_by_dst, eid_by_dst, rp_src, en, self_param, nch)
#define L2(EK, NCH) do { if (EK == GT_EDGE_LINEAR && en.kdim > 2) L2K(EK, NCH, 4); else L2K(EK, NCH, 2); } while (0)
#define L2N(EK) L2(EK, 1)
    // (node, 128-channel chunk) items: the global warp count must be a multiple of nch
    const int grid2 = (blocks_for((int64_t)N * nch, AGG_WARPS, kNumSMs * 8) + nch - 1) / nch * nch;
#define L(EK) k_agg_fwd<T, CONV, EK><<<grid, AGG_WARPS * 32, 0, st>>>(x, out, N, d, ld, rp_dst, src_by_dst, eid_by_dst, rp_src, en, self_param)
    static const int variant = getenv("GT_AGG_VARIANT") ? atoi(getenv("GT_AGG_VARIANT")) : 0;   // tuning knob: 1 = per-edge kernels, 2 = warp-per-chunk
    // measured (tools/agg_bench.py): v3 wins with table / no edge encoders (58-64 registers); with the Linear edge encoder
    // its per-thread weight registers (104-128 per thread) cost too much occupancy and the warp-per-chunk kernel stays
    const bool slots_ok = CONV != GT_CONV_GCN || en.norm_slot != nullptr;   // v3 reads the GCN norm per slot
    if (ld % 8 == 0 && ld <= 512 && slots_ok && (variant == 3 || (variant == 0 && ek != GT_EDGE_LINEAR))) {   // v3: thread per (node, 8-channel vector)
        const dim3 blk((unsigned)(ld / 8), (unsigned)max(1, 256 / (ld / 8)));
        const int grid3 = blocks_for(N, (int)blk.y, kNumSMs * 3);    // ~one resident wave (74-80 registers); threads pipeline over their nodes
#define L3K(EK, KD) k_agg_fwd3<T, CONV, EK, KD><<<grid3, blk, 0, st>>>(x, out, N, d, ld, rp_dst, src_by_dst, eid_by_dst, rp_src, en, self_param)
        if (ek == GT_EDGE_NONE) L3K(GT_EDGE_NONE, 1);
        else if (ek == GT_EDGE_LINEAR) { if (en.kdim > 2) L3K(GT_EDGE_LINEAR, 4); else L3K(GT_EDGE_LINEAR, 2); }
        else L3K(GT_EDGE_TABLE, 1);
#undef L3K
    } else if (nch <= 4 && variant != 1) {
        if (ek == GT_EDGE_NONE) L2N(GT_EDGE_NONE);
        else if (ek == GT_EDGE_LINEAR) L2N(GT_EDGE_LINEAR);
        else L2N(GT_EDGE_TABLE);
    } else if (ek == GT_EDGE_NONE) L(GT_EDGE_NONE);
    else if (ek == GT_EDGE_LINEAR) L(GT_EDGE_LINEAR);
    else L(GT_EDGE_TABLE);
#undef L
#undef L2N
#undef L2
#undef L2K
    return 0;
}

template <typename T, int CONV>
static int launch_bwd(int ek, const T* x, const T* dout, T* dx, int N, int d, int ld, const int32_t* rp_src,
                      const int32_t* dst_by_src, const int32_t* eid_by_src, EdgeEnc en,
                      const float* self_param, float* dw, float* db, float* dtab, float* dself, T* gm_out,
                      void* th_scratch, cudaStream_t st) {
    int grid = blocks_for(N, AGG_WARPS, kNumSMs * 8);
    const int nch = (ld + 127) / 128;
    const bool tab_grad = ek == GT_EDGE_TABLE && dtab != nullptr;
    const size_t smem2 = sizeof(float) * ((size_t)(2 + MAX_KDIM) * 128 * nch + (tab_grad ? (size_t)en.ntypes * ld : 0));
#define L2K(EK, NCH, KD) do { \
        static bool attr = false; \
        if (!attr) { cudaFuncSetAttribute(k_agg_bwd2<T, CONV, EK, NCH, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; } \
        k_agg_bwd2<T, CONV, EK, NCH, KD><<<grid, AGG_WARPS * 32, smem2, st>>>(x, dout, dx, N, d, ld, rp_src, dst_by_src, eid_by_src, en, self_param, dw, db, dtab, dself, nch); \
    } while (0)
#define L2(EK, NCH) do { if (EK == GT_EDGE_LINEAR && en.kdim > 2) L2K(EK, NCH, 4); else L2K(EK, NCH, 2); } while (0)
#define L2N(EK) L2(EK, 1)
#define L(EK) k_agg_bwd<T, CONV, EK><<<grid, AGG_WARPS * 32, 0, st>>>(x, dout, dx, N, d, ld, rp_src, dst_by_src, eid_by_src, en, self_param, dw, db, dtab, dself)
    static const int variant = getenv("GT_AGG_VARIANT") ? atoi(getenv("GT_AGG_VARIANT")) : 0;
    // measured (tools/agg_bench.py): with an edge-type table the block-private [ntypes][ld] shared-memory table of the
    // batched kernel costs too much occupancy (84 vs 53 us on the molpcba batch); the per-edge kernel keeps its
    // 32 KB per-chunk table.  Linear / no edge encoder: the batched kernel wins (62 vs 73 us on the code2 batch).
    // Without the table gradient (d_table == NULL: gt_aggregate_table_grad computes it) the batched kernel is used.
    const bool slots_ok = CONV != GT_CONV_GCN || en.norm_slot != nullptr;
    // opt-in (GT_AGG_TABLE_FUSED=1 / ops.TABLE_GRAD_FUSED): measured at config 4 the fused kernel takes 757 us against
    // 522 (k_agg_bwd3) + 203 (one-hot GEMM) + 71 (one-hot) = 796 us - both variants are issue bound at ~2.1 IPC per SM
    // (446 M vs 265 M warp instructions, profiles/r02_ncu_agg4_syn_v1.txt), so the 2.2 GB per layer it keeps out of HBM buy
    // 5 %; the default stays with the separate contraction.  fused_tab: -1 = not forced by the environment
    static const int fused_env = getenv("GT_AGG_TABLE_FUSED") ? atoi(getenv("GT_AGG_TABLE_FUSED")) : -1;
    const int fused_tab = fused_env != 0;   // the caller asks for it by passing d_table without gm_out
    if (std::is_same<T, bf16>::value && tab_grad && fused_tab && !gm_out && slots_ok && (ld == 128 || ld == 256) &&
        en.ntypes <= BW4_TYPES && variant == 0) {
        // adjoint + edge-table gradient in one kernel (tcgen05 contraction of the per-edge gradients with the one-hot types)
        int grid4 = blocks_for(N, BW4_WORKERS / (ld / 8), kNumSMs * 3);
        const int npb = (N + grid4 - 1) / grid4;
        grid4 = (N + npb - 1) / npb;
        const size_t smem4 = BW4_NBUF * ((size_t)(ld / 64) * 4096 + 4096) + 1024;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_agg_bwd4<CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
        k_agg_bwd4<CONV><<<grid4, BW4_THREADS, smem4, st>>>((const bf16*)x, (const bf16*)dout, (bf16*)dx, N, d, ld, rp_src, dst_by_src,
                                                            eid_by_src, en, self_param, dself, dtab, npb);
        return 0;
    }
    if (ld % 8 == 0 && ld <= 512 && slots_ok && (variant == 3 || (variant == 0 && ek != GT_EDGE_LINEAR)) && !tab_grad) {   // v3
        const dim3 blk((unsigned)(ld / 8), (unsigned)max(1, 256 / (ld / 8)));
        // ~one resident wave (threads pipeline over their nodes; every block ends with one global atomic per
        // parameter-gradient element)
        const int grid3 = blocks_for(N, (int)blk.y, kNumSMs * 3);
        const size_t smem3 = sizeof(float) * ((size_t)(2 + MAX_KDIM) * ld + 32);
        static const int packed = getenv("GT_AGG_PACKED") ? atoi(getenv("GT_AGG_PACKED")) : 1;
        if (packed && th_scratch && std::is_same<T, bf16>::value && CONV == GT_CONV_GIN && ek == GT_EDGE_TABLE &&
            ((uintptr_t)x | (uintptr_t)dout | (uintptr_t)dx | (uintptr_t)th_scratch | (uintptr_t)gm_out) % 16 == 0) {
            // packed-mask adjoint: thresholds of this layer's table first (61 x ld elements), then the walk
            const int nth = (en.ntypes + 1) * ld;
            k_edge_table_thresholds<<<(nth + 255) / 256, 256, 0, st>>>(en.table, en.ntypes, ld, (bf16*)th_scratch);
            k_agg_bwd3p<<<grid3, blk, sizeof(float) * 32, st>>>((const bf16*)x, (const bf16*)dout, (bf16*)dx, N, d, ld, rp_src, dst_by_src,
                                                              eid_by_src, en, (const bf16*)th_scratch, self_param, dself, (bf16*)gm_out);
            return 0;
        }
#define L3K(EK, KD) k_agg_bwd3<T, CONV, EK, KD><<<grid3, blk, smem3, st>>>(x, dout, dx, N, d, ld, rp_src, dst_by_src, eid_by_src, en, self_param, dw, db, dself, gm_out)
        if (ek == GT_EDGE_NONE) L3K(GT_EDGE_NONE, 1);
        else if (ek == GT_EDGE_LINEAR) { if (en.kdim > 2) L3K(GT_EDGE_LINEAR, 4); else L3K(GT_EDGE_LINEAR, 2); }
        else L3K(GT_EDGE_TABLE, 1);
#undef L3K
    } else if (gm_out) {
        set_error("gt_aggregate_bwd: gm_out needs the vector kernel (ld %% 8 == 0, ld <= 512, table / no edge encoder, d_table == NULL)");
        return -1;
    } else if (nch <= 4 && smem2 <= 100 * 1024 && variant != 1 && (!tab_grad || variant == 2)) {
        // every block flushes its private gradient tables once: keep the block count near the resident capacity
        const int cap = kNumSMs * (smem2 > 48 * 1024 ? 2 : 8);
        grid = blocks_for((int64_t)N * nch, AGG_WARPS, cap);
        grid = (grid + nch - 1) / nch * nch;     // global warp count must be a multiple of nch
        if (ek == GT_EDGE_NONE) L2N(GT_EDGE_NONE);
        else if (ek == GT_EDGE_LINEAR) L2N(GT_EDGE_LINEAR);
        else L2N(GT_EDGE_TABLE);
    } else if (ek == GT_EDGE_NONE) L(GT_EDGE_NONE);
    else if (ek == GT_EDGE_LINEAR) L(GT_EDGE_LINEAR);
    else L(GT_EDGE_TABLE);
#undef L
#undef L2N
#undef L2
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
                                const float* table, const float* self_param, const float* norm_slot,
                                const int32_t* etype_slot, const float* attr_slot, void* stream) {
    if (int r = check_common("gt_aggregate_fwd", conv, N, d, ld, edge_kind, kdim)) return r;
    EdgeEnc en{edge_attr, edge_w, edge_b, etype, table, kdim, 0, norm_slot, etype_slot, attr_slot};
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
                                float* d_edge_w, float* d_edge_b, float* d_table, float* d_self, const float* norm_slot,
                                const int32_t* etype_slot, const float* attr_slot, void* gm_out, void* th_scratch,
                                void* stream) {
    (void)rowptr_dst;
    if (int r = check_common("gt_aggregate_bwd", conv, N, d, ld, edge_kind, kdim)) return r;
    GT_CHECK_ARG(edge_kind != GT_EDGE_TABLE || ntypes > 0, "gt_aggregate_bwd: ntypes must be the row count of the edge table");
    EdgeEnc en{edge_attr, edge_w, edge_b, etype, table, kdim, ntypes, norm_slot, etype_slot, attr_slot};
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    GT_DISPATCH_DT(dt, {
        if (conv == GT_CONV_GCN)
            rc = launch_bwd<T, GT_CONV_GCN>(edge_kind, (const T*)x, (const T*)dout, (T*)dx, (int)N, d, ld, rowptr_src, dst_by_src, eid_by_src, en, self_param, d_edge_w, d_edge_b, d_table, d_self, (T*)gm_out, th_scratch, st);
        else
            rc = launch_bwd<T, GT_CONV_GIN>(edge_kind, (const T*)x, (const T*)dout, (T*)dx, (int)N, d, ld, rowptr_src, dst_by_src, eid_by_src, en, self_param, d_edge_w, d_edge_b, d_table, d_self, (T*)gm_out, th_scratch, st);
    });
    if (rc) return rc;
    GT_LAUNCH_CHECK("gt_aggregate_bwd");
    return 0;
}

extern "C" int gt_aggregate_table_grad(int dt, int conv, const void* x, const void* dout, int64_t N, int32_t d, int32_t ld,
                                       const int32_t* rowptr_src, int64_t E, const int32_t* src_t, const int32_t* dst_t,
                                       const int32_t* type_t, const float* table, int32_t ntypes, float* d_table,
                                       void* stream) {
    if (int r = check_common("gt_aggregate_table_grad", conv, N, d, ld, GT_EDGE_TABLE, 0)) return r;
    GT_CHECK_ARG(ntypes > 0 && E >= 0 && E < (1ll << 31), "gt_aggregate_table_grad: bad ntypes / E");
    if (E == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int nch = (ld + 127) / 128;
    const int64_t items = (E + 31) / 32 * nch;          // (32 sorted slots, 128-channel chunk) per warp
    const unsigned grid = (unsigned)((items + 7) / 8);
    GT_DISPATCH_DT(dt, {
        if (conv == GT_CONV_GCN)
            k_agg_table_grad<T, GT_CONV_GCN><<<grid, 256, 0, st>>>((const T*)x, (const T*)dout, d, ld, nch, rowptr_src, E, src_t, dst_t, type_t, table, d_table);
        else
            k_agg_table_grad<T, GT_CONV_GIN><<<grid, 256, 0, st>>>((const T*)x, (const T*)dout, d, ld, nch, rowptr_src, E, src_t, dst_t, type_t, table, d_table);
    });
    GT_LAUNCH_CHECK("gt_aggregate_table_grad");
    return 0;
}

// ------------------------------------------------------------------------------------------------ per-slot edge data
namespace gt {
__global__ void k_edge_slots(const int32_t* __restrict__ rp_src, const int32_t* __restrict__ nbr_a,
                             const int32_t* __restrict__ nbr_b_of_slot_rowptr, int64_t E, int64_t N,
                             const int32_t* __restrict__ eid_slot, const int32_t* __restrict__ etype,
                             const float* __restrict__ attr, int kdim, float* __restrict__ norm_slot,
                             int32_t* __restrict__ etype_slot, float* __restrict__ attr_slot,
                             const int32_t* __restrict__ rowptr_slot) {
    (void)nbr_b_of_slot_rowptr;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) {
        const int eid = eid_slot[p];
        if (etype_slot) etype_slot[p] = etype[eid];
        if (attr_slot)
            for (int k = 0; k < kdim; ++k) attr_slot[p * kdim + k] = attr[(int64_t)eid * kdim + k];
        if (norm_slot) {
            // owner row of slot p: binary search in the slot CSR row pointers; the other endpoint is nbr_a[p]
            int64_t lo = 0, hi = N;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (rowptr_slot[mid] <= p) lo = mid; else hi = mid;
            }
            const int a = (int)lo, b = nbr_a[p];
            norm_slot[p] = rsqrtf((float)(rp_src[a + 1] - rp_src[a] + 1)) * rsqrtf((float)(rp_src[b + 1] - rp_src[b] + 1));
        }
    }
}
}  // namespace gt

extern "C" int gt_edge_slots(const int32_t* rowptr_slot, const int32_t* nbr_slot, const int32_t* eid_slot,
                             const int32_t* rowptr_src, int64_t E, int64_t N, const int32_t* etype,
                             const float* edge_attr, int32_t kdim, float* norm_slot, int32_t* etype_slot,
                             float* attr_slot, void* stream) {
    GT_CHECK_ARG(E >= 0 && N > 0 && kdim >= 0 && kdim <= MAX_KDIM, "gt_edge_slots: bad arguments");
    GT_CHECK_ARG(!etype_slot || etype, "gt_edge_slots: etype_slot needs etype");
    GT_CHECK_ARG(!attr_slot || (edge_attr && kdim > 0), "gt_edge_slots: attr_slot needs edge_attr");
    if (E == 0) return 0;
    k_edge_slots<<<blocks_for(E, 256), 256, 0, (cudaStream_t)stream>>>(rowptr_src, nbr_slot, nullptr, E, N, eid_slot, etype,
                                                                     edge_attr, kdim, norm_slot, etype_slot, attr_slot,
                                                                     rowptr_slot);
    GT_LAUNCH_CHECK("gt_edge_slots");
    return 0;
}
