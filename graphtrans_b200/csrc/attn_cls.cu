// Pooled-query attention for the LAST encoder layer.  The model reads only the pooled row of the encoder output
// (reference models/gnn_transformer.py:114-115: `h_graph = transformer_out[-1]`, the <CLS> position), so in
// the last layer only ONE query per graph matters: its keys / values are still all token rows of the graph, but the
// attention, out-projection, both LayerNorms and the FFN of that layer run on B rows instead of n_rows.  Outputs and
// gradients of everything the loss depends on are unchanged (rows that nothing reads are simply never computed).
//
// One warp owns one (graph, head): lane = key for the score pass (32 keys in flight), lane = channel for the P.V /
// dQ accumulations (coalesced row reads).  fp32 math on fp32 or bf16 storage; memory-bound: K and V of every token
// are read exactly once (n_rows * 2d * s bytes).  Same dropout hash as the full kernels (row id = head * n_rows +
// query row, column = key row), so a full-layer run with the same salt drops the same probabilities.
#include "common.cuh"

namespace gt {

constexpr int CLS_WARPS = 4;
constexpr int CLS_MAXDH = 64;

template <typename T>
__device__ __forceinline__ float cls_dot(const float* __restrict__ a_sm, const T* __restrict__ row, int dh) {
    float s = 0.f;
    for (int c = 0; c < dh; c += 4) {
        float v[4];
        ld4(row + c, v);
        s = fmaf(a_sm[c], v[0], s);
        s = fmaf(a_sm[c + 1], v[1], s);
        s = fmaf(a_sm[c + 2], v[2], s);
        s = fmaf(a_sm[c + 3], v[3], s);
    }
    return s;
}

// q [B, d] (query of every graph, unscaled), kv [n_rows, 2d] (k | v), q_rows [B] = packed row of each query (dropout
// row id only), out [B, d], lse [B, nhead]
template <typename T>
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_mha_cls_fwd(const T* __restrict__ q, const T* __restrict__ kv, const int32_t* __restrict__ tok_off,
              const int32_t* __restrict__ q_rows, int64_t n_rows, int B, int nhead, int dh, float scale,
              T* __restrict__ out, float* __restrict__ lse, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    __shared__ float sq[CLS_WARPS][CLS_MAXDH];
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int item = blockIdx.x * CLS_WARPS + wid;
    if (item >= B * nhead) return;
    const int g = item / nhead, h = item - g * nhead;
    const int d = nhead * dh, ld2 = 2 * d;
    const int ks = tok_off[g], ke = tok_off[g + 1];
    for (int c = lane; c < dh; c += 32) sq[wid][c] = to_f(q[(int64_t)g * d + h * dh + c]) * scale;
    __syncwarp();
    const uint32_t rk = drop_row_key(dr, (uint64_t)h * (uint64_t)n_rows + (uint64_t)q_rows[g]);
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
    for (int kb = ks; kb < ke; kb += 32) {
        const int j = kb + lane;
        const bool valid = j < ke;
        const float s = valid ? cls_dot(sq[wid], kv + (int64_t)j * ld2 + h * dh, dh) : -INFINITY;
        const float mnew = fmaxf(m, warp_max(s));
        const float p = valid ? __expf(s - mnew) : 0.f;
        const float corr = __expf(m - mnew);
        l = l * corr + warp_sum(p);
        o0 *= corr;
        o1 *= corr;
        const float pd = (valid && dr.on) ? p * drop_elem(dr, rk, (uint32_t)j) : p;
        const int cnt = min(32, ke - kb);
        for (int jj = 0; jj < cnt; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, pd, jj);
            const T* vrow = kv + (int64_t)(kb + jj) * ld2 + d + h * dh;
            if (lane < dh) o0 = fmaf(pj, to_f(vrow[lane]), o0);
            if (lane + 32 < dh) o1 = fmaf(pj, to_f(vrow[lane + 32]), o1);
        }
        m = mnew;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;     // a graph without tokens cannot occur (every graph has its pooled row)
    T* orow = out + (int64_t)g * d + h * dh;
    if (lane < dh) orow[lane] = from_f<T>(o0 * inv);
    if (lane + 32 < dh) orow[lane + 32] = from_f<T>(o1 * inv);
    if (lane == 0) lse[(int64_t)g * nhead + h] = m + __logf(l);
}

// dq [B, d]; dkv [n_rows, 2d]: every row of a graph is written by the warp of (graph, head); rows beyond the last
// graph (unused tail of the static row bound) are zeroed by the trailing blocks
template <typename T>
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_mha_cls_bwd(const T* __restrict__ q, const T* __restrict__ kv, const T* __restrict__ out, const T* __restrict__ dout,
              const float* __restrict__ lse, const int32_t* __restrict__ tok_off, const int32_t* __restrict__ q_rows,
              int64_t n_rows, int B, int nhead, int dh, float scale, T* __restrict__ dq, T* __restrict__ dkv,
              float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    __shared__ float sq[CLS_WARPS][CLS_MAXDH];
    __shared__ float sdo[CLS_WARPS][CLS_MAXDH];
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int item = blockIdx.x * CLS_WARPS + wid;
    const int d = nhead * dh, ld2 = 2 * d;
    if (item >= B * nhead) {   // tail rows [tok_off[B], n_rows): zero gradient
        const int64_t first = tok_off[B];
        const int64_t n4 = (n_rows - first) * (ld2 / 4);
        const int64_t t0 = (int64_t)(item - B * nhead) * 32 + lane;
        const int64_t stride = ((int64_t)gridDim.x * CLS_WARPS - (int64_t)B * nhead) * 32;
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t i = t0; i < n4; i += stride) st4(dkv + first * ld2 + i * 4, z);
        return;
    }
    const int g = item / nhead, h = item - g * nhead;
    const int ks = tok_off[g], ke = tok_off[g + 1];
    float dl = 0.f;
    for (int c = lane; c < dh; c += 32) {
        sq[wid][c] = to_f(q[(int64_t)g * d + h * dh + c]);
        const float go = to_f(dout[(int64_t)g * d + h * dh + c]);
        sdo[wid][c] = go;
        dl = fmaf(go, to_f(out[(int64_t)g * d + h * dh + c]), dl);
    }
    dl = warp_sum(dl);
    __syncwarp();
    const float L = lse[(int64_t)g * nhead + h];
    const uint32_t rk = drop_row_key(dr, (uint64_t)h * (uint64_t)n_rows + (uint64_t)q_rows[g]);
    float a0 = 0.f, a1 = 0.f;     // dq (channel = lane, lane + 32)
    for (int kb = ks; kb < ke; kb += 32) {
        const int j = kb + lane;
        const bool valid = j < ke;
        float ds = 0.f, pd = 0.f;
        if (valid) {
            const T* krow = kv + (int64_t)j * ld2 + h * dh;
            const float s = cls_dot(sq[wid], krow, dh) * scale;
            const float p = __expf(s - L);
            const float keep = dr.on ? drop_elem(dr, rk, (uint32_t)j) : 1.f;
            const float dp = cls_dot(sdo[wid], krow + d, dh) * keep;     // d(loss)/d(P_j) through the dropped P
            ds = p * (dp - dl) * scale;
            pd = p * keep;
        }
        const int cnt = min(32, ke - kb);
        for (int jj = 0; jj < cnt; ++jj) {
            const float dsj = __shfl_sync(0xffffffffu, ds, jj);
            const float pj = __shfl_sync(0xffffffffu, pd, jj);
            const T* krow = kv + (int64_t)(kb + jj) * ld2 + h * dh;
            T* dk = dkv + (int64_t)(kb + jj) * ld2 + h * dh;
            if (lane < dh) {
                a0 = fmaf(dsj, to_f(krow[lane]), a0);
                dk[lane] = from_f<T>(dsj * sq[wid][lane]);
                dk[d + lane] = from_f<T>(pj * sdo[wid][lane]);
            }
            if (lane + 32 < dh) {
                a1 = fmaf(dsj, to_f(krow[lane + 32]), a1);
                dk[lane + 32] = from_f<T>(dsj * sq[wid][lane + 32]);
                dk[d + lane + 32] = from_f<T>(pj * sdo[wid][lane + 32]);
            }
        }
    }
    T* dqr = dq + (int64_t)g * d + h * dh;
    if (lane < dh) dqr[lane] = from_f<T>(a0);
    if (lane + 32 < dh) dqr[lane + 32] = from_f<T>(a1);
}

}  // namespace gt

using namespace gt;

extern "C" int gt_mha_cls_fwd(int dt, const void* q, const void* kv, const int32_t* tok_off, const int32_t* q_rows,
                              int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* out, float* lse,
                              float drop_p, const uint64_t* rng_state, uint64_t salt, void* stream) {
    GT_CHECK_ARG(B > 0 && n_rows > 0 && nhead > 0 && dh > 0 && dh <= CLS_MAXDH && dh % 4 == 0,
                 "gt_mha_cls_fwd: needs head dim %% 4 == 0 and <= %d (got %d)", CLS_MAXDH, dh);
    const int64_t items = B * nhead;
    GT_DISPATCH_DT(dt, (k_mha_cls_fwd<T><<<(unsigned)((items + CLS_WARPS - 1) / CLS_WARPS), CLS_WARPS * 32, 0, (cudaStream_t)stream>>>(
                           (const T*)q, (const T*)kv, tok_off, q_rows, n_rows, (int)B, nhead, dh, scale, (T*)out, lse, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_mha_cls_fwd");
    return 0;
}

extern "C" int gt_mha_cls_bwd(int dt, const void* q, const void* kv, const void* out, const void* dout, const float* lse,
                              const int32_t* tok_off, const int32_t* q_rows, int64_t n_rows, int64_t B, int32_t nhead,
                              int32_t dh, float scale, void* dq, void* dkv, float drop_p, const uint64_t* rng_state,
                              uint64_t salt, void* stream) {
    GT_CHECK_ARG(B > 0 && n_rows > 0 && nhead > 0 && dh > 0 && dh <= CLS_MAXDH && dh % 4 == 0,
                 "gt_mha_cls_bwd: needs head dim %% 4 == 0 and <= %d (got %d)", CLS_MAXDH, dh);
    const int64_t items = B * nhead;
    const int64_t blocks = (items + CLS_WARPS - 1) / CLS_WARPS + 64;     // + 64 blocks that clear the unused tail rows
    GT_DISPATCH_DT(dt, (k_mha_cls_bwd<T><<<(unsigned)blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>(
                           (const T*)q, (const T*)kv, (const T*)out, (const T*)dout, lse, tok_off, q_rows, n_rows, (int)B, nhead, dh,
                           scale, (T*)dq, (T*)dkv, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_mha_cls_bwd");
    return 0;
}
