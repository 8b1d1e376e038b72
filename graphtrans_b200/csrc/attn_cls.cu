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
#include <stdlib.h>

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

// ---- d = 256, bf16: one warp per GRAPH (all heads).  A token row's k (or v) part is 512 contiguous bytes = one 16-byte
// vector per lane, so every load of the pass is a fully coalesced row read, four rows of a warp in flight; the lanes of a
// head (DH / 8 of them) meet with xor shuffles for the score, then run the online softmax redundantly.  The kernels above
// read 8 bytes per lane from 32 different rows per instruction and walk V two bytes per lane (2 TB/s measured at config
// 4); same arithmetic, same dropout hash.
constexpr int CLSW_D = 256;

__device__ __forceinline__ void clsw_unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 clsw_pack(const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
    t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
    return t;
}
template <int LPH>
__device__ __forceinline__ float clsw_head_sum(float v) {      // sum over the LPH lanes of a head
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int DH>
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_mha_cls_fwd_w(const bf16* __restrict__ q, const bf16* __restrict__ kv, const int32_t* __restrict__ tok_off,
                const int32_t* __restrict__ q_rows, int64_t n_rows, int B, float scale, bf16* __restrict__ out,
                float* __restrict__ lse, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    constexpr int D = CLSW_D, LPH = DH / 8, NH = D / DH, U = 4;
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, g = blockIdx.x * CLS_WARPS + (threadIdx.x >> 5);
    if (g >= B) return;
    const int h = lane / LPH;
    const int ks = tok_off[g], ke = tok_off[g + 1];
    float qv[8];
    clsw_unpack(*reinterpret_cast<const uint4*>(q + (int64_t)g * D + lane * 8), qv);
#pragma unroll
    for (int c = 0; c < 8; ++c) qv[c] *= scale;
    const uint32_t rk = drop_row_key(dr, (uint64_t)h * (uint64_t)n_rows + (uint64_t)q_rows[g]);
    float m = -INFINITY, l = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = ks; j0 < ke; j0 += U) {
        uint4 kr[U], vr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bf16* row = kv + (int64_t)min(j0 + u, ke - 1) * (2 * D) + lane * 8;
            kr[u] = *reinterpret_cast<const uint4*>(row);
            vr[u] = *reinterpret_cast<const uint4*>(row + D);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j0 + u >= ke) break;          // warp-uniform
            float kf[8], vf[8];
            clsw_unpack(kr[u], kf);
            clsw_unpack(vr[u], vf);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) s = fmaf(qv[c], kf[c], s);
            s = clsw_head_sum<LPH>(s);
            const float mnew = fmaxf(m, s);
            const float corr = __expf(m - mnew), p = __expf(s - mnew);
            l = fmaf(l, corr, p);
            const float pd = dr.on ? p * drop_elem(dr, rk, (uint32_t)(j0 + u)) : p;
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = fmaf(o[c], corr, pd * vf[c]);
            m = mnew;
        }
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= inv;
    *reinterpret_cast<uint4*>(out + (int64_t)g * D + lane * 8) = clsw_pack(o);
    if (lane % LPH == 0) lse[(int64_t)g * NH + h] = m + __logf(l);
}

template <int DH>
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_mha_cls_bwd_w(const bf16* __restrict__ q, const bf16* __restrict__ kv, const bf16* __restrict__ out,
                const bf16* __restrict__ dout, const float* __restrict__ lse, const int32_t* __restrict__ tok_off,
                const int32_t* __restrict__ q_rows, int64_t n_rows, int B, float scale, bf16* __restrict__ dq,
                bf16* __restrict__ dkv, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    constexpr int D = CLSW_D, LPH = DH / 8, NH = D / DH, U = 4;
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, g = blockIdx.x * CLS_WARPS + (threadIdx.x >> 5);
    if (g >= B) {   // tail rows [tok_off[B], n_rows): zero gradient (the trailing blocks)
        const int64_t first = tok_off[B];
        const int64_t n16 = (n_rows - first) * (2 * D / 8);
        const int64_t t0 = (int64_t)(g - B) * 32 + lane;
        const int64_t stride = ((int64_t)gridDim.x * CLS_WARPS - (int64_t)B) * 32;
        for (int64_t i = t0; i < n16; i += stride)
            *reinterpret_cast<uint4*>(dkv + first * (2 * D) + i * 8) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int h = lane / LPH;
    const int ks = tok_off[g], ke = tok_off[g + 1];
    float qv[8], gv[8], ov[8];
    clsw_unpack(*reinterpret_cast<const uint4*>(q + (int64_t)g * D + lane * 8), qv);
    clsw_unpack(*reinterpret_cast<const uint4*>(dout + (int64_t)g * D + lane * 8), gv);
    clsw_unpack(*reinterpret_cast<const uint4*>(out + (int64_t)g * D + lane * 8), ov);
    float dl = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) dl = fmaf(gv[c], ov[c], dl);
    dl = clsw_head_sum<LPH>(dl);
    const float L = lse[(int64_t)g * NH + h];
    const uint32_t rk = drop_row_key(dr, (uint64_t)h * (uint64_t)n_rows + (uint64_t)q_rows[g]);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = ks; j0 < ke; j0 += U) {
        uint4 kr[U], vr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bf16* row = kv + (int64_t)min(j0 + u, ke - 1) * (2 * D) + lane * 8;
            kr[u] = *reinterpret_cast<const uint4*>(row);
            vr[u] = *reinterpret_cast<const uint4*>(row + D);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j0 + u >= ke) break;          // warp-uniform
            float kf[8], vf[8];
            clsw_unpack(kr[u], kf);
            clsw_unpack(vr[u], vf);
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                s = fmaf(qv[c], kf[c], s);
                dp = fmaf(gv[c], vf[c], dp);
            }
            s = clsw_head_sum<LPH>(s) * scale;
            dp = clsw_head_sum<LPH>(dp);
            const float p = __expf(s - L);
            const float keep = dr.on ? drop_elem(dr, rk, (uint32_t)(j0 + u)) : 1.f;
            const float ds = p * (dp * keep - dl) * scale, pd = p * keep;
            float dk[8], dv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                a[c] = fmaf(ds, kf[c], a[c]);
                dk[c] = ds * qv[c];
                dv[c] = pd * gv[c];
            }
            bf16* drow = dkv + (int64_t)(j0 + u) * (2 * D) + lane * 8;
            *reinterpret_cast<uint4*>(drow) = clsw_pack(dk);
            *reinterpret_cast<uint4*>(drow + D) = clsw_pack(dv);
        }
    }
    *reinterpret_cast<uint4*>(dq + (int64_t)g * D + lane * 8) = clsw_pack(a);
}

// One warp walks ALL token rows of its graph, four at a time: that needs enough graphs to fill the machine (>= 4 warps
// per SM) and short graphs (config 4: 129 rows on average).  Few long graphs (Code2: 128 graphs of up to 1001 rows) keep
// the warp-per-(graph, head) kernels with 32 keys in flight per warp (measured: 2.80 vs 2.91 ms per config-3 step).
static bool clsw_ok(int dt, int32_t nhead, int32_t dh, int64_t B, int64_t n_rows, const void* a, const void* b, const void* c,
                    const void* d4) {
    static const int on = getenv("GT_CLS_WIDE") ? atoi(getenv("GT_CLS_WIDE")) : 1;
    const bool shape = on == 2 || (B >= 4 * kNumSMs && n_rows <= 256 * B);
    return on && shape && dt == GT_BF16 && nhead * dh == CLSW_D && (dh == 32 || dh == 64) &&
           (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d4) % 16 == 0);
}

}  // namespace gt

using namespace gt;

extern "C" int gt_mha_cls_fwd(int dt, const void* q, const void* kv, const int32_t* tok_off, const int32_t* q_rows,
                              int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* out, float* lse,
                              float drop_p, const uint64_t* rng_state, uint64_t salt, void* stream) {
    GT_CHECK_ARG(B > 0 && n_rows > 0 && nhead > 0 && dh > 0 && dh <= CLS_MAXDH && dh % 4 == 0,
                 "gt_mha_cls_fwd: needs head dim %% 4 == 0 and <= %d (got %d)", CLS_MAXDH, dh);
    if (clsw_ok(dt, nhead, dh, B, n_rows, q, kv, out, nullptr)) {
        const unsigned blocks = (unsigned)((B + CLS_WARPS - 1) / CLS_WARPS);
        if (dh == 64)
            k_mha_cls_fwd_w<64><<<blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)kv, tok_off, q_rows, n_rows, (int)B, scale, (bf16*)out, lse, drop_p, rng_state, salt);
        else
            k_mha_cls_fwd_w<32><<<blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)kv, tok_off, q_rows, n_rows, (int)B, scale, (bf16*)out, lse, drop_p, rng_state, salt);
        GT_LAUNCH_CHECK("gt_mha_cls_fwd");
        return 0;
    }
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
    if (clsw_ok(dt, nhead, dh, B, n_rows, q, kv, dout, dkv) && ((uintptr_t)out | (uintptr_t)dq) % 16 == 0) {
        const unsigned blocks = (unsigned)((B + CLS_WARPS - 1) / CLS_WARPS + 64);     // + 64 blocks for the tail rows
        if (dh == 64)
            k_mha_cls_bwd_w<64><<<blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)kv, (const bf16*)out, (const bf16*)dout, lse, tok_off, q_rows, n_rows, (int)B, scale, (bf16*)dq, (bf16*)dkv, drop_p, rng_state, salt);
        else
            k_mha_cls_bwd_w<32><<<blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)kv, (const bf16*)out, (const bf16*)dout, lse, tok_off, q_rows, n_rows, (int)B, scale, (bf16*)dq, (bf16*)dkv, drop_p, rng_state, salt);
        GT_LAUNCH_CHECK("gt_mha_cls_bwd");
        return 0;
    }
    const int64_t items = B * nhead;
    const int64_t blocks = (items + CLS_WARPS - 1) / CLS_WARPS + 64;     // + 64 blocks that clear the unused tail rows
    GT_DISPATCH_DT(dt, (k_mha_cls_bwd<T><<<(unsigned)blocks, CLS_WARPS * 32, 0, (cudaStream_t)stream>>>(
                           (const T*)q, (const T*)kv, (const T*)out, (const T*)dout, lse, tok_off, q_rows, n_rows, (int)B, nhead, dh,
                           scale, (T*)dq, (T*)dkv, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_mha_cls_bwd");
    return 0;
}
