// Stage 2 reference path: masked multi-head self-attention over PACKED tokens on CUDA cores.
// Replaces F.multi_head_attention_forward's bmm/masked_fill/softmax/bmm chain (reference
// modules/transformer_encoder.py:28-32,59).  The packed layout has no
// padding rows, so the reference's -inf key-padding mask is implicit: the keys of a token are
// exactly the token rows of its graph.  Streaming (online) softmax; nothing of size T x T is
// materialised.  This file is the exact-fp32 parity path and the cross-check for the tcgen05
// kernel (attn_tc.cu); one warp owns one (token, head).
#include "common.cuh"

namespace gt {

int mha_meta_launch(const int32_t* tok_graph, const int32_t* tok_off, int64_t n_rows, int64_t B, int32_t* row_bounds,
                    int32_t* tile_bounds, cudaStream_t st);
int mha_tc_fwd_launch(int dt, const void* qkv, const int32_t* tok_graph, const int32_t* tok_off, const int32_t* key_start,
                      const int32_t* row_bounds, const int32_t* tile_bounds, int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* out, float* lse,
                      float drop_p, const uint64_t* rng, uint64_t salt, cudaStream_t st);  // attn_tc.cu; -2 = not eligible
int mha_tc_bwd_launch(int dt, const void* qkv, const void* out, const void* dout, const float* lse,
                      const int32_t* tok_graph, const int32_t* tok_off, const int32_t* key_start,
                      const int32_t* row_bounds, const int32_t* tile_bounds, int64_t n_rows,
                      int64_t B, int32_t nhead, int32_t dh, float scale, void* dqkv, float* delta, float drop_p,
                      const uint64_t* rng, uint64_t salt, cudaStream_t st);

// dropout row id of the probabilities P[h, query row, :] (shared by every MHA kernel, see common.cuh drop_elem)
__device__ __forceinline__ uint64_t att_row_id(int h, int64_t q, int64_t n_rows) {
    return (uint64_t)h * (uint64_t)n_rows + (uint64_t)q;
}

constexpr int ATT_WARPS = 4;
constexpr int ATT_MAXDH = 64;

template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ a_sm, const T* __restrict__ row, int dh) {
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

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_mha_fwd(const T* __restrict__ qkv, const int32_t* __restrict__ tok_graph, const int32_t* __restrict__ tok_off,
          const int32_t* __restrict__ key_start, int64_t n_rows, int nhead, int dh, float scale, T* __restrict__ out, float* __restrict__ lse,
          float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    __shared__ float sq[ATT_WARPS][ATT_MAXDH];
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t item = blockIdx.x * (int64_t)ATT_WARPS + wid;
    if (item >= n_rows * nhead) return;
    const int64_t t = item / nhead;
    const int h = (int)(item - t * nhead);
    const int d = nhead * dh, ld3 = 3 * d;
    const int g = tok_graph[t];
    T* orow = out + t * d + h * dh;
    if (g < 0) {
        for (int c = lane; c < dh; c += 32) orow[c] = from_f<T>(0.f);
        if (lane == 0) lse[(int64_t)h * n_rows + t] = 0.f;
        return;
    }
    const int ks = key_start ? key_start[g] : tok_off[g], ke = tok_off[g + 1];
    for (int c = lane; c < dh; c += 32) sq[wid][c] = to_f(qkv[t * ld3 + h * dh + c]) * scale;
    __syncwarp();
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
    const uint32_t rk = drop_row_key(dr, att_row_id(h, t, n_rows));
    for (int kb = ks; kb < ke; kb += 32) {
        const int j = kb + lane;
        const bool valid = j < ke;
        const float s = valid ? dot_row(sq[wid], qkv + (int64_t)j * ld3 + d + h * dh, dh) : -INFINITY;
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
            const T* vrow = qkv + (int64_t)(kb + jj) * ld3 + 2 * d + h * dh;
            if (lane < dh) o0 = fmaf(pj, to_f(vrow[lane]), o0);
            if (lane + 32 < dh) o1 = fmaf(pj, to_f(vrow[lane + 32]), o1);
        }
        m = mnew;
    }
    const float inv = 1.f / l;
    if (lane < dh) orow[lane] = from_f<T>(o0 * inv);
    if (lane + 32 < dh) orow[lane + 32] = from_f<T>(o1 * inv);
    if (lane == 0) lse[(int64_t)h * n_rows + t] = m + __logf(l);
}

// dQ (and delta = rowsum(dO * O)) : one warp per (query token, head)
template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_mha_bwd_dq(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
             const float* __restrict__ lse, const int32_t* __restrict__ tok_graph,
             const int32_t* __restrict__ tok_off, const int32_t* __restrict__ key_start, int64_t n_rows, int nhead,
             int dh, float scale, T* __restrict__ dqkv, float* __restrict__ delta, float drop_p,
             const uint64_t* __restrict__ rng, uint64_t salt) {
    __shared__ float sq[ATT_WARPS][ATT_MAXDH];
    __shared__ float sdo[ATT_WARPS][ATT_MAXDH];
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t item = blockIdx.x * (int64_t)ATT_WARPS + wid;
    if (item >= n_rows * nhead) return;
    const int64_t t = item / nhead;
    const int h = (int)(item - t * nhead);
    const int d = nhead * dh, ld3 = 3 * d;
    const int g = tok_graph[t];
    T* dq = dqkv + t * ld3 + h * dh;
    if (g < 0) {
        for (int c = lane; c < dh; c += 32) {
            dq[c] = from_f<T>(0.f);
            dq[d + c] = from_f<T>(0.f);
            dq[2 * d + c] = from_f<T>(0.f);
        }
        if (lane == 0) delta[(int64_t)h * n_rows + t] = 0.f;
        return;
    }
    const int ks = key_start ? key_start[g] : tok_off[g], ke = tok_off[g + 1];
    float dl = 0.f;
    for (int c = lane; c < dh; c += 32) {
        sq[wid][c] = to_f(qkv[t * ld3 + h * dh + c]) * scale;
        const float go = to_f(dout[t * d + h * dh + c]);
        sdo[wid][c] = go;
        dl = fmaf(go, to_f(out[t * d + h * dh + c]), dl);
    }
    dl = warp_sum(dl);
    __syncwarp();
    if (lane == 0) delta[(int64_t)h * n_rows + t] = dl;
    const float L = lse[(int64_t)h * n_rows + t];
    const uint32_t rk = drop_row_key(dr, att_row_id(h, t, n_rows));
    float a0 = 0.f, a1 = 0.f;
    for (int kb = ks; kb < ke; kb += 32) {
        const int j = kb + lane;
        float ds = 0.f;
        if (j < ke) {
            const float s = dot_row(sq[wid], qkv + (int64_t)j * ld3 + d + h * dh, dh);
            const float p = __expf(s - L);
            float dp = dot_row(sdo[wid], qkv + (int64_t)j * ld3 + 2 * d + h * dh, dh);
            if (dr.on) dp *= drop_elem(dr, rk, (uint32_t)j);
            ds = p * (dp - dl);
        }
        const int cnt = min(32, ke - kb);
        for (int jj = 0; jj < cnt; ++jj) {
            const float dj = __shfl_sync(0xffffffffu, ds, jj);
            const T* krow = qkv + (int64_t)(kb + jj) * ld3 + d + h * dh;
            if (lane < dh) a0 = fmaf(dj, to_f(krow[lane]), a0);
            if (lane + 32 < dh) a1 = fmaf(dj, to_f(krow[lane + 32]), a1);
        }
    }
    if (lane < dh) dq[lane] = from_f<T>(a0 * scale);
    if (lane + 32 < dh) dq[lane + 32] = from_f<T>(a1 * scale);
}

// dK, dV : one warp per (key token, head); lanes sweep the queries of the same graph
template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_mha_bwd_dkv(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ lse,
              const float* __restrict__ delta, const int32_t* __restrict__ tok_graph,
              const int32_t* __restrict__ tok_off, const int32_t* __restrict__ key_start, int64_t n_rows, int nhead,
              int dh, float scale, T* __restrict__ dqkv, float drop_p, const uint64_t* __restrict__ rng,
              uint64_t salt) {
    const Drop dr = make_drop(rng, salt, drop_p);
    __shared__ float sk[ATT_WARPS][ATT_MAXDH];
    __shared__ float sv[ATT_WARPS][ATT_MAXDH];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t item = blockIdx.x * (int64_t)ATT_WARPS + wid;
    if (item >= n_rows * nhead) return;
    const int64_t j = item / nhead;
    const int h = (int)(item - j * nhead);
    const int d = nhead * dh, ld3 = 3 * d;
    const int g = tok_graph[j];
    if (g < 0) return;  // dq kernel already zeroed the whole row
    const int qs = tok_off[g], qe = tok_off[g + 1];
    if (key_start && j < key_start[g]) {  // a padded row is never used as a key: zero gradient
        for (int c = lane; c < dh; c += 32) {
            dqkv[j * ld3 + d + h * dh + c] = from_f<T>(0.f);
            dqkv[j * ld3 + 2 * d + h * dh + c] = from_f<T>(0.f);
        }
        return;
    }
    for (int c = lane; c < dh; c += 32) {
        sk[wid][c] = to_f(qkv[j * ld3 + d + h * dh + c]) * scale;
        sv[wid][c] = to_f(qkv[j * ld3 + 2 * d + h * dh + c]);
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int qb = qs; qb < qe; qb += 32) {
        const int i = qb + lane;
        float p = 0.f, ds = 0.f;
        if (i < qe) {
            const float s = dot_row(sk[wid], qkv + (int64_t)i * ld3 + h * dh, dh);
            p = __expf(s - lse[(int64_t)h * n_rows + i]);
            const float m = dr.on ? drop_elem(dr, drop_row_key(dr, att_row_id(h, i, n_rows)), (uint32_t)j) : 1.f;
            const float dp = dot_row(sv[wid], dout + (int64_t)i * d + h * dh, dh) * m;
            ds = p * (dp - delta[(int64_t)h * n_rows + i]);
            p *= m;  // dV uses the dropped probabilities
        }
        const int cnt = min(32, qe - qb);
        for (int ii = 0; ii < cnt; ++ii) {
            const float pi = __shfl_sync(0xffffffffu, p, ii);
            const float di = __shfl_sync(0xffffffffu, ds, ii);
            const T* qrow = qkv + (int64_t)(qb + ii) * ld3 + h * dh;
            const T* grow = dout + (int64_t)(qb + ii) * d + h * dh;
            if (lane < dh) {
                v0 = fmaf(pi, to_f(grow[lane]), v0);
                k0 = fmaf(di, to_f(qrow[lane]), k0);
            }
            if (lane + 32 < dh) {
                v1 = fmaf(pi, to_f(grow[lane + 32]), v1);
                k1 = fmaf(di, to_f(qrow[lane + 32]), k1);
            }
        }
    }
    T* dk = dqkv + j * ld3 + d + h * dh;
    T* dv = dqkv + j * ld3 + 2 * d + h * dh;
    if (lane < dh) {
        dk[lane] = from_f<T>(k0 * scale);
        dv[lane] = from_f<T>(v0);
    }
    if (lane + 32 < dh) {
        dk[lane + 32] = from_f<T>(k1 * scale);
        dv[lane + 32] = from_f<T>(v1);
    }
}

static int check_mha(const char* fn, int64_t n_rows, int64_t B, int nhead, int dh) {
    GT_CHECK_ARG(n_rows > 0 && B > 0 && nhead > 0, "%s: bad shape", fn);
    GT_CHECK_ARG(dh >= 4 && dh <= ATT_MAXDH && dh % 4 == 0, "%s: head dim %d must be a multiple of 4 in 4..%d", fn, dh, ATT_MAXDH);
    return 0;
}

}  // namespace gt

using namespace gt;

extern "C" int gt_mha_meta(const int32_t* tok_graph, const int32_t* tok_off, int64_t n_rows, int64_t B,
                           int32_t* row_bounds, int32_t* tile_bounds, void* stream) {
    GT_CHECK_ARG(n_rows > 0 && B > 0 && row_bounds && tile_bounds, "gt_mha_meta: bad arguments");
    return mha_meta_launch(tok_graph, tok_off, n_rows, B, row_bounds, tile_bounds, (cudaStream_t)stream);
}

extern "C" int gt_mha_fwd(int dt, const void* qkv, const int32_t* tok_graph, const int32_t* tok_off,
                          const int32_t* key_start, const int32_t* row_bounds, const int32_t* tile_bounds, int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale,
                          void* out, float* lse, float drop_p, const uint64_t* rng_state, uint64_t salt, int impl,
                          void* stream) {
    if (int r = check_mha("gt_mha_fwd", n_rows, B, nhead, dh)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl != 1) {
        const int r = mha_tc_fwd_launch(dt, qkv, tok_graph, tok_off, key_start, row_bounds, tile_bounds, n_rows, B, nhead, dh, scale, out, lse, drop_p, rng_state, salt, st);
        if (r != -2) return r;
        GT_CHECK_ARG(impl != 2, "gt_mha_fwd: not eligible for the tcgen05 kernel (%s)", gt_last_error());
    }
    const int grid = (int)((n_rows * nhead + ATT_WARPS - 1) / ATT_WARPS);
    GT_DISPATCH_DT(dt, (k_mha_fwd<T><<<grid, ATT_WARPS * 32, 0, st>>>((const T*)qkv, tok_graph, tok_off, key_start, n_rows, nhead, dh, scale, (T*)out, lse, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_mha_fwd(simt)");
    return 0;
}

extern "C" int gt_mha_bwd(int dt, const void* qkv, const void* out, const void* dout, const float* lse,
                          const int32_t* tok_graph, const int32_t* tok_off, const int32_t* key_start,
                          const int32_t* row_bounds, const int32_t* tile_bounds, int64_t n_rows,
                          int64_t B, int32_t nhead, int32_t dh, float scale, void* dqkv, float* delta, float drop_p,
                          const uint64_t* rng_state, uint64_t salt, int impl, void* stream) {
    if (int r = check_mha("gt_mha_bwd", n_rows, B, nhead, dh)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl != 1) {
        const int r = mha_tc_bwd_launch(dt, qkv, out, dout, lse, tok_graph, tok_off, key_start, row_bounds, tile_bounds, n_rows, B, nhead, dh, scale, dqkv, delta, drop_p, rng_state, salt, st);
        if (r != -2) return r;
        GT_CHECK_ARG(impl != 2, "gt_mha_bwd: not eligible for the tcgen05 kernel (%s)", gt_last_error());
    }
    const int grid = (int)((n_rows * nhead + ATT_WARPS - 1) / ATT_WARPS);
    GT_DISPATCH_DT(dt, {
        k_mha_bwd_dq<T><<<grid, ATT_WARPS * 32, 0, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, tok_graph, tok_off, key_start, n_rows, nhead, dh, scale, (T*)dqkv, delta, drop_p, rng_state, salt);
        k_mha_bwd_dkv<T><<<grid, ATT_WARPS * 32, 0, st>>>((const T*)qkv, (const T*)dout, lse, delta, tok_graph, tok_off, key_start, n_rows, nhead, dh, scale, (T*)dqkv, drop_p, rng_state, salt);
    });
    GT_LAUNCH_CHECK("gt_mha_bwd(simt)");
    return 0;
}
