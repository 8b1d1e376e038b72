// Row-wise / column-wise helper kernels around the two hot stages: per-graph segment sums
// (PyG global_add_pool, reference modules/gnn_module.py:219), virtual-node broadcast (:199),
// BatchNorm1d train/eval forward + backward (gnn_module.py:58,84,164,167; conv.py:19),
// LayerNorm (+ residual, + token gather) forward/backward (modules/transformer_encoder.py:56-57
// and the norms inside nn.TransformerEncoderLayer), embedding-sum node encoders
// (dataset/utils.py:28-30; ogb AtomEncoder), pad_batch gather/scatter (modules/utils.py:5-29).
// All are HBM-bound elementwise/reduction passes with 8/16-byte vector accesses.
#include <stdlib.h>

#include <initializer_list>

#include "common.cuh"

namespace gt {

// ------------------------------------------------------------------ segment sum / broadcast
// block = 128 threads = 4 warps; the block owns 32 consecutive rows; thread t owns the channel
// vectors t, t+128, ...; consecutive rows of one graph are summed in registers and flushed
// with one atomicAdd per (graph run, channel).
template <typename T>
__global__ void k_segment_sum(const T* __restrict__ x, const int32_t* __restrict__ node_graph, int N, int ld,
                              float* __restrict__ out) {
    const int r0 = blockIdx.x * 32;
    const int r1 = min(r0 + 32, N);
    for (int c0 = threadIdx.x * 4; c0 < ld; c0 += blockDim.x * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int g = node_graph[r0];
        for (int r = r0; r < r1; ++r) {
            const int gr = node_graph[r];
            if (gr != g) {
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicAdd(out + (int64_t)g * ld + c0 + q, acc[q]), acc[q] = 0.f;
                g = gr;
            }
            float v[4];
            ld4(x + (int64_t)r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += v[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(out + (int64_t)g * ld + c0 + q, acc[q]);
    }
}

// sorted variant (graphs own contiguous row ranges [node_off[g], node_off[g+1])): one BLOCK per (graph, 128-channel
// chunk); its 8 warps take interleaved rows (a Code2 graph has up to 2000 rows: one warp walking them alone was the
// critical path of the virtual-node branch), partial sums meet in shared memory and warp 0 is the single writer of
// out[g, chunk] = init[g, chunk] + sum - no atomics, deterministic, `out` needs no zero fill
template <typename T>
__global__ void __launch_bounds__(256)
k_segment_sum_sorted(const T* __restrict__ x, const int32_t* __restrict__ node_off, int B, int ld, int nch,
                     const float* __restrict__ init, float* __restrict__ out) {
    __shared__ float4 part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x;
    const int g = item / nch, c0 = (item - g * nch) * 128 + lane * 4;
    const bool col_ok = c0 < ld;
    const int r0 = node_off[g], r1 = node_off[g + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (col_ok) {
#pragma unroll 4
        for (int r = r0 + warp; r < r1; r += 8) {
            float v[4];
            ld4(x + (int64_t)r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += v[q];
        }
    }
    part[warp][lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (warp == 0 && col_ok) {
        float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
        if (init) cur = *reinterpret_cast<const float4*>(init + (int64_t)g * ld + c0);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float4 p = part[w][lane];
            cur.x += p.x; cur.y += p.y; cur.z += p.z; cur.w += p.w;
        }
        *reinterpret_cast<float4*>(out + (int64_t)g * ld + c0) = cur;
    }
}

template <typename T>
__global__ void k_add_graph_vec(const T* __restrict__ x, const float* __restrict__ v,
                                const int32_t* __restrict__ node_graph, int64_t N, int ld, T* __restrict__ y) {
    const int vpr = ld / 4;
    const int64_t total = N * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4];
        if (x) ld4(x + r * ld + c0, a);
        const int g = node_graph[r];
        if (g >= 0) {      // g < 0: shape-bucket slack node (belongs to no graph)
            ld4(v + (int64_t)g * ld + c0, b);
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] += b[q];
        }
        st4(y + r * ld + c0, a);
    }
}

// ------------------------------------------------------------------ column statistics (BN)
// grid-stride over row blocks; thread owns one 4-channel vector column group; fp32 partials per
// thread over <= ROWS_PER_BLOCK rows, then one fp64 atomic per (block, channel).
// block = 256 threads = 8 row lanes x 32 column lanes; a column lane owns one 4-channel vector, so a warp reads
// 256/512 contiguous bytes of a row and a block covers 128 channels of a slab of rows.  fp32 partials per thread,
// shared-memory tree over the 8 row lanes, then ONE fp64 atomic per (block, channel) - deterministic enough for
// statistics and ~300 blocks keep every SM busy.
constexpr int STAT_TY = 8;
// rows [m_valid[0], M) of a matrix may be shape-bucket slack (graphtrans_b200.graphed: batches padded up to a bucket so
// that CUDA-graph signatures repeat): they take no part in the batch statistics and receive a zero gradient
__device__ __forceinline__ int64_t valid_rows(const int32_t* __restrict__ m_valid, int64_t M) {
    if (!m_valid) return M;
    const int64_t v = (int64_t)m_valid[0];
    return v < M ? (v > 0 ? v : 1) : M;
}
__device__ __forceinline__ void stat_flush(float (&s)[4], float (&s2)[4], int tx, int ty, int c0, int ld,
                                           double* __restrict__ out) {
    __shared__ float sh[2][STAT_TY][128 + 4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sh[0][ty][tx * 4 + q] = s[q];
        sh[1][ty][tx * 4 + q] = s2[q];
    }
    __syncthreads();
    const int t = ty * 32 + tx;   // 256 threads -> 2 x 128 (which, channel)
    const int which = t >> 7, ch = t & 127;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < STAT_TY; ++k) a += sh[which][k][ch];
    const int c = (c0 - tx * 4) + ch;
    if (c < ld) atomicAdd(out + which * ld + c, (double)a);
}

template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
k_colstats(const T* __restrict__ x, int64_t M, int ld, int64_t rows_per_block, double* __restrict__ stats,
           const int32_t* __restrict__ m_valid) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + tx * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(r0 + rows_per_block, valid_rows(m_valid, M));
    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < ld) {
        for (int64_t r = r0 + ty; r < r1; r += STAT_TY) {
            float v[4];
            ld4(x + r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s[q] += v[q];
                if (SQ) s2[q] = fmaf(v[q], v[q], s2[q]);
            }
        }
    }
    stat_flush(s, s2, tx, ty, c0, ld, stats);
}

__global__ void k_bn_finalize(const double* __restrict__ stats, int64_t M, int d, int ld,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps,
                              int training, float* __restrict__ ssmr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && nbt) *nbt += 1;
    if (c >= ld) return;
    float scale = 0.f, shift = 0.f, mean = 0.f, rstd = 0.f;
    if (c < d) {
        if (training) {
            const double mu = stats[c] / (double)M;
            double var = stats[ld + c] / (double)M - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            rstd = (float)(1.0 / sqrt(var + (double)eps));
            if (running_mean) {
                const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
            }
        } else {
            mean = running_mean[c];
            rstd = rsqrtf(running_var[c] + eps);
        }
        scale = gamma[c] * rstd;
        shift = beta[c] - mean * scale;
    }
    ssmr[c] = scale;
    ssmr[ld + c] = shift;
    ssmr[2 * ld + c] = mean;
    ssmr[3 * ld + c] = rstd;
}

template <typename T>
__global__ void k_bn_apply(const T* __restrict__ x, int64_t M, int ld, const float* __restrict__ ssmr, int relu,
                           const T* __restrict__ resid, const float* __restrict__ gvec,
                           const int32_t* __restrict__ node_graph, T* __restrict__ y, float drop_p,
                           const uint64_t* __restrict__ rng, uint64_t salt) {
    const int vpr = ld / 4;
    const int64_t total = M * vpr;
    const Drop dr = make_drop(rng, salt, drop_p);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        float v[4], sc[4], sh[4];
        ld4(x + r * ld + c0, v);
        ld4(ssmr + c0, sc);
        ld4(ssmr + ld + c0, sh);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = fmaf(v[q], sc[q], sh[q]);
            if (relu) v[q] = fmaxf(v[q], 0.f);
        }
        if (dr.on) {
            float ds[4];
            drop4(dr, (uint64_t)i, ds);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] *= ds[q];
        }
        if (resid) {
            float t[4];
            ld4(resid + r * ld + c0, t);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += t[q];
        }
        if (gvec && node_graph[r] >= 0) {
            float t[4];
            ld4(gvec + (int64_t)node_graph[r] * ld + c0, t);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += t[q];
        }
        st4(y + r * ld + c0, v);
    }
}

// fused finalize + apply: every thread owns one 4-channel vector of a 128-channel group, derives scale / shift of its
// channels from the fp64 batch statistics (train) or the running statistics (eval) once, then streams its row slab.
// Block (x, 0) also writes scale|shift|mean|rstd for the backward and updates the running statistics in place.
template <typename T>
__global__ void __launch_bounds__(256)
k_bn_norm_fwd(const T* __restrict__ x, int64_t M, int d, int ld, int64_t rows_per_block, const double* __restrict__ stats,
              const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
              int64_t* nbt, float momentum, float eps, int training, int relu, const T* __restrict__ resid,
              const float* __restrict__ gvec, const int32_t* __restrict__ node_graph, T* __restrict__ y,
              float* __restrict__ ssmr, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt,
              const int32_t* __restrict__ m_valid) {
    const Drop dr = make_drop(rng, salt, drop_p);
    const int vpr = ld / 4;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + tx * 4;
    const int64_t Mv = valid_rows(m_valid, M);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && training && nbt) *nbt += 1;
    if (c0 >= ld) return;
    float sc[4], sh[4];
    const bool writer = blockIdx.y == 0 && ty == 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = c0 + q;
        float scale = 0.f, shift = 0.f, mean = 0.f, rstd = 0.f;
        if (c < d) {
            if (training) {
                const double mu = stats[c] / (double)Mv;
                double var = stats[ld + c] / (double)Mv - mu * mu;
                if (var < 0) var = 0;
                mean = (float)mu;
                rstd = (float)(1.0 / sqrt(var + (double)eps));
                if (writer && running_mean) {
                    const double unb = Mv > 1 ? var * ((double)Mv / (double)(Mv - 1)) : var;
                    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
                }
            } else {
                mean = running_mean[c];
                rstd = rsqrtf(running_var[c] + eps);
            }
            scale = gamma[c] * rstd;
            shift = beta[c] - mean * scale;
        }
        sc[q] = scale, sh[q] = shift;
        if (writer) {
            ssmr[c] = scale;
            ssmr[ld + c] = shift;
            ssmr[2 * ld + c] = mean;
            ssmr[3 * ld + c] = rstd;
        }
    }
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    for (int64_t r = r0 + ty; r < r1; r += STAT_TY) {
        float v[4];
        ld4(x + r * ld + c0, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = fmaf(v[q], sc[q], sh[q]);
            if (relu) v[q] = fmaxf(v[q], 0.f);
        }
        if (dr.on) {
            float ds[4];
            drop4(dr, (uint64_t)(r * vpr + c0 / 4), ds);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] *= ds[q];
        }
        if (resid) {
            float t[4];
            ld4(resid + r * ld + c0, t);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += t[q];
        }
        if (gvec && node_graph[r] >= 0) {
            float t[4];
            ld4(gvec + (int64_t)node_graph[r] * ld + c0, t);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += t[q];
        }
        st4(y + r * ld + c0, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int ld, int64_t rows_per_block,
                const float* __restrict__ ssmr, int relu, double* __restrict__ red, float drop_p,
                const uint64_t* __restrict__ rng, uint64_t salt, const int32_t* __restrict__ m_valid) {
    const Drop dr = make_drop(rng, salt, drop_p);
    const int vpr = ld / 4;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + tx * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(r0 + rows_per_block, valid_rows(m_valid, M));
    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < ld) {
        float sc[4], sh[4], mu[4], rs[4];
        ld4(ssmr + c0, sc);
        ld4(ssmr + ld + c0, sh);
        ld4(ssmr + 2 * ld + c0, mu);
        ld4(ssmr + 3 * ld + c0, rs);
        auto acc_row = [&](int64_t r, float (&v)[4], float (&g)[4]) {
            float ds[4];
            drop4(dr, (uint64_t)(r * vpr + c0 / 4), ds);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                g[q] *= ds[q];
                if (relu && fmaf(v[q], sc[q], sh[q]) <= 0.f) g[q] = 0.f;
                s[q] += g[q];
                s2[q] = fmaf(g[q], (v[q] - mu[q]) * rs[q], s2[q]);
            }
        };
        int64_t r = r0 + ty;
        for (; r + 3 * STAT_TY < r1; r += 4 * STAT_TY) {       // four rows of a thread in flight (8-byte loads: the bytes
            float v[4][4], g[4][4];                             // in flight per SM, not the arithmetic, bound this pass)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ld4(x + (r + u * STAT_TY) * ld + c0, v[u]);
                ld4(dy + (r + u * STAT_TY) * ld + c0, g[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc_row(r + u * STAT_TY, v[u], g[u]);
        }
        for (; r < r1; r += STAT_TY) {
            float v[4], g[4];
            ld4(x + r * ld + c0, v);
            ld4(dy + r * ld + c0, g);
            acc_row(r, v, g);
        }
    }
    stat_flush(s, s2, tx, ty, c0, ld, red);
}

// thread = one 4-channel vector of a 128-channel group (per-channel constants stay in registers), rows strided by 8
template <typename T>
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int d, int ld, int64_t rows_per_block,
               const float* __restrict__ ssmr, const float* __restrict__ gamma, int relu, int training,
               const double* __restrict__ red, T* __restrict__ dx, float* __restrict__ dgamma,
               float* __restrict__ dbeta, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt,
               const int32_t* __restrict__ m_valid) {
    const Drop dr = make_drop(rng, salt, drop_p);
    const int vpr = ld / 4;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + tx * 4;
    if (c0 >= ld) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    const int64_t Mv = valid_rows(m_valid, M);
    const float invM = 1.f / (float)Mv;
    float sc[4], sh[4], mu[4], rs[4], m0[4], m1[4];
    ld4(ssmr + c0, sc);
    ld4(ssmr + ld + c0, sh);
    ld4(ssmr + 2 * ld + c0, mu);
    ld4(ssmr + 3 * ld + c0, rs);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        m0[q] = (float)red[c0 + q] * invM;
        m1[q] = (float)red[ld + c0 + q] * invM;
    }
    const int64_t r1v = min(r1, Mv);
    for (int64_t r = max(r0, Mv) + ty; r < r1; r += STAT_TY) {      // bucket slack rows: no gradient
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        st4(dx + r * ld + c0, z);
    }
    for (int64_t r = r0 + ty; r < r1v; r += STAT_TY) {
        float v[4], g[4], o[4], ds[4];
        ld4(x + r * ld + c0, v);
        ld4(dy + r * ld + c0, g);
        drop4(dr, (uint64_t)(r * vpr + c0 / 4), ds);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            g[q] *= ds[q];
            if (relu && fmaf(v[q], sc[q], sh[q]) <= 0.f) g[q] = 0.f;
            if (training) {
                const float xh = (v[q] - mu[q]) * rs[q];
                o[q] = sc[q] * (g[q] - m0[q] - xh * m1[q]);  // sc = gamma * rstd
            } else {
                o[q] = sc[q] * g[q];
            }
        }
        st4(dx + r * ld + c0, o);
    }
    if (blockIdx.y == 0 && ty == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c0 + q < d) {
                dgamma[c0 + q] += (float)red[ld + c0 + q];   // accumulate semantics (single writer per channel)
                dbeta[c0 + q] += (float)red[c0 + q];
            }
    }
    (void)gamma;
}

// ------------------------------------------------------------------ BatchNorm apply, row-slab kernel (fast path)
// A block owns a slab of consecutive rows and ALL columns: thread (rr, cv) = (t / tpr, t % tpr) owns the 16-byte
// column vector cv (8 bf16 / 4 fp32 channels) of the rows r0 + rr, r0 + rr + rpi, ...  (tpr = ld / V vectors per
// row, rpi = 256 / tpr rows in flight per block).  A warp reads 512 contiguous bytes, the per-channel constants stay
// in registers and ~88 % of the threads are active for any ld.  Measured in the training step (bench.py, config 2):
// the forward normalise+activate pass gains (2.83 -> 2.80 ms per step); the same mapping for the three REDUCING
// BatchNorm kernels loses (2.94 ms: every block ends with 2*ld fp64 atomics on the same 75 cache lines, ~5x more
// than the 128-channel x tall-slab kernels above), so those keep the column-group mapping.
constexpr int SLAB_THREADS = 256;
template <typename T> struct VecW { static constexpr int V = 16 / (int)sizeof(T); };

template <typename T, int V = VecW<T>::V>
__device__ __forceinline__ void ldw(const T* p, float (&v)[V]) {
    if constexpr (V == 4) {
        ld4(p, v);
    } else {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            v[2 * k] = f.x, v[2 * k + 1] = f.y;
        }
    }
}
template <typename T, int V = VecW<T>::V>
__device__ __forceinline__ void stw(T* p, const float (&v)[V]) {
    if constexpr (V == 4) {
        st4(p, v);
    } else {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// dropout scales of the V elements starting at element (r * ld + c0): same 4-element vector numbering as drop4 users
template <int V>
__device__ __forceinline__ void dropw(const Drop& d, uint64_t vec4_idx, float (&s)[V]) {
#pragma unroll
    for (int h = 0; h < V / 4; ++h) {
        float t[4];
        drop4(d, vec4_idx + h, t);
#pragma unroll
        for (int q = 0; q < 4; ++q) s[4 * h + q] = t[q];
    }
}

// mean / rstd / scale / shift of channel c from the fp64 batch statistics (train) or the running statistics (eval)
__device__ __forceinline__ void bn_channel(int c, int d, int ld, double invM, double unbias, const double* __restrict__ stats,
                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float* running_mean, float* running_var, float momentum, float eps,
                                           int training, bool writer, float* __restrict__ ssmr, float& scale, float& shift) {
    float mean = 0.f, rstd = 0.f;
    scale = 0.f, shift = 0.f;
    if (c < d) {
        if (training) {
            const double mu = stats[c] * invM;
            double var = stats[ld + c] * invM - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            const float ve = (float)(var + (double)eps);
            float r = rsqrtf(ve);
            r = r * (1.5f - 0.5f * ve * r * r);          // one Newton step: full fp32 accuracy without fp64 sqrt / div
            rstd = r;
            if (writer && running_mean) {
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
            }
        } else {
            mean = running_mean[c];
            rstd = rsqrtf(running_var[c] + eps);
        }
        scale = gamma[c] * rstd;
        shift = beta[c] - mean * scale;
    }
    if (writer) {
        ssmr[c] = scale;
        ssmr[ld + c] = shift;
        ssmr[2 * ld + c] = mean;
        ssmr[3 * ld + c] = rstd;
    }
}

template <typename T>
__global__ void __launch_bounds__(SLAB_THREADS)
k_bn_norm_fwd_slab(const T* __restrict__ x, int64_t M, int d, int ld, int tpr, int rpi, int rows_per_block,
                   double invM, double unbias, const double* __restrict__ stats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* running_mean, float* running_var, int64_t* nbt,
                   float momentum, float eps, int training, int relu, const T* __restrict__ resid,
                   const float* __restrict__ gvec, const int32_t* __restrict__ node_graph, T* __restrict__ y,
                   float* __restrict__ ssmr, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt,
                   const int32_t* __restrict__ m_valid) {
    constexpr int V = VecW<T>::V;
    const Drop dr = make_drop(rng, salt, drop_p);
    if (m_valid) {
        const int64_t Mv = valid_rows(m_valid, M);
        invM = 1.0 / (double)Mv;
        unbias = Mv > 1 ? (double)Mv / (double)(Mv - 1) : 1.0;
    }
    const int rr = threadIdx.x / tpr, c0 = (threadIdx.x - rr * tpr) * V;
    if (blockIdx.x == 0 && threadIdx.x == 0 && training && nbt) *nbt += 1;
    if (rr >= rpi) return;
    const bool writer = blockIdx.x == 0 && rr == 0;
    float sc[V], sh[V];
#pragma unroll
    for (int q = 0; q < V; ++q)
        bn_channel(c0 + q, d, ld, invM, unbias, stats, gamma, beta, running_mean, running_var, momentum, eps, training,
                   writer, ssmr, sc[q], sh[q]);
    const int vpr4 = ld / 4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, M);
#pragma unroll 4
    for (int64_t r = r0 + rr; r < r1; r += rpi) {
        float v[V];
        ldw(x + r * ld + c0, v);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            v[q] = fmaf(v[q], sc[q], sh[q]);
            if (relu) v[q] = fmaxf(v[q], 0.f);
        }
        if (dr.on) {
            float ds[V];
            dropw<V>(dr, (uint64_t)(r * vpr4 + c0 / 4), ds);
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] *= ds[q];
        }
        if (resid) {
            float t[V];
            ldw(resid + r * ld + c0, t);
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] += t[q];
        }
        if (gvec && node_graph[r] >= 0) {
            const float* gp = gvec + (int64_t)node_graph[r] * ld + c0;
#pragma unroll
            for (int h = 0; h < V / 4; ++h) {
                float t[4];
                ld4(gp + 4 * h, t);
#pragma unroll
                for (int q = 0; q < 4; ++q) v[4 * h + q] += t[q];
            }
        }
        stw(y + r * ld + c0, v);
    }
}

// backward apply on the row-slab mapping (no reduction in this pass: per-channel constants in registers, 16-byte
// accesses, four rows of the thread in flight): dx = scale * (g - mean(g) - xhat * mean(g * xhat)) (train) or scale * g
template <typename T>
__global__ void __launch_bounds__(SLAB_THREADS)
k_bn_bwd_apply_slab(const T* __restrict__ x, const T* __restrict__ dy, int64_t M, int d, int ld, int tpr, int rpi,
                    int rows_per_block, const float* __restrict__ ssmr, int relu, int training,
                    const double* __restrict__ red, T* __restrict__ dx, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt,
                    const int32_t* __restrict__ m_valid) {
    constexpr int V = VecW<T>::V;
    const Drop dr = make_drop(rng, salt, drop_p);
    const int rr = threadIdx.x / tpr, c0 = (threadIdx.x - rr * tpr) * V;
    if (rr >= rpi) return;
    const int64_t Mv = valid_rows(m_valid, M);
    const float invM = 1.f / (float)Mv;
    float sc[V], sh[V], mu[V], rs[V], m0[V], m1[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        sc[q] = ssmr[c0 + q];
        sh[q] = ssmr[ld + c0 + q];
        mu[q] = ssmr[2 * ld + c0 + q];
        rs[q] = ssmr[3 * ld + c0 + q];
        m0[q] = (float)red[c0 + q] * invM;
        m1[q] = (float)red[ld + c0 + q] * invM;
    }
    if (blockIdx.x == 0 && rr == 0) {
#pragma unroll
        for (int q = 0; q < V; ++q)
            if (c0 + q < d) {
                dgamma[c0 + q] += (float)red[ld + c0 + q];   // accumulate semantics (single writer per channel)
                dbeta[c0 + q] += (float)red[c0 + q];
            }
    }
    const int vpr4 = ld / 4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + (int64_t)rows_per_block, M);
    const int64_t r1v = min(r1, Mv);
    for (int64_t r = max(r0, Mv) + rr; r < r1; r += rpi) {       // bucket slack rows: no gradient
        float z[V];
#pragma unroll
        for (int q = 0; q < V; ++q) z[q] = 0.f;
        stw(dx + r * ld + c0, z);
    }
    // the block's rows are one contiguous byte range of x and dy: thread 0 keeps a window of them moving towards L2
    // (bulk prefetch of the rows 8 iterations ahead), the demand loads then mostly hit L2
    const bool pf = threadIdx.x == 0 && ((size_t)ld * sizeof(T)) % 16 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0);
#pragma unroll 4
    for (int64_t r = r0 + rr; r < r1v; r += rpi) {
        if (pf) {
            const int64_t rp = r + 8 * (int64_t)rpi;
            if (rp < r1v) {
                const uint32_t bytes = (uint32_t)(min((int64_t)rpi, r1v - rp) * ld * sizeof(T));
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x + rp * ld), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(dy + rp * ld), "r"(bytes) : "memory");
            }
        }
        float v[V], g[V], o[V];
        ldw(x + r * ld + c0, v);
        ldw(dy + r * ld + c0, g);
        if (dr.on) {
            float ds[V];
            dropw<V>(dr, (uint64_t)(r * vpr4 + c0 / 4), ds);
#pragma unroll
            for (int q = 0; q < V; ++q) g[q] *= ds[q];
        }
#pragma unroll
        for (int q = 0; q < V; ++q) {
            if (relu && fmaf(v[q], sc[q], sh[q]) <= 0.f) g[q] = 0.f;
            if (training) {
                const float xh = (v[q] - mu[q]) * rs[q];
                o[q] = sc[q] * (g[q] - m0[q] - xh * m1[q]);      // sc = gamma * rstd
            } else {
                o[q] = sc[q] * g[q];
            }
        }
        stw(dx + r * ld + c0, o);
    }
}

// ------------------------------------------------------------------ LayerNorm (+resid, +gather)
// one warp per row, d <= 1024; lane owns the VW-element vectors lane, lane+32, ...  VW = 8 (16-byte accesses) for bf16
// rows of d % 8 == 0 wider than one 4-element pass, VW = 4 otherwise.
constexpr int LN_MAXV = 8;  // d <= 1024 with VW = 4
template <int VW, typename T> __device__ __forceinline__ void ldx(const T* p, float (&v)[VW]) {
    if constexpr (VW == 8) ldw<T, 8>(p, v); else ld4(p, v);
}
template <int VW, typename T> __device__ __forceinline__ void stx(T* p, const float (&v)[VW]) {
    if constexpr (VW == 8) stw<T, 8>(p, v); else st4(p, v);
}
template <int VW> __device__ __forceinline__ void ldxf(const float* p, float (&v)[VW]) {
#pragma unroll
    for (int h = 0; h < VW / 4; ++h) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * h);
        v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
    }
}

template <typename T, int MAXV, int VW>
__global__ void k_layernorm_fwd(const T* __restrict__ x, const T* __restrict__ resid,
                                const int32_t* __restrict__ in_rows, const float* __restrict__ cls, int64_t M,
                                int d, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                T* __restrict__ y, T* __restrict__ presum, float* __restrict__ mean_rstd, float drop_p,
                                const uint64_t* __restrict__ rng, uint64_t salt) {
    const Drop dr = make_drop(rng, salt, drop_p);
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = d / VW, nv4 = d / 4;
    int64_t src = row;
    if (in_rows) src = in_rows[row];
    float v[MAXV][VW];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nv) {
            if (src >= 0) ldx<VW>(x + src * d + vi * VW, v[k]);
            else if (src == -1 && cls) ldxf<VW>(cls + vi * VW, v[k]);
            else {
#pragma unroll
                for (int q = 0; q < VW; ++q) v[k][q] = 0.f;
            }
            if (dr.on) {   // LN(drop(x) + resid): dropout on the sub-layer output before the residual add
                float ds[VW];
                dropw<VW>(dr, (uint64_t)(row * nv4 + vi * (VW / 4)), ds);
#pragma unroll
                for (int q = 0; q < VW; ++q) v[k][q] *= ds[q];
            }
            if (resid) {
                float t[VW];
                ldx<VW>(resid + row * d + vi * VW, t);
#pragma unroll
                for (int q = 0; q < VW; ++q) v[k][q] += t[q];
            }
            if (presum) stx<VW>(presum + row * d + vi * VW, v[k]);
#pragma unroll
            for (int q = 0; q < VW; ++q) s += v[k][q];
        }
    }
    const float mean = warp_sum(s) / (float)d;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
        if (lane + k * 32 < nv)
#pragma unroll
            for (int q = 0; q < VW; ++q) {
                const float t = v[k][q] - mean;
                s2 = fmaf(t, t, s2);
            }
    const float rstd = rsqrtf(warp_sum(s2) / (float)d + eps);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nv) {
            float g[VW], b[VW], o[VW];
            ldxf<VW>(gamma + vi * VW, g);
            ldxf<VW>(beta + vi * VW, b);
#pragma unroll
            for (int q = 0; q < VW; ++q) o[q] = fmaf((v[k][q] - mean) * rstd, g[q], b[q]);
            stx<VW>(y + row * d + vi * VW, o);
        }
    }
    if (lane == 0 && mean_rstd) {
        mean_rstd[2 * row] = mean;
        mean_rstd[2 * row + 1] = rstd;
    }
}

// block = 8 warps; each warp loops over rows (grid-stride), R rows per iteration with all their loads in flight before
// the first shuffle reduction (one row at a time is a load -> reduce -> store latency chain; rows per warp, not bytes,
// set the time); per-lane dgamma/dbeta partials kept in registers and flushed once per block through shared memory +
// atomics.
template <typename T, int MAXV, int VW, int R>
__global__ void __launch_bounds__(256, (MAXV * VW <= 16) ? 2 : 1)
k_layernorm_bwd(const T* __restrict__ dy, const T* __restrict__ presum,
                                const float* __restrict__ mean_rstd, const int32_t* __restrict__ out_rows,
                                int64_t M, int d, const float* __restrict__ gamma, T* __restrict__ dx,
                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcls,
                                T* __restrict__ dx_drop, float drop_p, const uint64_t* __restrict__ rng, uint64_t salt) {
    const Drop dr = make_drop(rng, salt, drop_p);
    extern __shared__ float sh[];  // [2*d]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nv = d / VW, nv4 = d / 4;
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    float ag[MAXV][VW], ab[MAXV][VW];
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
#pragma unroll
        for (int q = 0; q < VW; ++q) ag[k][q] = ab[k][q] = 0.f;
    float gmm[MAXV][VW];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nv) ldxf<VW>(gamma + vi * VW, gmm[k]);
    }
    const bool pf_ok = ((size_t)d * sizeof(T)) % 16 == 0 && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)presum % 16 == 0);
    for (int64_t row0 = (blockIdx.x * (int64_t)nw + wid) * R; row0 < M; row0 += (int64_t)gridDim.x * nw * R) {
        // the rows of this warp's NEXT iteration start moving towards L2 now (one bulk prefetch per operand): with 16
        // warps per SM and R rows each, the bytes in flight of the demand loads alone sit at the edge of what hides the HBM
        // latency, and the load phase is only part of an iteration
        if (pf_ok && lane == 0) {
            const int64_t nrow = row0 + (int64_t)gridDim.x * nw * R;
            if (nrow < M) {
                const uint32_t bytes = (uint32_t)(min((int64_t)R, M - nrow) * d * sizeof(T));
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(dy + nrow * d), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(presum + nrow * d), "r"(bytes) : "memory");
            }
        }
        float xh[R][MAXV][VW], gy[R][MAXV][VW], mean[R], rstd[R], s1[R], s2[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int64_t row = row0 + rr;
            const bool ok = row < M;
            mean[rr] = ok ? mean_rstd[2 * row] : 0.f;
            rstd[rr] = ok ? mean_rstd[2 * row + 1] : 0.f;
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int vi = lane + k * 32;
#pragma unroll
                for (int q = 0; q < VW; ++q) xh[rr][k][q] = gy[rr][k][q] = 0.f;
                if (ok && vi < nv) {
                    ldx<VW>(dy + row * d + vi * VW, gy[rr][k]);
                    ldx<VW>(presum + row * d + vi * VW, xh[rr][k]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            s1[rr] = s2[rr] = 0.f;
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int vi = lane + k * 32;
                if (vi < nv && row0 + rr < M) {
#pragma unroll
                    for (int q = 0; q < VW; ++q) {
                        const float t = gy[rr][k][q];
                        xh[rr][k][q] = (xh[rr][k][q] - mean[rr]) * rstd[rr];
                        ab[k][q] += t;
                        ag[k][q] = fmaf(t, xh[rr][k][q], ag[k][q]);
                        gy[rr][k][q] = t * gmm[k][q];
                        s1[rr] += gy[rr][k][q];
                        s2[rr] = fmaf(gy[rr][k][q], xh[rr][k][q], s2[rr]);
                    }
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            s1[rr] = warp_sum(s1[rr]) / (float)d;
            s2[rr] = warp_sum(s2[rr]) / (float)d;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int64_t row = row0 + rr;
            if (row >= M) break;
            int64_t dst = row;
            if (out_rows) dst = out_rows[row];
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int vi = lane + k * 32;
                if (vi < nv) {
                    float o[VW];
#pragma unroll
                    for (int q = 0; q < VW; ++q) o[q] = rstd[rr] * (gy[rr][k][q] - s1[rr] - xh[rr][k][q] * s2[rr]);
                    if (dst >= 0) {
                        stx<VW>(dx + dst * d + vi * VW, o);
                        if (dx_drop) {   // gradient of the dropped operand: same mask as the forward
                            float ds[VW];
                            dropw<VW>(dr, (uint64_t)(row * nv4 + vi * (VW / 4)), ds);
#pragma unroll
                            for (int q = 0; q < VW; ++q) o[q] *= ds[q];
                            stx<VW>(dx_drop + dst * d + vi * VW, o);
                        }
                    } else if (dst == -1 && dcls) {
#pragma unroll
                        for (int q = 0; q < VW; ++q) atomicAdd(dcls + vi * VW + q, o[q]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nv)
#pragma unroll
            for (int q = 0; q < VW; ++q) {
                atomicAdd(&sh[vi * VW + q], ag[k][q]);
                atomicAdd(&sh[d + vi * VW + q], ab[k][q]);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        atomicAdd(dgamma + i, sh[i]);
        atomicAdd(dbeta + i, sh[d + i]);
    }
}

// ------------------------------------------------------------------ row gather / scatter
template <typename T>
__global__ void k_gather_rows(const T* __restrict__ src, const int32_t* __restrict__ rows,
                              const float* __restrict__ cls, int64_t M, int ld, T* __restrict__ dst) {
    const int vpr = ld / 4;
    const int64_t total = M * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        const int64_t s = rows[r];
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (s >= 0) ld4(src + s * ld + c0, v);
        else if (s == -1 && cls) ld4(cls + c0, v);
        st4(dst + r * ld + c0, v);
    }
}

// ddst[rows[r]] = dsrc[r] for rows >= 0 (a permutation: no conflicts); rows == -1 accumulate into dcls
template <typename T>
__global__ void k_scatter_rows(const T* __restrict__ dsrc, const int32_t* __restrict__ rows, int64_t M, int ld,
                               T* __restrict__ ddst, float* __restrict__ dcls) {
    const int vpr = ld / 4;
    const int64_t total = M * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        const int64_t s = rows[r];
        if (s == -2) continue;
        float v[4];
        ld4(dsrc + r * ld + c0, v);
        if (s >= 0) st4(ddst + s * ld + c0, v);
        else if (dcls) {
#pragma unroll
            for (int q = 0; q < 4; ++q) atomicAdd(dcls + c0 + q, v[q]);
        }
    }
}

// public pad_batch layout (reference modules/utils.py:5-29): padded[p, g, :] = h[off_g + n_g - S + p]
// for p >= S - min(n_g, S), else 0; mask[g, p] = 1 where padded.
template <typename T>
__global__ void k_pad_fwd(const T* __restrict__ h, const int32_t* __restrict__ node_off, int64_t B, int64_t S,
                          int ld, T* __restrict__ padded, uint8_t* __restrict__ mask) {
    const int vpr = ld / 4;
    const int64_t total = S * B * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pg = i / vpr;
        const int c0 = (int)(i - pg * vpr) * 4;
        const int64_t p = pg / B, g = pg - p * B;
        const int32_t off = node_off[g], n = node_off[g + 1] - off;
        const int64_t k = n < S ? n : S;
        const bool valid = p >= S - k;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) ld4(h + ((int64_t)off + n - S + p) * ld + c0, v);
        st4(padded + pg * ld + c0, v);
        if (c0 == 0 && mask) mask[g * S + p] = valid ? 0 : 1;
    }
}

template <typename T>
__global__ void k_pad_bwd(const T* __restrict__ dpadded, const int32_t* __restrict__ node_off,
                          const int32_t* __restrict__ node_graph, int64_t B, int64_t S, int64_t N, int ld,
                          T* __restrict__ dh) {
    const int vpr = ld / 4;
    const int64_t total = N * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        const int g = node_graph[r];
        const int32_t off = node_off[g], n = node_off[g + 1] - off;
        const int64_t p = r - off - n + S;  // padded position of node r
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (p >= 0) ld4(dpadded + (p * B + g) * ld + c0, v);
        st4(dh + r * ld + c0, v);
    }
}

// ------------------------------------------------------------------ embedding-sum node encoders
constexpr int EMB_MAXCOL = 12;
struct EmbCols {
    const int64_t* idx[EMB_MAXCOL];
    int64_t stride[EMB_MAXCOL];
    int64_t clamp[EMB_MAXCOL];
    const float* table[EMB_MAXCOL];
    float* dtable[EMB_MAXCOL];
    int ncol;
};

template <typename T>
__global__ void k_embed_fwd(EmbCols cols, int64_t N, int d, int ld, T* __restrict__ out) {
    const int vpr = ld / 4;
    const int64_t total = N * vpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (c0 < d) {  // d % 4 == 0 is required for table rows
            for (int c = 0; c < cols.ncol; ++c) {
                int64_t id = cols.idx[c][r * cols.stride[c]];
                if (id > cols.clamp[c]) id = cols.clamp[c];
                float v[4];
                ld4(cols.table[c] + id * d + c0, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += v[q];
            }
        }
        st4(out + r * ld + c0, acc);
    }
}

// Gradient of the embedding sums.  Tables with few rows (atom / bond / node-type / depth vocabularies) would
// serialise global atomics on a handful of addresses, so a block owns (EMB_NB nodes x 32 channels), accumulates
// those tables in shared memory (a warp handles one node: 32 lanes = 32 distinct channels, no intra-warp
// conflicts) and flushes once; large tables (Code2 attribute vocabulary, 10030 rows) take global atomics directly.
constexpr int EMB_NB = 512;
constexpr int EMB_SMALL_ROWS = 128;      // per-table threshold
constexpr int EMB_SMEM_ROWS = 512;       // total rows staged per block (64 KB)
template <typename T>
__global__ void __launch_bounds__(256)
k_embed_bwd(EmbCols cols, int64_t N, int d, int ld, const T* __restrict__ dout) {
    extern __shared__ float sh_tab[];    // [rows_small][32] then the block's index tile int32 [ncol][EMB_NB]
    __shared__ int row_base[EMB_MAXCOL]; // first smem row of column c, -1 = global atomics
    __shared__ int rows_small_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < cols.ncol; ++c) {
            const int rows = (int)cols.clamp[c] + 1;
            if (rows <= EMB_SMALL_ROWS && acc + rows <= EMB_SMEM_ROWS) row_base[c] = acc, acc += rows;
            else row_base[c] = -1;
        }
        rows_small_s = acc;
    }
    __syncthreads();
    const int rows_small = rows_small_s;
    for (int i = threadIdx.x; i < rows_small * 32; i += blockDim.x) sh_tab[i] = 0.f;
    const int64_t n0 = (int64_t)blockIdx.x * EMB_NB, n1 = min(n0 + EMB_NB, N);
    // stage the (clamped) indices of this node block once: every thread issues independent loads, instead of each
    // warp walking node by node through dependent 8-byte index loads
    int32_t* sh_idx = reinterpret_cast<int32_t*>(sh_tab + EMB_SMEM_ROWS * 32);
    for (int i = threadIdx.x; i < cols.ncol * EMB_NB; i += blockDim.x) {
        const int c = i / EMB_NB;
        const int64_t r = n0 + (i - c * EMB_NB);
        int64_t id = 0;
        if (r < n1) {
            id = cols.idx[c][r * cols.stride[c]];
            if (id > cols.clamp[c]) id = cols.clamp[c];
        }
        sh_idx[i] = (int32_t)id;
    }
    __syncthreads();
    const int ch = blockIdx.y * 32 + lane;
    if (ch < d) {
        for (int64_t r = n0 + warp; r < n1; r += 8) {
            const float g = to_f(dout[r * ld + ch]);
            for (int c = 0; c < cols.ncol; ++c) {
                const int id = sh_idx[c * EMB_NB + (int)(r - n0)];
                if (row_base[c] >= 0) atomicAdd(&sh_tab[(row_base[c] + id) * 32 + lane], g);
                else atomicAdd(cols.dtable[c] + (int64_t)id * d + ch, g);
            }
        }
    }
    __syncthreads();
    if (ch < d) {
        for (int c = 0; c < cols.ncol; ++c) {
            if (row_base[c] < 0) continue;
            const int rows = (int)cols.clamp[c] + 1;
            for (int rr = warp; rr < rows; rr += 8) {
                const float v = sh_tab[(row_base[c] + rr) * 32 + lane];
                if (v != 0.f) atomicAdd(cols.dtable[c] + (int64_t)rr * d + ch, v);
            }
        }
    }
}

// dz = dy where y > 0 else 0 (backward of a ReLU fused into a producer's epilogue)
template <typename T>
__global__ void k_relu_bwd(const T* __restrict__ dy, const T* __restrict__ y, int64_t n4, T* __restrict__ dz, float scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float g[4], v[4];
        ld4(dy + i * 4, g);
        ld4(y + i * 4, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = v[q] > 0.f ? g[q] * scale : 0.f;
        st4(dz + i * 4, g);
    }
}

// y = x * keep_mask / (1 - p): standalone dropout (its own adjoint with the same salt)
template <typename T>
__global__ void k_dropout(const T* __restrict__ x, int64_t n4, T* __restrict__ y, float p,
                          const uint64_t* __restrict__ rng, uint64_t salt) {
    const Drop dr = make_drop(rng, salt, p);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float v[4], ds[4];
        ld4(x + i * 4, v);
        drop4(dr, (uint64_t)i, ds);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] *= ds[q];
        st4(y + i * 4, v);
    }
}

__global__ void k_rng_advance(uint64_t* rng) { rng[1] += 1; }

// one launch casting MANY fp32 parameter matrices to zero-padded bf16 operand copies.  desc (device, int64[8] per
// tensor): src, dst, rows, cols, ld_dst, first block, ld_src (0 = cols: contiguous source), width (0 = ld_dst: columns
// written per destination row; width < ld_dst writes a sub-block of a larger matrix, e.g. one diagonal block of the
// PNA tower operands); a block converts 2048 consecutive (row, column < width) elements.
constexpr int CASTM_PER_BLOCK = 2048;
constexpr int CASTM_FIELDS = 8;
__global__ void __launch_bounds__(256)
k_cast_multi(const int64_t* __restrict__ desc, int n) {
    int lo = 0, hi = n;   // largest i with desc[i].first_block <= blockIdx.x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (desc[mid * CASTM_FIELDS + 5] <= (int64_t)blockIdx.x) lo = mid; else hi = mid;
    }
    const int64_t* dsc = desc + lo * CASTM_FIELDS;
    const float* src = reinterpret_cast<const float*>(dsc[0]);
    bf16* dst = reinterpret_cast<bf16*>(dsc[1]);
    const bool keep_f32 = dsc[4] < 0;                       // ld < 0: fp32 destination (stacked bias vectors)
    const int64_t rows = dsc[2], cols = dsc[3], ld = keep_f32 ? -dsc[4] : dsc[4];
    const int64_t ld_src = dsc[6] ? dsc[6] : cols, width = dsc[7] ? dsc[7] : ld;
    const int64_t base = ((int64_t)blockIdx.x - dsc[5]) * CASTM_PER_BLOCK;
    const int64_t total = rows * width;
#pragma unroll
    for (int k = 0; k < CASTM_PER_BLOCK / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < total) {
            const int64_t r = i / width, c = i - r * width;
            const float v = c < cols ? src[r * ld_src + c] : 0.f;
            if (keep_f32) reinterpret_cast<float*>(dst)[r * ld + c] = v;
            else dst[r * ld + c] = __float2bfloat16_rn(v);
        }
    }
}

// dst_b[r, c] += src[r0_b + r, c0_b + c] for up to 16 rectangular blocks of one fp32 matrix (gradients of the diagonal
// blocks of a block-diagonal operand, added into the per-tower parameter gradients)
struct AddBlocks {
    float* dst[16];
    int ld_dst[16], r0[16], c0[16], rows[16], cols[16];
    int n;
};
__global__ void k_add_blocks(AddBlocks ab, const float* __restrict__ src, int ld_src) {
    const int b = blockIdx.y;
    const int total = ab.rows[b] * ab.cols[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / ab.cols[b], c = i - r * ab.cols[b];
        ab.dst[b][(int64_t)r * ab.ld_dst[b] + c] += src[(int64_t)(ab.r0[b] + r) * ld_src + ab.c0[b] + c];
    }
}

template <typename TI, typename TO>
__global__ void k_cast_pad(const TI* __restrict__ src, int64_t rows_in, int64_t cols_in, int64_t ld_in,
                           TO* __restrict__ dst, int64_t rows_out, int64_t cols_out, int64_t ld_out) {
    const int64_t total = rows_out * cols_out;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols_out, c = i - r * cols_out;
        float v = 0.f;
        if (r < rows_in && c < cols_in) v = to_f(src[r * ld_in + c]);
        dst[r * ld_out + c] = from_f<TO>(v);
    }
}

// out[n] = sum_m X[m, n]; block (32 x 8): 32 columns, 8 row lanes, grid.y row slabs + atomics
template <typename T>
__global__ void k_colsum(const T* __restrict__ X, int64_t M, int64_t N, int64_t ld, float* __restrict__ out) {
    __shared__ float sh[8][33];
    const int64_t c = blockIdx.x * 32 + threadIdx.x;
    const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, M);
    float s = 0.f;
    if (c < N)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += to_f(X[r * ld + c]);
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// vectorised variant (rows 8/16-byte aligned, ld % 4 == 0): 8 row lanes x 32 column lanes x 4 channels like k_colstats
template <typename T>
__global__ void __launch_bounds__(256)
k_colsum_v(const T* __restrict__ X, int64_t M, int N, int64_t ld, int64_t rows_per_block, float* __restrict__ out) {
    __shared__ float sh[STAT_TY][128 + 4];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + tx * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < N) {   // N % 4 may be non-zero: the row tail up to ld is readable (ld % 4 == 0) and masked at the flush
        for (int64_t r = r0 + ty; r < r1; r += STAT_TY) {
            float v[4];
            ld4(X + r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += v[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sh[ty][tx * 4 + q] = s[q];
    __syncthreads();
    if (threadIdx.x < 128) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < STAT_TY; ++k) a += sh[k][threadIdx.x];
        const int c = blockIdx.x * 128 + threadIdx.x;
        if (c < N) atomicAdd(out + c, a);
    }
}

// wide variant: 16-byte accesses (8 bf16 / 4 fp32 channels per lane), 16 row lanes x 16 column lanes, four rows of a
// thread in flight; one fp32 atomic per (block, channel)
template <typename T>
__global__ void __launch_bounds__(256)
k_colsum_w(const T* __restrict__ X, int64_t M, int N, int64_t ld, int64_t rows_per_block, float* __restrict__ out) {
    constexpr int V = VecW<T>::V, TX = 16, TY = 16;
    __shared__ float sh[TY][TX * V + 1];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = (blockIdx.x * TX + tx) * V;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s[V];
#pragma unroll
    for (int q = 0; q < V; ++q) s[q] = 0.f;
    if (c0 < N) {
        int64_t r = r0 + ty;
        for (; r + 3 * TY < r1; r += 4 * TY) {
            float v[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) ldw(X + (r + u * TY) * ld + c0, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < V; ++q) s[q] += v[u][q];
        }
        for (; r < r1; r += TY) {
            float v[V];
            ldw(X + r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < V; ++q) s[q] += v[q];
        }
    }
#pragma unroll
    for (int q = 0; q < V; ++q) sh[ty][tx * V + q] = s[q];
    __syncthreads();
    if (threadIdx.x < TX * V) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < TY; ++k) a += sh[k][threadIdx.x];
        const int c = blockIdx.x * TX * V + threadIdx.x;
        if (c < N) atomicAdd(out + c, a);
    }
}

// dz = dy * (y > 0) * scale and out[c] += sum over the rows of dz[:, c] in the same pass (the bias gradient of the Linear
// whose epilogue applied the ReLU): k_colsum_w's mapping (16 row lanes x 16 column lanes of 16 bytes, four rows of a
// thread in flight), one fp32 atomic per (block, channel)
template <typename T>
__global__ void __launch_bounds__(256)
k_relu_bwd_colsum(const T* __restrict__ dy, const T* __restrict__ y, int64_t M, int N, int64_t ld, int64_t rows_per_block,
                  T* __restrict__ dz, float scale, float* __restrict__ out) {
    constexpr int V = VecW<T>::V, TX = 16, TY = 16;
    __shared__ float sh[TY][TX * V + 1];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = (blockIdx.x * TX + tx) * V;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s[V];
#pragma unroll
    for (int q = 0; q < V; ++q) s[q] = 0.f;
    if (c0 < ld) {
        int64_t r = r0 + ty;
        for (; r + 3 * TY < r1; r += 4 * TY) {
            float g[4][V], v[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ldw(dy + (r + u * TY) * ld + c0, g[u]);
                ldw(y + (r + u * TY) * ld + c0, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    g[u][q] = v[u][q] > 0.f ? g[u][q] * scale : 0.f;
                    s[q] += g[u][q];
                }
                stw(dz + (r + u * TY) * ld + c0, g[u]);
            }
        }
        for (; r < r1; r += TY) {
            float g[V], v[V];
            ldw(dy + r * ld + c0, g);
            ldw(y + r * ld + c0, v);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                g[q] = v[q] > 0.f ? g[q] * scale : 0.f;
                s[q] += g[q];
            }
            stw(dz + r * ld + c0, g);
        }
    }
#pragma unroll
    for (int q = 0; q < V; ++q) sh[ty][tx * V + q] = s[q];
    __syncthreads();
    if (threadIdx.x < TX * V) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < TY; ++k) a += sh[k][threadIdx.x];
        const int c = blockIdx.x * TX * V + threadIdx.x;
        if (c < N) atomicAdd(out + c, a);
    }
}

}  // namespace gt

using namespace gt;

#define ST ((cudaStream_t)stream)

// column-reduction grid: x = 128-channel groups, y = row slabs sized for ~2 blocks per SM
static dim3 stat_grid(int64_t M, int ld, int64_t* rows_per_block) {
    const int gx = (ld + 127) / 128;
    int64_t gy = (2 * kNumSMs + gx - 1) / gx;
    const int64_t max_gy = (M + 4 * STAT_TY - 1) / (4 * STAT_TY);   // at least 32 rows per block
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    *rows_per_block = (M + gy - 1) / gy;
    gy = (M + *rows_per_block - 1) / *rows_per_block;
    return dim3((unsigned)gx, (unsigned)gy);
}

// row-slab mapping of the fast BatchNorm kernels; false -> use the generic kernels (odd ld / unaligned rows)
struct Slab { int tpr, rpi, rows_per_block, grid; };
static int env_int(const char* name, int dflt) {   // tuning knobs (tools/bn_bench.py); read per call, negligible
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
static bool slab_cfg(int dt, int64_t M, int ld, std::initializer_list<const void*> ptrs, Slab* s) {
    if (!env_int("GT_BN_SLAB", 1)) return false;   // tuning knob (tools/bn_bench.py)
    const int V = dt == GT_BF16 ? 8 : 4;
    if (ld % V) return false;
    for (const void* p : ptrs)
        if ((uintptr_t)p % 16) return false;
    const int nvec = ld / V;
    if (nvec > SLAB_THREADS || M >= (1ll << 31)) return false;
    s->tpr = nvec;
    s->rpi = SLAB_THREADS / nvec;
    // ONE wave of blocks (>= 2 blocks of 256 threads per SM are resident): a second, partial wave would cost a
    // full block latency
    const int bps = env_int("GT_BN_BPS", 2), minrows = env_int("GT_BN_MINROWS", 2);
    int64_t rpb = (M + bps * kNumSMs - 1) / (bps * kNumSMs);
    if (rpb < (int64_t)minrows * s->rpi) rpb = (int64_t)minrows * s->rpi;
    s->rows_per_block = (int)rpb;
    s->grid = (int)((M + rpb - 1) / rpb);
    return true;
}

extern "C" int gt_segment_sum(int dt, const void* x, const int32_t* node_graph, int64_t N, int32_t ld, float* out,
                              void* stream) {
    GT_CHECK_ARG(N > 0 && ld > 0 && ld % 4 == 0, "gt_segment_sum: bad shape");
    GT_DISPATCH_DT(dt, (k_segment_sum<T><<<(int)((N + 31) / 32), 128, 0, ST>>>((const T*)x, node_graph, (int)N, ld, out)));
    GT_LAUNCH_CHECK("gt_segment_sum");
    return 0;
}

extern "C" int gt_segment_sum_sorted(int dt, const void* x, const int32_t* node_off, int64_t B, int32_t ld,
                                     const float* init, float* out, void* stream) {
    GT_CHECK_ARG(B > 0 && ld > 0 && ld % 4 == 0, "gt_segment_sum_sorted: bad shape");
    const int nch = (ld + 127) / 128;
    GT_DISPATCH_DT(dt, (k_segment_sum_sorted<T><<<(unsigned)(B * nch), 256, 0, ST>>>((const T*)x, node_off, (int)B, ld, nch, init, out)));
    GT_LAUNCH_CHECK("gt_segment_sum_sorted");
    return 0;
}

extern "C" int gt_add_graph_vec(int dt, const void* x, const float* v, const int32_t* node_graph, int64_t N,
                                int32_t ld, void* y, void* stream) {
    GT_CHECK_ARG(N > 0 && ld > 0 && ld % 4 == 0, "gt_add_graph_vec: bad shape");
    GT_DISPATCH_DT(dt, (k_add_graph_vec<T><<<blocks_for(N * (ld / 4), 256), 256, 0, ST>>>((const T*)x, v, node_graph, N, ld, (T*)y)));
    GT_LAUNCH_CHECK("gt_add_graph_vec");
    return 0;
}

extern "C" int gt_colstats(int dt, const void* x, int64_t M, int32_t ld, double* stats, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(M > 0 && ld > 0 && ld % 4 == 0, "gt_colstats: bad shape");
    int64_t rpb;
    const dim3 grid = stat_grid(M, ld, &rpb);
    GT_DISPATCH_DT(dt, (k_colstats<T, true><<<grid, 256, 0, ST>>>((const T*)x, M, ld, rpb, stats, m_valid)));
    GT_LAUNCH_CHECK("gt_colstats");
    return 0;
}

extern "C" int gt_bn_finalize(const double* stats, int64_t M, int32_t d, int32_t ld, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, int64_t* nbt,
                              float momentum, float eps, int training, float* ssmr, void* stream) {
    GT_CHECK_ARG(M > 0 && d > 0 && ld >= d, "gt_bn_finalize: bad shape");
    GT_CHECK_ARG(training || (running_mean && running_var), "gt_bn_finalize: eval mode needs running stats");
    k_bn_finalize<<<(ld + 127) / 128, 128, 0, ST>>>(stats, M, d, ld, gamma, beta, running_mean, running_var, nbt,
                                                    momentum, eps, training, ssmr);
    GT_LAUNCH_CHECK("gt_bn_finalize");
    return 0;
}

extern "C" int gt_bn_apply_fwd(int dt, const void* x, int64_t M, int32_t d, int32_t ld, const float* ssmr, int relu,
                               const void* resid, const float* gvec, const int32_t* node_graph, void* y,
                               float drop_p, const uint64_t* rng_state, uint64_t salt, void* stream) {
    GT_CHECK_ARG(M > 0 && ld >= d && ld % 4 == 0, "gt_bn_apply_fwd: bad shape");
    GT_CHECK_ARG(!gvec || node_graph, "gt_bn_apply_fwd: gvec needs node_graph");
    GT_DISPATCH_DT(dt, (k_bn_apply<T><<<blocks_for(M * (ld / 4), 256), 256, 0, ST>>>((const T*)x, M, ld, ssmr, relu, (const T*)resid, gvec, node_graph, (T*)y, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_bn_apply_fwd");
    return 0;
}

extern "C" int gt_bn_norm_fwd(int dt, const void* x, int64_t M, int32_t d, int32_t ld, const double* stats,
                              const float* gamma, const float* beta, float* running_mean, float* running_var,
                              int64_t* nbt, float momentum, float eps, int training, int relu, const void* resid,
                              const float* gvec, const int32_t* node_graph, void* y, float* ssmr, float drop_p,
                              const uint64_t* rng_state, uint64_t salt, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(M > 0 && d > 0 && ld >= d && ld % 4 == 0, "gt_bn_norm_fwd: bad shape");
    GT_CHECK_ARG(training ? stats != nullptr : (running_mean && running_var), "gt_bn_norm_fwd: missing statistics");
    GT_CHECK_ARG(!gvec || node_graph, "gt_bn_norm_fwd: gvec needs node_graph");
    Slab sl;
    if (slab_cfg(dt, M, ld, {x, y, resid, gvec, ssmr}, &sl)) {
        const double invM = 1.0 / (double)M, unbias = M > 1 ? (double)M / (double)(M - 1) : 1.0;
        GT_DISPATCH_DT(dt, (k_bn_norm_fwd_slab<T><<<sl.grid, SLAB_THREADS, 0, ST>>>((const T*)x, M, d, ld, sl.tpr, sl.rpi, sl.rows_per_block, invM, unbias, stats, gamma, beta, running_mean, running_var, nbt, momentum, eps, training, relu, (const T*)resid, gvec, node_graph, (T*)y, ssmr, drop_p, rng_state, salt, m_valid)));
        GT_LAUNCH_CHECK("gt_bn_norm_fwd");
        return 0;
    }
    int64_t rpb;
    const dim3 grid = stat_grid(M, ld, &rpb);
    GT_DISPATCH_DT(dt, (k_bn_norm_fwd<T><<<grid, 256, 0, ST>>>((const T*)x, M, d, ld, rpb, stats, gamma, beta, running_mean, running_var, nbt, momentum, eps, training, relu, (const T*)resid, gvec, node_graph, (T*)y, ssmr, drop_p, rng_state, salt, m_valid)));
    GT_LAUNCH_CHECK("gt_bn_norm_fwd");
    return 0;
}

extern "C" int gt_bn_bwd_reduce(int dt, const void* x, const void* dy, int64_t M, int32_t d, int32_t ld,
                                const float* ssmr, int relu, double* red, float drop_p,
                                const uint64_t* rng_state, uint64_t salt, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(M > 0 && ld >= d && ld % 4 == 0, "gt_bn_bwd_reduce: bad shape");
    int64_t rpb;
    const dim3 grid = stat_grid(M, ld, &rpb);
    GT_DISPATCH_DT(dt, (k_bn_bwd_reduce<T><<<grid, 256, 0, ST>>>((const T*)x, (const T*)dy, M, ld, rpb, ssmr, relu, red, drop_p, rng_state, salt, m_valid)));
    GT_LAUNCH_CHECK("gt_bn_bwd_reduce");
    return 0;
}

extern "C" int gt_bn_bwd_apply(int dt, const void* x, const void* dy, int64_t M, int32_t d, int32_t ld,
                               const float* ssmr, const float* gamma, int relu, int training, const double* red,
                               void* dx, float* dgamma, float* dbeta, float drop_p, const uint64_t* rng_state,
                               uint64_t salt, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(M > 0 && ld >= d && ld % 4 == 0, "gt_bn_bwd_apply: bad shape");
    Slab sl;
    if (slab_cfg(dt, M, ld, {x, dy, dx, ssmr}, &sl)) {
        GT_DISPATCH_DT(dt, (k_bn_bwd_apply_slab<T><<<sl.grid, SLAB_THREADS, 0, ST>>>((const T*)x, (const T*)dy, M, d, ld, sl.tpr, sl.rpi, sl.rows_per_block, ssmr, relu, training, red, (T*)dx, dgamma, dbeta, drop_p, rng_state, salt, m_valid)));
        GT_LAUNCH_CHECK("gt_bn_bwd_apply");
        return 0;
    }
    int64_t rpb;
    const dim3 grid = stat_grid(M, ld, &rpb);
    GT_DISPATCH_DT(dt, (k_bn_bwd_apply<T><<<grid, 256, 0, ST>>>((const T*)x, (const T*)dy, M, d, ld, rpb, ssmr, gamma, relu, training, red, (T*)dx, dgamma, dbeta, drop_p, rng_state, salt, m_valid)));
    GT_LAUNCH_CHECK("gt_bn_bwd_apply");
    return 0;
}

extern "C" int gt_layernorm_fwd(int dt, const void* x, const void* resid, const int32_t* in_rows, const float* cls,
                                int64_t M, int32_t d, const float* gamma, const float* beta, float eps, void* y,
                                void* presum, float* mean_rstd, float drop_p, const uint64_t* rng_state, uint64_t salt,
                                void* stream) {
    GT_CHECK_ARG(M > 0 && d > 0 && d % 4 == 0 && d <= LN_MAXV * 128, "gt_layernorm_fwd: d=%d must be a multiple of 4 and <= %d", d, LN_MAXV * 128);
    const int grid_f = (int)((M + 7) / 8);
#define LNF(MAXV, VW) k_layernorm_fwd<T, MAXV, VW><<<grid_f, 256, 0, ST>>>((const T*)x, (const T*)resid, in_rows, cls, M, d, gamma, beta, eps, (T*)y, (T*)presum, mean_rstd, drop_p, rng_state, salt)
    // bf16 rows wider than one 4-element pass: 16-byte accesses (8 channels per lane and vector)
    const bool wide = dt == GT_BF16 && d % 8 == 0 && d > 128 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                      (!resid || (uintptr_t)resid % 16 == 0) && (!presum || (uintptr_t)presum % 16 == 0);
    if (wide) {
        if (d <= 256) k_layernorm_fwd<bf16, 1, 8><<<grid_f, 256, 0, ST>>>((const bf16*)x, (const bf16*)resid, in_rows, cls, M, d, gamma, beta, eps, (bf16*)y, (bf16*)presum, mean_rstd, drop_p, rng_state, salt);
        else if (d <= 512) k_layernorm_fwd<bf16, 2, 8><<<grid_f, 256, 0, ST>>>((const bf16*)x, (const bf16*)resid, in_rows, cls, M, d, gamma, beta, eps, (bf16*)y, (bf16*)presum, mean_rstd, drop_p, rng_state, salt);
        else k_layernorm_fwd<bf16, 4, 8><<<grid_f, 256, 0, ST>>>((const bf16*)x, (const bf16*)resid, in_rows, cls, M, d, gamma, beta, eps, (bf16*)y, (bf16*)presum, mean_rstd, drop_p, rng_state, salt);
    } else {
        GT_DISPATCH_DT(dt, {
            if (d <= 256) LNF(2, 4);
            else LNF(LN_MAXV, 4);
        });
    }
#undef LNF
    GT_LAUNCH_CHECK("gt_layernorm_fwd");
    return 0;
}

extern "C" int gt_layernorm_bwd(int dt, const void* dy, const void* presum, const float* mean_rstd,
                                const int32_t* out_rows, int64_t M, int32_t d, const float* gamma, void* dx,
                                float* dgamma, float* dbeta, float* dcls, void* dx_drop, float drop_p,
                                const uint64_t* rng_state, uint64_t salt, void* stream) {
    GT_CHECK_ARG(!dx_drop || !out_rows, "gt_layernorm_bwd: dropout and row scatter are exclusive");
    GT_CHECK_ARG(M > 0 && d > 0 && d % 4 == 0 && d <= LN_MAXV * 128, "gt_layernorm_bwd: bad d=%d", d);
    // every block ends with 2*d global atomics for dgamma / dbeta (<= 444 serialised L2 atomics per address); three
    // 8-warp blocks per SM keep a warp at <= ~3 sequential rows (each row is a load -> 2 shuffle reductions -> store
    // latency chain, so rows per warp, not bytes, set the duration at these sizes)
    const int br = d <= 128 ? 4 : (d <= 256 ? 2 : 1);                  // rows per warp iteration (kernel's R)
    const int grid = blocks_for((M + br - 1) / br, 8, 3 * kNumSMs);
    const bool wide = env_int("GT_LN_WIDE", 1) && dt == GT_BF16 && d % 8 == 0 && d > 128 && ((uintptr_t)dy % 16 == 0) &&
                      ((uintptr_t)presum % 16 == 0) && ((uintptr_t)dx % 16 == 0) && (!dx_drop || (uintptr_t)dx_drop % 16 == 0);
    // grid = exactly the blocks that are resident at once (the kernels loop grid-stride; a partial second wave costs a
    // full block duration: 444 blocks on 2 x 148 slots ran 1.5 waves)
#define LNB(TT, MAXV, VW, RR, G) do { \
        static int occ = 0; \
        if (!occ) { \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_layernorm_bwd<TT, MAXV, VW, RR>, 256, 2 * d * sizeof(float)); \
            if (occ < 1) occ = 1; \
        } \
        int g_ = (G); \
        if (g_ > occ * kNumSMs) g_ = occ * kNumSMs; \
        k_layernorm_bwd<TT, MAXV, VW, RR><<<g_, 256, 2 * d * sizeof(float), ST>>>((const TT*)dy, (const TT*)presum, mean_rstd, out_rows, M, d, gamma, (TT*)dx, dgamma, dbeta, dcls, (TT*)dx_drop, drop_p, rng_state, salt); \
    } while (0)
    if (wide) {      // 16-byte accesses, four (d <= 256) / two (d <= 512) / one row(s) of a warp in flight
        if (d <= 256) LNB(bf16, 1, 8, 4, blocks_for((M + 3) / 4, 8, 4 * kNumSMs));
        else if (d <= 512) LNB(bf16, 2, 8, 2, blocks_for((M + 1) / 2, 8, 4 * kNumSMs));
        else LNB(bf16, 4, 8, 1, blocks_for(M, 8, 4 * kNumSMs));
    } else {
        GT_DISPATCH_DT(dt, {
            if (d <= 128) LNB(T, 1, 4, 4, grid);
            else if (d <= 256) LNB(T, 2, 4, 2, grid);
            else LNB(T, LN_MAXV, 4, 1, grid);
        });
    }
#undef LNB
    GT_LAUNCH_CHECK("gt_layernorm_bwd");
    return 0;
}

extern "C" int gt_gather_rows(int dt, const void* src, const int32_t* rows, const float* cls, int64_t M, int32_t ld,
                              void* dst, void* stream) {
    GT_CHECK_ARG(M > 0 && ld > 0 && ld % 4 == 0, "gt_gather_rows: bad shape");
    GT_DISPATCH_DT(dt, (k_gather_rows<T><<<blocks_for(M * (ld / 4), 256), 256, 0, ST>>>((const T*)src, rows, cls, M, ld, (T*)dst)));
    GT_LAUNCH_CHECK("gt_gather_rows");
    return 0;
}

extern "C" int gt_scatter_rows(int dt, const void* dsrc, const int32_t* rows, int64_t M, int32_t ld, void* ddst,
                               float* dcls, void* stream) {
    GT_CHECK_ARG(M > 0 && ld > 0 && ld % 4 == 0, "gt_scatter_rows: bad shape");
    GT_DISPATCH_DT(dt, (k_scatter_rows<T><<<blocks_for(M * (ld / 4), 256), 256, 0, ST>>>((const T*)dsrc, rows, M, ld, (T*)ddst, dcls)));
    GT_LAUNCH_CHECK("gt_scatter_rows");
    return 0;
}

extern "C" int gt_pad_batch_fwd(int dt, const void* h, const int32_t* node_off, int64_t B, int64_t S, int32_t ld,
                                void* padded, uint8_t* mask, void* stream) {
    GT_CHECK_ARG(B > 0 && S > 0 && ld > 0 && ld % 4 == 0, "gt_pad_batch_fwd: bad shape");
    GT_DISPATCH_DT(dt, (k_pad_fwd<T><<<blocks_for(S * B * (ld / 4), 256), 256, 0, ST>>>((const T*)h, node_off, B, S, ld, (T*)padded, mask)));
    GT_LAUNCH_CHECK("gt_pad_batch_fwd");
    return 0;
}

extern "C" int gt_pad_batch_bwd(int dt, const void* dpadded, const int32_t* node_off, const int32_t* node_graph,
                                int64_t B, int64_t S, int64_t N, int32_t ld, void* dh, void* stream) {
    GT_CHECK_ARG(B > 0 && S > 0 && N > 0 && ld % 4 == 0, "gt_pad_batch_bwd: bad shape");
    GT_DISPATCH_DT(dt, (k_pad_bwd<T><<<blocks_for(N * (ld / 4), 256), 256, 0, ST>>>((const T*)dpadded, node_off, node_graph, B, S, N, ld, (T*)dh)));
    GT_LAUNCH_CHECK("gt_pad_batch_bwd");
    return 0;
}

static int fill_cols(EmbCols& c, int32_t ncol, const int64_t* const* idx, const int64_t* stride,
                     const int64_t* clamp, const float* const* table, float* const* dtable) {
    GT_CHECK_ARG(ncol >= 1 && ncol <= EMB_MAXCOL, "embed: ncol=%d not in 1..%d", ncol, EMB_MAXCOL);
    c.ncol = ncol;
    for (int i = 0; i < ncol; ++i) {
        c.idx[i] = idx[i];
        c.stride[i] = stride[i];
        c.clamp[i] = clamp[i];
        c.table[i] = table ? table[i] : nullptr;
        c.dtable[i] = dtable ? dtable[i] : nullptr;
    }
    return 0;
}

extern "C" int gt_embed_sum_fwd(int dt, void* out, int64_t N, int32_t d, int32_t ld, int32_t ncol,
                                const int64_t* const* idx_host, const int64_t* stride_host,
                                const int64_t* clamp_host, const float* const* table_host, void* stream) {
    GT_CHECK_ARG(N > 0 && d % 4 == 0 && ld >= d && ld % 4 == 0, "gt_embed_sum_fwd: d=%d must be a multiple of 4", d);
    EmbCols c;
    if (int r = fill_cols(c, ncol, idx_host, stride_host, clamp_host, table_host, nullptr)) return r;
    GT_DISPATCH_DT(dt, (k_embed_fwd<T><<<blocks_for(N * (ld / 4), 256), 256, 0, ST>>>(c, N, d, ld, (T*)out)));
    GT_LAUNCH_CHECK("gt_embed_sum_fwd");
    return 0;
}

extern "C" int gt_embed_sum_bwd(int dt, const void* dout, int64_t N, int32_t d, int32_t ld, int32_t ncol,
                                const int64_t* const* idx_host, const int64_t* stride_host,
                                const int64_t* clamp_host, float* const* dtable_host, void* stream) {
    GT_CHECK_ARG(N > 0 && d % 4 == 0 && ld >= d && ld % 4 == 0, "gt_embed_sum_bwd: bad shape");
    EmbCols c;
    if (int r = fill_cols(c, ncol, idx_host, stride_host, clamp_host, nullptr, dtable_host)) return r;
    dim3 grid((unsigned)((N + EMB_NB - 1) / EMB_NB), (unsigned)((d + 31) / 32));
    const size_t smem = (size_t)EMB_SMEM_ROWS * 32 * sizeof(float) + (size_t)EMB_MAXCOL * EMB_NB * sizeof(int32_t);
    GT_DISPATCH_DT(dt, {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_embed_bwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
        k_embed_bwd<T><<<grid, 256, smem, ST>>>(c, N, d, ld, (const T*)dout);
    });
    GT_LAUNCH_CHECK("gt_embed_sum_bwd");
    return 0;
}

// ---- embedding gradients of SMALL tables as a tensor-core contraction --------------------------------------------
// d_table[v] = sum over the nodes whose index is v of dout[node] is OneHot^T [R x N] . dout [N x d]: the few-row
// vocabularies (atom / node-type / depth tables) serialise atomics on a handful of rows, the contraction does not.
// gt_onehot writes the bf16 one-hot operand [N, R_pad] (row i has a 1 at base_c + idx_c[i] for every column c with
// base_c >= 0) once per batch; the backward is one gt_gemm (split-K over the nodes, fp32) + gt_embed_unpack.
namespace gt {
struct OneHotCols {
    const int64_t* idx[EMB_MAXCOL];
    int64_t stride[EMB_MAXCOL];
    int64_t clamp[EMB_MAXCOL];
    int32_t base[EMB_MAXCOL];
    int32_t rows[EMB_MAXCOL];
    float* dtable[EMB_MAXCOL];
    int ncol;
};
__global__ void k_onehot(OneHotCols cols, int64_t N, int groups, bf16* __restrict__ out) {
    // thread = (row, 16-byte group); four rows of a thread in flight (index loads first, then the stores): the pass is a
    // pure write stream whose only latency is the index fetch
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rows_per_pass = stride / groups;       // host sizes the grid so that groups divides the thread count
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int g = (int)(t % groups);
    for (int64_t r0 = t / groups; r0 < N; r0 += 4 * rows_per_pass) {
        uint32_t w[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0u;
        for (int c = 0; c < cols.ncol; ++c) {
            if (cols.base[c] < 0) continue;
            int64_t id[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = r0 + u * rows_per_pass;
                id[u] = r < N ? cols.idx[c][r * cols.stride[c]] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t v = id[u] > cols.clamp[c] ? cols.clamp[c] : id[u];
                const int pos = cols.base[c] + (int)v;
                if ((pos >> 3) == g) w[u][(pos & 7) >> 1] |= 0x3F80u << (16 * (pos & 1));   // bf16 1.0
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + u * rows_per_pass;
            if (r < N) *reinterpret_cast<uint4*>(out + (r * groups + g) * 8) = make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]);
        }
    }
}
__global__ void k_embed_unpack(OneHotCols cols, const float* __restrict__ temp, int ld_t, int d, int R) {
    const int vpr = d / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * vpr; i += gridDim.x * blockDim.x) {
        const int r = i / vpr, c0 = (i - r * vpr) * 4;
        for (int c = 0; c < cols.ncol; ++c) {
            if (cols.base[c] < 0 || r < cols.base[c] || r >= cols.base[c] + cols.rows[c]) continue;
            const float4 v = *reinterpret_cast<const float4*>(temp + (int64_t)r * ld_t + c0);
            float* dst = cols.dtable[c] + (int64_t)(r - cols.base[c]) * d + c0;
            dst[0] += v.x; dst[1] += v.y; dst[2] += v.z; dst[3] += v.w;     // single writer per element
        }
    }
}
}  // namespace gt

static int fill_onehot(OneHotCols& c, int32_t ncol, const int64_t* const* idx, const int64_t* stride, const int64_t* clamp,
                       const int32_t* base, const int32_t* rows, float* const* dtable) {
    GT_CHECK_ARG(ncol >= 1 && ncol <= EMB_MAXCOL, "onehot: ncol=%d not in 1..%d", ncol, EMB_MAXCOL);
    c.ncol = ncol;
    for (int i = 0; i < ncol; ++i) {
        c.idx[i] = idx ? idx[i] : nullptr;
        c.stride[i] = stride ? stride[i] : 0;
        c.clamp[i] = clamp ? clamp[i] : 0;
        c.base[i] = base[i];
        c.rows[i] = rows ? rows[i] : 0;
        c.dtable[i] = dtable ? dtable[i] : nullptr;
    }
    return 0;
}

extern "C" int gt_onehot(int64_t N, int32_t ncol, const int64_t* const* idx_host, const int64_t* stride_host,
                         const int64_t* clamp_host, const int32_t* base_host, int32_t r_pad, void* out, void* stream) {
    GT_CHECK_ARG(N > 0 && r_pad > 0 && r_pad % 8 == 0, "gt_onehot: bad shape");
    OneHotCols c;
    if (int r = fill_onehot(c, ncol, idx_host, stride_host, clamp_host, base_host, nullptr, nullptr)) return r;
    const int groups = r_pad / 8;
    // every thread keeps its 16-byte group for all its rows: the block size is a multiple of the groups per row
    const int bs = (256 / groups) * groups > 0 ? (256 / groups) * groups : groups;
    GT_CHECK_ARG(bs <= 1024, "gt_onehot: r_pad=%d too wide", r_pad);
    k_onehot<<<blocks_for((N * groups + 3) / 4, bs), bs, 0, ST>>>(c, N, groups, (bf16*)out);
    GT_LAUNCH_CHECK("gt_onehot");
    return 0;
}

extern "C" int gt_embed_unpack(const float* temp, int32_t ld_t, int32_t d, int32_t R, int32_t ncol, const int32_t* base_host,
                               const int32_t* rows_host, float* const* dtable_host, void* stream) {
    GT_CHECK_ARG(R > 0 && d % 4 == 0 && ld_t >= d && ld_t % 4 == 0, "gt_embed_unpack: bad shape");
    OneHotCols c;
    if (int r = fill_onehot(c, ncol, nullptr, nullptr, nullptr, base_host, rows_host, dtable_host)) return r;
    k_embed_unpack<<<blocks_for((int64_t)R * (d / 4), 256), 256, 0, ST>>>(c, temp, ld_t, d, R);
    GT_LAUNCH_CHECK("gt_embed_unpack");
    return 0;
}

template <typename TI>
static void cast_pad_out(int dt_out, const TI* src, int64_t ri, int64_t ci, int64_t li, void* dst, int64_t ro,
                         int64_t co, int64_t lo, cudaStream_t st) {
    const int grid = blocks_for(ro * co, 256);
    if (dt_out == GT_F32) k_cast_pad<TI, float><<<grid, 256, 0, st>>>(src, ri, ci, li, (float*)dst, ro, co, lo);
    else k_cast_pad<TI, bf16><<<grid, 256, 0, st>>>(src, ri, ci, li, (bf16*)dst, ro, co, lo);
}

extern "C" int gt_cast_pad(int dt_in, const void* src, int64_t rows_in, int64_t cols_in, int64_t ld_in, int dt_out,
                           void* dst, int64_t rows_out, int64_t cols_out, int64_t ld_out, void* stream) {
    GT_CHECK_ARG(rows_out >= rows_in && cols_out >= cols_in && ld_out >= cols_out && (ld_in >= cols_in || ld_in == 0), "gt_cast_pad: bad shape");
    GT_CHECK_ARG((dt_in == GT_F32 || dt_in == GT_BF16) && (dt_out == GT_F32 || dt_out == GT_BF16), "gt_cast_pad: bad dtype");
    if (rows_out * cols_out == 0) return 0;
    if (dt_in == GT_F32) cast_pad_out<float>(dt_out, (const float*)src, rows_in, cols_in, ld_in, dst, rows_out, cols_out, ld_out, ST);
    else cast_pad_out<bf16>(dt_out, (const bf16*)src, rows_in, cols_in, ld_in, dst, rows_out, cols_out, ld_out, ST);
    GT_LAUNCH_CHECK("gt_cast_pad");
    return 0;
}

extern "C" int gt_add_blocks(const float* src, int32_t ld_src, int32_t n, float* const* dst_host, const int32_t* ld_dst_host,
                             const int32_t* r0_host, const int32_t* c0_host, const int32_t* rows_host, const int32_t* cols_host,
                             void* stream) {
    GT_CHECK_ARG(src && n > 0 && n <= 16, "gt_add_blocks: 1..16 blocks");
    AddBlocks ab;
    ab.n = n;
    int maxel = 1;
    for (int b = 0; b < n; ++b) {
        ab.dst[b] = dst_host[b]; ab.ld_dst[b] = ld_dst_host[b]; ab.r0[b] = r0_host[b]; ab.c0[b] = c0_host[b];
        ab.rows[b] = rows_host[b]; ab.cols[b] = cols_host[b];
        GT_CHECK_ARG(ab.dst[b] && ab.rows[b] >= 0 && ab.cols[b] >= 0, "gt_add_blocks: bad block %d", b);
        if (ab.rows[b] * ab.cols[b] > maxel) maxel = ab.rows[b] * ab.cols[b];
    }
    dim3 grid((unsigned)blocks_for(maxel, 256, 64), (unsigned)n, 1);
    k_add_blocks<<<grid, 256, 0, ST>>>(ab, src, ld_src);
    GT_LAUNCH_CHECK("gt_add_blocks");
    return 0;
}

// fp32 -> three bf16 terms (x = p0 + p1 + p2 up to 2^-24 |x|): p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1)
namespace gt {
__global__ void k_split3(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src, bf16* __restrict__ dst,
                         int64_t ld_dst) {
    const int64_t total = rows * ld_dst, plane = rows * ld_dst;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld_dst, c = i - r * ld_dst;
        float x = c < cols ? src[r * ld_src + c] : 0.f;
        const bf16 p0 = __float2bfloat16_rn(x);
        x -= __bfloat162float(p0);
        const bf16 p1 = __float2bfloat16_rn(x);
        x -= __bfloat162float(p1);
        dst[i] = p0;
        dst[plane + i] = p1;
        dst[2 * plane + i] = __float2bfloat16_rn(x);
    }
}
}  // namespace gt
extern "C" int gt_split3(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst, int64_t ld_dst, void* stream) {
    GT_CHECK_ARG(rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "gt_split3: bad shape");
    gt::k_split3<<<blocks_for(rows * ld_dst, 256), 256, 0, ST>>>(src, rows, cols, ld_src, (gt::bf16*)dst, ld_dst);
    GT_LAUNCH_CHECK("gt_split3");
    return 0;
}

extern "C" int gt_cast_multi(const int64_t* desc_dev, int32_t n, int64_t total_blocks, void* stream) {
    GT_CHECK_ARG(desc_dev && n > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "gt_cast_multi: bad arguments");
    k_cast_multi<<<(unsigned)total_blocks, 256, 0, ST>>>(desc_dev, n);
    GT_LAUNCH_CHECK("gt_cast_multi");
    return 0;
}

extern "C" int gt_relu_bwd(int dt, const void* dy, const void* y, int64_t n, void* dz, float scale, void* stream) {
    GT_CHECK_ARG(n > 0 && n % 4 == 0, "gt_relu_bwd: element count must be a positive multiple of 4");
    GT_DISPATCH_DT(dt, (k_relu_bwd<T><<<blocks_for(n / 4, 256), 256, 0, ST>>>((const T*)dy, (const T*)y, n / 4, (T*)dz, scale)));
    GT_LAUNCH_CHECK("gt_relu_bwd");
    return 0;
}

extern "C" int gt_relu_bwd_colsum(int dt, const void* dy, const void* y, int64_t M, int64_t N, int64_t ld, void* dz,
                                  float scale, float* colsum, void* stream) {
    GT_CHECK_ARG(M > 0 && N > 0 && ld >= N && colsum, "gt_relu_bwd_colsum: bad shape");
    const int esz = dt == GT_BF16 ? 2 : 4;
    const int Vw = 16 / esz;
    GT_CHECK_ARG(ld % Vw == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dz) % 16 == 0),
                 "gt_relu_bwd_colsum: rows must be 16-byte aligned (ld %% %d == 0)", Vw);
    const int gx = (int)((ld + 16 * Vw - 1) / (16 * Vw));
    int64_t gy = (4 * kNumSMs + gx - 1) / gx;
    if (gy > (M + 63) / 64) gy = (M + 63) / 64;
    const int64_t rpb = (M + gy - 1) / gy;
    gy = (M + rpb - 1) / rpb;
    GT_DISPATCH_DT(dt, (k_relu_bwd_colsum<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ST>>>((const T*)dy, (const T*)y, M, (int)N, ld, rpb, (T*)dz, scale, colsum)));
    GT_LAUNCH_CHECK("gt_relu_bwd_colsum");
    return 0;
}

extern "C" int gt_dropout(int dt, const void* x, int64_t n, void* y, float drop_p, const uint64_t* rng_state,
                          uint64_t salt, void* stream) {
    GT_CHECK_ARG(n > 0 && n % 4 == 0, "gt_dropout: element count must be a positive multiple of 4");
    GT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gt_dropout: p=%f not in [0,1)", drop_p);
    GT_DISPATCH_DT(dt, (k_dropout<T><<<blocks_for(n / 4, 256), 256, 0, ST>>>((const T*)x, n / 4, (T*)y, drop_p, rng_state, salt)));
    GT_LAUNCH_CHECK("gt_dropout");
    return 0;
}

extern "C" int gt_rng_advance(uint64_t* rng_state, void* stream) {
    GT_CHECK_ARG(rng_state != nullptr, "gt_rng_advance: null state");
    k_rng_advance<<<1, 1, 0, ST>>>(rng_state);
    GT_LAUNCH_CHECK("gt_rng_advance");
    return 0;
}

extern "C" int gt_colsum(int dt, const void* X, int64_t M, int64_t N, int64_t ld, float* out, void* stream) {
    GT_CHECK_ARG(M > 0 && N > 0 && ld >= N, "gt_colsum: bad shape");
    const int esz = dt == GT_BF16 ? 2 : 4;
    const int Vw = 16 / esz;
    if (M >= 4096 && ld % Vw == 0 && (uintptr_t)X % 16 == 0 && (N + Vw - 1) / Vw * Vw <= ld) {   // tall matrices: 16-byte accesses
        const int gx = (int)((N + 16 * Vw - 1) / (16 * Vw));
        int64_t gy = (4 * kNumSMs + gx - 1) / gx;
        if (gy > (M + 255) / 256) gy = (M + 255) / 256;
        const int64_t rpb = (M + gy - 1) / gy;
        gy = (M + rpb - 1) / rpb;
        GT_DISPATCH_DT(dt, (k_colsum_w<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ST>>>((const T*)X, M, (int)N, ld, rpb, out)));
        GT_LAUNCH_CHECK("gt_colsum");
        return 0;
    }
    if (ld % 4 == 0 && (uintptr_t)X % (4 * esz) == 0 && (N + 3) / 4 * 4 <= ld) {
        int64_t rpb;
        const dim3 grid = stat_grid(M, (int)((N + 3) / 4 * 4), &rpb);
        GT_DISPATCH_DT(dt, (k_colsum_v<T><<<grid, 256, 0, ST>>>((const T*)X, M, (int)N, ld, rpb, out)));
        GT_LAUNCH_CHECK("gt_colsum");
        return 0;
    }
    int slabs = (int)((M + 255) / 256);
    if (slabs > 64) slabs = 64;
    dim3 grid((unsigned)((N + 31) / 32), slabs), block(32, 8);
    GT_DISPATCH_DT(dt, (k_colsum<T><<<grid, block, 0, ST>>>((const T*)X, M, N, ld, out)));
    GT_LAUNCH_CHECK("gt_colsum");
    return 0;
}
