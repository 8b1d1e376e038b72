// Stage 2 on tensor cores for SMALL graphs: when every graph of the batch has at most 128 tokens (molecule and TU
// workloads: <= 61 tokens) the block-diagonal attention of a graph never leaves one 128-row tile.  gt_mha_local_tiles
// packs consecutive graphs into graph-aligned tiles of <= 128 packed token rows; a CTA = (tile, head) then has the SAME
// rows as queries and as keys, so forward and backward are a fixed, loop-free sequence of tcgen05 MMAs over tiles that
// stay in shared memory / TMEM:
//   forward : S = Q K^T -> masked softmax (thread = row) -> P (bf16, smem) -> O = P V -> out, lse
//   backward: S = Q K^T, dP = dO V^T -> P, dS (smem) -> dV = P^T dO, dK = dS^T Q, dQ = dS K        (ONE kernel, no atomics,
//             no separate delta pass: delta = rowsum(dO o O) is a 64-128 byte dot product per thread)
// against the streamed kernels of attn_tc.cu (three launches per backward, 64-row streaming loop, key range up to two
// tiles per query tile).  Replaces the same reference ops (modules/transformer_encoder.py:28-32,59;
// F.multi_head_attention_forward) and uses the same dropout hash, so the two paths are interchangeable.
#include <stdlib.h>

#include "attn_common.cuh"

namespace gt {
namespace tc {

// warp 0: TMA, warp 1: MMA, warps 2..9: math - two warps per TMEM lane quarter (32 tile rows), each owning 64 of the 128
// key columns.  The math is a chain of dependent tcgen05.ld / MUFU / FADD per row, so the second warp per row (and the
// four-way split accumulators below) buys issue slots that one warp per row leaves idle (IPC 0.26 measured with one).
constexpr int LOC_THREADS = 320;

struct LocParams {
    const int2* row_bounds;    // [lo, hi) token rows of the graph of every row
    const int2* tiles;         // (first row, rows) of every graph-aligned tile; rows == 0: unused slot
    const int32_t* count;      // tiles used; < 0: the batch violated the host's bound (a graph > 128 tokens / too many tiles)
    void* out;                 // forward: written; backward: read (delta)
    const void* dout;
    float* lse;
    void* dqkv;
    const uint64_t* rng;
    uint64_t salt;
    int64_t n_rows;
    int nhead, d;
    float scale, scale_log2, drop_p;
    unsigned long long* trace;   // debug (GT_LOC_TRACE=1): %globaltimer stamps of the phases of CTA (0, 0), else NULL
};
#define LOC_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[slot] = t_; } } while (0)

// Rows that belong to no graph (the unused tail of the static row bound: bucket slack, truncated graphs) are in no tile.
// Every CTA (x, head) clears the head's column slice of such rows in its stripe of the row range, so the caller needs no
// zero-fill pass over the whole output: `parts` column blocks of width ld_cols/parts... (out: 1 part of d; dqkv: 3 of d)
template <int DH>
__device__ __forceinline__ void loc_clear_tail(const int2* __restrict__ row_bounds, int64_t n_rows, bf16* __restrict__ dst,
                                               int ld, int parts, int d, int h) {
    const int64_t chunk = (n_rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(r0 + chunk, n_rows);
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const int2 rb = row_bounds[r];
        if (rb.y > rb.x) continue;
        for (int part = 0; part < parts; ++part) {
            uint4* q = reinterpret_cast<uint4*>(dst + r * ld + part * d + h * DH);
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) q[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// store 8 consecutive bf16 of row r, columns [c, c+8) of a [128 x 128] two-block swizzled tile
__device__ __forceinline__ void st_p8(uint32_t tile, int r, int c, const float (&v)[8]) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + p_chunk_off(r, c)), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                 "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                 "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
}
__device__ __forceinline__ void st_zero32(uint32_t tile, int r, int c) {
#pragma unroll
    for (int i = 0; i < 32; i += 8)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile + p_chunk_off(r, c + i)), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_row_bf16(bf16* dst, const uint32_t* rr, int n, float mul) {
    for (int i = 0; i < n; i += 8) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(rr[i]) * mul, __uint_as_float(rr[i + 1]) * mul);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(rr[i + 2]) * mul, __uint_as_float(rr[i + 3]) * mul);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(rr[i + 4]) * mul, __uint_as_float(rr[i + 5]) * mul);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(rr[i + 6]) * mul, __uint_as_float(rr[i + 7]) * mul);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + i) = pk;
    }
}

// ------------------------------------------------------------------------------------------------------- forward
// TMEM: S [0,128); O reuses [0, DH) once every thread holds its probabilities in shared memory.  smem: the 32 KB P tile
// overlays Q and K (dead once S = Q K^T has completed) + the V tile: 40 KB at DH = 32, 48 KB at DH = 64 -> four CTAs
// per SM (the TMEM limit at 128 columns each).
template <int DH, bool DROP>
__global__ void __launch_bounds__(LOC_THREADS, 3)
k_mha_loc_fwd(const __grid_constant__ CUtensorMap tma_qkv, const LocParams p) {
    constexpr int PITCH = DH * 2, TILE = 128 * PITCH;
    constexpr uint32_t TMEM_COLS = 128;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ld_full, s_full, p_full, o_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ float xchg[2][128];                             // row max / row sum exchange between the two column halves
    if (threadIdx.x == 0) LOC_TRACE(0);
    if (p.count && p.count[0] < 0) {   // fail loudly: a NaN row poisons the loss instead of silently skipping graphs
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < p.d)
            reinterpret_cast<bf16*>(p.out)[threadIdx.x] = __float2bfloat16_rn(__int_as_float(0x7fc00000));
        return;
    }
    loc_clear_tail<DH>(p.row_bounds, p.n_rows, (bf16*)p.out, p.d, 1, p.d, (int)blockIdx.y);
    const int2 tile = p.tiles[blockIdx.x];
    const int row0 = tile.x, nrows = tile.y;
    if (nrows <= 0) return;                                   // unused slot of the (host-side) upper bound
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = blockIdx.y;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base, k_s = q_s + TILE, p_s = base, v_s = base + 32768;   // 2 * TILE <= 32 KB
    if (threadIdx.x == 0) {
        mbar_init(&ld_full, 1);
        mbar_init(&s_full, 1);
        mbar_init(&p_full, 256);
        mbar_init(&o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) LOC_TRACE(1);

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&ld_full, 3 * TILE);
            tma_load_2d(q_s, &tma_qkv, &ld_full, h * DH, row0);
            tma_load_2d(k_s, &tma_qkv, &ld_full, p.d + h * DH, row0);
            tma_load_2d(v_s, &tma_qkv, &ld_full, 2 * p.d + h * DH, row0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_s = idesc_f16(false, false, 128, 128);
            const uint32_t id_o = idesc_f16(false, true, 128, DH);
            mbar_wait(&ld_full, 0);
            LOC_TRACE(2);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
                umma_f16(tmem, desc_k(q_s + k * 32, PITCH), desc_k(k_s + k * 32, PITCH), id_s, k > 0);
            umma_commit(&s_full);
            mbar_wait(&p_full, 0);                             // every row's P is in smem, S is dead
            LOC_TRACE(5);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_f16(tmem, desc_k(p_s + (k >> 2) * 16384 + (k & 3) * 32, 128), desc_mn(v_s + k * 16 * PITCH, PITCH), id_o, k > 0);
            umma_commit(&o_full);
        }
    } else {
        const int qd = warp & 3, half = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const int cb = half * 64;                              // this warp's key columns: [cb, cb + 64)
        const int64_t row = (int64_t)row0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
        int lo = 0, hi = 0;                                    // key columns of this row inside the tile
        if (r < nrows) {
            const int2 rb = p.row_bounds[row];
            lo = rb.x - row0, hi = min(rb.y - row0, 128);
        }
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        const uint32_t rk = drop_row_key(dr, att_row_id_tc(h, row, p.n_rows));
        const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : 0x7fffffff);
        const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
        mbar_wait(&s_full, 0);
        if (threadIdx.x == 64) LOC_TRACE(3);
        tc_fence_after();
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c = cb; c < cb + 64; c += 32) {
            if (c + 32 <= wlo || c >= whi) continue;           // warp-uniform
            uint32_t rr[32];
            tmem_ld32(t_lane + c, rr);
            const uint32_t vm = range_mask32(lo, hi, c);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (vm & (1u << i)) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(rr[i]));
        }
        float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        xchg[half][r] = m;
        named_bar_sync(1, 256);
        m = fmaxf(xchg[0][r], xchg[1][r]);
        const float m_use = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        named_bar_sync(1, 256);                                // xchg is reused for the row sums
        if (threadIdx.x == 64) LOC_TRACE(4);
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        const float sl2 = p.scale_log2;
#pragma unroll 1
        for (int c = cb; c < cb + 64; c += 32) {
            if (c + 32 <= wlo || c >= whi) {
                st_zero32(p_s, r, c);
                continue;
            }
            uint32_t rr[32];
            tmem_ld32(t_lane + c, rr);
            const uint32_t vm = range_mask32(lo, hi, c);
            // branch-free: the exponential of every column is formed (masked columns hold finite scores of other
            // graphs; an overflow to +inf is discarded by the select), validity comes from the bit mask
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 8) {
                float pv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float e = ex2_approx(fmaf(__uint_as_float(rr[i0 + i]), sl2, -m_use));
                    e = (vm & (1u << (i0 + i))) ? e : 0.f;
                    ls[i & 3] += e;
                    if (DROP) e *= drop_elem(dr, rk, (uint32_t)(row0 + c + i0 + i));
                    pv[i] = e;
                }
                st_p8(p_s, r, c + i0, pv);
            }
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(&p_full);
        float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
        xchg[half][r] = l;
        named_bar_sync(1, 256);
        l = xchg[0][r] + xchg[1][r];
        mbar_wait(&o_full, 0);
        if (threadIdx.x == 64) LOC_TRACE(6);
        tc_fence_after();
        const float inv = l > 0.f ? 1.f / l : 0.f;
        constexpr int OC = DH / 2;                             // each warp of the pair stores half of the head dimension
        if (OC == 32) {
            uint32_t rr[32];
            tmem_ld32(t_lane + half * OC, rr);
            if (r < nrows) st_row_bf16((bf16*)p.out + row * p.d + h * DH + half * OC, rr, 32, inv);
        } else {
            uint32_t rr[16];
            tmem_ld16(t_lane + half * OC, rr);
            if (r < nrows) st_row_bf16((bf16*)p.out + row * p.d + h * DH + half * OC, rr, 16, inv);
        }
        tc_fence_before();
        if (r < nrows && half == 0) p.lse[(int64_t)h * p.n_rows + row] = l > 0.f ? (m_use + log2f(l)) * LN2 : 0.f;
        if (threadIdx.x == 64) LOC_TRACE(7);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        if (lane == 0) LOC_TRACE(8);
    }
}

// ------------------------------------------------------------------------------------------------------ backward
// TMEM: S [0,128) | dP [128,256); once P and dS sit in shared memory the accumulators reuse the columns:
// dV [0,DH) | dK [DH,2DH) | dQ [2DH,3DH).  smem: Q, K, V, dO tiles + P + dS = 4 * 128 * DH * 2 + 64 KB (96 KB at
// DH = 32, 128 KB at DH = 64).
template <int DH, bool DROP>
__global__ void __launch_bounds__(LOC_THREADS, 2)
k_mha_loc_bwd(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do, const LocParams p) {
    constexpr int PITCH = DH * 2, TILE = 128 * PITCH;
    constexpr uint32_t TMEM_COLS = 256;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ld_full, sdp_full, pds_full, acc_full;
    __shared__ uint32_t tmem_base_s;
    loc_clear_tail<DH>(p.row_bounds, p.n_rows, (bf16*)p.dqkv, 3 * p.d, 3, p.d, (int)blockIdx.y);
    const int2 tile = p.tiles[blockIdx.x];
    const int row0 = tile.x, nrows = tile.y;
    if (nrows <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = blockIdx.y;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base, k_s = q_s + TILE, v_s = k_s + TILE, do_s = v_s + TILE, ds_s = do_s + TILE, pp_s = ds_s + 32768;
    if (threadIdx.x == 0) {
        mbar_init(&ld_full, 1);
        mbar_init(&sdp_full, 1);
        mbar_init(&pds_full, 256);
        mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int colQ = h * DH, colK = p.d + h * DH, colV = 2 * p.d + h * DH;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&ld_full, 4 * TILE);
            tma_load_2d(q_s, &tma_qkv, &ld_full, colQ, row0);
            tma_load_2d(k_s, &tma_qkv, &ld_full, colK, row0);
            tma_load_2d(v_s, &tma_qkv, &ld_full, colV, row0);
            tma_load_2d(do_s, &tma_do, &ld_full, colQ, row0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_s = idesc_f16(false, false, 128, 128);   // S, dP: both operands K-major (k = dh)
            const uint32_t id_t = idesc_f16(true, true, 128, DH);      // dV, dK: A = [q x key] tile read MN-major (k = query)
            const uint32_t id_q = idesc_f16(false, true, 128, DH);     // dQ: A = dS K-major (k = key), B = K MN-major
            mbar_wait(&ld_full, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
                umma_f16(tmem, desc_k(q_s + k * 32, PITCH), desc_k(k_s + k * 32, PITCH), id_s, k > 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
                umma_f16(tmem + 128, desc_k(do_s + k * 32, PITCH), desc_k(v_s + k * 32, PITCH), id_s, k > 0);
            umma_commit(&sdp_full);
            mbar_wait(&pds_full, 0);                           // P and dS of all 128 rows are in smem; S / dP are dead
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)                        // dV[key, :] = sum_q P[q, key] dO[q, :]
                umma_f16(tmem, desc_mnmajor(pp_s + k * 2048, 16384), desc_mn(do_s + k * 16 * PITCH, PITCH), id_t, k > 0);
#pragma unroll
            for (int k = 0; k < 8; ++k)                        // dK[key, :] = sum_q dS[q, key] Q[q, :]
                umma_f16(tmem + DH, desc_mnmajor(ds_s + k * 2048, 16384), desc_mn(q_s + k * 16 * PITCH, PITCH), id_t, k > 0);
#pragma unroll
            for (int k = 0; k < 8; ++k)                        // dQ[q, :] = sum_key dS[q, key] K[key, :]
                umma_f16(tmem + 2 * DH, desc_k(ds_s + (k >> 2) * 16384 + (k & 3) * 32, 128), desc_mn(k_s + k * 16 * PITCH, PITCH), id_q, k > 0);
            umma_commit(&acc_full);
        }
    } else {
        const int qd = warp & 3, half = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const int cb = half * 64;                              // this warp's key columns
        const int64_t row = (int64_t)row0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
        int lo = 0, hi = 0;
        float lse2 = 0.f, dl = 0.f;
        if (r < nrows) {
            const int2 rb = p.row_bounds[row];
            lo = rb.x - row0, hi = min(rb.y - row0, 128);
            lse2 = p.lse[(int64_t)h * p.n_rows + row] * LOG2E;
            // delta = sum_c dO[row, c] O[row, c] over this head's columns
            const uint4* po = reinterpret_cast<const uint4*>((const bf16*)p.out + row * p.d + colQ);
            const uint4* pg = reinterpret_cast<const uint4*>((const bf16*)p.dout + row * p.d + colQ);
#pragma unroll
            for (int i = 0; i < DH / 8; ++i) {
                const uint4 a = po[i], b = pg[i];
                const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
                const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(bh[e]);
                    dl = fmaf(x.x, y.x, fmaf(x.y, y.y, dl));
                }
            }
        }
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        const uint32_t rk = drop_row_key(dr, att_row_id_tc(h, row, p.n_rows));
        const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : 0x7fffffff);
        const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
        mbar_wait(&sdp_full, 0);
        tc_fence_after();
        const float sl2 = p.scale_log2;
#pragma unroll 1
        for (int c = cb; c < cb + 64; c += 32) {
            if (c + 32 <= wlo || c >= whi) {                   // warp-uniform: nothing of these keys for this warp's rows
                st_zero32(ds_s, r, c);
                st_zero32(pp_s, r, c);
                continue;
            }
            uint32_t rs[32], rp[32];
            tmem_ld32(t_lane + c, rs);
            tmem_ld32(t_lane + 128 + c, rp);
            const uint32_t vm = range_mask32(lo, hi, c);
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 8) {
                float pv[8], dsv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float pr = ex2_approx(fmaf(__uint_as_float(rs[i0 + i]), sl2, -lse2));
                    pr = (vm & (1u << (i0 + i))) ? pr : 0.f;          // branch-free masking (see the forward)
                    float dp = __uint_as_float(rp[i0 + i]);
                    if (DROP) {
                        const float mk = drop_elem(dr, rk, (uint32_t)(row0 + c + i0 + i));
                        dp *= mk;
                        dsv[i] = pr * (dp - dl);
                        pr *= mk;
                    } else {
                        dsv[i] = pr * (dp - dl);
                    }
                    pv[i] = pr;
                }
                st_p8(ds_s, r, c + i0, dsv);
                st_p8(pp_s, r, c + i0, pv);
            }
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(&pds_full);
        mbar_wait(&acc_full, 0);
        tc_fence_after();
        bf16* gp = (bf16*)p.dqkv + row * (int64_t)(3 * p.d);
#pragma unroll
        for (int a = 0; a < 3; ++a) {                           // 0: dV, 1: dK, 2: dQ  (TMEM lane = key row / query row)
            const int col = a == 0 ? colV : (a == 1 ? colK : colQ);
            const float mul = a == 0 ? 1.f : p.scale;
            constexpr int OC = DH / 2;                         // each warp of the pair stores half of the head dimension
            if (OC == 32) {
                uint32_t rr[32];
                tmem_ld32(t_lane + a * DH + half * OC, rr);
                if (r < nrows) st_row_bf16(gp + col + half * OC, rr, 32, mul);
            } else {
                uint32_t rr[16];
                tmem_ld16(t_lane + a * DH + half * OC, rr);
                if (r < nrows) st_row_bf16(gp + col + half * OC, rr, 16, mul);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// Greedy graph-aligned tiling: consecutive graphs are packed while their token rows fit into 128.  One block; the
// graph offsets are staged through shared memory and thread 0 walks them (a few microseconds for thousands of graphs,
// on the plan-building branch stream).  tiles[t] = (first row, rows), unused slots (t >= count) = (0, 0).
__global__ void __launch_bounds__(256)
k_mha_local_tiles(const int32_t* __restrict__ tok_off, int B, int max_tiles, int2* __restrict__ tiles, int32_t* __restrict__ count) {
    constexpr int CH = 2048;
    __shared__ int32_t off[CH + 1];
    __shared__ int s_start, s_n, s_bad;
    for (int i = threadIdx.x; i < max_tiles; i += blockDim.x) tiles[i] = make_int2(0, 0);
    if (threadIdx.x == 0) s_start = tok_off[0], s_n = 0, s_bad = 0;
    __syncthreads();
    for (int g0 = 0; g0 < B; g0 += CH) {
        const int n = min(CH, B - g0);
        for (int i = threadIdx.x; i <= n; i += blockDim.x) off[i] = tok_off[g0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int start = s_start, cnt = s_n;
            for (int i = 0; i < n; ++i) {                      // graph g0 + i owns rows [off[i], off[i + 1])
                if (off[i + 1] - off[i] > 128) s_bad = 1;      // a graph that does not fit one tile: host bound violated
                if (off[i + 1] - start > 128 && off[i] > start) {
                    if (cnt < max_tiles) tiles[cnt] = make_int2(start, min(off[i] - start, 128));
                    ++cnt;
                    start = off[i];
                }
            }
            s_start = start, s_n = cnt;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int end = tok_off[B];
        int cnt = s_n;
        if (end > s_start) {
            if (cnt < max_tiles) tiles[cnt] = make_int2(s_start, min(end - s_start, 128));
            ++cnt;
        }
        *count = (s_bad || cnt > max_tiles) ? -1 : cnt;        // < 0: gt_mha_local_fwd poisons its output with NaN
    }
}

static unsigned long long* g_loc_trace = nullptr;
static unsigned long long* loc_trace_buf() {
    static const bool on = getenv("GT_LOC_TRACE") != nullptr;
    if (on && !g_loc_trace) cudaMalloc(&g_loc_trace, 16 * sizeof(unsigned long long));
    return on ? g_loc_trace : nullptr;
}

template <int DH, bool DROP>
static cudaError_t launch_loc_fwd2(const CUtensorMap& map, const LocParams& p, int max_tiles, cudaStream_t st) {
    const size_t smem = (size_t)128 * DH * 2 + 32768 + 1024;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_mha_loc_fwd<DH, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    k_mha_loc_fwd<DH, DROP><<<dim3((unsigned)max_tiles, (unsigned)p.nhead), LOC_THREADS, smem, st>>>(map, p);
    return cudaGetLastError();
}
template <int DH>
static cudaError_t launch_loc_fwd(const CUtensorMap& map, const LocParams& p, int max_tiles, cudaStream_t st) {
    return p.drop_p > 0.f ? launch_loc_fwd2<DH, true>(map, p, max_tiles, st) : launch_loc_fwd2<DH, false>(map, p, max_tiles, st);
}
template <int DH, bool DROP>
static cudaError_t launch_loc_bwd2(const CUtensorMap& mq, const CUtensorMap& md, const LocParams& p, int max_tiles, cudaStream_t st) {
    const size_t smem = (size_t)4 * 128 * DH * 2 + 65536 + 1024;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_mha_loc_bwd<DH, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    k_mha_loc_bwd<DH, DROP><<<dim3((unsigned)max_tiles, (unsigned)p.nhead), LOC_THREADS, smem, st>>>(mq, md, p);
    return cudaGetLastError();
}
template <int DH>
static cudaError_t launch_loc_bwd(const CUtensorMap& mq, const CUtensorMap& md, const LocParams& p, int max_tiles, cudaStream_t st) {
    return p.drop_p > 0.f ? launch_loc_bwd2<DH, true>(mq, md, p, max_tiles, st) : launch_loc_bwd2<DH, false>(mq, md, p, max_tiles, st);
}

}  // namespace tc
}  // namespace gt

using namespace gt;
using namespace gt::tc;

static int loc_check(const char* fn, int dt, const void* a, const void* b, int64_t n_rows, int32_t nhead, int32_t dh, int64_t max_tiles) {
    GT_CHECK_ARG(dt == GT_BF16, "%s: the tile-local attention takes bf16 activations", fn);
    GT_CHECK_ARG(dh == 32 || dh == 64, "%s: head dim %d not in {32, 64}", fn, dh);
    GT_CHECK_ARG(n_rows > 0 && n_rows < (1ll << 31) && nhead > 0 && max_tiles > 0 && max_tiles < (1 << 30), "%s: bad shape", fn);
    GT_CHECK_ARG(!(((uintptr_t)a | (uintptr_t)b) & 15), "%s: operands need 16-byte alignment", fn);
    return 0;
}

extern "C" int gt_mha_local_tiles(const int32_t* tok_off, int64_t B, int64_t max_tiles, int32_t* tiles, int32_t* count, void* stream) {
    GT_CHECK_ARG(B > 0 && max_tiles > 0 && max_tiles < (1 << 30), "gt_mha_local_tiles: bad shape");
    k_mha_local_tiles<<<1, 256, 0, (cudaStream_t)stream>>>(tok_off, (int)B, (int)max_tiles, (int2*)tiles, count);
    GT_LAUNCH_CHECK("gt_mha_local_tiles");
    return 0;
}

extern "C" int gt_mha_local_fwd(int dt, const void* qkv, const int32_t* row_bounds, const int32_t* tiles, const int32_t* count,
                                int64_t max_tiles, int64_t n_rows, int32_t nhead, int32_t dh, float scale, void* out, float* lse, float drop_p,
                                const uint64_t* rng_state, uint64_t salt, void* stream) {
    if (int r = loc_check("gt_mha_local_fwd", dt, qkv, out, n_rows, nhead, dh, max_tiles)) return r;
    const int d = nhead * dh;
    CUtensorMap map;
    if (!make_map(&map, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 128, dh * 2)) {
        set_error("gt_mha_local_fwd: cuTensorMapEncodeTiled failed or unavailable");
        return -1;
    }
    LocParams p{};
    p.row_bounds = (const int2*)row_bounds; p.tiles = (const int2*)tiles; p.count = count; p.out = out; p.lse = lse; p.rng = rng_state; p.salt = salt;
    p.n_rows = n_rows; p.nhead = nhead; p.d = d; p.scale = scale; p.scale_log2 = scale * LOG2E; p.drop_p = drop_p;
    p.trace = loc_trace_buf();
    const cudaError_t e = dh == 64 ? launch_loc_fwd<64>(map, p, (int)max_tiles, (cudaStream_t)stream)
                                   : launch_loc_fwd<32>(map, p, (int)max_tiles, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gt_mha_local_fwd");
    return 0;
}

extern "C" int gt_mha_local_bwd(int dt, const void* qkv, const void* out, const void* dout, const float* lse,
                                const int32_t* row_bounds, const int32_t* tiles, int64_t max_tiles, int64_t n_rows, int32_t nhead,
                                int32_t dh, float scale, void* dqkv, float drop_p, const uint64_t* rng_state, uint64_t salt,
                                void* stream) {
    if (int r = loc_check("gt_mha_local_bwd", dt, qkv, dqkv, n_rows, nhead, dh, max_tiles)) return r;
    GT_CHECK_ARG(!(((uintptr_t)out | (uintptr_t)dout) & 15), "gt_mha_local_bwd: operands need 16-byte alignment");
    const int d = nhead * dh;
    CUtensorMap mq, md;
    if (!make_map(&mq, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 128, dh * 2) ||
        !make_map(&md, dout, (uint64_t)d, (uint64_t)n_rows, (uint64_t)d, (uint32_t)dh, 128, dh * 2)) {
        set_error("gt_mha_local_bwd: cuTensorMapEncodeTiled failed or unavailable (CUresult %d, qkv %p dout %p n_rows %lld d %d dh %d)",
                  g_last_map_result(), qkv, dout, (long long)n_rows, d, dh);
        return -1;
    }
    LocParams p{};
    p.row_bounds = (const int2*)row_bounds; p.tiles = (const int2*)tiles; p.out = const_cast<void*>(out); p.dout = dout; p.lse = const_cast<float*>(lse);
    p.dqkv = dqkv; p.rng = rng_state; p.salt = salt;
    p.n_rows = n_rows; p.nhead = nhead; p.d = d; p.scale = scale; p.scale_log2 = scale * LOG2E; p.drop_p = drop_p;
    const cudaError_t e = dh == 64 ? launch_loc_bwd<64>(mq, md, p, (int)max_tiles, (cudaStream_t)stream)
                                   : launch_loc_bwd<32>(mq, md, p, (int)max_tiles, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "gt_mha_local_bwd");
    return 0;
}

// profiling hook (not part of the product ABI): phase stamps (ns) of CTA (0, 0) of the last traced gt_mha_local_fwd
extern "C" int gtdbg_loc_trace_read(unsigned long long* out16) {
    if (!gt::tc::g_loc_trace) return -1;
    return (int)cudaMemcpy(out16, gt::tc::g_loc_trace, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
}
