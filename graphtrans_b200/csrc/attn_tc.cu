// Stage 2 on tensor cores: masked multi-head self-attention over PACKED tokens with tcgen05 / TMEM / TMA
// (bf16 operands, fp32 accumulate).  Replaces F.multi_head_attention_forward's bmm / masked_fill / softmax / bmm
// (reference modules/transformer_encoder.py:28-32,59).
//
// Work item = (128 consecutive packed token rows, head).  Token rows are sorted by graph and the keys of a row
// are exactly the rows of its graph, so the keys a 128-row query tile can see form ONE contiguous row range
// [tok_off[g_first], tok_off[g_last + 1]): the kernel streams that range in 128-key tiles and masks with the
// per-row bounds [lo, hi) (block-diagonal attention; the reference's -inf key-padding mask is implicit).
//
//   warp 0     : TMA producer - Q tile once, K / V tiles through small rings (cp.async.bulk.tensor.2d, swizzled)
//   warp 1     : MMA issuer   - S = Q K^T (128x128xdh) into TMEM, then O_j = P V_j (128 x dh x 128) into TMEM
//   warps 2..5 : softmax      - one thread per query row: tcgen05.ld S -> scale/mask -> online max/sum (exp2) ->
//                               dropout -> P (bf16) into swizzled smem as the A operand of the PV MMA; the running
//                               output lives in registers: o = o * corr + O_j (O_j read back with tcgen05.ld)
// The backward follows FlashAttention-2's recompute scheme with the same roles (see k_mha_tc_bwd_*).
#include "attn_common.cuh"

namespace gt {

namespace tc {

struct AttnParams {
    const int32_t* tok_graph;
    const int32_t* tok_off;
    const int2* row_bounds;    // optional (gt_mha_meta): [lo, hi) key rows of every token row
    const int2* tile_bounds;   // optional: (first interacting row, number of 128-row tiles) of every 128-row tile
    void* out;
    float* lse;
    const uint64_t* rng;
    uint64_t salt;
    int64_t n_rows;
    int B, nhead, d;
    float scale_log2, drop_p;
};

// Forward, two-phase ("max first") streaming softmax.  Phase 1 streams the key tiles once and only takes the row
// maxima of S = Q K^T; phase 2 streams them again, forms P = exp2(S * scale*log2e - m) with the FINAL maximum and
// accumulates O += P V directly in TMEM.  With the maximum known up front there is no running rescale: the output
// accumulator never leaves TMEM until the end, the softmax threads never wait for the P V MMA, and S is released as
// soon as it sits in registers so that the next Q K^T overlaps the exponentials.  The extra Q K^T costs tensor-core
// time that is idle anyway (this kernel is bound by the softmax threads, not by the MMAs).
template <int DH, bool DROP>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 2)
k_mha_tc_fwd(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
    constexpr int PITCH = DH * 2;                 // bytes per operand row
    constexpr int TILE = 128 * PITCH;             // Q / K / V tile bytes
    constexpr int KST = 2, VST = (DH == 64 ? 1 : 2);
    constexpr uint32_t TMEM_COLS = 256;           // S: [0,128), O: [128, 128+DH)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full, k_full[KST], k_empty[KST], v_full[VST], v_empty[VST], s_full, s_free, p_full, p_empty, o_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ int kv_lo_s, nkv_s;
    __shared__ float xchg[2][128];                // row max / row sum exchange between the two column halves

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base, k_s = q_s + TILE, v_s = k_s + KST * TILE, p_s = v_s + VST * TILE;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        for (int s = 0; s < KST; ++s) mbar_init(&k_full[s], 1), mbar_init(&k_empty[s], 1);
        for (int s = 0; s < VST; ++s) mbar_init(&v_full[s], 1), mbar_init(&v_empty[s], 1);
        mbar_init(&s_full, 1);
        mbar_init(&s_free, 256);
        mbar_init(&p_full, 256);
        mbar_init(&p_empty, 1);
        mbar_init(&o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        // contiguous key-row range visible from this query tile
        int lo = 0, n = 0;
        if (p.tile_bounds) {
            const int2 tb = p.tile_bounds[blockIdx.x];
            lo = tb.x, n = (tb.y + BKV - 1) / BKV;
        } else if (q0 < p.tok_off[p.B]) {
            const int n_tok = p.tok_off[p.B];
            const int g0 = p.tok_graph[q0];
            const int last = min(q0 + BQ - 1, n_tok - 1);
            const int g1 = p.tok_graph[last];
            lo = p.tok_off[g0];
            n = (p.tok_off[g1 + 1] - lo + BKV - 1) / BKV;
        }
        kv_lo_s = lo;
        nkv_s = n;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int kv_lo = kv_lo_s, nkv = nkv_s;

    if (warp == 0) {
        if (lane == 0 && nkv > 0) {  // ===== TMA producer: K tiles twice (phase 1, phase 2), V tiles in phase 2 =====
            mbar_expect_tx(&q_full, TILE);
            tma_load_2d(q_s, &tma_qkv, &q_full, h * DH, q0);
            for (int it = 0; it < 2 * nkv; ++it) {
                const int j = it < nkv ? it : it - nkv;
                const int ks = it % KST;
                mbar_wait(&k_empty[ks], ((uint32_t)(it / KST) & 1u) ^ 1u);
                mbar_expect_tx(&k_full[ks], TILE);
                tma_load_2d(k_s + ks * TILE, &tma_qkv, &k_full[ks], p.d + h * DH, kv_lo + j * BKV);
                if (it >= nkv) {
                    const int vs = j % VST;
                    mbar_wait(&v_empty[vs], ((uint32_t)(j / VST) & 1u) ^ 1u);
                    mbar_expect_tx(&v_full[vs], TILE);
                    tma_load_2d(v_s + vs * TILE, &tma_qkv, &v_full[vs], 2 * p.d + h * DH, kv_lo + j * BKV);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkv > 0) {  // ===== MMA issuer =====
            const uint32_t id_s = idesc_f16(false, false, BQ, BKV);   // S = Q K^T   : A, B K-major (k = dh)
            const uint32_t id_o = idesc_f16(false, true, BQ, DH);     // O += P V    : A K-major (k = key), B MN-major
            mbar_wait(&q_full, 0);
            for (int it = 0; it < 2 * nkv; ++it) {
                const int ks = it % KST;
                mbar_wait(&k_full[ks], (uint32_t)(it / KST) & 1u);
                if (it > 0) mbar_wait(&s_free, (uint32_t)(it - 1) & 1u);   // previous S sits in the softmax registers
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_f16(tmem, desc_k(q_s + k * 32, PITCH), desc_k(k_s + ks * TILE + k * 32, PITCH), id_s, k > 0);
                umma_commit(&k_empty[ks]);
                umma_commit(&s_full);
                // P V of the previous phase-2 tile: issued after the next Q K^T so the softmax threads get S first
                const int jp = it - nkv - 1;
                if (jp >= 0) {
                    const int vs = jp % VST;
                    mbar_wait(&p_full, (uint32_t)jp & 1u);
                    mbar_wait(&v_full[vs], (uint32_t)(jp / VST) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        umma_f16(tmem + 128, desc_k(p_s + (k >> 2) * 16384 + (k & 3) * 32, 128),
                                 desc_mn(v_s + vs * TILE + k * 16 * PITCH, PITCH), id_o, (jp > 0 || k > 0));
                    umma_commit(&v_empty[vs]);
                    umma_commit(&p_empty);
                }
            }
            {   // P V of the last tile
                const int jp = nkv - 1, vs = jp % VST;
                mbar_wait(&p_full, (uint32_t)jp & 1u);
                mbar_wait(&v_full[vs], (uint32_t)(jp / VST) & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    umma_f16(tmem + 128, desc_k(p_s + (k >> 2) * 16384 + (k & 3) * 32, 128),
                             desc_mn(v_s + vs * TILE + k * 16 * PITCH, PITCH), id_o, (jp > 0 || k > 0));
                umma_commit(&o_full);
            }
        }
    } else {  // ===== softmax warps 2..9: thread = (query row, half of the 128 key columns of every tile) =====
        // Two warps share a TMEM lane group (= 32 query rows) and split the key columns of each tile; with the
        // two-phase scheme they only have to meet twice: once to combine the row maxima, once for the row sums.
        const int qd = warp & 3, half = (warp - 2) >> 2;
        const int r = qd * 32 + lane;              // row inside the tile = TMEM lane
        const int64_t row = (int64_t)q0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
        const int cb = half * 64;                  // this warp's columns: [cb, cb + 64) of every key tile
        int lo = 0, hi = 0;
        if (row < p.n_rows) {
            if (p.row_bounds) {
                const int2 rb = p.row_bounds[row];
                lo = rb.x, hi = rb.y;
            } else {
                const int g = p.tok_graph[row];
                if (g >= 0) lo = p.tok_off[g], hi = p.tok_off[g + 1];
            }
        }
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        const uint32_t rk = drop_row_key(dr, att_row_id_tc(h, row, p.n_rows));
        // key range any row of this WARP can see: 32-key groups outside it are skipped (graphs are short compared
        // with the 128-key tile in the molecule workloads, so most of a tile is fully masked for a given warp)
        const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : 0x7fffffff);
        const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
        // ---- phase 1: row maxima
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int j = 0; j < nkv; ++j) {
            const int kv0 = kv_lo + j * BKV;
            mbar_wait(&s_full, (uint32_t)j & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = cb; c < cb + 64; c += 32) {
                if (kv0 + c + 32 <= wlo || kv0 + c >= whi) continue;   // warp-uniform
                uint32_t rr[32];
                tmem_ld32(t_lane + c, rr);
                const uint32_t vm = range_mask32(lo, hi, kv0 + c);     // branch-free: predicated maxima, 4 chains
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (vm & (1u << i)) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(rr[i]));
            }
            tc_fence_before();
            mbar_arrive(&s_free);
        }
        float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        xchg[half][r] = m;
        named_bar_sync(1, 256);
        m = fmaxf(xchg[0][r], xchg[1][r]);
        const float m_use = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        named_bar_sync(1, 256);                     // xchg is reused for the row sums
        // ---- phase 2: probabilities with the final maximum; O accumulates in TMEM
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        const float sl2 = p.scale_log2;
        for (int j = 0; j < nkv; ++j) {
            const int kv0 = kv_lo + j * BKV;
            mbar_wait(&s_full, (uint32_t)(nkv + j) & 1u);
            tc_fence_after();
            if (j > 0) mbar_wait(&p_empty, (uint32_t)(j - 1) & 1u);   // P V of the previous tile has read the P buffer
#pragma unroll 1
            for (int c = cb; c < cb + 64; c += 32) {
                const bool last = c + 32 >= cb + 64;
                if (kv0 + c + 32 <= wlo || kv0 + c >= whi) {           // fully masked for this warp: P = 0
#pragma unroll
                    for (int i = 0; i < 32; i += 8)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(p_s + p_chunk_off(r, c + i)), "r"(0u) : "memory");
                    if (last) {
                        tc_fence_before();
                        mbar_arrive(&s_free);
                    }
                    continue;
                }
                uint32_t rr[32];
                tmem_ld32(t_lane + c, rr);
                if (last) {                                            // this thread's share of S sits in registers
                    tc_fence_before();
                    mbar_arrive(&s_free);
                }
                const uint32_t vm = range_mask32(lo, hi, kv0 + c);
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    float pv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // every column's exponential is formed; masked columns (finite scores of other graphs, possibly
                        // overflowing to +inf) are discarded by the select
                        float e = ex2_approx(fmaf(__uint_as_float(rr[i0 + i]), sl2, -m_use));
                        e = (vm & (1u << (i0 + i))) ? e : 0.f;
                        ls[i & 3] += e;
                        if (DROP) e *= drop_elem(dr, rk, (uint32_t)(kv0 + c + i0 + i));
                        pv[i] = e;
                    }
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_s + p_chunk_off(r, c + i0)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                }
            }
            fence_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&p_full);
        }
        float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
        xchg[half][r] = l;
        named_bar_sync(1, 256);
        l = xchg[0][r] + xchg[1][r];
        // ---- output: each of the two warps of a lane group normalises and stores half of the head dimension
        constexpr int OC = DH / 2;
        float o[OC];
        if (nkv > 0) {
            mbar_wait(&o_full, 0);
            tc_fence_after();
            if (OC == 32) {
                uint32_t rr[32];
                tmem_ld32(t_lane + 128 + half * OC, rr);
#pragma unroll
                for (int i = 0; i < OC; ++i) o[i] = __uint_as_float(rr[i]);
            } else {
                uint32_t rr[16];
                tmem_ld16(t_lane + 128 + half * OC, rr);
#pragma unroll
                for (int i = 0; i < OC; ++i) o[i] = __uint_as_float(rr[i]);
            }
            tc_fence_before();
        } else {
#pragma unroll
            for (int i = 0; i < OC; ++i) o[i] = 0.f;
        }
        if (row < p.n_rows) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            bf16* op = (bf16*)p.out + row * p.d + h * DH + half * OC;
#pragma unroll
            for (int i = 0; i < OC; i += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv), h1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv), h3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(op + i) = pk;
            }
            if (half == 0) p.lse[(int64_t)h * p.n_rows + row] = l > 0.f ? (m_use + log2f(l)) * LN2 : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int DH, bool DROP>
static cudaError_t launch_fwd2(const CUtensorMap& map, const AttnParams& p, cudaStream_t st) {
    constexpr int TILE = 128 * DH * 2;
    constexpr int KST = 2, VST = (DH == 64 ? 1 : 2);
    const size_t smem = (size_t)(1 + KST + VST) * TILE + 32768 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_mha_tc_fwd<DH, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((unsigned)((p.n_rows + BQ - 1) / BQ), (unsigned)p.nhead);
    k_mha_tc_fwd<DH, DROP><<<grid, ATT_FWD_THREADS, smem, st>>>(map, p);
    return cudaGetLastError();
}
template <int DH>
static cudaError_t launch_fwd(const CUtensorMap& map, const AttnParams& p, cudaStream_t st) {
    return p.drop_p > 0.f ? launch_fwd2<DH, true>(map, p, st) : launch_fwd2<DH, false>(map, p, st);
}


// per-step attention metadata (once per batch instead of three dependent loads in every CTA of every layer):
// row_bounds[row] = [lo, hi) token rows of the row's graph (0,0 for tail rows); tile_bounds[t] = (first row, number
// of rows) of the contiguous row range that interacts with rows [128 t, 128 t + 128): the forward streams it in
// ceil(n / 128) key tiles, the backward in ceil(n / 64) tiles
__global__ void k_mha_meta(const int32_t* __restrict__ tok_graph, const int32_t* __restrict__ tok_off, int64_t n_rows, int B,
                           int2* __restrict__ row_bounds, int2* __restrict__ tile_bounds) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int g = tok_graph[row];
    int2 rb = make_int2(0, 0);
    if (g >= 0) rb = make_int2(tok_off[g], tok_off[g + 1]);
    row_bounds[row] = rb;
    if ((row & 127) == 0) {
        const int n_tok = tok_off[B];
        int2 tb = make_int2(0, 0);
        if (row < n_tok) {
            const int g1 = tok_graph[min((int64_t)row + 127, (int64_t)n_tok - 1)];
            tb.x = rb.x;
            tb.y = tok_off[g1 + 1] - rb.x;
        }
        tile_bounds[row >> 7] = tb;
    }
}

// ------------------------------------------------------------------------------------------ backward
// delta[h, row] = sum_c dO[row, h*dh + c] * O[row, h*dh + c].  A thread owns one 16-byte vector (8 channels) of a row,
// the dh / 8 lanes of a (row, head) meet with xor shuffles: the two [n_rows, d] matrices are streamed once with fully
// coalesced 16-byte loads, two vectors per thread in flight (dh in {32, 64}: 4 or 8 lanes per head).
__global__ void __launch_bounds__(256)
k_mha_delta(const bf16* __restrict__ out, const bf16* __restrict__ dout, int64_t n_rows, int nhead, int dh,
            float* __restrict__ delta) {
    const int d = nhead * dh, gl = dh >> 3;          // lanes per (row, head)
    const int64_t nvec = n_rows * (int64_t)(d >> 3);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v0 - threadIdx.x % 32 < nvec; v0 += 2 * stride) {
        float s[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t v = v0 + u * stride;
            s[u] = 0.f;
            if (v < nvec) {
                const uint4 a = *reinterpret_cast<const uint4*>(out + v * 8);
                const uint4 b = *reinterpret_cast<const uint4*>(dout + v * 8);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    s[u] = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(bw[i] << 16), s[u]);
                    s[u] = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(bw[i] & 0xffff0000u), s[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            for (int o = gl >> 1; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
            const int64_t v = v0 + u * stride;
            if (v < nvec && (v & (gl - 1)) == 0) {
                const int64_t e = v * 8, row = e / d;
                const int h = (int)(e - row * d) / dh;
                delta[(int64_t)h * n_rows + row] = s[u];
            }
        }
    }
}

struct AttnBwdParams {
    const int32_t* tok_graph;
    const int32_t* tok_off;
    const int2* row_bounds;
    const int2* tile_bounds;
    const float* lse;
    const float* delta;
    void* dqkv;
    const uint64_t* rng;
    uint64_t salt;
    int64_t n_rows;
    int B, nhead, d;
    float scale, scale_log2, drop_p;
};

// FlashAttention-2 style recompute backward in two kernels that share one body.  The CTA OWNS 128 rows (their
// accumulators live on the 128 TMEM lanes, thread = owned row) and STREAMS the interacting rows of the other kind in
// tiles of 64:
//   DKV = false : owns 128 QUERY rows (Q, dO resident), streams key tiles (K, V):
//                   S  = Q K^T,  dP  = dO V^T   [128 q x 64 keys];   dQ += dS K
//   DKV = true  : owns 128 KEY rows (K, V resident), streams query tiles (Q, dO), transposed formulation:
//                   S^T = K Q^T, dP^T = V dO^T  [128 keys x 64 q];   dK += dS^T Q,  dV += P^T dO
// with P = exp2(S*scale*log2e - lse*log2e), dS = P o (dP*keep - delta).  In both modes the per-element math tile
// [128 owned x 64 streamed] is written as one 128B-swizzled K-major smem operand (k = streamed index) and the streamed
// tile is the MN-major B operand of the accumulation MMA, so the two modes differ only in which index is the query.
// TMEM: S 64 + dP 64 + 2 x 64 accumulator columns = 256, smem 96 KB -> two CTAs per SM.
struct ColMeta {     // DKV mode: per streamed QUERY column
    float lse2[64];
    float delta[64];
    uint32_t rk[64];
};

template <int DH, bool DKV, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 2)
k_mha_tc_bwd(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
             const __grid_constant__ CUtensorMap tma_qkv64, const __grid_constant__ CUtensorMap tma_do64, const AttnBwdParams p) {
    constexpr int PITCH = DH * 2;
    constexpr int TILE = 128 * PITCH;             // owned tile bytes
    constexpr int STILE = 64 * PITCH;             // streamed tile bytes
    constexpr int STG = 2;
    constexpr uint32_t TMEM_COLS = 256;           // S [0,64) | dP [64,128) | acc0 [128,128+DH) | acc1 [192,192+DH)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t own_full, st_full[STG], st_empty[STG], sdp_full, sdp_free, pds_full, acc_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int st_lo_s, nst_s;
    __shared__ ColMeta cmeta[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, h = blockIdx.y;      // first owned row (query rows or key rows)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t own_x = base, own_y = own_x + TILE;    // dQ: Q, dO   | dKV: K, V
    const uint32_t st_s = own_y + TILE;                   // ring: stage s -> X at st_s + 2*s*STILE, Y right after
    const uint32_t ds_s = st_s + 2 * STG * STILE;         // dS (or dS^T) [128 x 64] bf16, one swizzled 16 KB block
    const uint32_t pp_s = ds_s + 16384;                   // P^T (DKV only)

    if (threadIdx.x == 0) {
        mbar_init(&own_full, 1);
        for (int s = 0; s < STG; ++s) mbar_init(&st_full[s], 1), mbar_init(&st_empty[s], 1);
        mbar_init(&sdp_full, 1);
        mbar_init(&sdp_free, 256);
        mbar_init(&pds_full, 256);
        mbar_init(&acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        // rows of the OTHER kind that interact with the owned tile: the rows of the graphs it touches (symmetric)
        int lo = 0, n = 0;
        if (p.tile_bounds) {
            const int2 tb = p.tile_bounds[blockIdx.x];
            lo = tb.x, n = tb.y;
        } else if (t0 < p.tok_off[p.B]) {
            const int n_tok = p.tok_off[p.B];
            const int g0 = p.tok_graph[t0];
            const int g1 = p.tok_graph[min(t0 + 127, n_tok - 1)];
            lo = p.tok_off[g0];
            n = p.tok_off[g1 + 1] - lo;
        }
        st_lo_s = lo;
        nst_s = (n + 63) / 64;      // 64-row streamed tiles covering the n interacting rows
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int st_lo = st_lo_s, nst = nst_s;
    const int colQ = h * DH, colK = p.d + h * DH, colV = 2 * p.d + h * DH;

    if (warp == 0) {
        if (lane == 0 && nst > 0) {  // ===== TMA producer =====
            mbar_expect_tx(&own_full, 2 * TILE);
            if (!DKV) {
                tma_load_2d(own_x, &tma_qkv, &own_full, colQ, t0);
                tma_load_2d(own_y, &tma_do, &own_full, colQ, t0);
            } else {
                tma_load_2d(own_x, &tma_qkv, &own_full, colK, t0);
                tma_load_2d(own_y, &tma_qkv, &own_full, colV, t0);
            }
            for (int t = 0; t < nst; ++t) {
                const int s = t % STG;
                mbar_wait(&st_empty[s], ((uint32_t)(t / STG) & 1u) ^ 1u);
                mbar_expect_tx(&st_full[s], 2 * STILE);
                const uint32_t x = st_s + 2 * s * STILE, y = x + STILE;
                const int r0 = st_lo + t * 64;
                if (!DKV) {
                    tma_load_2d(x, &tma_qkv64, &st_full[s], colK, r0);
                    tma_load_2d(y, &tma_qkv64, &st_full[s], colV, r0);
                } else {
                    tma_load_2d(x, &tma_qkv64, &st_full[s], colQ, r0);
                    tma_load_2d(y, &tma_do64, &st_full[s], colQ, r0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nst > 0) {  // ===== MMA issuer =====
            const uint32_t id_s = idesc_f16(false, false, 128, 64);     // [128 owned x 64 streamed], k = dh
            const uint32_t id_acc = idesc_f16(false, true, 128, DH);    // A = math tile K-major (k = streamed), B MN-major
            auto issue_sdp = [&](int t) {
                const int s = t % STG;
                const uint32_t x = st_s + 2 * s * STILE, y = x + STILE;
                mbar_wait(&st_full[s], (uint32_t)(t / STG) & 1u);
                if (t > 0) mbar_wait(&sdp_free, (uint32_t)(t - 1) & 1u);   // previous S / dP sit in registers
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_f16(tmem, desc_k(own_x + k * 32, PITCH), desc_k(x + k * 32, PITCH), id_s, k > 0);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_f16(tmem + 64, desc_k(own_y + k * 32, PITCH), desc_k(y + k * 32, PITCH), id_s, k > 0);
                umma_commit(&sdp_full);
            };
            mbar_wait(&own_full, 0);
            issue_sdp(0);
            for (int t = 0; t < nst; ++t) {
                if (t + 1 < nst) issue_sdp(t + 1);     // next S / dP run while the math warps work on tile t
                const int s = t % STG;
                const uint32_t x = st_s + 2 * s * STILE, y = x + STILE;
                mbar_wait(&pds_full, (uint32_t)t & 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)   // dQ += dS K   |   dK += dS^T Q      (k = streamed index)
                    umma_f16(tmem + 128, desc_k(ds_s + k * 32, 128), desc_mn(x + k * 16 * PITCH, PITCH), id_acc, (t > 0 || k > 0));
                if (DKV) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // dV += P^T dO
                        umma_f16(tmem + 192, desc_k(pp_s + k * 32, 128), desc_mn(y + k * 16 * PITCH, PITCH), id_acc, (t > 0 || k > 0));
                }
                umma_commit(&st_empty[s]);
                umma_commit(&acc_done);
            }
        }
    } else {  // ===== math warps 2..9: thread = (owned row, 32 of the 64 streamed columns of every tile) =====
        // Two warps share a TMEM lane group and split the streamed columns: per row the math is a chain of tcgen05.ld ->
        // MUFU -> FMUL with little ILP, so the second warp per row (8 instead of 4 math warps per CTA, 2 CTAs per SM)
        // fills issue slots that were idle; masking is branch-free (bit mask of the row's index range), dropout is a
        // template flag.
        const int qd = warp & 3, half = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const int c = half * 32;                  // this warp's columns of every streamed tile: [c, c + 32)
        const int64_t orow = (int64_t)t0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        // owned row: its graph's row range [lo, hi) is both "the keys a query sees" and "the queries a key is seen by"
        int lo = 0, hi = 0;
        if (orow < p.n_rows) {
            if (p.row_bounds) {
                const int2 rb = p.row_bounds[orow];
                lo = rb.x, hi = rb.y;
            } else {
                const int g = p.tok_graph[orow];
                if (g >= 0) lo = p.tok_off[g], hi = p.tok_off[g + 1];
            }
        }
        float lse2 = 0.f, dl = 0.f;       // dQ mode: per owned query row
        uint32_t rk = 0;
        if (!DKV && hi > lo) {
            lse2 = p.lse[(int64_t)h * p.n_rows + orow] * LOG2E;
            dl = p.delta[(int64_t)h * p.n_rows + orow];
            rk = drop_row_key(dr, att_row_id_tc(h, orow, p.n_rows));
        }
        const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : 0x7fffffff);
        const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
        const int tid = threadIdx.x - 64;     // 0..255 among the math warps
        const float sl2 = p.scale_log2;
        for (int t = 0; t < nst; ++t) {
            const int c0 = st_lo + t * 64;    // first streamed row of this tile
            if (DKV) {                         // per-column query metadata of this tile
                ColMeta& cm = cmeta[t & 1];
                if (tid < 64) {
                    const int64_t q = (int64_t)c0 + tid;
                    float a = 0.f, b = 0.f;
                    uint32_t k = 0;
                    if (q < p.n_rows) {
                        a = p.lse[(int64_t)h * p.n_rows + q] * LOG2E;
                        b = p.delta[(int64_t)h * p.n_rows + q];
                        k = drop_row_key(dr, att_row_id_tc(h, q, p.n_rows));
                    }
                    cm.lse2[tid] = a, cm.delta[tid] = b, cm.rk[tid] = k;
                }
                named_bar_sync(1, 256);
            }
            mbar_wait(&sdp_full, (uint32_t)t & 1u);
            tc_fence_after();
            if (t > 0) mbar_wait(&acc_done, (uint32_t)(t - 1) & 1u);   // previous dS / P tiles have been consumed
            const int s0 = c0 + c;            // streamed index of this warp's first column
            if (s0 + 32 <= wlo || s0 >= whi) {   // fully masked for this warp (warp-uniform): dS = P = 0
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const uint32_t off = p_chunk_off(r, c + i);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ds_s + off), "r"(0u) : "memory");
                    if (DKV) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(pp_s + off), "r"(0u) : "memory");
                }
                tc_fence_before();
                mbar_arrive(&sdp_free);
            } else {
                uint32_t rs[32], rp[32];
                tmem_ld32(t_lane + c, rs);
                tmem_ld32(t_lane + 64 + c, rp);
                tc_fence_before();
                mbar_arrive(&sdp_free);
                const uint32_t vm = range_mask32(lo, hi, s0);
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    float pv[8], dsv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float l2 = DKV ? cmeta[t & 1].lse2[c + i0 + i] : lse2;
                        const float de = DKV ? cmeta[t & 1].delta[c + i0 + i] : dl;
                        // every column's exponential is formed (an overflow on a masked column is discarded by the select)
                        float pr = ex2_approx(fmaf(__uint_as_float(rs[i0 + i]), sl2, -l2));
                        pr = (vm & (1u << (i0 + i))) ? pr : 0.f;
                        float dp = __uint_as_float(rp[i0 + i]);
                        if (DROP) {
                            // mask element (query, key): dQ mode query = owned row, key = streamed; dKV the reverse
                            const float mk = DKV ? drop_elem(dr, cmeta[t & 1].rk[c + i0 + i], (uint32_t)orow)
                                                 : drop_elem(dr, rk, (uint32_t)(s0 + i0 + i));
                            dp *= mk;
                            dsv[i] = pr * (dp - de);
                            pr *= mk;
                        } else {
                            dsv[i] = pr * (dp - de);
                        }
                        pv[i] = pr;
                    }
                    const uint32_t off = p_chunk_off(r, c + i0);
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(dsv[0], dsv[1]), h1 = __floats2bfloat162_rn(dsv[2], dsv[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(dsv[4], dsv[5]), h3 = __floats2bfloat162_rn(dsv[6], dsv[7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_s + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                    if (DKV) {
                        h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
                        h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pp_s + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                                     "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                                     "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                    }
                }
            }
            fence_async_smem();
            mbar_arrive(&pds_full);
        }
        // epilogue: the accumulators' TMEM lanes are the owned rows; the two warps of a lane group split the columns
        if (nst > 0) {
            mbar_wait(&acc_done, (uint32_t)(nst - 1) & 1u);
            tc_fence_after();
        }
        bf16* gp = (bf16*)p.dqkv + orow * (int64_t)(3 * p.d);
        constexpr int OC = DH / 2;
#pragma unroll
        for (int a = 0; a < (DKV ? 2 : 1); ++a) {
            const int col = (!DKV ? colQ : (a == 0 ? colK : colV)) + half * OC;
            const float mul = (DKV && a == 1) ? 1.f : p.scale;
            uint32_t rr[32];
            if (nst > 0) {
                if (OC == 32) {
                    tmem_ld32(t_lane + 128 + a * 64 + half * OC, rr);
                } else {
                    uint32_t r16[16];
                    tmem_ld16(t_lane + 128 + a * 64 + half * OC, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) rr[i] = r16[i];
                }
            }
            if (orow < p.n_rows) {
#pragma unroll
                for (int i = 0; i < OC; i += 8) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = nst > 0 ? __uint_as_float(rr[i + e]) * mul : 0.f;
                    uint4 pk;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(gp + col + i) = pk;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int DH, bool DKV, bool DROP>
static cudaError_t launch_bwd2(const CUtensorMap& mq, const CUtensorMap& md, const CUtensorMap& mq64, const CUtensorMap& md64,
                               const AttnBwdParams& p, cudaStream_t st) {
    constexpr int TILE = 128 * DH * 2, STILE = 64 * DH * 2;
    const size_t smem = (size_t)2 * TILE + (size_t)4 * STILE + 16384 * (DKV ? 2 : 1) + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_mha_tc_bwd<DH, DKV, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((unsigned)((p.n_rows + 127) / 128), (unsigned)p.nhead);
    k_mha_tc_bwd<DH, DKV, DROP><<<grid, ATT_THREADS, smem, st>>>(mq, md, mq64, md64, p);
    return cudaGetLastError();
}
template <int DH, bool DKV>
static cudaError_t launch_bwd(const CUtensorMap& mq, const CUtensorMap& md, const CUtensorMap& mq64, const CUtensorMap& md64,
                              const AttnBwdParams& p, cudaStream_t st) {
    return p.drop_p > 0.f ? launch_bwd2<DH, DKV, true>(mq, md, mq64, md64, p, st) : launch_bwd2<DH, DKV, false>(mq, md, mq64, md64, p, st);
}

}  // namespace tc

int mha_meta_launch(const int32_t* tok_graph, const int32_t* tok_off, int64_t n_rows, int64_t B, int32_t* row_bounds,
                    int32_t* tile_bounds, cudaStream_t st) {
    tc::k_mha_meta<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(tok_graph, tok_off, n_rows, (int)B, (int2*)row_bounds,
                                                                   (int2*)tile_bounds);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "gt_mha_meta");
}

int mha_tc_fwd_launch(int dt, const void* qkv, const int32_t* tok_graph, const int32_t* tok_off, const int32_t* key_start,
                      const int32_t* row_bounds, const int32_t* tile_bounds,
                      int64_t n_rows, int64_t B, int32_t nhead, int32_t dh, float scale, void* out, float* lse, float drop_p,
                      const uint64_t* rng, uint64_t salt, cudaStream_t st) {
    using namespace tc;
    if (dt != GT_BF16) { set_error("tcgen05 attention takes bf16 activations"); return -2; }
    if (dh != 32 && dh != 64) { set_error("tcgen05 attention is built for head dims 32 and 64"); return -2; }
    if (key_start) { set_error("dense left-padded layout (key_start) runs on the CUDA-core kernel"); return -2; }
    const int d = nhead * dh;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || n_rows >= (1ll << 31)) { set_error("alignment"); return -2; }
    CUtensorMap map;
    if (!make_map(&map, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 128, dh * 2)) {
        set_error("cuTensorMapEncodeTiled failed or unavailable");
        return -2;
    }
    AttnParams p;
    p.tok_graph = tok_graph; p.tok_off = tok_off; p.out = out; p.lse = lse; p.rng = rng; p.salt = salt;
    p.row_bounds = (const int2*)row_bounds; p.tile_bounds = (const int2*)tile_bounds;
    p.n_rows = n_rows; p.B = (int)B; p.nhead = nhead; p.d = d;
    p.scale_log2 = scale * LOG2E; p.drop_p = drop_p;
    const cudaError_t e = dh == 64 ? launch_fwd<64>(map, p, st) : launch_fwd<32>(map, p, st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_mha_fwd(tcgen05)");
    return 0;
}

int mha_tc_bwd_launch(int dt, const void* qkv, const void* out, const void* dout, const float* lse, const int32_t* tok_graph,
                      const int32_t* tok_off, const int32_t* key_start, const int32_t* row_bounds, const int32_t* tile_bounds,
                      int64_t n_rows, int64_t B, int32_t nhead, int32_t dh,
                      float scale, void* dqkv, float* delta, float drop_p, const uint64_t* rng, uint64_t salt, cudaStream_t st) {
    using namespace tc;
    if (dt != GT_BF16) { set_error("tcgen05 attention takes bf16 activations"); return -2; }
    if (dh != 32 && dh != 64) { set_error("tcgen05 attention is built for head dims 32 and 64"); return -2; }
    if (key_start) { set_error("dense left-padded layout (key_start) runs on the CUDA-core kernel"); return -2; }
    const int d = nhead * dh;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)dqkv & 15) || n_rows >= (1ll << 31)) { set_error("alignment"); return -2; }
    CUtensorMap mq, md, mq64, md64;   // owned tiles: 128-row boxes; streamed tiles: 64-row boxes
    if (!make_map(&mq, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 128, dh * 2) ||
        !make_map(&md, dout, (uint64_t)d, (uint64_t)n_rows, (uint64_t)d, (uint32_t)dh, 128, dh * 2) ||
        !make_map(&mq64, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 64, dh * 2) ||
        !make_map(&md64, dout, (uint64_t)d, (uint64_t)n_rows, (uint64_t)d, (uint32_t)dh, 64, dh * 2)) {
        set_error("cuTensorMapEncodeTiled failed or unavailable");
        return -2;
    }
    {
        const int64_t nvec = n_rows * (int64_t)(d >> 3);
        const int blocks = blocks_for((nvec + 1) / 2, 256, kNumSMs * 8);
        k_mha_delta<<<blocks, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, n_rows, nhead, dh, delta);
    }
    AttnBwdParams p;
    p.tok_graph = tok_graph; p.tok_off = tok_off; p.lse = lse; p.delta = delta; p.dqkv = dqkv; p.rng = rng; p.salt = salt;
    p.row_bounds = (const int2*)row_bounds; p.tile_bounds = (const int2*)tile_bounds;
    p.n_rows = n_rows; p.B = (int)B; p.nhead = nhead; p.d = d;
    p.scale = scale; p.scale_log2 = scale * LOG2E; p.drop_p = drop_p;
    cudaError_t e;
    if (dh == 64) {
        e = launch_bwd<64, false>(mq, md, mq64, md64, p, st);
        if (e == cudaSuccess) e = launch_bwd<64, true>(mq, md, mq64, md64, p, st);
    } else {
        e = launch_bwd<32, false>(mq, md, mq64, md64, p, st);
        if (e == cudaSuccess) e = launch_bwd<32, true>(mq, md, mq64, md64, p, st);
    }
    if (e != cudaSuccess) return cuda_fail(e, "gt_mha_bwd(tcgen05)");
    return 0;
}
}  // namespace gt
