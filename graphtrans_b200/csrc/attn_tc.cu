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
#include "tc_common.cuh"

namespace gt {

__device__ __forceinline__ uint64_t att_row_id_tc(int h, int64_t q, int64_t n_rows) {   // == attn_simt.cu att_row_id
    return (uint64_t)h * (uint64_t)n_rows + (uint64_t)q;
}

namespace tc {

constexpr int ATT_THREADS = 192;
constexpr int BQ = 128, BKV = 128;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// K-major operand tile with `pitch`-byte rows (pitch = 64: SWIZZLE_64B, 128: SWIZZLE_128B)
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t pitch) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * pitch) >> 4) << 32) | (1ull << 46) |
           ((pitch == 128 ? 2ull : 4ull) << 61);
}
// MN-major operand tile: k rows of `pitch` bytes holding pitch/2 contiguous m|n elements (one chunk wide)
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t pitch) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * pitch) >> 4) << 32) | (1ull << 46) |
           ((pitch == 128 ? 2ull : 4ull) << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(bool a_mn, bool b_mn, int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of the 16-byte chunk holding elements [c, c+8) of row r in a [128 x 128] bf16 K-major operand stored
// as two 64-column SWIZZLE_128B blocks of 16 KB
__device__ __forceinline__ uint32_t p_chunk_off(int r, int c) {
    return (uint32_t)(c >> 6) * 16384u + (uint32_t)r * 128u + ((((uint32_t)(c & 63) >> 3) ^ ((uint32_t)r & 7u)) << 4);
}

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
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS)
k_mha_tc_fwd(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
    constexpr int PITCH = DH * 2;                 // bytes per operand row
    constexpr int TILE = 128 * PITCH;             // Q / K / V tile bytes
    constexpr int KST = 2, VST = (DH == 64 ? 1 : 2);
    constexpr uint32_t TMEM_COLS = 256;           // S: [0,128), O: [128, 128+DH)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full, k_full[KST], k_empty[KST], v_full[VST], v_empty[VST], s_full, s_free, p_full, p_empty, o_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ int kv_lo_s, nkv_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base, k_s = q_s + TILE, v_s = k_s + KST * TILE, p_s = v_s + VST * TILE;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        for (int s = 0; s < KST; ++s) mbar_init(&k_full[s], 1), mbar_init(&k_empty[s], 1);
        for (int s = 0; s < VST; ++s) mbar_init(&v_full[s], 1), mbar_init(&v_empty[s], 1);
        mbar_init(&s_full, 1);
        mbar_init(&s_free, 128);
        mbar_init(&p_full, 128);
        mbar_init(&p_empty, 1);
        mbar_init(&o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        // contiguous key-row range visible from this query tile
        int lo = 0, n = 0;
        if (p.tile_bounds) {
            const int2 tb = p.tile_bounds[blockIdx.x];
            lo = tb.x, n = tb.y;
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
    } else {  // ===== softmax warps: one thread per query row =====
        const int qd = warp & 3;
        const int r = qd * 32 + lane;              // row inside the tile = TMEM lane
        const int64_t row = (int64_t)q0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
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
        float m = -INFINITY;
        for (int j = 0; j < nkv; ++j) {
            const int kv0 = kv_lo + j * BKV;
            mbar_wait(&s_full, (uint32_t)j & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                if (kv0 + c + 32 <= wlo || kv0 + c >= whi) continue;   // warp-uniform
                uint32_t rr[32];
                tmem_ld32(t_lane + c, rr);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int key = kv0 + c + i;
                    if (key >= lo && key < hi) m = fmaxf(m, __uint_as_float(rr[i]));
                }
            }
            tc_fence_before();
            mbar_arrive(&s_free);
        }
        const float m_use = (m == -INFINITY) ? 0.f : m * p.scale_log2;
        // ---- phase 2: probabilities with the final maximum; O accumulates in TMEM
        float l = 0.f;
        for (int j = 0; j < nkv; ++j) {
            const int kv0 = kv_lo + j * BKV;
            mbar_wait(&s_full, (uint32_t)(nkv + j) & 1u);
            tc_fence_after();
            if (j > 0) mbar_wait(&p_empty, (uint32_t)(j - 1) & 1u);   // P V of the previous tile has read the P buffer
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                if (kv0 + c + 32 <= wlo || kv0 + c >= whi) {           // fully masked for this warp: P = 0
#pragma unroll
                    for (int i = 0; i < 32; i += 8)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(p_s + p_chunk_off(r, c + i)), "r"(0u) : "memory");
                    continue;
                }
                uint32_t rr[32];
                tmem_ld32(t_lane + c, rr);
                if (c + 32 >= BKV || kv0 + c + 32 >= whi) {            // last group this warp reads: S may be overwritten
                    tc_fence_before();
                    mbar_arrive(&s_free);
                }
#pragma unroll
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    float pv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int key = kv0 + c + i0 + i;
                        const bool valid = key >= lo && key < hi;
                        float e = valid ? exp2f(fmaf(__uint_as_float(rr[i0 + i]), p.scale_log2, -m_use)) : 0.f;
                        l += e;
                        if (dr.on && valid) e *= drop_elem(dr, rk, (uint32_t)key);
                        pv[i] = e;
                    }
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_s + p_chunk_off(r, c + i0)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                }
            }
            if (kv0 >= whi || kv0 + BKV <= wlo) {   // this warp read nothing of the tile: it still owes the S release
                tc_fence_before();
                mbar_arrive(&s_free);
            }
            fence_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&p_full);
        }
        float o[DH];
        if (nkv > 0) {
            mbar_wait(&o_full, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < DH; c += 32) {
                uint32_t rr[32];
                tmem_ld32(t_lane + 128 + c, rr);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(rr[i]);
            }
            tc_fence_before();
        } else {
#pragma unroll
            for (int i = 0; i < DH; ++i) o[i] = 0.f;
        }
        if (row < p.n_rows) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            bf16* op = (bf16*)p.out + row * p.d + h * DH;
#pragma unroll
            for (int i = 0; i < DH; i += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv), h1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv), h3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(op + i) = pk;
            }
            p.lse[(int64_t)h * p.n_rows + row] = l > 0.f ? (m_use + log2f(l)) * LN2 : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int DH>
static cudaError_t launch_fwd(const CUtensorMap& map, const AttnParams& p, cudaStream_t st) {
    constexpr int TILE = 128 * DH * 2;
    constexpr int KST = 2, VST = (DH == 64 ? 1 : 2);
    const size_t smem = (size_t)(1 + KST + VST) * TILE + 32768 + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_mha_tc_fwd<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((unsigned)((p.n_rows + BQ - 1) / BQ), (unsigned)p.nhead);
    k_mha_tc_fwd<DH><<<grid, ATT_THREADS, smem, st>>>(map, p);
    return cudaGetLastError();
}


// per-step attention metadata (once per batch instead of three dependent loads in every CTA of every layer):
// row_bounds[row] = [lo, hi) token rows of the row's graph (0,0 for tail rows); tile_bounds[t] = (first row, number
// of 128-row tiles) of the contiguous row range that interacts with rows [128 t, 128 t + 128)
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
            tb.y = (tok_off[g1 + 1] - rb.x + 127) / 128;
        }
        tile_bounds[row >> 7] = tb;
    }
}

// ------------------------------------------------------------------------------------------ backward
// delta[h, row] = sum_c dO[row, h*dh + c] * O[row, h*dh + c]   (one warp per (row, head))
__global__ void __launch_bounds__(256)
k_mha_delta(const bf16* __restrict__ out, const bf16* __restrict__ dout, int64_t n_rows, int nhead, int dh,
            float* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    const int64_t item = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_rows * nhead) return;
    const int64_t row = item / nhead;
    const int h = (int)(item - row * nhead);
    const int64_t off = row * (int64_t)(nhead * dh) + h * dh;
    float s = 0.f;
    for (int c = lane * 2; c < dh; c += 64) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(out + off + c));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dout + off + c));
        s = fmaf(a.x, b.x, fmaf(a.y, b.y, s));
    }
    s = warp_sum(s);
    if (lane == 0) delta[(int64_t)h * n_rows + row] = s;
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

// FlashAttention-2 style recompute backward.  S = Q K^T and dP = dO V^T are formed with query rows on the TMEM
// lanes (thread = query row) in both modes:
//   DKV = false : the CTA owns a 128-row QUERY tile (Q, dO resident), streams the key tiles (K, V) and
//                 accumulates dQ += dS K                                     (A = dS K-major from smem, B = K MN-major)
//   DKV = true  : the CTA owns a 128-row KEY tile (K, V resident), streams the query tiles (Q, dO) and
//                 accumulates dV += P^T dO, dK += dS^T Q                     (A = P|dS MN-major from smem, B MN-major)
// with P = exp2(S*scale*log2e - lse*log2e), dS = P o (dP*keep - delta).
template <int DH, bool DKV>
__global__ void __launch_bounds__(ATT_THREADS)
k_mha_tc_bwd(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do, const AttnBwdParams p) {
    constexpr int PITCH = DH * 2;
    constexpr int TILE = 128 * PITCH;
    constexpr int STG = 2;
    constexpr uint32_t TMEM_COLS = 512;           // S [0,128) | dP [128,256) | acc0 [256,256+DH) | acc1 [320,320+DH)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t own_full, st_full[STG], st_empty[STG], sdp_full, pds_full, acc_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int st_lo_s, nst_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, h = blockIdx.y;      // first row of the owned tile (query rows or key rows)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t own_x = base, own_y = own_x + TILE;    // dQ: Q, dO   | dKV: K, V
    const uint32_t st_s = own_y + TILE;                   // ring: stage s -> X at st_s + 2*s*TILE, Y right after
    const uint32_t ds_s = st_s + 2 * STG * TILE;          // dS  [128 q x 128 keys] bf16, two swizzled 64-key blocks
    const uint32_t pp_s = ds_s + 32768;                   // P   (DKV only)

    if (threadIdx.x == 0) {
        mbar_init(&own_full, 1);
        for (int s = 0; s < STG; ++s) mbar_init(&st_full[s], 1), mbar_init(&st_empty[s], 1);
        mbar_init(&sdp_full, 1);
        mbar_init(&pds_full, 128);
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
            n = (p.tok_off[g1 + 1] - lo + 127) / 128;
        }
        st_lo_s = lo;
        nst_s = n;
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
                mbar_expect_tx(&st_full[s], 2 * TILE);
                const uint32_t x = st_s + 2 * s * TILE, y = x + TILE;
                const int r0 = st_lo + t * 128;
                if (!DKV) {
                    tma_load_2d(x, &tma_qkv, &st_full[s], colK, r0);
                    tma_load_2d(y, &tma_qkv, &st_full[s], colV, r0);
                } else {
                    tma_load_2d(x, &tma_qkv, &st_full[s], colQ, r0);
                    tma_load_2d(y, &tma_do, &st_full[s], colQ, r0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nst > 0) {  // ===== MMA issuer =====
            const uint32_t id_s = idesc_f16(false, false, 128, 128);
            const uint32_t id_acc = idesc_f16(DKV, true, 128, DH);
            mbar_wait(&own_full, 0);
            for (int t = 0; t < nst; ++t) {
                const int s = t % STG;
                const uint32_t x = st_s + 2 * s * TILE, y = x + TILE;
                const uint32_t q_a = DKV ? x : own_x, k_a = DKV ? own_x : x, do_a = DKV ? y : own_y, v_a = DKV ? own_y : y;
                mbar_wait(&st_full[s], (uint32_t)(t / STG) & 1u);
                if (t > 0) mbar_wait(&acc_done, (uint32_t)(t - 1) & 1u);   // S/dP columns are free again
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_f16(tmem, desc_k(q_a + k * 32, PITCH), desc_k(k_a + k * 32, PITCH), id_s, k > 0);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_f16(tmem + 128, desc_k(do_a + k * 32, PITCH), desc_k(v_a + k * 32, PITCH), id_s, k > 0);
                umma_commit(&sdp_full);
                mbar_wait(&pds_full, (uint32_t)t & 1u);
                tc_fence_after();
                if (!DKV) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // dQ += dS K : k = key
                        umma_f16(tmem + 256, desc_k(ds_s + (k >> 2) * 16384 + (k & 3) * 32, 128),
                                 desc_mn(k_a + k * 16 * PITCH, PITCH), id_acc, (t > 0 || k > 0));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // dK += dS^T Q : m = key, k = query row
                        umma_f16(tmem + 256, desc_mnmajor(ds_s + k * 2048, 16384), desc_mn(q_a + k * 16 * PITCH, PITCH), id_acc,
                                 (t > 0 || k > 0));
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // dV += P^T dO
                        umma_f16(tmem + 320, desc_mnmajor(pp_s + k * 2048, 16384), desc_mn(do_a + k * 16 * PITCH, PITCH), id_acc,
                                 (t > 0 || k > 0));
                }
                umma_commit(&st_empty[s]);
                umma_commit(&acc_done);
            }
        }
    } else {  // ===== math warps: thread = query row of the current (S, dP) tile =====
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(qd * 32) << 16);
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        for (int t = 0; t < nst; ++t) {
            const int64_t qrow = DKV ? (int64_t)st_lo + t * 128 + r : (int64_t)t0 + r;
            const int kv0 = DKV ? t0 : st_lo + t * 128;
            int lo = 0, hi = 0;
            float lse2 = 0.f, dl = 0.f;
            if (qrow < p.n_rows) {
                if (p.row_bounds) {
                    const int2 rb = p.row_bounds[qrow];
                    lo = rb.x, hi = rb.y;
                } else {
                    const int g = p.tok_graph[qrow];
                    if (g >= 0) lo = p.tok_off[g], hi = p.tok_off[g + 1];
                }
                if (hi > lo) {
                    lse2 = p.lse[(int64_t)h * p.n_rows + qrow] * LOG2E;
                    dl = p.delta[(int64_t)h * p.n_rows + qrow];
                }
            }
            const uint32_t rk = drop_row_key(dr, att_row_id_tc(h, qrow, p.n_rows));
            const int wlo = __reduce_min_sync(0xffffffffu, hi > lo ? lo : 0x7fffffff);
            const int whi = __reduce_max_sync(0xffffffffu, hi > lo ? hi : 0);
            if (t > 0) mbar_wait(&acc_done, (uint32_t)(t - 1) & 1u);   // previous P / dS tiles have been consumed
            mbar_wait(&sdp_full, (uint32_t)t & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 128; c += 16) {
                if (kv0 + c + 16 <= wlo || kv0 + c >= whi) {   // fully masked for this warp (warp-uniform): dS = P = 0
#pragma unroll
                    for (int i = 0; i < 16; i += 8) {
                        const uint32_t off = p_chunk_off(r, c + i);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ds_s + off), "r"(0u) : "memory");
                        if (DKV) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(pp_s + off), "r"(0u) : "memory");
                    }
                    continue;
                }
                uint32_t rs[16], rp[16];
                tmem_ld16(t_lane + c, rs);
                tmem_ld16(t_lane + 128 + c, rp);
                float pv[16], dsv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int key = kv0 + c + i;
                    const bool valid = key >= lo && key < hi;
                    float pr = valid ? exp2f(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse2)) : 0.f;
                    float dp = __uint_as_float(rp[i]);
                    if (dr.on && valid) {
                        const float mk = drop_elem(dr, rk, (uint32_t)key);
                        dp *= mk;
                        dsv[i] = pr * (dp - dl);
                        pr *= mk;
                    } else {
                        dsv[i] = valid ? pr * (dp - dl) : 0.f;
                    }
                    pv[i] = pr;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    const uint32_t off = p_chunk_off(r, c + i);
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(dsv[i], dsv[i + 1]), h1 = __floats2bfloat162_rn(dsv[i + 2], dsv[i + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(dsv[i + 4], dsv[i + 5]), h3 = __floats2bfloat162_rn(dsv[i + 6], dsv[i + 7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_s + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                                 "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                    if (DKV) {
                        h0 = __floats2bfloat162_rn(pv[i], pv[i + 1]), h1 = __floats2bfloat162_rn(pv[i + 2], pv[i + 3]);
                        h2 = __floats2bfloat162_rn(pv[i + 4], pv[i + 5]), h3 = __floats2bfloat162_rn(pv[i + 6], pv[i + 7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pp_s + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                                     "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                                     "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&pds_full);
        }
        // epilogue: the accumulators' TMEM lanes are the owned tile's rows
        const int64_t row = (int64_t)t0 + r;
        if (nst > 0) {
            mbar_wait(&acc_done, (uint32_t)(nst - 1) & 1u);
            tc_fence_after();
        }
        bf16* gp = (bf16*)p.dqkv + row * (int64_t)(3 * p.d);
#pragma unroll
        for (int a = 0; a < (DKV ? 2 : 1); ++a) {
            const int col = !DKV ? colQ : (a == 0 ? colK : colV);
            const float mul = (DKV && a == 1) ? 1.f : p.scale;
#pragma unroll
            for (int c = 0; c < DH; c += 16) {
                uint32_t rr[16];
                if (nst > 0) tmem_ld16(t_lane + 256 + a * 64 + c, rr);
                if (row < p.n_rows) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = nst > 0 ? __uint_as_float(rr[i + e]) * mul : 0.f;
                        uint4 pk;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(gp + col + c + i) = pk;
                    }
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

template <int DH, bool DKV>
static cudaError_t launch_bwd(const CUtensorMap& mq, const CUtensorMap& md, const AttnBwdParams& p, cudaStream_t st) {
    constexpr int TILE = 128 * DH * 2;
    const size_t smem = (size_t)(2 + 2 * 2) * TILE + 32768 * (DKV ? 2 : 1) + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_mha_tc_bwd<DH, DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid((unsigned)((p.n_rows + 127) / 128), (unsigned)p.nhead);
    k_mha_tc_bwd<DH, DKV><<<grid, ATT_THREADS, smem, st>>>(mq, md, p);
    return cudaGetLastError();
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
    CUtensorMap mq, md;
    if (!make_map(&mq, qkv, (uint64_t)3 * d, (uint64_t)n_rows, (uint64_t)3 * d, (uint32_t)dh, 128, dh * 2) ||
        !make_map(&md, dout, (uint64_t)d, (uint64_t)n_rows, (uint64_t)d, (uint32_t)dh, 128, dh * 2)) {
        set_error("cuTensorMapEncodeTiled failed or unavailable");
        return -2;
    }
    const int64_t items = n_rows * nhead;
    k_mha_delta<<<(unsigned)((items + 7) / 8), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, n_rows, nhead, dh, delta);
    AttnBwdParams p;
    p.tok_graph = tok_graph; p.tok_off = tok_off; p.lse = lse; p.delta = delta; p.dqkv = dqkv; p.rng = rng; p.salt = salt;
    p.row_bounds = (const int2*)row_bounds; p.tile_bounds = (const int2*)tile_bounds;
    p.n_rows = n_rows; p.B = (int)B; p.nhead = nhead; p.d = d;
    p.scale = scale; p.scale_log2 = scale * LOG2E; p.drop_p = drop_p;
    cudaError_t e;
    if (dh == 64) {
        e = launch_bwd<64, false>(mq, md, p, st);
        if (e == cudaSuccess) e = launch_bwd<64, true>(mq, md, p, st);
    } else {
        e = launch_bwd<32, false>(mq, md, p, st);
        if (e == cudaSuccess) e = launch_bwd<32, true>(mq, md, p, st);
    }
    if (e != cudaSuccess) return cuda_fail(e, "gt_mha_bwd(tcgen05)");
    return 0;
}
}  // namespace gt
