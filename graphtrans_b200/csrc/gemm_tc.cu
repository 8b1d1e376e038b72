// tcgen05 / TMEM / TMA dense contraction for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//   C[m,n] = sum_k A(m,k) * B(n,k)  (+bias[n]) (+resid[m,n]) (relu)      - same contract as gemm_simt.cu
// Replaces nn.Linear -> cuBLAS (reference modules/conv.py:18-20,44; modules/gnn_module.py:161-170;
// models/gnn_transformer.py:70,85-88; in/out projections + FFN of nn.TransformerEncoderLayer built at
// modules/transformer_encoder.py:28-32) and the three GEMMs of its backward (dX, dW split-K, bias).
//
// PERSISTENT kernel: one CTA per SM walks the (m-tile, n-tile, k-split) work list; a work item is one 128 x BN
// output tile (BN <= 256, runtime) over a K range.  The accumulator is double-buffered in TMEM so the epilogue of
// tile i overlaps the TMA + MMA main loop of tile i+1 (these GEMMs have only 2-10 k-blocks per tile, so the
// epilogue would otherwise dominate):
//   warp 0      : TMA producer  - cp.async.bulk.tensor.2d tiles (128B swizzle) into a 4-stage smem ring that runs
//                                 ahead across tile boundaries
//   warp 1      : MMA issuer    - one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16) into
//                                 TMEM buffer (tile & 1); tcgen05.commit releases smem stages / signals the epilogue
//   warps 2..5  : epilogue      - tcgen05.ld (32 lanes x 16 columns per warp and step) -> bias / residual / ReLU ->
//                                 16-byte global stores (or red.global.add.v4.f32 for split-K weight gradients),
//                                 then hands the TMEM buffer back
// Both operands may be K-major (row = m or n, k contiguous) or MN-major (row = k, m or n contiguous); the
// major-ness goes into the UMMA instruction descriptor and the TMA box shape, so forward (K,K), dX (K,MN) and
// dW (MN,MN) all run here without transposes.  Ragged M/N/K edges are zero-filled by TMA.
#include <stdlib.h>

#include "tc_common.cuh"

namespace gt {

namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int MAX_STAGES = 4;
constexpr int EPI_WARPS = 8;                                   // two per TMEM lane group: they alternate column slabs
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int STG_PITCH = 68;                                  // floats per staged row: 64 + 4 pad (bank-conflict free)
constexpr int STG_WARP_BYTES = 9216;                           // per epilogue warp: >= 32 x 68 fp32 (8704 B), 1 KB multiple so the
                                                               // two 4 KB TMA-store slabs inside it are 1024-byte aligned (swizzle)
constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;

struct Params {
    void* C;
    const float* bias;
    const void* resid;
    int64_t ldc, ldr;
    int M, N, n_fill, flags;
    int vec_c, vec_r;       // 16-byte vector access allowed on C / resid rows
    int kb_total, kb_per_split;
    int BN, stages;
    float drop_p;                        // epilogue dropout after the activation (vector index m*(ldc/4) + n/4)
    const uint64_t* rng;
    uint64_t salt;
    int tma_store;                       // epilogue writes C through TMA bulk stores / reductions (fast path)
    int m_tiles, n_tiles, total_tiles;   // work list: tile t -> n = t % n_tiles, m = (t / n_tiles) % m_tiles, split = rest
    uint32_t idesc, tmem_cols, acc_stride;
    unsigned long long* trace;           // debug (GT_GEMM_TRACE=1): clock64 stamps of CTA 0's phases, else NULL
    double* col_stats;                   // optional [2][ldc]: += column sums / sums of squares of the stored C (BatchNorm)
    const int32_t* m_valid;              // optional: only rows < m_valid[0] enter col_stats (shape-bucket slack rows)
};

#define GT_TRACE(slot) do { if (p.trace && blockIdx.x == 0) p.trace[slot] = (unsigned long long)clock64(); } while (0)

template <bool A_MN, bool B_MN, bool OUT_BF16>
__global__ void __launch_bounds__(THREADS)
k_gemm_tc(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
          const __grid_constant__ CUtensorMap tma_c, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float stat_sm[2][2][128];   // [column-slab parity][slab buffer][64 columns x (sum, sumsq)]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) GT_TRACE(0);
    for (int i = threadIdx.x; i < 2 * 2 * 128; i += blockDim.x) (&stat_sm[0][0][0])[i] = 0.f;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // B tile: K-major = BN rows of 128 B; MN-major = ceil(BN/64) TMA boxes of [64 k-rows x 64 columns] (8 KB each)
    const int b_boxes = (p.BN + 63) / 64;
    const uint32_t b_bytes = B_MN ? (uint32_t)b_boxes * 8192u : (uint32_t)p.BN * (BK * 2);
    const uint32_t stage_bytes = A_TILE_BYTES + b_bytes;

    if (threadIdx.x == 32) {   // descriptor fetch overlaps barrier init / TMEM allocation instead of the first TMA issue
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_c)) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], EPI_WARPS);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) GT_TRACE(1);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int n0 = (t % p.n_tiles) * p.BN, m0 = ((t / p.n_tiles) % p.m_tiles) * BM;
                const int kb0 = (t / (p.n_tiles * p.m_tiles)) * p.kb_per_split;
                const int nkb = min(p.kb_total, kb0 + p.kb_per_split) - kb0;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], stage_bytes);
                    const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes, b_dst = a_dst + A_TILE_BYTES;
                    const int k0 = (kb0 + kb) * BK;
                    if (!A_MN) {
                        tma_load_2d(a_dst, &tma_a, &full_bar[s], k0, m0);
                    } else {
                        tma_load_2d(a_dst, &tma_a, &full_bar[s], m0, k0);
                        tma_load_2d(a_dst + 8192, &tma_a, &full_bar[s], m0 + 64, k0);
                    }
                    if (!B_MN) {
                        tma_load_2d(b_dst, &tma_b, &full_bar[s], k0, n0);
                    } else {
                        for (int j = 0; j < b_boxes; ++j) tma_load_2d(b_dst + j * 8192, &tma_b, &full_bar[s], n0 + j * 64, k0);
                    }
                    if (it == 0) GT_TRACE(2);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            int it = 0, ti = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
                const int kb0 = (t / (p.n_tiles * p.m_tiles)) * p.kb_per_split;
                const int nkb = min(p.kb_total, kb0 + p.kb_per_split) - kb0;
                const int buf = ti & 1;
                mbar_wait(&tmem_empty_bar[buf], ((uint32_t)(ti >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)buf * p.acc_stride;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    if (it == 0) GT_TRACE(3);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes, b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = A_MN ? desc_mnmajor(a_addr + k * 2048, 8192) : desc_kmajor(a_addr + k * 32);
                        const uint64_t bd = B_MN ? desc_mnmajor(b_addr + k * 2048, 8192) : desc_kmajor(b_addr + k * 32);
                        umma_f16(acc, ad, bd, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
                }
                umma_commit(&tmem_full_bar[buf]);  // accumulator of this tile complete
                if (ti == 0) GT_TRACE(4);
            }
        }
    } else {  // ===== epilogue warps 2..9: TMEM lane group = warp % 4, column-slab parity = (warp - 2) / 4 =====
        // TMEM gives each thread one accumulator ROW (32 lanes x 16 columns per tcgen05.ld); storing rows straight
        // from registers would scatter 16-byte pieces over 32 rows per instruction.  Each warp therefore stages a
        // [32 rows x 64 columns] fp32 slab in its private shared-memory region and streams it out with 16 lanes per
        // row (256 B contiguous fp32 / 128 B bf16), applying bias / residual / ReLU / conversion on the way out
        // with coalesced reads.
        const int q = warp & 3, sub = (warp - 2) >> 2;
        const bool accum = p.flags & GT_EPI_ACCUM;
        const bool relu = p.flags & GT_EPI_RELU;
        const bool resid_f32 = p.flags & GT_EPI_RESID_F32;
        const int ncols = max(p.N, p.n_fill);
        float* stg = reinterpret_cast<float*>(smem_raw + ((smem_base - smem_u32(smem_raw)) + (uint32_t)p.stages * stage_bytes) + (warp - 2) * STG_WARP_BYTES);
        const int seg = lane & 15, half = lane >> 4;
        const Drop dr = make_drop(p.rng, p.salt, p.drop_p);
        int ti = 0, slab_i = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
            const int n0 = (t % p.n_tiles) * p.BN, m0 = ((t / p.n_tiles) % p.m_tiles) * BM;
            const bool first = (t / (p.n_tiles * p.m_tiles)) == 0;
            const int buf = ti & 1;
            const uint32_t acc = tmem_base + (uint32_t)buf * p.acc_stride + ((uint32_t)(q * 32) << 16);
            mbar_wait(&tmem_full_bar[buf], (uint32_t)(ti >> 1) & 1u);
            if (ti == 0 && warp == 2 && lane == 0) GT_TRACE(5);
            tc_fence_after();
            if (p.tma_store) {
                // fast path: thread = accumulator row.  TMEM -> registers -> (bias, ReLU, convert) -> this warp's
                // [32 rows x 128 B] 128B-swizzled staging slab -> one TMA bulk store (or fp32 reduce-add for split-K)
                // per slab, double-buffered so the next slab is produced while the previous one drains.
                constexpr int SLAB_COLS = OUT_BF16 ? 64 : 32;
                uint8_t* sbase = reinterpret_cast<uint8_t*>(stg);          // 2 x 4 KB of this warp's region
                float* bias_sm = reinterpret_cast<float*>(sbase + 8192);     // + 1 KB spare: bias of the current slab
                const uint32_t s_u32 = smem_u32(sbase);
                const uint32_t rsw = (uint32_t)lane & 7u;
                for (int c0 = sub * SLAB_COLS; c0 < p.BN; c0 += 2 * SLAB_COLS) {   // BN % SLAB_COLS == 0 on this path
                    const int n_slab = n0 + c0;
                    if (n_slab >= ncols) break;   // warp-uniform
                    const uint32_t sb = s_u32 + (uint32_t)(slab_i & 1) * 4096u;
                    // bias of this slab: two coalesced loads per lane into the warp-private spare KB of the staging
                    // region, read back below as broadcast 16-byte shared loads.  (One scalar global load per
                    // (thread, column) serialises on its destination registers: 2.6 us per slab, measured with
                    // tools/gemm_trace.py - the whole kernel was epilogue bound.)
                    const bool use_bias = first && p.bias;
                    // bf16 residual operand: the [32 rows x 128 B] box of this slab is fetched with coalesced 16-byte loads
                    // (8 lanes per row), parked in the staging slab in the store's swizzled layout and picked up below by
                    // the thread that owns the row (in place: a chunk is read by its owner right before it is rewritten)
                    const bool use_res = OUT_BF16 && first && p.resid != nullptr;
                    uint4 rv[8];
                    if (use_res) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = i * 4 + (lane >> 3), m = m0 + q * 32 + row, n = n_slab + (lane & 7) * 8;
                            rv[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (m < p.M && n + 8 <= p.N)
                                rv[i] = __ldg(reinterpret_cast<const uint4*>((const bf16*)p.resid + (int64_t)m * p.ldr + n));
                        }
                    }
                    if (use_bias) {
#pragma unroll
                        for (int i = lane; i < SLAB_COLS; i += 32) bias_sm[i] = (n_slab + i < p.N) ? __ldg(p.bias + n_slab + i) : 0.f;
                    }
                    if (lane == 0) tma_wait_group_read<1>();   // the store that last used this buffer has read it
                    __syncwarp();
                    if (use_res) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t row = (uint32_t)(i * 4 + (lane >> 3)), ch = (uint32_t)lane & 7u;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + row * 128u + ((ch ^ (row & 7u)) << 4)),
                                         "r"(rv[i].x), "r"(rv[i].y), "r"(rv[i].z), "r"(rv[i].w) : "memory");
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int c32 = 0; c32 < SLAB_COLS; c32 += 32) {
                      uint32_t r32[32];
                      tmem_ld32(acc + (uint32_t)(c0 + c32), r32);
#pragma unroll
                      for (int cc = 0; cc < 32; cc += 16) {
                        const int c = c32 + cc;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r32[cc + i]);
                        if (use_bias) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(bias_sm + c + i);
                                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                            }
                        }
                        if (use_res) {
#pragma unroll
                            for (int i = 0; i < 16; i += 8) {
                                const uint32_t ch = (uint32_t)(c + i) >> 3;
                                uint32_t w0, w1, w2, w3;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                             : "r"(sb + (uint32_t)lane * 128u + ((ch ^ rsw) << 4)) : "memory");
                                const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    v[i + 2 * k] += __uint_as_float(w[k] << 16);
                                    v[i + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
                                }
                            }
                        }
                        if (relu) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (dr.on) {
                            const uint64_t vbase = (uint64_t)(m0 + q * 32 + lane) * (uint64_t)(p.ldc >> 2) + (uint64_t)((n_slab + c) >> 2);
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                float dsc[4];
                                drop4(dr, vbase + (i >> 2), dsc);
                                v[i] *= dsc[0]; v[i + 1] *= dsc[1]; v[i + 2] *= dsc[2]; v[i + 3] *= dsc[3];
                            }
                        }
                        if (OUT_BF16) {
#pragma unroll
                            for (int i = 0; i < 16; i += 8) {
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i], v[i + 1]), h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                                const uint32_t ch = (uint32_t)(c + i) >> 3;
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (uint32_t)lane * 128u + ((ch ^ rsw) << 4)),
                                             "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)),
                                             "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const uint32_t ch = (uint32_t)(c + i) >> 2;
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (uint32_t)lane * 128u + ((ch ^ rsw) << 4)),
                                             "r"(__float_as_uint(v[i])), "r"(__float_as_uint(v[i + 1])), "r"(__float_as_uint(v[i + 2])),
                                             "r"(__float_as_uint(v[i + 3])) : "memory");
                            }
                        }
                      }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (OUT_BF16 && p.col_stats) {
                        // BatchNorm statistics of the tile while it sits in shared memory: lane = 2 adjacent columns of
                        // the [32 rows x 64 columns] bf16 slab (one conflict-free 128-byte row per shared load), fp32
                        // partials over the 32 rows; the four lane-group warps of a slab (128 rows) meet in shared
                        // memory and ONE warp issues the fp64 atomics (one per (tile, column) instead of four: the
                        // L2 atomics on ~1200 addresses are what this costs) - replaces the separate gt_colstats
                        // pass over C.  Rows beyond M hold bias-only garbage and are skipped.
                        const int m_lim = p.m_valid ? min(p.M, max(__ldg(p.m_valid), 1)) : p.M;
                        const int rows_valid = min(32, m_lim - (m0 + q * 32));      // warp-uniform
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                        for (int rr = 0; rr < rows_valid; ++rr) {
                            uint32_t u;
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(sb + (uint32_t)rr * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)rr & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u));
                            const float lo = __uint_as_float(u << 16), hi = __uint_as_float(u & 0xffff0000u);
                            s0 += lo; q0 = fmaf(lo, lo, q0);
                            s1 += hi; q1 = fmaf(hi, hi, q1);
                        }
                        float* sm = &stat_sm[sub][slab_i & 1][0];
                        if (rows_valid > 0) {
                            atomicAdd(sm + 4 * lane, s0);
                            atomicAdd(sm + 4 * lane + 1, q0);
                            atomicAdd(sm + 4 * lane + 2, s1);
                            atomicAdd(sm + 4 * lane + 3, q1);
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + sub) : "memory");   // the 4 warps of this column-slab parity
                        if (q == 0) {
                            const float4 t = *reinterpret_cast<const float4*>(sm + 4 * lane);
                            *reinterpret_cast<float4*>(sm + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
                            const int col = n_slab + 2 * lane;
                            if (col < p.N) {
                                atomicAdd(p.col_stats + col, (double)t.x);
                                atomicAdd(p.col_stats + p.ldc + col, (double)t.y);
                                if (col + 1 < p.N) {
                                    atomicAdd(p.col_stats + col + 1, (double)t.z);
                                    atomicAdd(p.col_stats + p.ldc + col + 1, (double)t.w);
                                }
                            }
                        }
                    }
                    if (lane == 0) {
                        if (accum) tma_reduce_add_2d(&tma_c, sb, n_slab, m0 + q * 32);
                        else tma_store_2d(&tma_c, sb, n_slab, m0 + q * 32);
                        tma_commit_group();
                    }
                    ++slab_i;
                }
                // this warp's share of the TMEM buffer has been read (wait::ld inside tmem_ld32): hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
                if (ti == 0 && warp == 2 && lane == 0) GT_TRACE(6);
                continue;
            }
            bool released = false;
            for (int c0 = sub * 64; c0 < p.BN; c0 += 128) {
                const int n_slab = n0 + c0;
                if (n_slab >= ncols) break;               // warp-uniform
                const int w_slab = min(64, p.BN - c0);    // multiple of 16
                for (int c = 0; c < w_slab; c += 16) {    // TMEM -> this thread's row of the slab
                    uint32_t r[16];
                    tmem_ld16(acc + (uint32_t)(c0 + c), r);
                    float* dst = stg + lane * STG_PITCH + c;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                        __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
                if (c0 + 128 >= p.BN || n_slab + 128 >= ncols) {   // this warp's last slab of the tile: its TMEM reads are done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
                    released = true;
                } else {
                    __syncwarp();
                }
                // ---- stream the slab out: lane -> 4 consecutive columns, 2 rows per pass
                const int n = n_slab + seg * 4;
                const bool col_in = seg * 4 < w_slab && n < ncols;
                const bool full4 = n + 4 <= p.N;
                float b4[4] = {0.f, 0.f, 0.f, 0.f};
                if (first && p.bias && col_in) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n + i < p.N) b4[i] = __ldg(p.bias + n + i);
                }
#pragma unroll 4
                for (int it = 0; it < 16; ++it) {
                    const int row = it * 2 + half;
                    const int m = m0 + q * 32 + row;
                    if (!col_in || m >= p.M) continue;
                    const float4 a = *reinterpret_cast<const float4*>(stg + row * STG_PITCH + seg * 4);
                    float v[4] = {a.x + b4[0], a.y + b4[1], a.z + b4[2], a.w + b4[3]};
                    if (first && p.resid) {
                        if (resid_f32) {
                            const float* rr = (const float*)p.resid + (int64_t)m * p.ldr + n;
                            if (full4 && p.vec_r) {
                                const float4 tr = *reinterpret_cast<const float4*>(rr);
                                v[0] += tr.x; v[1] += tr.y; v[2] += tr.z; v[3] += tr.w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (n + i < p.N) v[i] += rr[i];
                            }
                        } else {
                            const bf16* rr = (const bf16*)p.resid + (int64_t)m * p.ldr + n;
                            if (full4 && p.vec_r) {
                                float tr[4];
                                ld4(rr, tr);
                                v[0] += tr[0]; v[1] += tr[1]; v[2] += tr[2]; v[3] += tr[3];
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (n + i < p.N) v[i] += to_f(rr[i]);
                            }
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (dr.on) {
                        float dsc[4];
                        drop4(dr, (uint64_t)m * (uint64_t)(p.ldc >> 2) + (uint64_t)(n >> 2), dsc);
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] *= dsc[i];
                    }
                    if (OUT_BF16) {
                        bf16* cp = (bf16*)p.C + (int64_t)m * p.ldc + n;
                        if (full4 && p.vec_c) {
                            st4(cp, v);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (n + i < p.N) cp[i] = __float2bfloat16_rn(v[i]);
                                else if (n + i < p.n_fill) cp[i] = __float2bfloat16_rn(0.f);
                            }
                        }
                    } else {
                        float* cp = (float*)p.C + (int64_t)m * p.ldc + n;
                        if (accum) {
                            if (full4 && p.vec_c) {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (n + i < p.N) atomicAdd(cp + i, v[i]);
                            }
                        } else if (full4 && p.vec_c) {
                            *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (n + i < p.N) cp[i] = v[i];
                                else if (n + i < p.n_fill) cp[i] = 0.f;
                            }
                        }
                    }
                }
                __syncwarp();   // the slab region is reused by the next slab / tile
            }
            if (!released) {   // a warp without a slab in this tile still owes its arrival
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
        }
        if (warp == 2 && lane == 0) GT_TRACE(7);
        if (p.tma_store && lane == 0) tma_wait_group_all();   // staged slabs must be drained before the CTA exits
        if (warp == 2 && lane == 0) GT_TRACE(8);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
    if (threadIdx.x == 32) GT_TRACE(9);
}

// ---------------------------------------------------------------------------------------- host side
template <bool A_MN, bool B_MN>
static cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const Params& p, dim3 grid, size_t smem,
                          bool out_bf16, cudaStream_t st) {
    if (out_bf16) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024); attr = true; }
        k_gemm_tc<A_MN, B_MN, true><<<grid, THREADS, smem, st>>>(ma, mb, mc, p);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024); attr = true; }
        k_gemm_tc<A_MN, B_MN, false><<<grid, THREADS, smem, st>>>(ma, mb, mc, p);
    }
    return cudaGetLastError();
}

}  // namespace tc

static unsigned long long* g_trace_buf = nullptr;

// returns 0 ok, -2 = shape/layout/dtype not eligible (caller falls back to the CUDA-core kernel), >0 CUDA error
int gemm_tc_launch(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C, int64_t ldc,
                   int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias, const void* resid, int64_t ldr, int flags,
                   float drop_p, const uint64_t* rng, uint64_t salt, double* col_stats, const int32_t* m_valid, int* stats_fused, cudaStream_t st) {
    if (stats_fused) *stats_fused = 0;
    using namespace tc;
    if (dt != GT_BF16) { set_error("tcgen05 GEMM takes bf16 operands"); return -2; }
    if (((uintptr_t)A | (uintptr_t)B) & 15 || lda % 8 || ldb % 8) { set_error("operands need 16-byte aligned rows"); return -2; }
    if (M > (1ll << 30) || N > (1ll << 30) || K > (1ll << 30)) { set_error("extent too large"); return -2; }
    const bool out_bf16 = !(flags & GT_EPI_OUT_F32);
    const bool accum = flags & GT_EPI_ACCUM;
    const int64_t ncols = n_fill > N ? n_fill : N;
    // output tile width: the fewest tiles of <= 256 columns, rounded to the UMMA N granularity of 16 (wide tiles keep
    // the operand re-read factor, i.e. L2 traffic, low; an MN-major B tile is loaded as ceil(BN/64) TMA boxes)
    // The TMA-store epilogue (rows 16-byte aligned, no fp32 residual operand) moves whole [32 x 128 B] boxes, so there BN
    // is a multiple of the slab width (64 bf16 / 32 fp32 columns); padded columns only cost MMA issue slots.
    const int csz0 = out_bf16 ? 2 : 4;
    // a bf16 residual operand with 16-byte aligned rows rides in the TMA-store epilogue too (staged through the slab)
    const bool res_fast = resid && out_bf16 && !(flags & GT_EPI_RESID_F32) && !accum && ((uintptr_t)resid % 16 == 0) &&
                          ((ldr * 2) % 16 == 0) && N % 8 == 0;
    const bool tma_ok = ((uintptr_t)C % 16 == 0) && ((ldc * csz0) % 16 == 0) && (!resid || res_fast);
    const int gran = tma_ok ? (out_bf16 ? 64 : 32) : 16;
    const int64_t nt = (ncols + 255) / 256;
    const int BN = (int)(((ncols + nt - 1) / nt + gran - 1) / gran * gran);
    Params p;
    p.C = C; p.bias = bias; p.resid = resid; p.ldc = ldc; p.ldr = ldr;
    p.M = (int)M; p.N = (int)N; p.n_fill = (int)n_fill; p.flags = flags;
    p.drop_p = drop_p; p.rng = rng; p.salt = salt;
    const int csz = out_bf16 ? 2 : 4;
    p.vec_c = ((uintptr_t)C % 16 == 0) && ((ldc * csz) % 16 == 0);
    const int rsz = (flags & GT_EPI_RESID_F32) ? 4 : 2;
    p.vec_r = resid && ((uintptr_t)resid % 16 == 0) && ((ldr * rsz) % 16 == 0);
    p.kb_total = (int)((K + BK - 1) / BK);
    p.n_tiles = (int)((ncols + BN - 1) / BN);
    p.m_tiles = (int)((M + BM - 1) / BM);
    int splits = 1;
    if (accum) {   // split-K: enough work items for ~2 per SM (weight gradients: few output tiles, very long K)
        const int64_t tiles = (int64_t)p.n_tiles * p.m_tiles;
        splits = (int)(kNumSMs / tiles);   // ONE wave (<= 148 work items): an item more than the SM count doubles the duration
        // every split pays the CTA prologue and one fp32 reduction of its tile: at least `min_kb` k-blocks per split keep
        // the SM-seconds of these (side-stream) launches down, which is what the overlapped main stream feels
        // (measured in the overlapped step, config 2: 1 -> 2234 us, 8 -> 2200, 16 -> 2150, 32 -> 2180, 64 -> 2600 us/step)
        static const int min_kb = getenv("GT_SPLITK_MIN_KB") ? atoi(getenv("GT_SPLITK_MIN_KB")) : 16;
        if (min_kb > 1 && splits > p.kb_total / min_kb) splits = p.kb_total / min_kb;
        if (splits > p.kb_total) splits = p.kb_total;
        if (splits < 1) splits = 1;
    }
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    const int64_t total = (int64_t)p.n_tiles * p.m_tiles * splits;
    if (total > (1ll << 30)) { set_error("too many tiles"); return -2; }
    p.total_tiles = (int)total;
    dim3 grid((unsigned)(total < kNumSMs ? total : kNumSMs), 1, 1);   // persistent: one CTA per SM
    p.BN = BN;
    const size_t stage_bytes = A_TILE_BYTES + (b_mn ? (size_t)((BN + 63) / 64) * 8192 : (size_t)BN * BK * 2);
    p.stages = (int)((223 * 1024 - 1024 - STG_BYTES) / stage_bytes);   // ring + epilogue staging + alignment slack
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    const size_t smem = p.stages * stage_bytes + STG_BYTES + 1024;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
              ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    p.acc_stride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;   // two accumulator buffers
    p.tmem_cols = 2 * p.acc_stride;

    CUtensorMap ma, mb;
    bool ok = a_mn ? make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64)
                   : make_map(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM);
    ok = ok && (b_mn ? make_map(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64)
                     : make_map(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)BN));
    if (!ok) { set_error("cuTensorMapEncodeTiled failed or unavailable"); return -2; }
    // output through TMA when rows are 16-byte aligned (residual operand: bf16 only); [32 x 128 B] boxes per warp.
    // The map's column extent is max(N, n_fill): columns N..n_fill-1 receive the (zero) accumulators of B's OOB rows.
    CUtensorMap mc = ma;
    p.tma_store = tma_ok;
    if (p.tma_store && !make_map(&mc, C, (uint64_t)ncols, (uint64_t)M, (uint64_t)ldc, out_bf16 ? 64 : 32, 32, 128, !out_bf16))
        p.tma_store = 0;

    static unsigned long long* trace_buf = nullptr;
    static const bool trace_on = getenv("GT_GEMM_TRACE") != nullptr;
    if (trace_on && !trace_buf) { cudaMalloc(&trace_buf, 16 * sizeof(unsigned long long)); g_trace_buf = trace_buf; }
    p.trace = trace_on ? trace_buf : nullptr;
    // column statistics ride in the TMA-store epilogue of a bf16, non-accumulating output
    p.col_stats = (col_stats && p.tma_store && out_bf16 && !accum) ? col_stats : nullptr;
    p.m_valid = m_valid;
    if (stats_fused) *stats_fused = p.col_stats != nullptr;

    cudaError_t e;
    if (a_mn && b_mn) e = launch<true, true>(ma, mb, mc, p, grid, smem, out_bf16, st);
    else if (a_mn) e = launch<true, false>(ma, mb, mc, p, grid, smem, out_bf16, st);
    else if (b_mn) e = launch<false, true>(ma, mb, mc, p, grid, smem, out_bf16, st);
    else e = launch<false, false>(ma, mb, mc, p, grid, smem, out_bf16, st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_gemm(tcgen05)");
    return 0;
}

}  // namespace gt

// profiling hook (not part of the product ABI): phase stamps of CTA 0 of the last traced gt_gemm launch
extern "C" int gtdbg_gemm_trace_read(unsigned long long* out16) {
    if (!gt::g_trace_buf) return -1;
    return (int)cudaMemcpy(out16, gt::g_trace_buf, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
}
