// tcgen05 / TMEM / TMA dense contraction for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//   C[m,n] = sum_k A(m,k) * B(n,k)  (+bias[n]) (+resid[m,n]) (relu)      - same contract as gemm_simt.cu
// Replaces nn.Linear -> cuBLAS (reference modules/conv.py:18-20,44; modules/gnn_module.py:161-170;
// models/gnn_transformer.py:70,85-88; in/out projections + FFN of nn.TransformerEncoderLayer built at
// modules/transformer_encoder.py:28-32) and the three GEMMs of its backward (dX, dW split-K, bias).
//
// One CTA computes one 128 x BN output tile (BN <= 256, runtime) over a K range:
//   warp 0      : TMA producer  - cp.async.bulk.tensor.2d tiles (128B swizzle) into a 3-4 stage smem ring
//   warp 1      : MMA issuer    - one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16),
//                                 accumulator in TMEM; tcgen05.commit releases smem stages / signals the epilogue
//   warps 2..5  : epilogue      - tcgen05.ld (32 lanes x 16 columns per warp and step) -> bias / residual / ReLU ->
//                                 16-byte global stores (or red.global.add.v4.f32 for split-K weight gradients)
// Both operands may be K-major (row = m or n, k contiguous) or MN-major (row = k, m or n contiguous); the
// major-ness goes into the UMMA instruction descriptor and the TMA box shape, so forward (K,K), dX (K,MN) and
// dW (MN,MN) all run here without transposes.  Ragged M/N/K edges are zero-filled by TMA.
#include "tc_common.cuh"

namespace gt {

namespace tc {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int MAX_STAGES = 4;
constexpr int THREADS = 192;

struct Params {
    void* C;
    const float* bias;
    const void* resid;
    int64_t ldc, ldr;
    int M, N, n_fill, flags;
    int vec_c, vec_r;       // 16-byte vector access allowed on C / resid rows
    int kb_total, kb_per_split;
    int BN, stages;
    uint32_t idesc, tmem_cols;
};

template <bool A_MN, bool B_MN, bool OUT_BF16>
__global__ void __launch_bounds__(THREADS)
k_gemm_tc(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * p.BN;
    const int kb0 = blockIdx.z * p.kb_per_split;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
    const int nkb = kb1 - kb0;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.BN * (BK * 2);
    const uint32_t stage_bytes = A_TILE_BYTES + b_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
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

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int it = 0; it < nkb; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], stage_bytes);
                const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes, b_dst = a_dst + A_TILE_BYTES;
                const int k0 = (kb0 + it) * BK;
                if (!A_MN) {
                    tma_load_2d(a_dst, &tma_a, &full_bar[s], k0, m0);
                } else {
                    tma_load_2d(a_dst, &tma_a, &full_bar[s], m0, k0);
                    tma_load_2d(a_dst + 8192, &tma_a, &full_bar[s], m0 + 64, k0);
                }
                if (!B_MN) {
                    tma_load_2d(b_dst, &tma_b, &full_bar[s], k0, n0);
                } else {
                    for (int j = 0; j < p.BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tma_b, &full_bar[s], n0 + j * 64, k0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            for (int it = 0; it < nkb; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes, b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t ad = A_MN ? desc_mnmajor(a_addr + k * 2048, 8192) : desc_kmajor(a_addr + k * 32);
                    const uint64_t bd = B_MN ? desc_mnmajor(b_addr + k * 2048, 8192) : desc_kmajor(b_addr + k * 32);
                    umma_f16(tmem_base, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);     // accumulator complete
        }
    } else {  // ===== epilogue warps 2..5: TMEM lane group = warp % 4 =====
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const bool first = blockIdx.z == 0;
        const bool accum = p.flags & GT_EPI_ACCUM;
        const bool relu = p.flags & GT_EPI_RELU;
        const bool resid_f32 = p.flags & GT_EPI_RESID_F32;
        const int ncols = max(p.N, p.n_fill);
        for (int c = 0; c < p.BN; c += 16) {
            const int n = n0 + c;
            if (n >= ncols) break;  // warp-uniform
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            if (m >= p.M) continue;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            const bool full16 = n + 16 <= p.N;
            if (first) {
                if (p.bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n + i < p.N) v[i] += __ldg(p.bias + n + i);
                }
                if (p.resid) {
                    if (resid_f32) {
                        const float* rr = (const float*)p.resid + (int64_t)m * p.ldr + n;
                        if (full16 && p.vec_r) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 t = *reinterpret_cast<const float4*>(rr + i);
                                v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n + i < p.N) v[i] += rr[i];
                        }
                    } else {
                        const bf16* rr = (const bf16*)p.resid + (int64_t)m * p.ldr + n;
                        if (full16 && p.vec_r) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                float t[4];
                                ld4(rr + i, t);
                                v[i] += t[0]; v[i + 1] += t[1]; v[i + 2] += t[2]; v[i + 3] += t[3];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n + i < p.N) v[i] += to_f(rr[i]);
                        }
                    }
                }
            }
            if (relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (OUT_BF16) {
                bf16* cp = (bf16*)p.C + (int64_t)m * p.ldc + n;
                if (full16 && p.vec_c) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8) {
                        uint4 pk;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i], v[i + 1]), h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(cp + i) = pk;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (n + i < p.N) cp[i] = __float2bfloat16_rn(v[i]);
                        else if (n + i < p.n_fill) cp[i] = __float2bfloat16_rn(0.f);
                    }
                }
            } else {
                float* cp = (float*)p.C + (int64_t)m * p.ldc + n;
                if (accum) {
                    if (full16 && p.vec_c) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n + i < p.N) atomicAdd(cp + i, v[i]);
                    }
                } else if (full16 && p.vec_c) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (n + i < p.N) cp[i] = v[i];
                        else if (n + i < p.n_fill) cp[i] = 0.f;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------- host side
template <bool A_MN, bool B_MN>
static cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, dim3 grid, size_t smem, bool out_bf16, cudaStream_t st) {
    if (out_bf16) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
        k_gemm_tc<A_MN, B_MN, true><<<grid, THREADS, smem, st>>>(ma, mb, p);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
        k_gemm_tc<A_MN, B_MN, false><<<grid, THREADS, smem, st>>>(ma, mb, p);
    }
    return cudaGetLastError();
}

}  // namespace tc

// returns 0 ok, -2 = shape/layout/dtype not eligible (caller falls back to the CUDA-core kernel), >0 CUDA error
int gemm_tc_launch(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C, int64_t ldc,
                   int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias, const void* resid, int64_t ldr, int flags,
                   cudaStream_t st) {
    using namespace tc;
    if (dt != GT_BF16) { set_error("tcgen05 GEMM takes bf16 operands"); return -2; }
    if (((uintptr_t)A | (uintptr_t)B) & 15 || lda % 8 || ldb % 8) { set_error("operands need 16-byte aligned rows"); return -2; }
    if (M > (1ll << 30) || N > (1ll << 30) || K > (1ll << 30)) { set_error("extent too large"); return -2; }
    const bool out_bf16 = !(flags & GT_EPI_OUT_F32);
    const bool accum = flags & GT_EPI_ACCUM;
    const int64_t ncols = n_fill > N ? n_fill : N;
    // output tile width: K-major B may use any multiple of 16, MN-major B whole 64-column TMA boxes
    int BN;
    if (!b_mn) {
        const int64_t nt = (ncols + 255) / 256;
        BN = (int)(((ncols + nt - 1) / nt + 15) / 16 * 16);
    } else {
        int best = 64;
        int64_t best_pad = -1;
        for (int cand = 64; cand <= 256; cand += 64) {
            const int64_t pad = (ncols + cand - 1) / cand * cand;
            if (best_pad < 0 || pad < best_pad || (pad == best_pad && cand > best)) best = cand, best_pad = pad;
        }
        BN = best;
    }
    Params p;
    p.C = C; p.bias = bias; p.resid = resid; p.ldc = ldc; p.ldr = ldr;
    p.M = (int)M; p.N = (int)N; p.n_fill = (int)n_fill; p.flags = flags;
    const int csz = out_bf16 ? 2 : 4;
    p.vec_c = ((uintptr_t)C % 16 == 0) && ((ldc * csz) % 16 == 0);
    const int rsz = (flags & GT_EPI_RESID_F32) ? 4 : 2;
    p.vec_r = resid && ((uintptr_t)resid % 16 == 0) && ((ldr * rsz) % 16 == 0);
    p.kb_total = (int)((K + BK - 1) / BK);
    dim3 grid((unsigned)((ncols + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), 1);
    int splits = 1;
    if (accum) {
        const int64_t tiles = (int64_t)grid.x * grid.y;
        splits = (int)((kNumSMs * 2 + tiles - 1) / tiles);
        if (splits > p.kb_total) splits = p.kb_total;
        if (splits < 1) splits = 1;
    }
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    grid.z = (unsigned)((p.kb_total + p.kb_per_split - 1) / p.kb_per_split);
    p.BN = BN;
    const size_t stage_bytes = A_TILE_BYTES + (size_t)BN * BK * 2;
    // ring depth: enough to cover TMA latency, but small enough that two CTAs share an SM (<= ~112 KB each) so one
    // CTA's prologue / epilogue overlaps the other's main loop; these GEMMs have only 1-10 k-blocks per CTA
    int fit = (int)((112 * 1024 - 1024) / stage_bytes);
    if (fit < 2) fit = 2;
    if (fit > MAX_STAGES) fit = MAX_STAGES;
    p.stages = p.kb_per_split < fit ? p.kb_per_split : fit;
    const size_t smem = p.stages * stage_bytes + 1024;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
              ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    p.tmem_cols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

    CUtensorMap ma, mb;
    bool ok = a_mn ? make_map(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64)
                   : make_map(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM);
    ok = ok && (b_mn ? make_map(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64)
                     : make_map(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)BN));
    if (!ok) { set_error("cuTensorMapEncodeTiled failed or unavailable"); return -2; }

    cudaError_t e;
    if (a_mn && b_mn) e = launch<true, true>(ma, mb, p, grid, smem, out_bf16, st);
    else if (a_mn) e = launch<true, false>(ma, mb, p, grid, smem, out_bf16, st);
    else if (b_mn) e = launch<false, true>(ma, mb, p, grid, smem, out_bf16, st);
    else e = launch<false, false>(ma, mb, p, grid, smem, out_bf16, st);
    if (e != cudaSuccess) return cuda_fail(e, "gt_gemm(tcgen05)");
    return 0;
}

}  // namespace gt
