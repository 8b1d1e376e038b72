// CUDA-core reference GEMM (fp32 accumulate) for every operand layout the hot path needs.
// It is (a) the exact-fp32 parity path (GT_F32 activations), (b) the fallback for shapes the
// tcgen05 kernel (gemm_tc.cu) does not take, and (c) the on-device cross-check for that kernel.
//   C[m,n] = sum_k A(m,k) * B(n,k)  (+bias[n]) (+resid[m,n]) (relu)
// Replaces nn.Linear -> cuBLAS SGEMM (reference modules/conv.py:18-20,44;
// modules/gnn_module.py:161-170; models/gnn_transformer.py:70,85-88).
#include "common.cuh"

namespace gt {

int gemm_tc_launch(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias,
                   const void* resid, int64_t ldr, int flags, float drop_p, const uint64_t* rng, uint64_t salt,
                   double* col_stats, const int32_t* m_valid, int* stats_fused, cudaStream_t st);  // gemm_tc.cu; -2 = not eligible

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256)
k_gemm_simt(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, void* __restrict__ Cv,
            int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* __restrict__ bias,
            const void* __restrict__ resid, int64_t ldr, int flags, int64_t k_per_split, float drop_p,
            const uint64_t* __restrict__ rng, uint64_t salt) {
    const Drop dr = make_drop(rng, salt, drop_p);
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kb; k0 < ke; k0 += SBK) {
        if (A_MN) {
            const int k = t >> 4, m4 = (t & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t m = m0 + m4 + q, kk = k0 + k;
                As[k][m4 + q] = (m < M && kk < ke) ? to_f(A[kk * lda + m]) : 0.f;
            }
        } else {
            const int m = t >> 2, k4 = (t & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t mm = m0 + m, kk = k0 + k4 + q;
                As[k4 + q][m] = (mm < M && kk < ke) ? to_f(A[mm * lda + kk]) : 0.f;
            }
        }
        if (B_MN) {
            const int k = t >> 4, n4 = (t & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t n = n0 + n4 + q, kk = k0 + k;
                Bs[k][n4 + q] = (n < N && kk < ke) ? to_f(B[kk * ldb + n]) : 0.f;
            }
        } else {
            const int n = t >> 2, k4 = (t & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t nn = n0 + n, kk = k0 + k4 + q;
                Bs[k4 + q][n] = (nn < N && kk < ke) ? to_f(B[nn * ldb + kk]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool out_f32 = (flags & GT_EPI_OUT_F32) || DType<T>::id == GT_F32;
    const bool accum = flags & GT_EPI_ACCUM;
    const bool first = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float dsc[4];   // epilogue dropout: element (m, n) belongs to vector m*(ldc/4) + n/4 (n0 + tx*4 is 4-aligned)
        drop4(dr, (uint64_t)(m * (ldc >> 2) + ((n0 + tx * 4) >> 2)), dsc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n < N) {
                float v = acc[i][j];
                if (first) {
                    if (bias) v += bias[n];
                    if (resid) v += (flags & GT_EPI_RESID_F32) ? ((const float*)resid)[m * ldr + n] : to_f(((const T*)resid)[m * ldr + n]);
                }
                if (flags & GT_EPI_RELU) v = fmaxf(v, 0.f);
                v *= dsc[j];
                if (out_f32) {
                    float* c = (float*)Cv + m * ldc + n;
                    if (accum) atomicAdd(c, v); else *c = v;
                } else {
                    ((T*)Cv)[m * ldc + n] = from_f<T>(v);
                }
            } else if (n < n_fill && first && !accum) {
                if (out_f32) ((float*)Cv)[m * ldc + n] = 0.f; else ((T*)Cv)[m * ldc + n] = from_f<T>(0.f);
            }
        }
    }
}

template <typename T>
static int launch_simt(const T* A, int a_mn, int64_t lda, const T* B, int b_mn, int64_t ldb, void* C, int64_t ldc,
                       int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias, const void* resid,
                       int64_t ldr, int flags, float drop_p, const uint64_t* rng, uint64_t salt, cudaStream_t st) {
    const int64_t ncols = n_fill > N ? n_fill : N;
    dim3 grid((unsigned)((ncols + SBN - 1) / SBN), (unsigned)((M + SBM - 1) / SBM), 1);
    int64_t kps = K;
    if ((flags & GT_EPI_ACCUM) && ((flags & GT_EPI_OUT_F32) || DType<T>::id == GT_F32)) {
        // split-K: fill the machine when the output tile grid is small (weight gradients)
        const int64_t tiles = (int64_t)grid.x * grid.y;
        int64_t splits = (kNumSMs * 4 + tiles - 1) / tiles;
        const int64_t max_splits = (K + 255) / 256;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        kps = ((K + splits - 1) / splits + SBK - 1) / SBK * SBK;
        grid.z = (unsigned)((K + kps - 1) / kps);
    }
#define L(AM, BM_) k_gemm_simt<T, AM, BM_><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, kps, drop_p, rng, salt)
    if (a_mn && b_mn) L(true, true);
    else if (a_mn) L(true, false);
    else if (b_mn) L(false, true);
    else L(false, false);
#undef L
    return 0;
}

}  // namespace gt

using namespace gt;

extern "C" int gt_colstats(int dt, const void* x, int64_t M, int32_t ld, double* stats, const int32_t* m_valid, void* stream);

static int gemm_dispatch(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C,
                         int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias,
                         const void* resid, int64_t ldr, int flags, float drop_p, const uint64_t* rng_state, uint64_t salt,
                         int impl, double* col_stats, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(M > 0 && N > 0 && K > 0, "gt_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    GT_CHECK_ARG(!(flags & GT_EPI_ACCUM) || (flags & GT_EPI_OUT_F32) || dt == GT_F32, "gt_gemm: ACCUM needs an fp32 C");
    GT_CHECK_ARG(!((flags & GT_EPI_ACCUM) && (flags & GT_EPI_RELU)), "gt_gemm: ACCUM and RELU are exclusive");
    GT_CHECK_ARG(!(drop_p > 0.f) || (!(flags & GT_EPI_ACCUM) && ldc % 4 == 0), "gt_gemm: epilogue dropout needs ldc %% 4 == 0 and no ACCUM");
    GT_CHECK_ARG(!col_stats || (!(flags & GT_EPI_ACCUM) && ldc % 4 == 0 && n_fill >= ldc), "gt_gemm_stats: needs a non-accumulating C with all ldc columns written");
    cudaStream_t st = (cudaStream_t)stream;
    const int dt_out = (flags & GT_EPI_OUT_F32) ? GT_F32 : dt;
    if (impl != 1) {
        int fused = 0;
        const int r = gemm_tc_launch(dt, A, a_mn, lda, B, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng_state, salt, col_stats, m_valid, &fused, st);
        if (r == 0 && col_stats && !fused) return gt_colstats(dt_out, C, M, (int32_t)ldc, col_stats, m_valid, stream);
        if (r != -2) return r;
        GT_CHECK_ARG(impl != 2, "gt_gemm: shape/layout not eligible for the tcgen05 kernel (%s)", gt_last_error());
    }
    GT_DISPATCH_DT(dt, launch_simt<T>((const T*)A, a_mn, lda, (const T*)B, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng_state, salt, st));
    GT_LAUNCH_CHECK("gt_gemm(simt)");
    if (col_stats) return gt_colstats(dt_out, C, M, (int32_t)ldc, col_stats, m_valid, stream);
    return 0;
}

extern "C" int gt_gemm(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C,
                       int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias,
                       const void* resid, int64_t ldr, int flags, float drop_p, const uint64_t* rng_state, uint64_t salt,
                       int impl, void* stream) {
    return gemm_dispatch(dt, A, a_mn, lda, B, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng_state, salt,
                         impl, nullptr, nullptr, stream);
}

extern "C" int gt_gemm_stats(int dt, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, void* C,
                             int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t n_fill, const float* bias,
                             const void* resid, int64_t ldr, int flags, float drop_p, const uint64_t* rng_state, uint64_t salt,
                             int impl, double* col_stats, const int32_t* m_valid, void* stream) {
    GT_CHECK_ARG(col_stats != nullptr, "gt_gemm_stats: col_stats is NULL");
    return gemm_dispatch(dt, A, a_mn, lda, B, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng_state, salt,
                         impl, col_stats, m_valid, stream);
}
