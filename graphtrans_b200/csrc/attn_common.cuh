// Helpers shared by the tcgen05 attention kernels (attn_tc.cu: streamed block-diagonal tiles; attn_local.cu:
// graph-aligned single tiles for small graphs): operand descriptors, the P-tile smem layout, fast exp2.
#pragma once
#include "tc_common.cuh"

namespace gt {

__device__ __forceinline__ uint64_t att_row_id_tc(int h, int64_t q, int64_t n_rows) {   // == attn_simt.cu att_row_id
    return (uint64_t)h * (uint64_t)n_rows + (uint64_t)q;
}

namespace tc {

constexpr int ATT_THREADS = 320;          // backward kernels: TMA warp, MMA warp, 8 math warps (2 per TMEM lane group)
constexpr int ATT_FWD_THREADS = 320;      // forward: TMA warp, MMA warp, 8 softmax warps (2 per TMEM lane group)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
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

// bit i set <=> lo <= c + i < hi   (validity of the 32 columns [c, c + 32) for a row with the index range [lo, hi))
__device__ __forceinline__ uint32_t range_mask32(int lo, int hi, int c) {
    const int a = min(max(lo - c, 0), 32), b = min(max(hi - c, 0), 32);
    const uint32_t below_b = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
    const uint32_t below_a = a >= 32 ? 0xffffffffu : ((1u << a) - 1u);
    return below_b & ~below_a;
}

}  // namespace tc
}  // namespace gt
