// Shared helpers for the graphtrans_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/graphtrans_b200.h"

namespace gt {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define GT_CHECK_ARG(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            gt::set_error(__VA_ARGS__);        \
            return -1;                         \
        }                                      \
    } while (0)

#define GT_LAUNCH_CHECK(name)                                   \
    do {                                                        \
        cudaError_t e__ = cudaGetLastError();                   \
        if (e__ != cudaSuccess) return gt::cuda_fail(e__, name); \
    } while (0)

constexpr int kNumSMs = 148;
constexpr int VEC = 4;  // elements per vector access for feature rows (ld % 4 == 0)

using bf16 = __nv_bfloat16;

template <typename T> struct DType;
template <> struct DType<float> { static constexpr int id = GT_F32; };
template <> struct DType<bf16> { static constexpr int id = GT_BF16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store converting to/from fp32 registers
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
}

// V-wide (4 or 1) variants for rows whose logical sub-blocks are not multiples of 4
template <int V, typename T> __device__ __forceinline__ void ldv(const T* p, float (&v)[V]) {
    if constexpr (V == 4) ld4(p, v); else v[0] = to_f(p[0]);
}
template <int V, typename T> __device__ __forceinline__ void stv(T* p, const float (&v)[V]) {
    if constexpr (V == 4) st4(p, v); else p[0] = from_f<T>(v[0]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


// ---- counter-based dropout RNG (replaces torch's Philox stream; reference nn.Dropout / F.dropout
// sites gnn_module.py:86-90,205-209,227-229 and the 4 dropout sites of nn.TransformerEncoderLayer).
// rng_state is a DEVICE array {seed, step}: reading it on the device keeps every kernel
// CUDA-graph replayable (gt_rng_advance bumps `step` inside the graph); `salt` identifies the
// call site inside a step, `idx` the element (vector) inside the call.  splitmix64 finaliser.
struct Drop {
    uint64_t key;
    uint32_t thresh16;  // keep iff 16-bit field >= thresh16
    float inv_keep;
    bool on;
};
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ Drop make_drop(const uint64_t* __restrict__ rng_state, uint64_t salt, float p) {
    Drop d;
    d.on = p > 0.f && rng_state != nullptr;
    d.key = 0;
    d.thresh16 = 0;
    d.inv_keep = 1.f;
    if (d.on) {
        d.key = mix64(rng_state[0] ^ mix64(rng_state[1] * 0x9E3779B97F4A7C15ull + salt));
        d.thresh16 = (uint32_t)(p * 65536.f);
        d.inv_keep = 1.f / (1.f - p);
    }
    return d;
}
// scale factors (0 or 1/(1-p)) of the 4 elements of vector `vec_idx`
__device__ __forceinline__ void drop4(const Drop& d, uint64_t vec_idx, float (&s)[4]) {
    if (!d.on) { s[0] = s[1] = s[2] = s[3] = 1.f; return; }
    const uint64_t r = mix64(d.key + vec_idx * 0x9E3779B97F4A7C15ull);
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = ((uint32_t)(r >> (16 * q)) & 0xffffu) >= d.thresh16 ? d.inv_keep : 0.f;
}
// attention probabilities: element (row_id = head * n_rows + query row, col = key row).  The 64-bit part of the hash
// is paid once per row, the per-element part is a 32-bit murmur3 finaliser (7 integer ops) - shared by the CUDA-core
// and the tcgen05 attention kernels so that forward and backward of either kind see the same mask.
__device__ __forceinline__ uint32_t drop_row_key(const Drop& d, uint64_t row_id) {
    return d.on ? (uint32_t)(mix64(d.key + row_id * 0x9E3779B97F4A7C15ull) >> 32) : 0u;
}
__device__ __forceinline__ float drop_elem(const Drop& d, uint32_t row_key, uint32_t col) {
    uint32_t x = row_key ^ (col * 0x9E3779B1u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return (x >> 16) >= d.thresh16 ? d.inv_keep : 0.f;
}
__device__ __forceinline__ float drop1(const Drop& d, uint64_t idx) {
    if (!d.on) return 1.f;
    const uint64_t r = mix64(d.key + idx * 0x9E3779B97F4A7C15ull);
    return ((uint32_t)(r >> 24) & 0xffffu) >= d.thresh16 ? d.inv_keep : 0.f;
}

inline int blocks_for(int64_t work_items, int per_block, int max_blocks = kNumSMs * 16) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

#define GT_DISPATCH_DT(dt, ...)                                      \
    do {                                                             \
        if ((dt) == GT_F32) { using T = float; __VA_ARGS__; }        \
        else if ((dt) == GT_BF16) { using T = gt::bf16; __VA_ARGS__; } \
        else { gt::set_error("bad dtype %d", (int)(dt)); return -1; } \
    } while (0)

}  // namespace gt
