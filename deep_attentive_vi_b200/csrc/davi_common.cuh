// Shared helpers for the davi C-ABI kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "davi.h"

namespace davi {

// thread-local last-error text behind davi_last_error()
void set_error(const char* fmt, ...);
// cudaPeekAtLastError() after a launch; records text, returns DAVI_OK / DAVI_ECUDA
int check_launch(const char* what);
// current value of a davi_set_option() switch (0 = default)
int get_option(int option);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool mult4(int64_t v) { return (v & 3) == 0; }

}  // namespace davi

#define DAVI_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            davi::set_error("%s: %s (%s)", __func__, msg, #cond); \
            return DAVI_EINVAL;                                   \
        }                                                         \
    } while (0)

#define DAVI_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            davi::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
            return DAVI_ECUDA;                                                     \
        }                                                                          \
    } while (0)

namespace davi {

// ---- device helpers --------------------------------------------------------

__device__ __forceinline__ float4 ld_stream4(const float* p) {
    // read-once data: bypass L1 allocation
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x);
    acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z);
    acc.w = fmaf(s, v.w, acc.w);
}

// reductions over aligned groups of G lanes; `mask` must name (at least) the
// caller's whole group and only lanes that execute the call.
template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned mask = 0xffffffffu) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v, unsigned mask = 0xffffffffu) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ unsigned group_mask() {
    const unsigned lane = threadIdx.x & 31u;
    return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
}

// block-wide sum; result valid in thread 0.  `red` needs >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = group_sum<32>(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (warp == 0) v = group_sum<32>(v);
    return v;
}

// tfa.activations.gelu (tanh form) and its derivative
__device__ __forceinline__ float gelu_tanh(float x) {
    const float c = 0.7978845608028654f;
    return 0.5f * x * (1.f + tanhf(c * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
    const float c = 0.7978845608028654f;
    const float u = c * (x + 0.044715f * x * x * x);
    const float t = tanhf(u);
    return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * x * x);
}
__device__ __forceinline__ float softplusf(float x) {
    return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace davi
