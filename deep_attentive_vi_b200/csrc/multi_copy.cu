// Row f4 (trainer plumbing): gather the per-parameter gradient tensors that autograd produced into the
// flat gradient array in a handful of launches (one per 64 tensors) instead of one accumulate kernel
// per parameter (~2 600 launches of ~3 us per step for the 118.97 M-parameter model).  The flat array
// is what the single NCCL all-reduce and the fused Adamax / weight-norm-gradient kernels consume
// (replaces the implicit per-variable gradient aggregation of Keras under MirroredStrategy,
// util/trainer.py:160-187 of the reference).
#include "davi_common.cuh"

namespace davi {

constexpr int kMtMax = 64;           // tensors per launch (pointer table travels in the launch arguments)
constexpr int kMtChunk = 16384;      // elements per CTA
constexpr int kMtThreads = 256;

struct MtTable {
    const float* src[kMtMax];
    long long dst_off[kMtMax];
    long long size[kMtMax];
};

__global__ void __launch_bounds__(kMtThreads)
multi_copy_kernel(const MtTable tb, float* __restrict__ dst) {
    const int t = blockIdx.y;
    const long long n = tb.size[t];
    const long long i0 = (long long)blockIdx.x * kMtChunk;
    if (i0 >= n) return;
    const long long i1 = i0 + kMtChunk < n ? i0 + kMtChunk : n;
    const float* s = tb.src[t];
    float* d = dst + tb.dst_off[t];
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15u) == 0;
    if (vec) {
        const long long v1 = i0 + ((i1 - i0) & ~3ll);
        for (long long i = i0 + 4 * threadIdx.x; i < v1; i += 4 * kMtThreads)
            *reinterpret_cast<float4*>(d + i) = ld_stream4(s + i);
        for (long long i = v1 + threadIdx.x; i < i1; i += kMtThreads) d[i] = s[i];
    } else {
        for (long long i = i0 + threadIdx.x; i < i1; i += kMtThreads) d[i] = s[i];
    }
}

}  // namespace davi

using namespace davi;

extern "C" int davi_multi_copy(const float* const* srcs, const int64_t* dst_offsets, const int64_t* sizes,
                               int count, float* dst, davi_stream_t stream) {
    DAVI_REQUIRE(srcs && dst_offsets && sizes && dst && count >= 0, "null pointer");
    for (int base = 0; base < count; base += kMtMax) {
        const int m = count - base < kMtMax ? count - base : kMtMax;
        MtTable tb{};
        long long mx = 0;
        for (int i = 0; i < m; ++i) {
            DAVI_REQUIRE(srcs[base + i] && sizes[base + i] >= 0 && dst_offsets[base + i] >= 0, "bad table entry");
            tb.src[i] = srcs[base + i];
            tb.dst_off[i] = dst_offsets[base + i];
            tb.size[i] = sizes[base + i];
            if (tb.size[i] > mx) mx = tb.size[i];
        }
        if (mx == 0) continue;
        const unsigned gx = (unsigned)((mx + kMtChunk - 1) / kMtChunk);
        multi_copy_kernel<<<dim3(gx, m), kMtThreads, 0, (cudaStream_t)stream>>>(tb, dst);
        int rc = check_launch(__func__);
        if (rc) return rc;
    }
    return DAVI_OK;
}
