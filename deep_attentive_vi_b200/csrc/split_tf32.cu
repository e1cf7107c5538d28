// Row f1 (host-model glue): operand split for fp32-grade convolutions on the TF32 tensor cores.
// The reference computes its Conv2D layers in fp32 (layers/resnet.py:589-601 under TF 2.3); cuDNN's
// fp32 convolutions run on the CUDA cores (~240 ms of a 283 ms step), its TF32 ones drop 13 mantissa
// bits of every operand (gradient error 5e-4).  Error compensation ("3xTF32") keeps the tensor cores:
//     v ~= hi + lo,  hi = rna_tf32(v),   lo = rna_tf32(v - hi)   (both exactly representable in TF32, so the
//                                                          tensor core's operand truncation changes nothing)
//     a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b          (error ~2^-21 relative, fp32 accumulation)
// and the three products are ONE convolution over a three times longer contraction axis: this kernel
// writes the concatenated operand.  src is [rows, C] contiguous; dst is
//     layout 0: [rows, 3C]   (concatenation along the channel axis: forward / input gradient)
//     layout 1: [3, rows, C] (three planes = concatenation along the batch axis: weight gradient)
// with the part order (hi, lo, hi) for `pattern` 0 and (hi, hi, lo) for pattern 1 — the two operands
// of a product take opposite patterns.  HBM-bound: reads 4 B and writes 12 B per element.
#include "davi_common.cuh"

namespace davi {

constexpr int kSplitThreads = 256;

__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

template <int PATTERN>
__device__ __forceinline__ void parts(float v, float& p0, float& p1, float& p2) {
    const float hi = rna_tf32(v), lo = rna_tf32(v - hi);
    p0 = hi;
    p1 = PATTERN == 0 ? lo : hi;
    p2 = PATTERN == 0 ? hi : lo;
}

// one thread per float4 of src; C % 4 == 0
template <int PATTERN>
__global__ void __launch_bounds__(kSplitThreads)
split_tf32_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n4, int C4, int64_t part_stride,
                      int64_t row_stride) {
    // element (r, c) -> dst[r * row_stride + k * part_stride + c], k = 0..2
    for (int64_t i = (int64_t)blockIdx.x * kSplitThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSplitThreads) {
        const int64_t r = i / C4;
        const int c = (int)(i - r * C4) * 4;
        const float4 v = ld_stream4(src + 4 * i);
        float4 a, b, d;
        parts<PATTERN>(v.x, a.x, b.x, d.x);
        parts<PATTERN>(v.y, a.y, b.y, d.y);
        parts<PATTERN>(v.z, a.z, b.z, d.z);
        parts<PATTERN>(v.w, a.w, b.w, d.w);
        float* o = dst + r * row_stride + c;
        *reinterpret_cast<float4*>(o) = a;
        *reinterpret_cast<float4*>(o + part_stride) = b;
        *reinterpret_cast<float4*>(o + 2 * part_stride) = d;
    }
}

template <int PATTERN>
__global__ void __launch_bounds__(kSplitThreads)
split_tf32_scalar_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int C, int64_t part_stride,
                         int64_t row_stride) {
    for (int64_t i = (int64_t)blockIdx.x * kSplitThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSplitThreads) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        float a, b, d;
        parts<PATTERN>(src[i], a, b, d);
        float* o = dst + r * row_stride + c;
        o[0] = a;
        o[part_stride] = b;
        o[2 * part_stride] = d;
    }
}

// Both layouts from one read of src: dst_ch [rows, 3C] with part order `pat_ch`, dst_pl [3, rows, C] with
// `pat_pl` (an activation feeds the forward conv in one layout and the weight gradient in the other; an
// output gradient feeds the input gradient and the weight gradient).
__global__ void __launch_bounds__(kSplitThreads)
split_tf32_dual_vec_kernel(const float* __restrict__ src, float* __restrict__ dst_ch, float* __restrict__ dst_pl,
                           int64_t n4, int C4, int pat_ch, int pat_pl) {
    const int64_t C = 4 * (int64_t)C4, n = 4 * n4;
    for (int64_t i = (int64_t)blockIdx.x * kSplitThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kSplitThreads) {
        const int64_t r = i / C4;
        const int c = (int)(i - r * C4) * 4;
        const float4 v = ld_stream4(src + 4 * i);
        float4 hi, lo;
        hi.x = rna_tf32(v.x); lo.x = rna_tf32(v.x - hi.x);
        hi.y = rna_tf32(v.y); lo.y = rna_tf32(v.y - hi.y);
        hi.z = rna_tf32(v.z); lo.z = rna_tf32(v.z - hi.z);
        hi.w = rna_tf32(v.w); lo.w = rna_tf32(v.w - hi.w);
        float* o = dst_ch + r * 3 * C + c;
        *reinterpret_cast<float4*>(o) = hi;
        *reinterpret_cast<float4*>(o + C) = pat_ch == 0 ? lo : hi;
        *reinterpret_cast<float4*>(o + 2 * C) = pat_ch == 0 ? hi : lo;
        float* q = dst_pl + 4 * i;
        *reinterpret_cast<float4*>(q) = hi;
        *reinterpret_cast<float4*>(q + n) = pat_pl == 0 ? lo : hi;
        *reinterpret_cast<float4*>(q + 2 * n) = pat_pl == 0 ? hi : lo;
    }
}

}  // namespace davi

using namespace davi;

extern "C" int davi_split_tf32_dual(const float* src, float* dst_ch, float* dst_pl, int64_t rows, int C,
                                    int pattern_ch, int pattern_pl, davi_stream_t stream) {
    DAVI_REQUIRE(src && dst_ch && dst_pl, "null pointer");
    DAVI_REQUIRE(rows > 0 && C > 0, "empty tensor");
    DAVI_REQUIRE((pattern_ch == 0 || pattern_ch == 1) && (pattern_pl == 0 || pattern_pl == 1), "pattern");
    if (C % 4 == 0 && aligned16(src) && aligned16(dst_ch) && aligned16(dst_pl)) {
        const int64_t n4 = rows * C / 4;
        int64_t blocks = (n4 + kSplitThreads - 1) / kSplitThreads;
        if (blocks > 148 * 16) blocks = 148 * 16;
        split_tf32_dual_vec_kernel<<<(unsigned)blocks, kSplitThreads, 0, (cudaStream_t)stream>>>(
            src, dst_ch, dst_pl, n4, C / 4, pattern_ch, pattern_pl);
        return check_launch(__func__);
    }
    int rc = davi_split_tf32(src, dst_ch, rows, C, 0, pattern_ch, stream);      // odd channel counts: two launches
    if (rc) return rc;
    return davi_split_tf32(src, dst_pl, rows, C, 1, pattern_pl, stream);
}


extern "C" int davi_split_tf32(const float* src, float* dst, int64_t rows, int C, int layout, int pattern,
                               davi_stream_t stream) {
    DAVI_REQUIRE(src && dst, "null pointer");
    DAVI_REQUIRE(rows > 0 && C > 0, "empty tensor");
    DAVI_REQUIRE((layout == 0 || layout == 1) && (pattern == 0 || pattern == 1), "layout / pattern");
    const int64_t n = rows * C;
    const int64_t part_stride = layout == 0 ? C : n;
    const int64_t row_stride = layout == 0 ? 3 * (int64_t)C : C;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cap = 148 * 16;                      // grid-stride beyond 16 CTAs per SM
    if (C % 4 == 0 && aligned16(src) && aligned16(dst)) {
        const int64_t n4 = n / 4;
        int64_t blocks = (n4 + kSplitThreads - 1) / kSplitThreads;
        if (blocks > cap) blocks = cap;
        if (pattern == 0)
            split_tf32_vec_kernel<0><<<(unsigned)blocks, kSplitThreads, 0, st>>>(src, dst, n4, C / 4, part_stride, row_stride);
        else
            split_tf32_vec_kernel<1><<<(unsigned)blocks, kSplitThreads, 0, st>>>(src, dst, n4, C / 4, part_stride, row_stride);
    } else {
        int64_t blocks = (n + kSplitThreads - 1) / kSplitThreads;
        if (blocks > cap) blocks = cap;
        if (pattern == 0)
            split_tf32_scalar_kernel<0><<<(unsigned)blocks, kSplitThreads, 0, st>>>(src, dst, n, C, part_stride, row_stride);
        else
            split_tf32_scalar_kernel<1><<<(unsigned)blocks, kSplitThreads, 0, st>>>(src, dst, n, C, part_stride, row_stride);
    }
    return check_launch(__func__);
}
