// Row f1 (host-model glue): column sums of a [rows, C] channels-last tensor -- the bias gradient of
// every convolution (Keras Conv2D use_bias=True, layers/resnet.py:589-601: d bias = sum over batch and
// pixels of d out).  torch's generic reduction reaches ~0.5 TB/s on this shape (rows = 10 240,
// C = 256..316); this kernel streams the tensor once with 16-byte loads: a CTA owns a strip of rows,
// threads are laid out [row lanes][float4 columns], partial sums meet in shared memory and leave with
// per-CTA partial rows, which a second small kernel adds up (same-address atomics from ~600 CTAs cost
// ~20 us here: ncu showed the first version stalled in "drain" on its REDs).
#include "davi_common.cuh"

namespace davi {

constexpr int kCsThreads = 256;

__global__ void __launch_bounds__(kCsThreads)
colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int C4, int64_t x_s,
              int rows_per_cta) {
    extern __shared__ __align__(16) float cs_smem[];            // [RL][C4] float4
    const int RL = kCsThreads / C4;                              // row lanes (>= 1 since C4 <= 256)
    const int rl = threadIdx.x / C4, c = threadIdx.x - rl * C4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < RL) {
        int64_t r = r0 + rl;
        for (; r + 3 * RL < r1; r += 4 * RL) {                  // 4 independent loads in flight
            const float4 a0 = ld_stream4(x + r * x_s + 4 * c);
            const float4 a1 = ld_stream4(x + (r + RL) * x_s + 4 * c);
            const float4 a2 = ld_stream4(x + (r + 2 * RL) * x_s + 4 * c);
            const float4 a3 = ld_stream4(x + (r + 3 * RL) * x_s + 4 * c);
            acc.x += (a0.x + a1.x) + (a2.x + a3.x);
            acc.y += (a0.y + a1.y) + (a2.y + a3.y);
            acc.z += (a0.z + a1.z) + (a2.z + a3.z);
            acc.w += (a0.w + a1.w) + (a2.w + a3.w);
        }
        for (; r < r1; r += RL) {
            const float4 a0 = ld_stream4(x + r * x_s + 4 * c);
            acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
        }
        reinterpret_cast<float4*>(cs_smem)[rl * C4 + c] = acc;
    }
    __syncthreads();
    if (threadIdx.x < C4) {
        float4 t = reinterpret_cast<float4*>(cs_smem)[threadIdx.x];
        for (int k = 1; k < RL; ++k) {
            const float4 u = reinterpret_cast<float4*>(cs_smem)[k * C4 + threadIdx.x];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4*>(out)[(int64_t)blockIdx.x * C4 + threadIdx.x] = t;   // out = partials [ctas][C]
    }
}

// out[c] = sum_k partial[k][c]: 32 columns x 32 k-lanes per CTA (the partial rows are read
// coalesced, every thread has ~ctas/32 independent loads), reduced across the k-lanes in shared memory
__global__ void __launch_bounds__(1024)
colsum_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int ctas, int C) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float a = 0.f;
    if (c < C)
        for (int k = kl; k < ctas; k += 32) a += partial[(int64_t)k * C + c];
    red[kl][lane] = a;
    __syncthreads();
    if (kl == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][lane];
        out[c] = t;
    }
}

}  // namespace davi

using namespace davi;

static void colsum_plan(int64_t rows, int C, int* ctas, int* rows_per_cta) {
    const int RL = kCsThreads / (C / 4);
    int n = 2 * 148;
    int per = (int)((rows + n - 1) / n);
    if (per < 4 * RL) per = 4 * RL;
    *rows_per_cta = per;
    *ctas = (int)((rows + per - 1) / per);
}

extern "C" size_t davi_colsum_workspace_bytes(int64_t rows, int C) {
    if (rows <= 0 || C <= 0 || (C & 3) || C > 1024) return 0;
    int ctas, per;
    colsum_plan(rows, C, &ctas, &per);
    return (size_t)ctas * C * sizeof(float);
}

extern "C" int davi_colsum(const float* x, float* out, int64_t rows, int C, int64_t x_s, void* workspace,
                           size_t workspace_bytes, davi_stream_t stream) {
    DAVI_REQUIRE(x && out && rows > 0, "null pointer / no rows");
    DAVI_REQUIRE(C > 0 && (C & 3) == 0 && C <= 1024 && mult4(x_s) && x_s >= C && aligned16(x) && aligned16(out),
                 "need C % 4 == 0 (<= 1024), 16-byte aligned rows");
    if (!workspace || !aligned16(workspace) || workspace_bytes < davi_colsum_workspace_bytes(rows, C)) {
        set_error("%s: workspace too small or misaligned", __func__);
        return DAVI_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = C / 4;
    const int RL = kCsThreads / C4;
    int ctas, rows_per_cta;
    colsum_plan(rows, C, &ctas, &rows_per_cta);
    const size_t smem = (size_t)RL * C4 * sizeof(float4);
    colsum_kernel<<<ctas, kCsThreads, smem, st>>>(x, (float*)workspace, rows, C4, x_s, rows_per_cta);
    int rc = check_launch(__func__);
    if (rc) return rc;
    colsum_finish_kernel<<<(C + 31) / 32, 1024, 0, st>>>((const float*)workspace, out, ctas, C);
    return check_launch(__func__);
}
