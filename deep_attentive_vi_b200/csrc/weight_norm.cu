// Row f1 (host-model glue): weight normalisation of EVERY convolution kernel of the model in one
// launch.  Replaces WeightNorm._compute_weights (layers/weight_norm.py:126-135):
//     kernel[o,:] = l2_normalize(v[o,:]) * exp(log_g[o]),   l2_normalize(x) = x * rsqrt(max(sum x^2, 1e-12))
// and its gradient.  The reference evaluates this per layer with ~6 small TF ops (and ~10 in the
// gradient); here all parameters live in one flat fp32 array (util/trainer.py), a device table names
// every output-channel row (offset of v, of log_g, of the row in the effective-weight array, width),
// and one warp normalises one row.  HBM-bound: reads v once (+ once more from L1/L2), writes w once.
#include "davi_common.cuh"

namespace davi {

constexpr int kWnThreads = 256;

__global__ void __launch_bounds__(kWnThreads)
weight_norm_fwd_kernel(const float* __restrict__ params, float* __restrict__ w,
                       const int64_t* __restrict__ row_v, const int64_t* __restrict__ row_g,
                       const int64_t* __restrict__ row_w, const int* __restrict__ row_cols, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (kWnThreads / 32) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* v = params + row_v[r];
    float* out = w + row_w[r];
    const int n = row_cols[r];
    float ss = 0.f;
    for (int i = lane; i < n; i += 32) { const float t = v[i]; ss = fmaf(t, t, ss); }
    ss = group_sum<32>(ss);
    const float c = expf(params[row_g[r]]) * rsqrtf(fmaxf(ss, 1e-12f));
    for (int i = lane; i < n; i += 32) out[i] = v[i] * c;
}

// dv = c * dw - (c * <dw, v> / ss) * v  (second term only where ss > 1e-12);  dlog_g = c * <dw, v>
__global__ void __launch_bounds__(kWnThreads)
weight_norm_bwd_kernel(const float* __restrict__ params, const float* __restrict__ dw, float* __restrict__ grads,
                       const int64_t* __restrict__ row_v, const int64_t* __restrict__ row_g,
                       const int64_t* __restrict__ row_w, const int* __restrict__ row_cols, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (kWnThreads / 32) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* v = params + row_v[r];
    const float* g = dw + row_w[r];
    float* dv = grads + row_v[r];
    const int n = row_cols[r];
    float ss = 0.f, dot = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float t = v[i];
        ss = fmaf(t, t, ss);
        dot = fmaf(g[i], t, dot);
    }
    ss = group_sum<32>(ss);
    dot = group_sum<32>(dot);
    const float c = expf(params[row_g[r]]) * rsqrtf(fmaxf(ss, 1e-12f));
    const float proj = ss > 1e-12f ? c * dot / ss : 0.f;
    for (int i = lane; i < n; i += 32) dv[i] = fmaf(c, g[i], -proj * v[i]);
    if (lane == 0) grads[row_g[r]] = c * dot;
}

}  // namespace davi

using namespace davi;

extern "C" int davi_weight_norm_fwd(const float* params, float* w, const int64_t* row_v, const int64_t* row_g,
                                    const int64_t* row_w, const int* row_cols, int64_t rows,
                                    davi_stream_t stream) {
    DAVI_REQUIRE(params && w && row_v && row_g && row_w && row_cols && rows > 0, "null pointer / no rows");
    const int64_t per = kWnThreads / 32;
    weight_norm_fwd_kernel<<<(unsigned)((rows + per - 1) / per), kWnThreads, 0, (cudaStream_t)stream>>>(
        params, w, row_v, row_g, row_w, row_cols, rows);
    return check_launch(__func__);
}

extern "C" int davi_weight_norm_bwd(const float* params, const float* dw, float* grads, const int64_t* row_v,
                                    const int64_t* row_g, const int64_t* row_w, const int* row_cols, int64_t rows,
                                    davi_stream_t stream) {
    DAVI_REQUIRE(params && dw && grads && row_v && row_g && row_w && row_cols && rows > 0, "null pointer / no rows");
    const int64_t per = kWnThreads / 32;
    weight_norm_bwd_kernel<<<(unsigned)((rows + per - 1) / per), kWnThreads, 0, (cudaStream_t)stream>>>(
        params, dw, grads, row_v, row_g, row_w, row_cols, rows);
    return check_launch(__func__);
}
