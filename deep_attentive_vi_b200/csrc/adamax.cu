// fused Adamax update with the Keras l2 regulariser gradient
// folded in -- replaces tf.keras.optimizers.Adamax (util/trainer.py:156-157,
// 226-231) plus the kernel/bias regulariser losses (layers/resnet.py:564-566).
// One pass over flat parameter / gradient / moment arrays (HBM-bound:
// 4 reads + 3 writes of 4 bytes per parameter).
#include "davi_common.cuh"

namespace davi {

__global__ void __launch_bounds__(256)
adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
              float* __restrict__ u, int64_t n4, int64_t n, float lr_t, float b1, float b2,
              float eps, float l2x2, float gscale, const float* __restrict__ hyper) {
    if (hyper) {   // device-resident step size: a captured CUDA graph replays with fresh values
        lr_t = __ldg(hyper);
        gscale = __ldg(hyper + 1);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 uv = reinterpret_cast<float4*>(u)[i];
#define DAVI_AX(c)                                                   \
    {                                                                \
        const float gg = fmaf(l2x2, pv.c, gscale * gv.c);            \
        mv.c = fmaf(b1, mv.c, (1.f - b1) * gg);                      \
        uv.c = fmaxf(b2 * uv.c, fabsf(gg));                          \
        pv.c -= lr_t * mv.c / (uv.c + eps);                          \
    }
        DAVI_AX(x) DAVI_AX(y) DAVI_AX(z) DAVI_AX(w)
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(u)[i] = uv;
    }
    // ragged tail (n % 4 elements)
    if (blockIdx.x == 0 && threadIdx.x < (n - 4 * n4)) {
        const int64_t i = 4 * n4 + threadIdx.x;
        const float gg = fmaf(l2x2, p[i], gscale * g[i]);
        m[i] = fmaf(b1, m[i], (1.f - b1) * gg);
        u[i] = fmaxf(b2 * u[i], fabsf(gg));
        p[i] -= lr_t * m[i] / (u[i] + eps);
    }
#undef DAVI_AX
}

}  // namespace davi

extern "C" int davi_adamax_step(float* param, const float* grad, float* m, float* u, int64_t n,
                                float lr, float beta1, float beta2, float eps, int64_t step,
                                float l2, float grad_scale, davi_stream_t stream) {
    using namespace davi;
    DAVI_REQUIRE(param && grad && m && u && n > 0 && step >= 1, "arguments");
    DAVI_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(u), "alignment");
    const float lr_t = lr / (1.f - powf(beta1, (float)step));   // Keras Adamax bias correction
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    adamax_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, u, n4, n, lr_t, beta1,
                                                                 beta2, eps, 2.f * l2, grad_scale, nullptr);
    return check_launch(__func__);
}

extern "C" int davi_adamax_step_dev(float* param, const float* grad, float* m, float* u, int64_t n,
                                    const float* hyper, float beta1, float beta2, float eps, float l2,
                                    davi_stream_t stream) {
    using namespace davi;
    DAVI_REQUIRE(param && grad && m && u && hyper && n > 0, "arguments");
    DAVI_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(u), "alignment");
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    adamax_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, u, n4, n, 0.f, beta1, beta2,
                                                                 eps, 2.f * l2, 1.f, hyper);
    return check_launch(__func__);
}
