// R2/R3 glue: LayerNormalization(eps) over the channel axis, optionally fused
// with the "x + gelu(LN(x))" residual of model/variational_layer.py:394,409,
// 502,503 (mode 1); mode 0 is the plain LN of vl.py:399 and vae.py:347-348.
// One warp per row, the row lives in registers (float4 x ITERS per lane), one
// read + one write of the tensor (HBM-bound).  Backward: dx in one pass, the
// per-channel dgamma/dbeta are accumulated per CTA into a workspace and reduced
// by a second tiny kernel (deterministic, no atomics).
#include "davi_common.cuh"

namespace davi {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;
constexpr int kLnMaxGrid = 592;  // 4 CTAs x 148 SMs

template <int ITERS>
__global__ void __launch_bounds__(kLnThreads)
ln_gelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ y, int64_t rows, int C4,
                   int64_t x_s, int64_t y_s, float eps, int mode, const float* __restrict__ base,
                   int64_t b_s, const float* __restrict__ res_gamma) {
    const float rg = (mode == 2) ? __ldg(res_gamma) : 0.f;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
    const float invC = 1.f / (4.f * C4);

    float4 gm[ITERS], bt[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int c = lane + 32 * it;
        gm[it] = c < C4 ? ld4(gamma + 4 * c) : make_float4(0, 0, 0, 0);
        bt[it] = c < C4 ? ld4(beta + 4 * c) : make_float4(0, 0, 0, 0);
    }
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float* xp = x + row * x_s;
        float4 v[ITERS];
        float s = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            v[it] = c < C4 ? ld_stream4(xp + 4 * c) : make_float4(0, 0, 0, 0);
            s += v[it].x + v[it].y + v[it].z + v[it].w;
        }
        const float mean = group_sum<32>(s) * invC;
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            if (c < C4) {
                const float a = v[it].x - mean, b = v[it].y - mean, cc = v[it].z - mean,
                            d = v[it].w - mean;
                ss += a * a + b * b + cc * cc + d * d;
            }
        }
        const float rstd = rsqrtf(group_sum<32>(ss) * invC + eps);
        float* yp = y + row * y_s;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            if (c < C4) {
                float4 o;
                o.x = (v[it].x - mean) * rstd * gm[it].x + bt[it].x;
                o.y = (v[it].y - mean) * rstd * gm[it].y + bt[it].y;
                o.z = (v[it].z - mean) * rstd * gm[it].z + bt[it].z;
                o.w = (v[it].w - mean) * rstd * gm[it].w + bt[it].w;
                if (mode >= 1) {
                    o.x = v[it].x + gelu_tanh(o.x);
                    o.y = v[it].y + gelu_tanh(o.y);
                    o.z = v[it].z + gelu_tanh(o.z);
                    o.w = v[it].w + gelu_tanh(o.w);
                }
                if (mode == 2) {                       // y = base + res_gamma * (x + gelu(LN(x)))
                    const float4 bs = ld_stream4(base + row * b_s + 4 * c);
                    o.x = fmaf(rg, o.x, bs.x); o.y = fmaf(rg, o.y, bs.y);
                    o.z = fmaf(rg, o.z, bs.z); o.w = fmaf(rg, o.w, bs.w);
                }
                st_stream4(yp + 4 * c, o);
            }
        }
    }
}

// workspace layout: [gridDim.x][2C + 4] (dgamma partials, dbeta partials, d res_gamma partial + pad)
template <int ITERS>
__global__ void __launch_bounds__(kLnThreads)
ln_gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ dy,
                   float* __restrict__ dx, float* __restrict__ part, int64_t rows, int C4,
                   int64_t x_s, int64_t dy_s, int64_t dx_s, float eps, int mode,
                   const float* __restrict__ res_gamma) {
    extern __shared__ float4 s_acc[];  // [kLnWarps][2][C4] + [kLnWarps] floats
    const float rg = (mode == 2) ? __ldg(res_gamma) : 1.f;
    float dres = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + warp;
    const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
    const float invC = 1.f / (4.f * C4);

    float4 gm[ITERS], bt[ITERS], ag[ITERS], ab[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int c = lane + 32 * it;
        gm[it] = c < C4 ? ld4(gamma + 4 * c) : make_float4(0, 0, 0, 0);
        bt[it] = c < C4 ? ld4(beta + 4 * c) : make_float4(0, 0, 0, 0);
        ag[it] = make_float4(0, 0, 0, 0);
        ab[it] = make_float4(0, 0, 0, 0);
    }
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float* xp = x + row * x_s;
        const float* gp = dy + row * dy_s;
        float4 v[ITERS], g[ITERS];
        float s = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            v[it] = c < C4 ? ld_stream4(xp + 4 * c) : make_float4(0, 0, 0, 0);
            g[it] = c < C4 ? ld_stream4(gp + 4 * c) : make_float4(0, 0, 0, 0);
            s += v[it].x + v[it].y + v[it].z + v[it].w;
        }
        const float mean = group_sum<32>(s) * invC;
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            if (c < C4) {
                const float a = v[it].x - mean, b = v[it].y - mean, cc = v[it].z - mean,
                            d = v[it].w - mean;
                ss += a * a + b * b + cc * cc + d * d;
            }
        }
        const float rstd = rsqrtf(group_sum<32>(ss) * invC + eps);
        // xhat in v, dln (gradient w.r.t. the LN output) in g
        float s1 = 0.f, s2 = 0.f;
        float4 dxh[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            dxh[it] = make_float4(0, 0, 0, 0);
            if (c < C4) {
                float xh[4] = {(v[it].x - mean) * rstd, (v[it].y - mean) * rstd,
                               (v[it].z - mean) * rstd, (v[it].w - mean) * rstd};
                float gg[4] = {g[it].x, g[it].y, g[it].z, g[it].w};
                const float gmv[4] = {gm[it].x, gm[it].y, gm[it].z, gm[it].w};
                const float btv[4] = {bt[it].x, bt[it].y, bt[it].z, bt[it].w};
                float dh[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (mode == 2) {                   // y = base + rg * (x + gelu(ln)):  d rg, then dy <- rg * dy
                        const float xk = k == 0 ? v[it].x : k == 1 ? v[it].y : k == 2 ? v[it].z : v[it].w;
                        dres = fmaf(gg[k], xk + gelu_tanh(xh[k] * gmv[k] + btv[k]), dres);
                        gg[k] *= rg;
                    }
                    if (mode >= 1) gg[k] *= gelu_tanh_grad(xh[k] * gmv[k] + btv[k]);
                    dh[k] = gg[k] * gmv[k];
                    s1 += dh[k];
                    s2 += dh[k] * xh[k];
                }
                ag[it].x += gg[0] * xh[0]; ag[it].y += gg[1] * xh[1];
                ag[it].z += gg[2] * xh[2]; ag[it].w += gg[3] * xh[3];
                ab[it].x += gg[0]; ab[it].y += gg[1]; ab[it].z += gg[2]; ab[it].w += gg[3];
                dxh[it] = make_float4(dh[0], dh[1], dh[2], dh[3]);
                v[it] = make_float4(xh[0], xh[1], xh[2], xh[3]);
            }
        }
        s1 = group_sum<32>(s1) * invC;
        s2 = group_sum<32>(s2) * invC;
        float* dp = dx + row * dx_s;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int c = lane + 32 * it;
            if (c < C4) {
                float4 o;
                o.x = rstd * (dxh[it].x - s1 - v[it].x * s2);
                o.y = rstd * (dxh[it].y - s1 - v[it].y * s2);
                o.z = rstd * (dxh[it].z - s1 - v[it].z * s2);
                o.w = rstd * (dxh[it].w - s1 - v[it].w * s2);
                if (mode >= 1) {                       // the identity branch of x + gelu(LN(x))
                    o.x = fmaf(rg, g[it].x, o.x); o.y = fmaf(rg, g[it].y, o.y);
                    o.z = fmaf(rg, g[it].z, o.z); o.w = fmaf(rg, g[it].w, o.w);
                }
                st_stream4(dp + 4 * c, o);
            }
        }
    }
    // per-CTA reduction of the parameter gradients
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int c = lane + 32 * it;
        if (c < C4) {
            s_acc[(warp * 2 + 0) * C4 + c] = ag[it];
            s_acc[(warp * 2 + 1) * C4 + c] = ab[it];
        }
    }
    float* s_res = reinterpret_cast<float*>(s_acc + kLnWarps * 2 * C4);
    dres = group_sum<32>(dres);
    if (lane == 0) s_res[warp] = dres;
    __syncthreads();
    float4* prow = reinterpret_cast<float4*>(part) + (int64_t)blockIdx.x * (2 * C4 + 1);
    for (int i = threadIdx.x; i < 2 * C4; i += kLnThreads) {
        float4 t = make_float4(0, 0, 0, 0);
        for (int w = 0; w < kLnWarps; ++w) {
            const float4 u = s_acc[w * 2 * C4 + i];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        prow[i] = t;
    }
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kLnWarps; ++w) t += s_res[w];
        prow[2 * C4] = make_float4(t, 0.f, 0.f, 0.f);
    }
}

// [nblk][2C] partials -> dgamma | dbeta: 32 columns x 32 partial-lanes per CTA (coalesced reads, ~nblk/32
// independent loads per thread instead of nblk dependent ones), combined in shared memory
__global__ void __launch_bounds__(1024)
ln_param_reduce_kernel(const float* __restrict__ part, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dres, int nblk, int C) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int W = 2 * C + 4;                            // floats per partial row
    float s = 0.f;
    if (i <= 2 * C)
        for (int b = kl; b < nblk; b += 32) s += part[(int64_t)b * W + i];
    red[kl][lane] = s;
    __syncthreads();
    if (kl == 0 && i <= 2 * C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][lane];
        if (i < C) dgamma[i] = t;
        else if (i < 2 * C) dbeta[i - C] = t;
        else if (dres) *dres = t;
    }
}

static int ln_grid(int64_t rows) {
    int64_t g = (rows + kLnWarps - 1) / kLnWarps;
    if (g > kLnMaxGrid) g = kLnMaxGrid;
    if (g < 1) g = 1;
    return (int)g;
}

static int ln_fwd_impl(const char* fn, const float* x, const float* gamma, const float* beta, float* y,
                       int64_t rows, int C, int64_t x_s, int64_t y_s, float eps, int mode, const float* base,
                       int64_t b_s, const float* res_gamma, cudaStream_t st) {
    if (!x || !gamma || !beta || !y || !aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y)) {
        set_error("%s: null or misaligned pointer", fn);
        return DAVI_EINVAL;
    }
    if (rows <= 0 || C <= 0 || (C & 3) || C > 512 || !mult4(x_s) || !mult4(y_s) || x_s < C || y_s < C) {
        set_error("%s: need rows>0, C%%4==0, C<=512, row strides %%4 and >= C", fn);
        return DAVI_EINVAL;
    }
    if (mode == 2 && (!base || !res_gamma || !aligned16(base) || !mult4(b_s) || b_s < C)) {
        set_error("%s: residual base / gamma missing or misaligned", fn);
        return DAVI_EINVAL;
    }
    const int C4 = C / 4, grid = ln_grid(rows);
#define LN_FWD(I) ln_gelu_fwd_kernel<I><<<grid, kLnThreads, 0, st>>>(x, gamma, beta, y, rows, C4, x_s, y_s, eps, mode, base, b_s, res_gamma)
    switch ((C4 + 31) / 32) {
        case 1: LN_FWD(1); break;
        case 2: LN_FWD(2); break;
        case 3: LN_FWD(3); break;
        default: LN_FWD(4); break;
    }
#undef LN_FWD
    return check_launch(fn);
}

static int ln_bwd_impl(const char* fn, const float* x, const float* gamma, const float* beta, const float* dy,
                       float* dx, float* dgamma, float* dbeta, float* dres, int64_t rows, int C, int64_t x_s,
                       int64_t dy_s, int64_t dx_s, float eps, int mode, const float* res_gamma, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
    if (!x || !gamma || !beta || !dy || !dx || !dgamma || !dbeta || !aligned16(x) || !aligned16(gamma) ||
        !aligned16(beta) || !aligned16(dy) || !aligned16(dx) || !aligned16(workspace)) {
        set_error("%s: null or misaligned pointer", fn);
        return DAVI_EINVAL;
    }
    if (rows <= 0 || C <= 0 || (C & 3) || C > 512 || !mult4(x_s) || !mult4(dy_s) || !mult4(dx_s) || x_s < C ||
        dy_s < C || dx_s < C || (mode == 2 && (!res_gamma || !dres))) {
        set_error("%s: bad shape / strides / residual arguments", fn);
        return DAVI_EINVAL;
    }
    if (!workspace || workspace_bytes < davi_ln_gelu_bwd_workspace_bytes(rows, C)) {
        set_error("%s: workspace too small", fn);
        return DAVI_EWORKSPACE;
    }
    const int C4 = C / 4, grid = ln_grid(rows);
    const size_t smem = (size_t)kLnWarps * 2 * C4 * sizeof(float4) + kLnWarps * sizeof(float);
    float* part = (float*)workspace;
#define LN_BWD(I) ln_gelu_bwd_kernel<I><<<grid, kLnThreads, smem, st>>>(x, gamma, beta, dy, dx, part, rows, C4, x_s, dy_s, dx_s, eps, mode, res_gamma)
    switch ((C4 + 31) / 32) {
        case 1: LN_BWD(1); break;
        case 2: LN_BWD(2); break;
        case 3: LN_BWD(3); break;
        default: LN_BWD(4); break;
    }
#undef LN_BWD
    int rc = check_launch(fn);
    if (rc) return rc;
    ln_param_reduce_kernel<<<(2 * C + 1 + 31) / 32, 1024, 0, st>>>(part, dgamma, dbeta, mode == 2 ? dres : nullptr,
                                                                  grid, C);
    return check_launch(fn);
}

}  // namespace davi

using namespace davi;

extern "C" size_t davi_ln_gelu_bwd_workspace_bytes(int64_t rows, int C) {
    (void)rows;
    return (size_t)kLnMaxGrid * (2 * (size_t)C + 4) * sizeof(float);
}

extern "C" int davi_ln_gelu_fwd(const float* x, const float* gamma, const float* beta, float* y,
                                int64_t rows, int C, int64_t x_s, int64_t y_s, float eps, int mode,
                                davi_stream_t stream) {
    DAVI_REQUIRE(mode == 0 || mode == 1, "mode");
    return ln_fwd_impl(__func__, x, gamma, beta, y, rows, C, x_s, y_s, eps, mode, nullptr, 0, nullptr,
                       (cudaStream_t)stream);
}

extern "C" int davi_ln_gelu_bwd(const float* x, const float* gamma, const float* beta,
                                const float* dy, float* dx, float* dgamma, float* dbeta,
                                int64_t rows, int C, int64_t x_s, int64_t dy_s, int64_t dx_s,
                                float eps, int mode, void* workspace, size_t workspace_bytes,
                                davi_stream_t stream) {
    DAVI_REQUIRE(mode == 0 || mode == 1, "mode");
    return ln_bwd_impl(__func__, x, gamma, beta, dy, dx, dgamma, dbeta, nullptr, rows, C, x_s, dy_s, dx_s, eps,
                       mode, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int davi_ln_gelu_res_fwd(const float* x, const float* gamma, const float* beta, const float* base,
                                    const float* res_gamma, float* y, int64_t rows, int C, int64_t x_s,
                                    int64_t base_s, int64_t y_s, float eps, davi_stream_t stream) {
    return ln_fwd_impl(__func__, x, gamma, beta, y, rows, C, x_s, y_s, eps, 2, base, base_s, res_gamma,
                       (cudaStream_t)stream);
}

extern "C" int davi_ln_gelu_res_bwd(const float* x, const float* gamma, const float* beta,
                                    const float* res_gamma, const float* dy, float* dx, float* dgamma,
                                    float* dbeta, float* dres_gamma, int64_t rows, int C, int64_t x_s,
                                    int64_t dy_s, int64_t dx_s, float eps, void* workspace,
                                    size_t workspace_bytes, davi_stream_t stream) {
    return ln_bwd_impl(__func__, x, gamma, beta, dy, dx, dgamma, dbeta, dres_gamma, rows, C, x_s, dy_s, dx_s, eps,
                       2, res_gamma, workspace, workspace_bytes, (cudaStream_t)stream);
}
