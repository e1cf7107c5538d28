// R5-R8, R10: the ELBO reductions.
//   gauss_kl : GaussianDiag bounds + residual posterior + sample + KL (+ log q,
//              log p)  -- stat_tools/gaussian_diag.py:121-240,
//              model/variational_layer.py:511-542
//   dlm_nll  : DiscretizedLogisticMixture.create_params/log_prob fused with
//              its gradient -- stat_tools/discretized_logistic_mixture.py:
//              161-211,311-412
//   bernoulli_nll : stat_tools/bernoulli.py:155-179
//   is_logsumexp  : model/vae.py:480-482
// The reference runs each as dozens of elementwise TF kernels; here each is one
// launch over (pixel tile x image): the CTAs of one image form a THREAD-BLOCK CLUSTER, each
// reduces its tile and rank 0 adds the partial sums of its peers out of their shared memory
// (DSMEM) in rank order -- one launch, no workspace, no atomics, deterministic.
#include <cooperative_groups.h>

#include "davi_common.cuh"

namespace cg = cooperative_groups;

namespace davi {

constexpr int kMaxTiles = 8;     // portable cluster size

// Per-image totals of NV per-thread partial sums across the cluster of CTAs that share blockIdx.y.
// `red` >= 32 floats, `part` NV floats, both shared.  Results are valid in thread 0 of rank 0.
template <int NV>
__device__ __forceinline__ void image_reduce(float (&v)[NV], float* red, float* part) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float t = block_sum(v[i], red);
        if (threadIdx.x == 0) part[i] = t;
    }
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (cl.block_rank() == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = part[i];
            for (unsigned r = 1; r < cl.num_blocks(); ++r) t += *cl.map_shared_rank(part + i, r);
            v[i] = t;
        }
    }
    cl.sync();                   // peers keep their shared memory alive until rank 0 has read it
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_tiles(void (*kern)(KArgs...), int tiles, int B, int threads, cudaStream_t st,
                                Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles, (unsigned)B);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)tiles;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int tiles_for(int64_t items, int per_cta) {
    int64_t t = (items + per_cta - 1) / per_cta;
    int r = 1;
    while (r < kMaxTiles && r < t) r *= 2;
    return r;
}

constexpr float kHalfLog2Pi = 0.9189385332046727f;

struct Bound {
    float lo, hi;
    bool sym;  // |lo| == |hi| -> tanh form (gd.py:182), else smoothstep (gd.py:180)
};

__device__ __forceinline__ float bound_fwd(float x, const Bound& b, float* deriv) {
    if (b.sym) {
        const float m = fabsf(b.lo);
        const float t = tanhf(x / (2.f * m));
        if (deriv) *deriv = 0.5f * (1.f - t * t);
        return m * t;
    }
    const float w = b.hi - b.lo;
    const float t = (x - b.lo) / w;
    float s, ds;
    if (t < 0.f) { s = 0.f; ds = 0.f; }
    else if (t <= 1.f) { s = t * t * (3.f - 2.f * t); ds = 6.f * t * (1.f - t); }
    else { s = 1.f; ds = 0.f; }
    if (deriv) *deriv = ds;  // d(lo + w*S(t))/dx = S'(t)
    return b.lo + w * s;
}

struct GaussElem {
    float loc_q, sq, loc_p, sp;        // bounded parameters
    float dlocq_dmu, dsq_dls;          // d loc_q / d mu_q_raw, d sigma_q / d ls_q_raw
    float dlocp_dmu, dsp_dls;          // same for the prior
};

__device__ __forceinline__ GaussElem gauss_elem(float d_mu, float d_ls, bool has_prior, float mu_p,
                                                float ls_p, const Bound& bq, const Bound& bp) {
    GaussElem e;
    const float mu_q = has_prior ? mu_p + d_mu : d_mu;      // gd.py:214
    const float ls_q = has_prior ? d_ls + ls_p : d_ls;      // gd.py:219
    float t = tanhf(mu_q * 0.1f);
    e.loc_q = 5.f * t;                                      // bound(mu,-5,5) = 5 tanh(mu/10)
    e.dlocq_dmu = 0.5f * (1.f - t * t);
    float db;
    float ex = expf(bound_fwd(ls_q, bq, &db));
    e.sq = ex + 1e-2f;                                      // gd.py:185
    e.dsq_dls = ex * db;
    if (has_prior) {
        t = tanhf(mu_p * 0.1f);
        e.loc_p = 5.f * t;
        e.dlocp_dmu = 0.5f * (1.f - t * t);
        ex = expf(bound_fwd(ls_p, bp, &db));
        e.sp = ex + 1e-2f;
        e.dsp_dls = ex * db;
    } else {                                                // vl.py:535-537: N(0, I)
        e.loc_p = 0.f; e.sp = 1.f; e.dlocp_dmu = 0.f; e.dsp_dls = 0.f;
    }
    return e;
}

constexpr int kGaussThreads = 128;

// grad (nullable) [B,P,6z]: with it the forward also emits, per element, everything the gradient needs so that
// the backward sweep is one multiply-add pass without transcendentals (the KL's own upstream gradient is the
// known scalar beta / B_global; it is applied in the backward so any upstream works):
//   grad[row + 0z + c] = dKL / d mu_post_raw      grad[row + 1z + c] = dKL / d ls_post_raw
//   grad[row + 2z + c] = dKL / d mu_prior_raw     grad[row + 3z + c] = dKL / d ls_prior_raw   (both paths: p and q)
//   grad[row + 4z + c] = d xi / d mu_raw          grad[row + 5z + c] = d xi / d ls_raw         (sample path)
__global__ void __launch_bounds__(kGaussThreads)
gauss_kl_fwd_kernel(const float* __restrict__ post, const float* __restrict__ prior,
                    const float* __restrict__ eps, const float* __restrict__ noise_post,
                    const float* __restrict__ noise_prior, float* __restrict__ xi,
                    float* __restrict__ kl, float* __restrict__ logq, float* __restrict__ logp,
                    float* __restrict__ grad, int P, int z, Bound bq, Bound bp, float std_q, float std_p) {
    __shared__ float red[32];
    __shared__ float part[3];
    const int b = blockIdx.y;
    const int64_t ne = (int64_t)P * z;
    const float* po = post + (int64_t)b * ne * 2;
    const float* pr = prior ? prior + (int64_t)b * ne * 2 : nullptr;
    float acc[3] = {0.f, 0.f, 0.f};                                      // kl, log q, log p
    for (int64_t e = (int64_t)blockIdx.x * kGaussThreads + threadIdx.x; e < ne;
         e += (int64_t)gridDim.x * kGaussThreads) {
        const int64_t p = e / z;
        const int c = (int)(e - p * z);
        const int64_t row = p * 2 * z;
        const int64_t ge = (int64_t)b * ne + e;
        float d_mu = po[row + c], d_ls = po[row + z + c];
        if (noise_post) d_ls += std_q * noise_post[ge];                  // gd.py:165-166
        float mu_p = 0.f, ls_p = 0.f;
        if (pr) {
            mu_p = pr[row + c];
            ls_p = pr[row + z + c];
            if (noise_prior) ls_p += std_p * noise_prior[ge];
        }
        const GaussElem g = gauss_elem(d_mu, d_ls, pr != nullptr, mu_p, ls_p, bq, bp);
        const float ep = eps[ge];
        const float x = fmaf(g.sq, ep, g.loc_q);                         // gd.py:240
        xi[ge] = x;
        const float lsq = logf(g.sq), lsp = logf(g.sp);
        const float dm = (g.loc_p - g.loc_q) / g.sp, rs = g.sq / g.sp;
        acc[0] += lsp - lsq + 0.5f * (rs * rs + dm * dm - 1.f);          // vl.py:540
        if (logq) {
            acc[1] += -0.5f * ep * ep - lsq - kHalfLog2Pi;               // vl.py:515
            const float zp = (x - g.loc_p) / g.sp;
            acc[2] += -0.5f * zp * zp - lsp - kHalfLog2Pi;               // vl.py:526
        }
        if (grad) {
            float* gr = grad + ((int64_t)b * P + p) * 6 * z;
            const float isp2 = 1.f / (g.sp * g.sp);
            const float diff = g.loc_q - g.loc_p;
            const float kq_mu = diff * isp2 * g.dlocq_dmu;
            const float kq_ls = (g.sq * isp2 - 1.f / g.sq) * g.dsq_dls;
            gr[c] = kq_mu;
            gr[z + c] = kq_ls;
            gr[2 * z + c] = -diff * isp2 * g.dlocp_dmu + kq_mu;          // raw prior feeds p AND q (gd.py:214-221)
            gr[3 * z + c] = (1.f / g.sp - (g.sq * g.sq + diff * diff) * isp2 / g.sp) * g.dsp_dls + kq_ls;
            gr[4 * z + c] = g.dlocq_dmu;
            gr[5 * z + c] = ep * g.dsq_dls;
        }
    }
    image_reduce<3>(acc, red, part);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        kl[b] = acc[0];
        if (logq) { logq[b] = acc[1]; logp[b] = acc[2]; }
    }
}

// backward from the forward's saved `grad`: dpost = dkl[b] * dKL/dpost + dxi * dxi/dpost (same for the raw prior)
__global__ void __launch_bounds__(256)
gauss_kl_bwd_saved_kernel(const float* __restrict__ grad, const float* __restrict__ dxi,
                          const float* __restrict__ dkl, float* __restrict__ dpost,
                          float* __restrict__ dprior, int64_t total, int P, int z) {
    const int64_t ge = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ge >= total) return;
    const int64_t pix = ge / z;
    const int c = (int)(ge - pix * z);
    const int64_t b = pix / P;
    const float* gr = grad + pix * 6 * z;
    const float gx = dxi ? dxi[ge] : 0.f;
    const float gk = dkl ? dkl[b] : 0.f;
    const float sm = gx * gr[4 * z + c], sl = gx * gr[5 * z + c];
    const int64_t row = pix * 2 * z;
    dpost[row + c] = fmaf(gk, gr[c], sm);
    dpost[row + z + c] = fmaf(gk, gr[z + c], sl);
    if (dprior) {
        dprior[row + c] = fmaf(gk, gr[2 * z + c], sm);
        dprior[row + z + c] = fmaf(gk, gr[3 * z + c], sl);
    }
}

__global__ void __launch_bounds__(256)
gauss_kl_bwd_kernel(const float* __restrict__ post, const float* __restrict__ prior,
                    const float* __restrict__ eps, const float* __restrict__ noise_post,
                    const float* __restrict__ noise_prior, const float* __restrict__ dxi,
                    const float* __restrict__ dkl, float* __restrict__ dpost,
                    float* __restrict__ dprior, int64_t total, int P, int z, Bound bq, Bound bp,
                    float std_q, float std_p) {
    const int64_t ge = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ge >= total) return;
    const int64_t ne = (int64_t)P * z;
    const int64_t b = ge / ne;
    const int64_t e = ge - b * ne;
    const int64_t p = e / z;
    const int c = (int)(e - p * z);
    const int64_t row = (b * P + p) * 2 * z;
    float d_mu = post[row + c], d_ls = post[row + z + c];
    if (noise_post) d_ls += std_q * noise_post[ge];
    float mu_p = 0.f, ls_p = 0.f;
    if (prior) {
        mu_p = prior[row + c];
        ls_p = prior[row + z + c];
        if (noise_prior) ls_p += std_p * noise_prior[ge];
    }
    const GaussElem g = gauss_elem(d_mu, d_ls, prior != nullptr, mu_p, ls_p, bq, bp);
    const float gx = dxi ? dxi[ge] : 0.f;
    const float gk = dkl ? dkl[b] : 0.f;
    const float isp2 = 1.f / (g.sp * g.sp);
    const float diff = g.loc_q - g.loc_p;
    // SURVEY.md appendix C
    const float dlocq = gx + gk * diff * isp2;
    const float dsq = gx * eps[ge] + gk * (g.sq * isp2 - 1.f / g.sq);
    const float dmu_q = dlocq * g.dlocq_dmu;
    const float dls_q = dsq * g.dsq_dls;
    dpost[row + c] = dmu_q;
    dpost[row + z + c] = dls_q;
    if (prior) {
        const float dlocp = -gk * diff * isp2;
        const float dsp = gk * (1.f / g.sp - (g.sq * g.sq + diff * diff) * isp2 / g.sp);
        dprior[row + c] = dlocp * g.dlocp_dmu + dmu_q;        // raw prior feeds p AND q (gd.py:214-221)
        dprior[row + z + c] = dsp * g.dsp_dls + dls_q;
    }
}

// ---- discretized logistic mixture -----------------------------------------

constexpr int kMix = 5;
constexpr int kDlmCh = 10 * kMix;  // 5 logits + 5*(3 coeffs + 3 locs + 3 log-scales)
constexpr int kDlmThreads = 128;

// channel of loc[c][k] / log_scale[c][k] inside the 50-channel parameter row,
// obtained by replaying the reshapes of dlm.py:175-184:
//   rest = ch[5:50] viewed [mix m][9]; coeffs = rest[m][0:3];
//   rest[m][3:9] flattened over (m, 6) is re-viewed as [c][10] = loc[c][0:5] | ls[c][0:5]
__device__ __forceinline__ int dlm_ch(int flat) { return kMix + 9 * (flat / 6) + 3 + (flat % 6); }

__global__ void __launch_bounds__(kDlmThreads)
dlm_nll_kernel(const float* __restrict__ params, const float* __restrict__ x,
               float* __restrict__ nll, float* __restrict__ dparams, int P, float ls_low) {
    __shared__ float red[32];
    __shared__ float part[1];
    const int b = blockIdx.y;
    float acc = 0.f;
    for (int p = blockIdx.x * kDlmThreads + threadIdx.x; p < P; p += gridDim.x * kDlmThreads) {
        const int64_t pix = (int64_t)b * P + p;
        float pr[kDlmCh];
        {
            const float2* pp = reinterpret_cast<const float2*>(params + pix * kDlmCh);
#pragma unroll
            for (int i = 0; i < kDlmCh / 2; ++i) {
                const float2 t = __ldg(pp + i);
                pr[2 * i] = t.x;
                pr[2 * i + 1] = t.y;
            }
        }
        const float xv[3] = {x[pix * 3 + 0], x[pix * 3 + 1], x[pix * 3 + 2]};
        float gr[kDlmCh];

        // log-softmax of the mixture logits (dlm.py:398-412)
        float mx = pr[0];
#pragma unroll
        for (int k = 1; k < kMix; ++k) mx = fmaxf(mx, pr[k]);
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < kMix; ++k) se += expf(pr[k] - mx);
        const float lse_logits = mx + logf(se);

        float S[kMix];
        float dmean[3][kMix], dls[3][kMix], coef[3][kMix];
#pragma unroll
        for (int k = 0; k < kMix; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) coef[r][k] = tanhf(pr[kMix + 9 * k + r]);   // dlm.py:177
            float sum_lp = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float mean = pr[dlm_ch(c * 10 + k)];
                if (c == 1) mean += coef[0][k] * xv[0];                             // dlm.py:199
                if (c == 2) mean += coef[1][k] * xv[0] + coef[2][k] * xv[1];        // dlm.py:204
                const float ls_raw = pr[dlm_ch(c * 10 + 5 + k)];
                const float ls = fmaxf(ls_raw, ls_low);                             // dlm.py:211
                const float ls_pass = (ls_raw >= ls_low) ? 1.f : 0.f;               // tf.maximum grad
                const float inv = expf(-ls);
                const float u = xv[c] - mean;
                const float plus = inv * (u + 1.f / 255.f);
                const float mn = inv * (u - 1.f / 255.f);
                float lp, dm, dl;
                if (xv[c] < -0.999f) {                                              // dlm.py:365
                    lp = plus - softplusf(plus);
                    const float w = sigmoidf(-plus);
                    dm = -inv * w; dl = -plus * w;
                } else if (xv[c] > 0.99f) {
                    lp = -softplusf(mn);
                    const float w = -sigmoidf(mn);
                    dm = -inv * w; dl = -mn * w;
                } else {
                    const float cp = sigmoidf(plus), cm = sigmoidf(mn);
                    // sigma(plus) - sigma(mn) without cancellation:
                    //   = sigma(plus) * sigma(-mn) * (1 - exp(mn - plus))
                    const float delta = cp * sigmoidf(-mn) * (-expm1f(mn - plus));
                    if (delta > 1e-5f) {
                        lp = logf(fmaxf(delta, 1e-12f));
                        // (gp - gm) / delta == 1 - cp - cm exactly (g = c (1 - c))
                        const float gsum = cp * (1.f - cp) + cm * (1.f - cm);
                        const float hw = inv * (1.f / 255.f), mid = inv * u;
                        dm = -inv * (1.f - cp - cm);
                        dl = -mid * (1.f - cp - cm) - hw * gsum / delta;
                    } else {
                        const float mid = inv * u;
                        lp = mid - ls - 2.f * softplusf(mid) - 4.848116405963898f;   // log(127.5)
                        const float w = 1.f - 2.f * sigmoidf(mid);
                        dm = -inv * w; dl = -mid * w - 1.f;
                    }
                }
                sum_lp += lp;
                dmean[c][k] = dm;
                dls[c][k] = dl * ls_pass;
            }
            S[k] = sum_lp + pr[k] - lse_logits;                                     // dlm.py:373
        }
        float ms = S[0];
#pragma unroll
        for (int k = 1; k < kMix; ++k) ms = fmaxf(ms, S[k]);
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kMix; ++k) ss += expf(S[k] - ms);
        const float L = ms + logf(ss);                                              // dlm.py:387-395
        acc -= L;

        if (dparams) {
#pragma unroll
            for (int k = 0; k < kMix; ++k) {
                const float w = expf(S[k] - L);            // responsibility of mixture k
                const float pi = expf(pr[k] - lse_logits);
                gr[k] = pi - w;                            // d(-L)/d logit_k
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    gr[dlm_ch(c * 10 + k)] = -w * dmean[c][k];
                    gr[dlm_ch(c * 10 + 5 + k)] = -w * dls[c][k];
                }
                const float c0 = coef[0][k], c1 = coef[1][k], c2 = coef[2][k];
                gr[kMix + 9 * k + 0] = -w * dmean[1][k] * xv[0] * (1.f - c0 * c0);
                gr[kMix + 9 * k + 1] = -w * dmean[2][k] * xv[0] * (1.f - c1 * c1);
                gr[kMix + 9 * k + 2] = -w * dmean[2][k] * xv[1] * (1.f - c2 * c2);
            }
            float2* gp = reinterpret_cast<float2*>(dparams + pix * kDlmCh);
#pragma unroll
            for (int i = 0; i < kDlmCh / 2; ++i) gp[i] = make_float2(gr[2 * i], gr[2 * i + 1]);
        }
    }
    float tot[1] = {acc};
    image_reduce<1>(tot, red, part);
    if (blockIdx.x == 0 && threadIdx.x == 0) nll[b] = tot[0];
}

// ---- Bernoulli ------------------------------------------------------------

__global__ void __launch_bounds__(256)
bernoulli_nll_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                     float* __restrict__ nll, float* __restrict__ dlogits, int H, int W, int crop) {
    __shared__ float red[32];
    __shared__ float part[1];
    const int b = blockIdx.y;
    const int P = H * W;
    float acc = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const int h = p / W, w = p - h * W;
        const bool in = h >= crop && h < H - crop && w >= crop && w < W - crop;   // bern.py:177
        const int64_t i = (int64_t)b * P + p;
        const float l = logits[i], xv = x[i];
        if (in) acc += softplusf(l) - xv * l;                                     // -(x l - softplus(l))
        if (dlogits) dlogits[i] = in ? sigmoidf(l) - xv : 0.f;
    }
    float tot[1] = {acc};
    image_reduce<1>(tot, red, part);
    if (blockIdx.x == 0 && threadIdx.x == 0) nll[b] = tot[0];
}

__global__ void scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                  float* __restrict__ out, int64_t per_image, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = in[i] * w[i / per_image];
}

__global__ void is_logsumexp_kernel(const float* __restrict__ lw, float* __restrict__ out, int S, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float m = -INFINITY;
    for (int s = 0; s < S; ++s) m = fmaxf(m, lw[(int64_t)s * B + b]);
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += expf(lw[(int64_t)s * B + b] - m);
    out[b] = m + logf(acc);
}

static Bound make_bound(float lo, float hi) {
    Bound b; b.lo = lo; b.hi = hi; b.sym = fabsf(lo) == fabsf(hi);
    return b;
}

}  // namespace davi

using namespace davi;

extern "C" int davi_gauss_kl_fwd(const float* post, const float* prior, const float* eps,
                                 const float* noise_post, const float* noise_prior, float* xi,
                                 float* kl, float* logq, float* logp, int B, int P, int z,
                                 float post_lo, float post_hi, float prior_lo, float prior_hi,
                                 float noise_std_post, float noise_std_prior, float* grad,
                                 davi_stream_t stream) {
    DAVI_REQUIRE(post && eps && xi && kl, "null pointer");
    DAVI_REQUIRE(B > 0 && P > 0 && z > 0, "shape");
    DAVI_REQUIRE((logq == nullptr) == (logp == nullptr), "logq and logp go together");
    DAVI_REQUIRE(post_hi > post_lo && prior_hi > prior_lo, "bounds");
    const int tiles = tiles_for((int64_t)P * z, 4 * kGaussThreads);
    DAVI_CUDA(launch_tiles(gauss_kl_fwd_kernel, tiles, B, kGaussThreads, (cudaStream_t)stream,
                           post, prior, eps, noise_post, noise_prior, xi, kl, logq, logp, grad, P, z,
                           make_bound(post_lo, post_hi), make_bound(prior_lo, prior_hi), noise_std_post,
                           noise_std_prior));
    return check_launch(__func__);
}

extern "C" int davi_gauss_kl_bwd_saved(const float* grad, const float* dxi, const float* dkl, float* dpost,
                                       float* dprior, int B, int P, int z, davi_stream_t stream) {
    DAVI_REQUIRE(grad && dpost, "null pointer");
    DAVI_REQUIRE(B > 0 && P > 0 && z > 0, "shape");
    const int64_t total = (int64_t)B * P * z;
    gauss_kl_bwd_saved_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad, dxi, dkl, dpost, dprior, total, P, z);
    return check_launch(__func__);
}

extern "C" int davi_gauss_kl_bwd(const float* post, const float* prior, const float* eps,
                                 const float* noise_post, const float* noise_prior,
                                 const float* dxi, const float* dkl, float* dpost, float* dprior,
                                 int B, int P, int z, float post_lo, float post_hi, float prior_lo,
                                 float prior_hi, float noise_std_post, float noise_std_prior,
                                 davi_stream_t stream) {
    DAVI_REQUIRE(post && eps && dpost, "null pointer");
    DAVI_REQUIRE((prior == nullptr) == (dprior == nullptr), "dprior iff prior");
    DAVI_REQUIRE(B > 0 && P > 0 && z > 0, "shape");
    DAVI_REQUIRE(post_hi > post_lo && prior_hi > prior_lo, "bounds");
    const int64_t total = (int64_t)B * P * z;
    const int grid = (int)((total + 255) / 256);
    gauss_kl_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        post, prior, eps, noise_post, noise_prior, dxi, dkl, dpost, dprior, total, P, z,
        make_bound(post_lo, post_hi), make_bound(prior_lo, prior_hi), noise_std_post, noise_std_prior);
    return check_launch(__func__);
}

extern "C" int davi_dlm_nll_fwd(const float* params, const float* x, float* nll, float* dparams,
                                int B, int P, float log_scale_low_bound, davi_stream_t stream) {
    DAVI_REQUIRE(params && x && nll, "null pointer");
    DAVI_REQUIRE(B > 0 && P > 0, "shape");
    DAVI_REQUIRE((reinterpret_cast<uintptr_t>(params) & 7u) == 0 &&
                     (!dparams || (reinterpret_cast<uintptr_t>(dparams) & 7u) == 0), "alignment");
    DAVI_CUDA(launch_tiles(dlm_nll_kernel, tiles_for(P, kDlmThreads), B, kDlmThreads, (cudaStream_t)stream,
                           params, x, nll, dparams, P, log_scale_low_bound));
    return check_launch(__func__);
}

extern "C" int davi_bernoulli_nll_fwd(const float* logits, const float* x, float* nll,
                                      float* dlogits, int B, int H, int W, int crop,
                                      davi_stream_t stream) {
    DAVI_REQUIRE(logits && x && nll, "null pointer");
    DAVI_REQUIRE(B > 0 && H > 0 && W > 0 && crop >= 0 && 2 * crop < H && 2 * crop < W, "shape");
    DAVI_CUDA(launch_tiles(bernoulli_nll_kernel, tiles_for((int64_t)H * W, 256), B, 256, (cudaStream_t)stream,
                           logits, x, nll, dlogits, H, W, crop));
    return check_launch(__func__);
}

extern "C" int davi_scale_rows(const float* in, const float* w, float* out, int B, int64_t per_image,
                               davi_stream_t stream) {
    DAVI_REQUIRE(in && w && out && B > 0 && per_image > 0, "arguments");
    const int64_t total = (int64_t)B * per_image;
    scale_rows_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, w, out,
                                                                                    per_image, total);
    return check_launch(__func__);
}

extern "C" int davi_is_logsumexp(const float* lw, float* out, int S, int B, davi_stream_t stream) {
    DAVI_REQUIRE(lw && out && S > 0 && B > 0, "arguments");
    is_logsumexp_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lw, out, S, B);
    return check_launch(__func__);
}
