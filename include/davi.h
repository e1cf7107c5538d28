/*
 * davi.h -- C ABI of the B200 (sm_100a) attention + ELBO hot path of the
 * Deep Attentive VAE (reference: ifiaposto/Deep_Attentive_VI, paths below are
 * relative to its src/ directory).
 *
 * The reference has no FFI: its seams are Python call signatures
 * (DepthWiseAttention.apply, NonlocalResNetBlock.call, GaussianDiag /
 * DiscretizedLogisticMixture / Bernoulli .log_prob, tfp kl_divergence).  Each
 * entry point below replaces the body behind one of those seams and cites it.
 * A TensorFlow custom-op shim (tf_ops/) or the in-tree ctypes binding
 * (deep_attentive_vi_b200/_lib.py) binds exactly these symbols; see
 * INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer to fp32 data, 16-byte aligned, owned by
 *     the caller; the library never allocates, frees or retains pointers;
 *   - tensors are channels-last (NHWC); "P" is the number of pixels H*W of one
 *     image, flattened row-major exactly like the reference's [B,W,H,.] tensors;
 *   - channel counts (d, C, z ...) and all strides are multiples of 4 floats
 *     where the argument says "stride"; strides are in FLOATS;
 *   - calls only enqueue work on `stream` (a cudaStream_t) on the calling
 *     thread's current device and return; no host synchronisation, no default
 *     stream use, no mutable global state (re-entrant across host threads);
 *   - return 0 on success or a negative DAVI_E* code; never abort, never throw.
 *     davi_last_error() returns the calling thread's last message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns DAVI_ECUDA.
 */
#ifndef DAVI_H_
#define DAVI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAVI_VERSION 200          /* major*100 + minor */
#define DAVI_MAX_SLOTS 32         /* max stacked layers n of the depth-wise attention */

#define DAVI_OK 0
#define DAVI_EINVAL (-1)          /* bad shape / stride / alignment / null pointer */
#define DAVI_EWORKSPACE (-2)      /* workspace too small */
#define DAVI_ECUDA (-3)           /* CUDA runtime error (text via davi_last_error) */

typedef void* davi_stream_t;      /* cudaStream_t */

int davi_version(void);
/* copies the calling thread's last error text (NUL terminated) into buf */
int davi_last_error(char* buf, size_t buf_bytes);

/* Implementation switches for tests and profiling (A/B of kernel variants; results are identical
 * within the stated tolerances).  Process-wide relaxed atomics read on every call: safe to set
 * from any thread, never read from the environment.  value 0 = default. */
#define DAVI_OPT_DW_IMPL 0        /* 1: register-streaming depth-wise kernels for every shape   */
#define DAVI_OPT_NONLOCAL_IMPL 1  /* 1: exact-fp32 CUDA-core nonlocal kernels instead of tcgen05 */
#define DAVI_OPT_NL_DS 2          /* 1: per-block dQ/dK variant also for N = 256                */
#define DAVI_OPT_NL_BWD 3         /* 1: nonlocal gradient as three launches instead of one      */
#define DAVI_OPT_CONV_DEBUG 4      /* profiling only (WRONG results): bits: 1 no lo arithmetic, 2 one MMA of three, 4 no stores, 8 256+rest column split, 16 / 32 no A / B loads */
#define DAVI_OPT_CONV_SM_RESERVE 5 /* SMs the convolution grids leave free (for concurrently running NCCL kernels)        */
#define DAVI_OPT_COUNT_ 8
int davi_set_option(int option, int value);
int davi_get_option(int option);

/* ------------------------------------------------------------------------
 * R1  Depth-wise attention, stacked tensors.
 * Replaces DepthWiseAttention.apply (layers/attention.py:135-164) together
 * with SoftmaxAlignment.call (layers/attention.py:98-116):
 *     s_j = scale * <q_p, K_{j,p}>,  a = softmax_j(s),  out_p = sum_j a_j V_{j,p}
 * q   [B,P,d]   pixel stride q_sp
 * K   [B,n,P,d] strides k_sb (batch), k_sn (layer), k_sp (pixel)
 * V   [B,n,P,C] strides v_sb, v_sn, v_sp
 * out [B,P,C]   pixel stride o_sp
 * attn (nullable) [B,P,n] dense: the softmax weights, for inspection/tests.
 * scale: DEVICE pointer to the learnable scalar of SoftmaxAlignment
 * (layers/attention.py:84-92), or NULL for 1.0 -- a device scalar so that no
 * host synchronisation is needed to read a trained value.
 * 1 <= n <= DAVI_MAX_SLOTS.  No 1/sqrt(d) factor (the reference has none).
 * ---------------------------------------------------------------------- */
int davi_dw_attn_fwd(const float* q, const float* keys, const float* values,
                     float* out, float* attn,
                     int B, int n, int P, int d, int C,
                     int64_t q_sp,
                     int64_t k_sb, int64_t k_sn, int64_t k_sp,
                     int64_t v_sb, int64_t v_sn, int64_t v_sp,
                     int64_t o_sp,
                     const float* scale, davi_stream_t stream);

/* Hand-written gradient of the above; the softmax is
 * recomputed from q and K.  dq/dkeys/dvalues use the SAME strides as their
 * forward tensors.  dscale (nullable): ONE float, overwritten with
 * sum_{pixels} sum_j ds_j * <q,K_j>. */
int davi_dw_attn_bwd(const float* q, const float* keys, const float* values,
                     const float* dout,
                     float* dq, float* dkeys, float* dvalues, float* dscale,
                     int B, int n, int P, int d, int C,
                     int64_t q_sp,
                     int64_t k_sb, int64_t k_sn, int64_t k_sp,
                     int64_t v_sb, int64_t v_sn, int64_t v_sp,
                     int64_t o_sp,
                     const float* scale, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R1/R2/R3  Depth-wise attention over a LIST of context slots (no tf.stack).
 * Replaces the stack -> split -> unstack/stack -> apply sequence of
 * model/vae.py:339-350 + model/variational_layer.py:357-360,396-406,472-499:
 * slot j contributes keys k_ptrs[j] [B,P,d] and values v_ptrs[j] [B,P,C] with its
 * OWN pixel strides k_sps[j] / v_sps[j] (floats; batch stride = P * pixel stride),
 * so a slot may be a channel slice of a wider per-layer tensor (values || key)
 * or a dense fresh tensor -- nothing is ever copied into a stacked layout.
 * k_ptrs / v_ptrs / k_sps / v_sps are HOST arrays of n entries (copied into the
 * launch arguments, not retained).
 * ---------------------------------------------------------------------- */
int davi_dw_attn_slots_fwd(const float* q, const float* const* k_ptrs,
                           const float* const* v_ptrs,
                           const int64_t* k_sps, const int64_t* v_sps,
                           float* out, float* attn,
                           int B, int n, int P, int d, int C,
                           int64_t q_sp, int64_t o_sp,
                           const float* scale, davi_stream_t stream);

/* Gradient.  dk_ptrs[j] / dv_ptrs[j] (pixel strides dk_sps[j] / dv_sps[j]) are
 * OVERWRITTEN, or read-modify-written (+=) when bit j of accumulate_mask is set:
 * one gradient buffer per context then collects the contributions of every
 * later layer directly (replaces TF's add_n over the L-1 consumers and the
 * zero-filled slice gradients).  A NULL entry skips that slot.
 * attn_out / ds_out (both or neither, nullable) [B,P,n] dense: the recomputed softmax
 * weights a_j and scale * ds_j.  With them the gradient of a context that MANY layers share
 * is not accumulated at all: its slot is passed as NULL here and davi_dw_attn_gather below
 * writes it once from the saved weights of all its consumers. */
int davi_dw_attn_slots_bwd(const float* q, const float* const* k_ptrs,
                           const float* const* v_ptrs,
                           const int64_t* k_sps, const int64_t* v_sps,
                           const float* dout,
                           float* dq, float* const* dk_ptrs, float* const* dv_ptrs,
                           const int64_t* dk_sps, const int64_t* dv_sps,
                           float* dscale,
                           int B, int n, int P, int d, int C,
                           int64_t q_sp, int64_t o_sp,
                           const float* scale, uint32_t accumulate_mask,
                           float* attn_out, float* ds_out, davi_stream_t stream);

/* Gradient gather of shared contexts (the lazy gather of SURVEY.md section 7 / 8f3; replaces TF's
 * add_n over the consumers of c[k] / e[j], model/vae.py:339-350).  Problem p (one context) has
 * m[p] consumers; consumer t of the concatenated tables (problem-major) contributes
 *     out_v[p][pix,:C] += wa_ptrs[t][pix * w_sps[t]] * v_ptrs[t][pix * v_sps[t] + :C]   (dOut of that layer)
 *     out_k[p][pix,:d] += wd_ptrs[t][pix * w_sps[t]] * k_ptrs[t][pix * k_sps[t] + :d]   (its query)
 * where wa / wd point at column j of that layer's attn_out / ds_out (j = the context's slot index
 * there, w_sps = its n).  Every output row is written exactly once (no read-modify-write, no
 * memset); out_k[p] may be NULL.  All tables are HOST arrays.  nprob <= 32, sum m <= 320. */
int davi_dw_attn_gather(int nprob, const int* m,
                        float* const* out_v, float* const* out_k,
                        const int64_t* out_v_sps, const int64_t* out_k_sps,
                        const float* const* v_ptrs, const int64_t* v_sps,
                        const float* const* k_ptrs, const int64_t* k_sps,
                        const float* const* wa_ptrs, const float* const* wd_ptrs,
                        const int64_t* w_sps,
                        int B, int P, int d, int C, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R2/R3 glue: LayerNormalization(epsilon) over the last axis, optionally with
 * the "x + gelu(LN(x))" residual of model/variational_layer.py:394,409,502,503
 * (tfa gelu, tanh form).  mode 0: y = LN(x) (vl.py:399, vae.py:347-348);
 * mode 1: y = x + gelu(LN(x)).  x,y [rows,C] with row strides x_s,y_s.
 * y may alias x.  gamma/beta [C].
 * ---------------------------------------------------------------------- */
int davi_ln_gelu_fwd(const float* x, const float* gamma, const float* beta, float* y,
                     int64_t rows, int C, int64_t x_s, int64_t y_s,
                     float eps, int mode, davi_stream_t stream);
/* dgamma/dbeta [C] are OVERWRITTEN.  workspace: davi_ln_gelu_bwd_workspace_bytes. */
size_t davi_ln_gelu_bwd_workspace_bytes(int64_t rows, int C);
int davi_ln_gelu_bwd(const float* x, const float* gamma, const float* beta,
                     const float* dy, float* dx, float* dgamma, float* dbeta,
                     int64_t rows, int C, int64_t x_s, int64_t dy_s, int64_t dx_s,
                     float eps, int mode, void* workspace, size_t workspace_bytes,
                     davi_stream_t stream);

/* The same glue with the gamma-residual of model/variational_layer.py:408-412 / 501-505 folded in:
 *     y = base + res_gamma * (x + gelu(LN(x)))
 * base, y [rows,C] (row strides base_s, y_s); res_gamma: DEVICE scalar (the zero-initialised
 * `gamma` / `enc_gamma` weights, vl.py:226-230,241-245).  The backward returns dx, the LN
 * parameter gradients and d res_gamma (ONE float, overwritten); d base == dy is left to the caller. */
int davi_ln_gelu_res_fwd(const float* x, const float* gamma, const float* beta, const float* base,
                         const float* res_gamma, float* y,
                         int64_t rows, int C, int64_t x_s, int64_t base_s, int64_t y_s,
                         float eps, davi_stream_t stream);
int davi_ln_gelu_res_bwd(const float* x, const float* gamma, const float* beta,
                         const float* res_gamma, const float* dy,
                         float* dx, float* dgamma, float* dbeta, float* dres_gamma,
                         int64_t rows, int C, int64_t x_s, int64_t dy_s, int64_t dx_s,
                         float eps, void* workspace, size_t workspace_bytes,
                         davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R4  Nonlocal spatial self-attention core.
 * Replaces NonlocalResNetBlock.call lines layers/attention.py:318-389 (after
 * the three 1x1 convs, which stay host-side):
 *     O = softmax_rows(scale * Q K^T) V,   y = gamma * O + x
 * Q,K [B,N,d] (row strides q_s,k_s), V [B,N,C] (v_s), x,y [B,N,C] (x_s,y_s);
 * batch strides are N*row_stride.  lse [B,N] dense: row log-sum-exp of the
 * scaled logits, saved for the backward.  o_save (nullable) [B,N,C] row stride
 * o_s: a copy of O, which the tensor-core backward needs (E_i = <dy_i, O_i>).
 * d % 4 == 0, C % 4 == 0.  N % 128 == 0, d <= 32, C <= 320 run on the tensor
 * cores (tcgen05, 3xTF32 error-compensated = fp32-grade results); other shapes
 * (N % 16 == 0, d <= 64, C <= 512) on an exact-fp32 CUDA-core kernel.
 * scale (nullable = 1.0) and gamma are DEVICE scalars: the trained weights
 * `scale` (layers/attention.py:84-92) and `gamma` (layers/attention.py:256-261).
 * ---------------------------------------------------------------------- */
int davi_nonlocal_attn_fwd(const float* Q, const float* K, const float* V,
                           const float* x, float* y, float* lse, float* o_save,
                           int B, int N, int d, int C,
                           int64_t q_s, int64_t k_s, int64_t v_s, int64_t x_s, int64_t y_s,
                           int64_t o_s,
                           const float* scale, const float* gamma, davi_stream_t stream);

/* Gradient: dy [B,N,C] -> dQ,dK,dV (same strides as
 * Q,K,V); dx == dy is left to the caller.  O (nullable): the o_save copy of the
 * forward; without it the CUDA-core kernels are used.  dgamma_dscale: TWO
 * floats, overwritten with { sum(O*dy), sum(dS * QK^T) }.
 * workspace: davi_nonlocal_attn_bwd_workspace_bytes. */
size_t davi_nonlocal_attn_bwd_workspace_bytes(int B, int N, int d, int C);
int davi_nonlocal_attn_bwd(const float* Q, const float* K, const float* V,
                           const float* lse, const float* O, const float* dy,
                           float* dQ, float* dK, float* dV, float* dgamma_dscale,
                           int B, int N, int d, int C,
                           int64_t q_s, int64_t k_s, int64_t v_s, int64_t o_s, int64_t dy_s,
                           const float* scale, const float* gamma,
                           void* workspace, size_t workspace_bytes, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R5/R6  Bounded residual Gaussians: sample + KL (+ log q, log p).
 * Replaces GaussianDiag.call/create_params/sample (stat_tools/gaussian_diag.py:
 * 121-240), tfp kl_divergence at model/variational_layer.py:533-542 and the
 * importance-sampling log-probs of vl.py:513-527.
 *   post  [B,P,2z] raw (d_mu || d_logsigma) from the posterior net
 *   prior [B,P,2z] raw (mu_p || logsigma_p) = prior.params(), or NULL for
 *         layer 0 (p = N(0,I), posterior not residual)
 *   eps   [B,P,z]  standard normal;  xi [B,P,z] = loc_q + sigma_q * eps
 *   noise_post / noise_prior (nullable) [B,P,z] standard normal, scaled by
 *         noise_std_* and added to the log-sigma halves (gd.py:165-166)
 *   kl [B]; logq, logp (nullable) [B]
 *   bounds: loc = 5 tanh(mu/10); sigma = exp(bound(ls, lo, hi)) + 1e-2 with
 *         tanh form when |lo| == |hi| else smoothstep (gd.py:175-185).
 *   grad  (nullable) [B,P,6z]: the fused forward+backward form (north_star (3)).  The KL's own
 *         upstream gradient is the known scalar beta / B_global, so the
 *         forward pass emits the whole gradient in the same sweep, per pixel row:
 *         [dKL/dmu_post | dKL/dls_post | dKL/dmu_prior | dKL/dls_prior | dxi/dmu | dxi/dls]
 *         (the prior entries include the residual path through q); davi_gauss_kl_bwd_saved then
 *         needs one multiply-add pass and no transcendental.
 * One launch over (pixel tile x image); the tiles of an image are a thread-block cluster whose
 * partial sums are combined through distributed shared memory (deterministic, no workspace).
 * ---------------------------------------------------------------------- */
int davi_gauss_kl_fwd(const float* post, const float* prior, const float* eps,
                      const float* noise_post, const float* noise_prior,
                      float* xi, float* kl, float* logq, float* logp,
                      int B, int P, int z,
                      float post_lo, float post_hi, float prior_lo, float prior_hi,
                      float noise_std_post, float noise_std_prior,
                      float* grad, davi_stream_t stream);

/* Gradient from the forward's `grad`: dpost = dkl[b] * dKL/dpost + dxi * dxi/dpost, likewise for
 * the raw prior (dprior NULL for layer 0).  dxi [B,P,z] / dkl [B] nullable = 0. */
int davi_gauss_kl_bwd_saved(const float* grad, const float* dxi, const float* dkl,
                            float* dpost, float* dprior, int B, int P, int z, davi_stream_t stream);

/* Gradient by recomputation (no `grad` saved): dxi [B,P,z] (nullable = 0) and dkl [B]
 * (nullable = 0) -> dpost [B,P,2z], dprior [B,P,2z] (NULL iff prior is NULL).  The raw prior
 * parameters receive gradient through BOTH p and the residual q. */
int davi_gauss_kl_bwd(const float* post, const float* prior, const float* eps,
                      const float* noise_post, const float* noise_prior,
                      const float* dxi, const float* dkl,
                      float* dpost, float* dprior,
                      int B, int P, int z,
                      float post_lo, float post_hi, float prior_lo, float prior_hi,
                      float noise_std_post, float noise_std_prior,
                      davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R7  Discretized-logistic-mixture negative log-likelihood, forward (+ fused
 * gradient).  Replaces DiscretizedLogisticMixture.create_params/log_prob
 * (stat_tools/discretized_logistic_mixture.py:161-211,311-412) and the negation
 * of model/data_layer.py:179-181.  5 mixtures, 3 channels (RGB autoregression),
 * params [B,P,50], x [B,P,3] in [-1,1]; nll [B] = -log p(x).
 * dparams (nullable) [B,P,50] = d nll[b] / d params (per image, UNWEIGHTED);
 * upstream weights are applied with davi_scale_rows.
 * ---------------------------------------------------------------------- */
int davi_dlm_nll_fwd(const float* params, const float* x, float* nll, float* dparams,
                     int B, int P, float log_scale_low_bound, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R8  Bernoulli negative log-likelihood with the border crop of
 * stat_tools/bernoulli.py:155-179: logits,x [B,H,W] (1 channel), pixels with
 * crop <= h < H-crop and crop <= w < W-crop contribute.  nll [B];
 * dlogits (nullable) [B,H,W] = sigmoid(l) - x inside the crop, 0 outside.
 * ---------------------------------------------------------------------- */
int davi_bernoulli_nll_fwd(const float* logits, const float* x, float* nll, float* dlogits,
                           int B, int H, int W, int crop, davi_stream_t stream);

/* out[b, :] = w[b] * in[b, :]  (per-image upstream weight for the fused
 * gradients above).  in/out [B,per_image], may alias. */
int davi_scale_rows(const float* in, const float* w, float* out,
                    int B, int64_t per_image, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * R10 Importance-sampled log-likelihood combine (model/vae.py:465-482,539):
 *   l[s,b] = sum_layers(logp - logq)[s*B+b] - nll[s*B+b]   (sample-major)
 *   out[b] = logsumexp_s l[s,b]
 * lw [S*B] is the per-sample log-weight already summed over layers.
 * ---------------------------------------------------------------------- */
int davi_is_logsumexp(const float* lw, float* out, int S, int B, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * 8(f4)  Fused Adamax step over flat arrays (replaces tf.keras.optimizers.Adamax
 * of util/trainer.py:156-157,226-231 with the Keras l2 regulariser gradient
 * 2*l2*w of layers/resnet.py:564-566 folded in):
 *   g' = grad_scale*g + 2*l2*p;  m = b1 m + (1-b1) g';  u = max(b2 u, |g'|);
 *   p -= lr / (1 - b1^step) * m / (u + eps).        step counts from 1.
 * ---------------------------------------------------------------------- */
int davi_adamax_step(float* param, const float* grad, float* m, float* u, int64_t n,
                     float lr, float beta1, float beta2, float eps, int64_t step,
                     float l2, float grad_scale, davi_stream_t stream);

/* Same update with the step size read from DEVICE memory, so that a captured
 * CUDA graph of the whole training step can be replayed while the learning-rate
 * schedule and the bias correction advance: hyper[0] = lr / (1 - beta1^step),
 * hyper[1] = grad_scale (two floats, written by the host before each replay). */
int davi_adamax_step_dev(float* param, const float* grad, float* m, float* u, int64_t n,
                         const float* hyper, float beta1, float beta2, float eps, float l2,
                         davi_stream_t stream);

/* ------------------------------------------------------------------------
 * 8(f1)  Weight normalisation of every convolution kernel of the model in one
 * launch.  Replaces WeightNorm._compute_weights (layers/weight_norm.py:126-135)
 *   kernel[o,:] = v[o,:] * exp(log_g[o]) * rsqrt(max(sum_k v[o,k]^2, 1e-12))
 * over the flat parameter array of util/trainer.py.  The four DEVICE tables
 * name each of the `rows` output-channel rows: offset of its v row and of its
 * log_g scalar inside params, offset of the row inside the effective-weight
 * array w, and its width (in floats).
 * ---------------------------------------------------------------------- */
int davi_weight_norm_fwd(const float* params, float* w,
                         const int64_t* row_v, const int64_t* row_g, const int64_t* row_w,
                         const int* row_cols, int64_t rows, davi_stream_t stream);
/* Gradient: dw (laid out like w) -> grads (laid out like params): the v rows and
 * the log_g scalars named by the tables are OVERWRITTEN. */
int davi_weight_norm_bwd(const float* params, const float* dw, float* grads,
                         const int64_t* row_v, const int64_t* row_g, const int64_t* row_w,
                         const int* row_cols, int64_t rows, davi_stream_t stream);

/* Column sums of a channels-last tensor x [rows, C] (row stride x_s): out[c] =
 * sum_r x[r,c], OVERWRITTEN.  The bias gradient of the Keras Conv2D layers
 * (layers/resnet.py:589-601, use_bias=True).  C % 4 == 0.  Two launches: per-CTA
 * partial rows into the workspace, then their sum (no same-address atomics). */
size_t davi_colsum_workspace_bytes(int64_t rows, int C);
int davi_colsum(const float* x, float* out, int64_t rows, int C, int64_t x_s,
                void* workspace, size_t workspace_bytes, davi_stream_t stream);

/* 8(f4)  dst[dst_offsets[i] : dst_offsets[i] + sizes[i]] = srcs[i][0 : sizes[i]] for `count`
 * device tensors, 64 tensors per launch.  srcs / dst_offsets / sizes are HOST arrays
 * (copied into the launch arguments).  Gathers the per-variable gradients into the flat
 * gradient array of the data-parallel step (reference: Keras' per-variable gradient
 * aggregation under MirroredStrategy, util/trainer.py:160-187). */
int davi_multi_copy(const float* const* srcs, const int64_t* dst_offsets, const int64_t* sizes,
                    int count, float* dst, davi_stream_t stream);

/* 8(f1)  Operand split for fp32-grade convolutions on the TF32 tensor cores ("3xTF32").
 * src [rows, C] contiguous -> the three-part operand hi = rna_tf32(v), lo = rna_tf32(v - hi):
 *   layout 0: dst [rows, 3C]   (parts concatenated along the channel axis)
 *   layout 1: dst [3, rows, C] (three planes: concatenation along the batch axis)
 *   pattern 0: parts (hi, lo, hi);  pattern 1: parts (hi, hi, lo)
 * A TF32 convolution of a pattern-0 operand with a pattern-1 operand over the tripled axis
 * accumulates hi*hi' + lo*hi' + hi*lo' in fp32: the reference's fp32 Conv2D
 * (layers/resnet.py:589-601) to ~2^-21 per product. */
int davi_split_tf32(const float* src, float* dst, int64_t rows, int C, int layout, int pattern,
                    davi_stream_t stream);
/* Both layouts from one read: dst_ch = layout 0 with pattern_ch, dst_pl = layout 1 with pattern_pl. */
int davi_split_tf32_dual(const float* src, float* dst_ch, float* dst_pl, int64_t rows, int C,
                         int pattern_ch, int pattern_pl, davi_stream_t stream);

/* ------------------------------------------------------------------------
 * 8(f1)  Dense convolution of the ResNet cells on the tensor cores.
 * Replaces the tf.keras.layers.Conv2D calls behind WeightNorm
 * (layers/resnet.py:589-601,613-633; layers/weight_norm.py:126-135) for
 * stride 1, padding 'same', odd kernel sizes:
 *     y[b,h,w,co] = bias[co] + sum_{r,s,ci} x[b, h+r-(kh-1)/2, w+s-(kw-1)/2, ci] * w[co,r,s,ci]
 * x [B,H,W,Cin] pixel stride x_sp; y [B,H,W,Cout] pixel stride y_sp;
 * w [Cout][kh][kw][Cin] dense (OHWI: the channels-last memory of a [Cout,Cin,kh,kw] kernel);
 * bias [Cout] or NULL.  W in {8,16,32}, H*W % 256 == 0, Cin % 4 == Cout % 4 == 0, Cin, Cout <= 384.
 * fp32-grade: implicit GEMM on tcgen05 (TF32 operands, fp32 accumulation in TMEM) with the
 * 3xTF32 error compensation computed in shared memory (csrc/conv_tc.cu).
 * workspace: davi_conv2d_workspace_bytes() bytes, ZERO-FILLED ONCE by the caller before its first
 * use and then left alone (the kernel keeps partial accumulators and hand-shake flags of its
 * stream-K split there and leaves the flags zero); one workspace per stream.
 * Input gradient:  gx = davi_conv2d_fwd(gy, wt, NULL) with wt from davi_conv2d_wt and Cin/Cout swapped.
 * Weight gradient: gw[co][r][s][ci] = sum_{b,h,w} gy[b,h,w,co] * x[b, h+r-(kh-1)/2, w+s-(kw-1)/2, ci],
 * OVERWRITTEN.  (The bias gradient is davi_colsum of gy.)
 * davi_conv2d_supported returns 1 when the shape is inside the kernel's envelope, else 0.
 * ---------------------------------------------------------------------- */
size_t davi_conv2d_workspace_bytes(void);
/* dst[0 : n] = 0 on `stream` (the one-time zero fill of the workspace for callers without a memset of their own) */
int davi_fill_zero(float* dst, int64_t n, davi_stream_t stream);
int davi_conv2d_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw);
/* Host-only: how a launch on a GPU with `sms` SMs would be laid out (no device needed; for tests and capacity planning).
 * mode 0 = forward / input gradient, 1 = weight gradient.  plan[0..11] = tiles, k-chunks per tile, CTA pairs, accumulator
 * columns, columns of MMA 0, of MMA 1, TMA ring stages, lo-ring slots, reduce mode (0/1), dynamic shared memory bytes,
 * workspace bytes needed, staged-tile epilogue available (0/1). */
int davi_conv2d_plan(int mode, int B, int H, int W, int Cin, int Cout, int kh, int kw, int sms, int64_t* plan);
int davi_conv2d_fwd(const float* x, const float* w, const float* bias, float* y,
                    int B, int H, int W, int Cin, int Cout, int kh, int kw,
                    int64_t x_sp, int64_t y_sp, void* workspace, size_t workspace_bytes,
                    davi_stream_t stream);
int davi_conv2d_wgrad(const float* x, const float* gy, float* gw,
                      int B, int H, int W, int Cin, int Cout, int kh, int kw,
                      int64_t x_sp, int64_t gy_sp, void* workspace, size_t workspace_bytes,
                      davi_stream_t stream);
/* The same kernels with the activation in front of the convolution inside a ResNet node fused in
 * (layers/resnet.py:624-627: conv(activation(t))):
 *   act = 1   the input goes through swish (x * sigmoid(x)) on its way into the tensor core: y = conv(swish(x)) + bias
 *             (forward), gw = wgrad(swish(x), gy) (weight gradient); swish(0) = 0 keeps the zero padding.  The input
 *             gradient of such a convolution is davi_conv2d_fwd(gy, wt) times swish'(x), applied by the caller. */
int davi_conv2d_fwd_ex(const float* x, const float* w, const float* bias, float* y,
                       int B, int H, int W, int Cin, int Cout, int kh, int kw,
                       int64_t x_sp, int64_t y_sp, int act,
                       void* workspace, size_t workspace_bytes, davi_stream_t stream);
int davi_conv2d_wgrad_ex(const float* x, const float* gy, float* gw,
                         int B, int H, int W, int Cin, int Cout, int kh, int kw,
                         int64_t x_sp, int64_t gy_sp, int act,
                         void* workspace, size_t workspace_bytes, davi_stream_t stream);
/* Weights that live across many calls (a trainer's parameters between two optimiser steps): the lo half of their
 * 3xTF32 split can be computed once -- dst[i] = lo(src[i]), v = trunc_tf32(v) + lo(v), lo rounded to tf32, the value the
 * kernel otherwise derives from every weight tile it stages -- and handed to the convolution, which then loads it by TMA
 * beside the weights themselves.  davi_conv2d_fwd_pre(x, w, w_lo = davi_tf32_lo(w), ...) == davi_conv2d_fwd_ex(x, w, ...)
 * bit for bit; w_lo has the layout of w.  (Input gradient: wt and davi_tf32_lo(wt).)  The kernel takes the image where
 * that is the faster path (up to 320 output channels: three ring slots) and derives it on chip otherwise. */
int davi_tf32_lo(const float* src, float* dst, int64_t n, davi_stream_t stream);
int davi_conv2d_fwd_pre(const float* x, const float* w, const float* w_lo, const float* bias, float* y,
                        int B, int H, int W, int Cin, int Cout, int kh, int kw,
                        int64_t x_sp, int64_t y_sp, int act,
                        void* workspace, size_t workspace_bytes, davi_stream_t stream);
/* wt[ci][taps-1-tap][co] = w[co][tap][ci] (taps = kh*kw): the weights of the input-gradient convolution */
int davi_conv2d_wt(const float* w, float* wt, int Cout, int taps, int Cin, davi_stream_t stream);
/* The same for `count` kernels in ONE launch (a trainer calls it once per step, after davi_weight_norm_fwd): kernel i
 * lives at w_flat + offs[i], its transpose goes to wt_flat + offs[i]; couts / taps / cins give its shape.  DEVICE
 * tables; tile_prefix[i] = sum over j < i of taps[j] * ceil(couts[j] / 32) * ceil(cins[j] / 32), total_tiles = the
 * full sum (the grid size). */
int davi_conv2d_wt_batched(const float* w_flat, float* wt_flat, const int64_t* offs, const int* couts,
                           const int* taps, const int* cins, const int64_t* tile_prefix, int count,
                           int64_t total_tiles, davi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DAVI_H_ */
