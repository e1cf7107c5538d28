"""Op-for-op CPU restatement of the reference hot-path functions (torch CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY PINNED: the reference ships no
golden vectors, so the pins are fixtures produced by executing the reference's own source files on
a float64 TensorFlow stand-in (tests/golden/make_golden.py -> tests/golden/*.npz); tests/test_golden.py
holds every function below to them at 1e-10.

Every function follows the cited reference lines literally, including the
transposes / stacks / reshapes the reference materialises, so that (a) autograd
through it reproduces what TF autodiff computes for the reference and (b) timing
it on the host cores is a fair stand-in for the reference's unfused CPU path.
Citations are ``path:line`` under /root/reference/src.

Tensors are NHWC like the reference.  dtype follows the inputs (tests use
float64 for the checker and float32 for the CPU-baseline timing).
"""
import math

import torch

LOG_2PI = math.log(2.0 * math.pi)


# ---------------------------------------------------------------------------
# library semantics the reference relies on
# ---------------------------------------------------------------------------

def layer_norm(x, gamma, beta, eps=1e-5):
    """tf.keras.layers.LayerNormalization(epsilon=1e-5): last axis, biased
    variance, non-fused ``moments`` path (vl.py:214-224, vae.py:205-207)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def gelu(x):
    """tfa.activations.gelu default (approximate=True, tanh form); call sites
    vl.py:394,409,502,503."""
    c = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1.0 + torch.tanh(c * (x + 0.044715 * x ** 3)))


def softplus(x):
    """tf.nn.softplus = log(1 + exp(x)), evaluated stably."""
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-x.abs()))


# ---------------------------------------------------------------------------
# R1: depth-wise attention
# ---------------------------------------------------------------------------

def softmax_alignment(proj_query, proj_key, scale=None):
    """SoftmaxAlignment.call, layers/attention.py:98-116.

    scores = matmul(proj_query^T, proj_key) [* scale]; softmax over last axis."""
    scores = torch.matmul(proj_query.transpose(-1, -2), proj_key)  # att.py:105
    if scale is not None:
        scores = scores * scale                                   # att.py:107-108
    return torch.softmax(scores, dim=-1)                          # att.py:116


def depthwise_attention_apply(query, keys, values, scale=None):
    """DepthWiseAttention.apply, layers/attention.py:135-164.

    query [B,W,H,d]; keys [B,n,W,H,d]; values [B,n,W,H,C] -> [B,W,H,C]."""
    q = query.unsqueeze(-1)                                       # att.py:142
    k = keys.permute(0, 2, 3, 4, 1).contiguous()                  # att.py:145 (materialised)
    a = softmax_alignment(q, k, scale)                            # att.py:149 -> [B,W,H,1,n]
    a = a[:, :, :, 0, :]                                          # att.py:151
    v = values.permute(0, 2, 3, 4, 1).contiguous()                # att.py:155 (materialised)
    out = torch.matmul(v, a.unsqueeze(-1)).squeeze(-1)            # att.py:161 (matvec)
    return out.reshape(-1, v.shape[1], v.shape[2], v.shape[3])    # att.py:162


# R2: decoder-side glue, model/variational_layer.py:390-412
def decoder_attention_glue(new_context, query, keys, values, ln_ctx, ln_attn, ln_attn2,
                           gamma, use_layer_norm=True, scale=None):
    """Returns (new_context_after_ln, prior_condition).

    ln_* are (gamma, beta) pairs: ln_ctx=_ctx_layer_norm_layer (vl.py:222),
    ln_attn=_attention_layer_norm_layer (vl.py:214), ln_attn2 (vl.py:218)."""
    if use_layer_norm:
        new_context = new_context + gelu(layer_norm(new_context, *ln_ctx))       # vl.py:393-394
    vals = list(values.unbind(dim=1))                                            # vl.py:396
    if use_layer_norm:
        vals[-1] = layer_norm(vals[-1], *ln_attn)                                # vl.py:398-399
    vals = torch.stack(vals, dim=1)                                              # vl.py:401
    att = depthwise_attention_apply(query, keys, vals, scale)                    # vl.py:403-406
    if use_layer_norm:
        att = att + gelu(layer_norm(att, *ln_attn2))                             # vl.py:408-409
    prior_condition = new_context + gamma * att                                  # vl.py:412
    return new_context, prior_condition


# R3: encoder-side glue, model/variational_layer.py:480-507
def encoder_attention_glue(condition, new_key, enc_query, keys, values, ln_att, ln_cond,
                           gamma_enc, use_layer_norm=True, replace_first=True, scale=None):
    """condition [B,W,H,C] (already split from new_key when replace_first)."""
    if replace_first:
        vals = list(values.unbind(dim=1)); vals[0] = condition                   # vl.py:486-488
        values = torch.stack(vals, dim=1)
        ks = list(keys.unbind(dim=1)); ks[0] = new_key                           # vl.py:492-494
        keys = torch.stack(ks, dim=1)
    att = depthwise_attention_apply(enc_query, keys, values, scale)              # vl.py:496-499
    if use_layer_norm:
        att = att + gelu(layer_norm(att, *ln_att))                               # vl.py:502
        condition = condition + gelu(layer_norm(condition, *ln_cond))            # vl.py:503
    condition = condition + gamma_enc * att                                      # vl.py:505
    return torch.cat([condition, keys[:, 0]], dim=-1)                            # vl.py:507


# ---------------------------------------------------------------------------
# R4: nonlocal spatial self-attention core
# ---------------------------------------------------------------------------

def nonlocal_core(proj_query, proj_key, proj_value, inputs, gamma, scale=None):
    """NonlocalResNetBlock.call core, layers/attention.py:318-389
    (use_layer_norm=False, subsample=False).

    proj_query/proj_key [B,W,H,kd], proj_value/inputs [B,W,H,C] (outputs of the
    1x1 convs att.py:310,346,362)."""
    B, W, H, kd = proj_query.shape
    C = proj_value.shape[-1]
    q = proj_query.permute(0, 3, 1, 2).reshape(B, kd, W * H)       # att.py:318-321
    k = proj_key.permute(0, 3, 1, 2).reshape(B, kd, W * H)         # att.py:350-356
    a = softmax_alignment(q, k, scale)                             # att.py:358 -> [B,Nq,Nk]
    v = proj_value.permute(0, 3, 1, 2).reshape(B, C, W * H)        # att.py:366-369
    o = torch.matmul(v, a.transpose(-1, -2))                       # att.py:371 -> [B,C,Nq]
    o = o.reshape(B, C, W, H).permute(0, 2, 3, 1)                  # att.py:378-382
    return gamma * o + inputs                                      # att.py:389


# ---------------------------------------------------------------------------
# f1: the dense convolution of the ResNet cells
# ---------------------------------------------------------------------------

def conv2d_same(x, kernel, bias=None):
    """tf.keras.layers.Conv2D(kernel_size=k, strides=1, padding='same') as the cells build it
    (layers/resnet.py:589-601 behind WeightNorm, layers/weight_norm.py:126-135), written out tap by tap
    (library semantics: cross-correlation, zero padding (k-1)//2 on each side for odd k).

    x [B,H,W,Cin] NHWC, kernel [kh,kw,Cin,Cout] (the Keras HWIO layout), bias [Cout] -> [B,H,W,Cout]."""
    B, H, W, _ = x.shape
    kh, kw, _, cout = kernel.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = torch.nn.functional.pad(x, (0, 0, pl, kw - 1 - pl, pt, kh - 1 - pt))
    y = x.new_zeros(B, H, W, cout)
    for r in range(kh):
        for s in range(kw):
            y = y + torch.einsum("bhwi,io->bhwo", xp[:, r:r + H, s:s + W, :], kernel[r, s])
    return y if bias is None else y + bias


# ---------------------------------------------------------------------------
# R5/R6: bounded residual Gaussians, sample, KL, log-probs
# ---------------------------------------------------------------------------

def bound_param(x, mi, mx):
    """GaussianDiag.call.bound_param, stat_tools/gaussian_diag.py:175-182."""
    if abs(mi) != abs(mx):
        t = (x - mi) / (mx - mi)
        s = torch.where(t < 0.0, torch.zeros_like(t),
                        torch.where(t <= 1.0, 3 * t ** 2 - 2 * t ** 3, torch.ones_like(t)))
        return mi + (mx - mi) * s                                  # gd.py:180
    return abs(mi) * torch.tanh(x / (2 * abs(mi)))                 # gd.py:182


def gaussian_dist(mu_raw, ls_raw, lo, hi):
    """(loc, scale_diag) of gd.py:185."""
    return bound_param(mu_raw, -5.0, 5.0), torch.exp(bound_param(ls_raw, lo, hi)) + 1e-2


def kl_mvn_diag(loc_q, sc_q, loc_p, sc_p):
    """tfp.distributions.kl_divergence(MVNDiag q, MVNDiag p) (tfp 0.11
    mvn_linear_operator._kl_brute_force), event = last axis."""
    return (torch.log(sc_p) - torch.log(sc_q)
            + 0.5 * (-1.0 + (sc_q / sc_p) ** 2 + ((loc_p - loc_q) / sc_p) ** 2)).sum(dim=-1)


def mvn_diag_log_prob(x, loc, sc):
    """MultivariateNormalDiag.log_prob, event = last axis."""
    z = (x - loc) / sc
    return (-0.5 * z * z - torch.log(sc) - 0.5 * LOG_2PI).sum(dim=-1)


def variational_sample_kl(post_delta, eps, prior_params, post_bounds, prior_bounds,
                          compute_logp=False):
    """Posterior sample + KL of one top-down layer.

    post_delta  [B,W,H,2z] = output of the posterior param net (noise already
                added to its log-sigma half, gd.py:165-166)
    prior_params [B,W,H,2z] or None = GaussianDiag.params() of the prior
                (raw mu || raw log-sigma incl. its noise, gd.py:267-269,
                vl.py:466-467); None = layer 0, p = N(0, I) (vl.py:535-537)
    eps         [B,W,H,z] standard normal (TFP draws it inside sample()).

    Follows gd.py:146-155,198-221 (residual on RAW params), gd.py:185 (bounds
    of the *owning* distribution), gd.py:240 (sample), vl.py:533-542 (KL,
    summed over W,H), vl.py:513-527 + gd.py:263 (log q, log p)."""
    z = eps.shape[-1]
    d_mu, d_ls = post_delta[..., :z], post_delta[..., z:]          # gd.py:160
    if prior_params is not None:
        mu_p, ls_p = prior_params[..., :z], prior_params[..., z:]  # gd.py:149
        mu_q, ls_q = mu_p + d_mu, d_ls + ls_p                      # gd.py:214-219
        loc_p, sc_p = gaussian_dist(mu_p, ls_p, *prior_bounds)     # prior._p[0]
    else:
        mu_q, ls_q = d_mu, d_ls                                    # gd.py:208-211
        loc_p, sc_p = torch.zeros_like(d_mu), torch.ones_like(d_ls)
    loc_q, sc_q = gaussian_dist(mu_q, ls_q, *post_bounds)          # posterior._p[-1]
    xi = loc_q + sc_q * eps                                        # gd.py:240
    kl = kl_mvn_diag(loc_q, sc_q, loc_p, sc_p).sum(dim=(1, 2))     # vl.py:540-542
    if not compute_logp:
        return xi, kl, None, None
    q_logp = mvn_diag_log_prob(xi, loc_q, sc_q).sum(dim=(1, 2))    # vl.py:515, gd.py:263
    p_logp = mvn_diag_log_prob(xi, loc_p, sc_p).sum(dim=(1, 2))    # vl.py:522-527
    return xi, kl, q_logp, p_logp


# ---------------------------------------------------------------------------
# R7: discretized logistic mixture
# ---------------------------------------------------------------------------

def dlm_create_params(params, x, num_mix=5, ls_low=-7.0):
    """DiscretizedLogisticMixture.create_params, dlm.py:161-211, literal replay
    of every reshape.  params [B,W,H,10*num_mix], x [B,W,H,3]."""
    B, W, H, _ = params.shape
    C = x.shape[-1]
    logits = params[:, :, :, :num_mix]                                        # dlm.py:172
    p = params[:, :, :, num_mix:].reshape(B, W, H, num_mix, 2 * C + 3)        # dlm.py:175
    coeffs = torch.tanh(p[:, :, :, :, :3]).permute(0, 1, 2, 4, 3)             # dlm.py:177-179
    p = p[:, :, :, :, 3:].reshape(B, W, H, C, num_mix * 2)                    # dlm.py:182
    locs, log_scales = p[..., :num_mix], p[..., num_mix:]                     # dlm.py:184
    xt = x.reshape(B, W, H, C, 1).repeat(1, 1, 1, 1, num_mix)                 # dlm.py:188-191
    m2 = locs[:, :, :, 1, :] + coeffs[:, :, :, 0, :] * xt[:, :, :, 0, :]      # dlm.py:199
    m3 = (locs[:, :, :, 2, :] + coeffs[:, :, :, 1, :] * xt[:, :, :, 0, :]
          + coeffs[:, :, :, 2, :] * xt[:, :, :, 1, :])                        # dlm.py:204
    locs = torch.stack([locs[:, :, :, 0, :], m2, m3], dim=3)                  # dlm.py:208
    log_scales = torch.clamp(log_scales, min=ls_low)                          # dlm.py:211
    return logits, locs, log_scales


def dlm_log_prob(params, x, num_mix=5, ls_low=-7.0):
    """DiscretizedLogisticMixture.log_prob, dlm.py:311-383 -> [B]."""
    logits, locs, log_scales = dlm_create_params(params, x, num_mix, ls_low)
    xb = x.unsqueeze(-1) + torch.zeros_like(locs)                             # dlm.py:336
    centered = xb - locs                                                      # dlm.py:349
    inv_stdv = torch.exp(-log_scales)                                         # dlm.py:351
    plus_in = inv_stdv * (centered + 1.0 / 255.0)
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_stdv * (centered - 1.0 / 255.0)
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - softplus(plus_in)                                # dlm.py:356
    log_one_minus_cdf_min = -softplus(min_in)                                 # dlm.py:357
    cdf_delta = cdf_plus - cdf_min
    mid_in = inv_stdv * centered
    log_pdf_mid = mid_in - log_scales - 2.0 * softplus(mid_in)                # dlm.py:362
    log_probs = torch.where(
        xb < -0.999, log_cdf_plus,
        torch.where(xb > 0.99, log_one_minus_cdf_min,
                    torch.where(cdf_delta > 1e-5,
                                torch.log(torch.clamp(cdf_delta, min=1e-12)),
                                log_pdf_mid - math.log(127.5))))              # dlm.py:365
    lp = log_probs.sum(dim=3) + torch.log_softmax(logits, dim=-1)             # dlm.py:373,398-412
    return torch.logsumexp(lp, dim=-1).sum(dim=(1, 2))                        # dlm.py:375,387-395


# ---------------------------------------------------------------------------
# R8: Bernoulli
# ---------------------------------------------------------------------------

def bernoulli_log_prob(logits, x, crop_size):
    """Bernoulli.log_prob, stat_tools/bernoulli.py:155-179 -> [B].
    tfp Bernoulli(logits).log_prob(x) = -sigmoid_cross_entropy(labels=x, logits)."""
    logp = x * logits - softplus(logits)
    if crop_size > 0:
        logp = logp[:, crop_size:-crop_size, crop_size:-crop_size, :]         # bern.py:177
    else:
        # python slice [0:-0] is empty; the reference only calls this with
        # crop_size=pad_size=2 (run.py:49, data_layer.py:179).
        logp = logp[:, 0:0, 0:0, :]
    return logp.sum(dim=(1, 2, 3))


# ---------------------------------------------------------------------------
# R9/R10: NELBO assembly and importance-sampled estimate
# ---------------------------------------------------------------------------

def nelbo(nll, kls):
    """vae.py:262-265."""
    kl = sum(kls)
    return nll, kl, nll + kl


def is_log_likelihood(p_logps, q_logps, nll, num_samples):
    """infer_worker_body + predict_step for ONE chunk, vae.py:465-482,539.
    p_logps/q_logps: lists (one per layer) of [k*B] (sample-major), nll [k*B]."""
    l = sum(p - q for p, q in zip(p_logps, q_logps))                          # vae.py:465
    l = l - nll                                                               # vae.py:478
    l = l.reshape(num_samples, -1)                                            # vae.py:480
    return torch.logsumexp(l, dim=0)                                          # vae.py:482


def is_combine_chunks(chunk_lse, num_samples_total):
    """predict_step, vae.py:539: logsumexp over chunks - log K."""
    return torch.logsumexp(torch.stack(chunk_lse, 0), dim=0) - math.log(num_samples_total)


# ---------------------------------------------------------------------------
# 3xTF32 operand split of the convolution / nonlocal kernels (no reference counterpart: the reference convolves in
# fp32, layers/resnet.py:589-601; this restates what csrc/tc_common.cuh lo_tf32 must produce so that
# hi*hi' + lo*hi' + hi*lo' recovers the fp32 product)
# ---------------------------------------------------------------------------

def tf32_split(v):
    """v (float32 tensor) -> (hi, lo): hi = v truncated to the 10 explicit mantissa bits a TF32 operand keeps (what the
    tensor core reads of an fp32 bit pattern), lo = (v - hi) rounded to 10 explicit mantissa bits, nearest, ties away
    from zero (cvt.rna.tf32.f32).  Written with frexp / ldexp arithmetic, not bit masks, so that it is independent of
    the integer formulation the GPU tests use."""
    import numpy as np
    a = v.detach().cpu().numpy().astype(np.float32)
    m, e = np.frexp(a.astype(np.float64))                      # a = m * 2^e, 0.5 <= |m| < 1
    hi = np.ldexp(np.trunc(m * 2048.0), e - 11)                # 11 significant bits, toward zero
    r = a.astype(np.float64) - hi                              # exact: both are fp32 values
    m, e = np.frexp(r)
    q = m * 2048.0
    lo = np.ldexp(np.sign(q) * np.floor(np.abs(q) + 0.5), e - 11)
    return torch.from_numpy(hi.astype(np.float32)), torch.from_numpy(lo.astype(np.float32))

