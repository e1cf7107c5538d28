#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/src: layers/, stat_tools/, model/) on top of the TensorFlow stand-in in
tf_shim.py (float64).  Run in the build container only (the reference tree does not travel):

    python tests/golden/make_golden.py

Outputs (committed):
  ref_ops.npz            op-level vectors: DepthWiseAttention.apply, NonlocalResNetBlock.call,
                         GaussianDiag (bounds, residual, KL, log-probs), DiscretizedLogisticMixture
                         and Bernoulli log_prob
  ref_small_cifar.npz    whole-model [nll, kl, nelbo] of DeepVAE.call for a 4-layer CIFAR-like
  ref_small_omniglot.npz config / a 3-layer OMNIGLOT-like config / a 3-layer NVAE-like config without
  ref_small_nvae.npz     attention: weights (by reference attribute path), input batch, every random
                         draw in call order, outputs; plus one chunk of the importance-sampled
                         estimator (DeepVAE.infer_worker_body, 3 samples per image) with its draws
What this pins: the reference's Python logic (reshape scrambles, channel maps, stack/split order,
LN placement, residual parametrisation, loss assembly).  What it assumes: the TF/TFP/tfa kernels
behave as tf_shim.py restates them (documented semantics of the pinned versions).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))          # tests/ (helpers.py)
REF = "/root/reference/src"

import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, REF)
_u = types.ModuleType("util")                       # skip util/__init__ (imports tfds, matplotlib, ...)
_u.__path__ = [os.path.join(REF, "util")]
sys.modules["util"] = _u

import helpers  # noqa: E402
from layers import LAYER, DepthWiseAttention, NonlocalResNetBlock  # noqa: E402  (reference code)
from layers.resnet import Conv2D as RefConv2D  # noqa: E402
from model import DeepVAE  # noqa: E402
from stat_tools import DISTRIBUTION  # noqa: E402
import tensorflow_probability as tfp  # noqa: E402  (the shim)

IMAGE_INFO = {"cifar10": {"shape": (32, 32, 3), "high": 255, "low": 0, "binary": False, "pad_size": 0},
              "omniglot": {"shape": (28, 28, 1), "high": 1, "low": 0, "binary": True, "pad_size": 2}}


def f32(t):
    """round to fp32-representable values (the fp32 product then starts from identical numbers)"""
    return t.to(torch.float32).to(torch.float64)


def build_reference_model(flags):
    """run.py:59-109 without the flags / strategy plumbing."""
    g = lambda k: flags.get(k, "")
    hp = DeepVAE.get_default_hparams().update_config(g("hparams"))
    net = DeepVAE.get_default_network_hparams
    att = lambda s: None if not s else DepthWiseAttention.get_default_hparams().update_config(s)
    return DeepVAE(
        hparams=hp,
        data_dist_hparams=DISTRIBUTION[hp.data_distribution].get_default_hparams().update_config(g("data_dist_hparams")),
        preproc_encoder_hparams=net(hp).preproc_encoder.update_config(g("preproc_encoder_hparams")),
        postproc_decoder_hparams=net(hp).postproc_decoder.update_config(g("postproc_decoder_hparams")),
        encoder_hparams=net(hp).encoder.update_config(g("encoder_hparams")),
        merge_encoder_hparams=LAYER[hp.merge_encoder_type].get_default_hparams().update_config(g("merge_encoder_hparams")),
        decoder_hparams=net(hp).decoder.update_config(g("decoder_hparams")),
        merge_decoder_hparams=LAYER[hp.merge_decoder_type].get_default_hparams().update_config(g("merge_decoder_hparams")),
        posterior_hparams=DISTRIBUTION[hp.posterior_type].get_default_hparams().update_config(g("posterior_hparams")),
        prior_hparams=DISTRIBUTION[hp.prior_type].get_default_hparams().update_config(g("prior_hparams")),
        image_info=dict(IMAGE_INFO[flags["dataset"]]),
        decoder_attention_hparams=att(g("decoder_attention_hparams")),
        encoder_attention_hparams=att(g("encoder_attention_hparams")))


def randomize(model, seed):
    """Make every path count: the zero-initialised gamma scalars, LN / BN affine weights, logit
    scales and weight-norm gains get non-trivial values; everything is rounded to fp32."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda shape: torch.randn(tuple(shape), generator=g, dtype=torch.float64)
    with torch.no_grad():
        for path, name, w, _, _ in tf_shim.walk_weights(model):
            if name in ("initialized", "moving_mean", "moving_variance"):
                continue
            if name in ("gamma", "enc_gamma") and w.dim() == 0:
                w.copy_(0.5 * rn(()))
            elif name == "scale" and w.dim() == 0:
                w.copy_(1.0 + 0.3 * rn(()))
            elif name == "gamma":                       # LN / BN scale
                w.copy_(1.0 + 0.2 * rn(w.shape))
            elif name == "beta":
                w.copy_(0.2 * rn(w.shape))
            elif name == "g":                           # weight-norm log-gain
                w.add_(0.1 * rn(w.shape))
            w.copy_(f32(w))


def model_case(flags, fname, batch, seed):
    torch.manual_seed(seed)
    tf_shim.REC.gen.manual_seed(seed)
    m = build_reference_model(flags)
    kw = {"image_info": IMAGE_INFO[flags["dataset"]]}
    x = helpers.synthetic_images(kw, batch, seed=seed).double()
    m(x, training=True)                                 # builds the lazily created weights (+ weight-norm init)
    randomize(m, seed + 1)
    out = {}
    for mode, training in (("train", True), ("eval", False)):
        tf_shim.REC.normal_draws.clear()
        tf_shim.REC.training_noise = training
        nll, kl, nelbo = m(x, training=training)
        out[f"{mode}/nll"], out[f"{mode}/kl"], out[f"{mode}/nelbo"] = nll.numpy(), kl.numpy(), nelbo.numpy()
        for i, (tag, z) in enumerate(tf_shim.REC.normal_draws):
            out[f"{mode}/draw/{i:03d}/{tag}"] = z.numpy()
    # R10: one worker chunk of the importance-sampled estimator (vae.py:452-482) with S = 3 samples per image:
    # tiled batch (index = s * B + b), per-layer log p - log q, minus nll, logsumexp over the samples
    S = 3
    tf_shim.REC.normal_draws.clear()
    tf_shim.REC.training_noise = False
    out["is/num_samples"] = np.int64(S)
    out["is/logsumexp"] = m.infer_worker_body(x, S).numpy()
    for i, (tag, z) in enumerate(tf_shim.REC.normal_draws):
        out[f"is/draw/{i:03d}/{tag}"] = z.numpy()
    out["x"] = x.numpy()
    for path, name, w, _, _ in tf_shim.walk_weights(m):
        if name == "initialized":
            continue
        out[f"w/{path}/{name}"] = w.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, {k: v for k, v in out.items() if k.endswith(("nll", "kl"))})


def ops_case(fname):
    out = {}
    g = torch.Generator().manual_seed(7)
    rn = lambda *s: f32(torch.randn(*s, generator=g, dtype=torch.float64))
    HP = LAYER["resnet"].get_default_hparams()

    # ---- R1: DepthWiseAttention.apply (att.py:135-164) ----
    for tag, (B, n, H, d, C) in {"dw_a": (2, 5, 8, 20, 36), "dw_b": (3, 17, 4, 8, 12), "dw_c": (1, 1, 4, 4, 8)}.items():
        ahp = DepthWiseAttention.get_default_hparams().update_config(f"key_dim={d},query_dim={d}")
        att = DepthWiseAttention(HP, ahp)
        q, k, v = rn(B, H, H, d), rn(B, n, H, H, d), rn(B, n, H, H, C)
        y = att.apply(q, k, v)                                       # scale is created lazily = 1.0
        out.update({f"{tag}/q": q, f"{tag}/k": k, f"{tag}/v": v, f"{tag}/out": y})
        att._alignment_function.scale.assign(0.7)
        out[f"{tag}/out_scale0.7"] = att.apply(q, k, v)

    # ---- R4: NonlocalResNetBlock.call (att.py:285-398), q/k/v captured after the reference's own 1x1 convs ----
    for tag, (B, H, C, d) in {"nl_a": (2, 16, 24, 8), "nl_b": (1, 8, 12, 4)}.items():
        nhp = NonlocalResNetBlock.get_default_hparams().update_config(f"key_dim={d},query_dim={d}")
        blk = NonlocalResNetBlock(HP, nhp)
        x = rn(B, H, H, C)
        blk(x)                                                       # build
        blk.gamma.assign(0.6)
        blk._alignment_function.scale.assign(0.8)
        y = blk(x)
        q, k, v = (c.layer.last_output for c in (blk.query_conv, blk.key_conv, blk.value_conv))
        out.update({f"{tag}/x": x, f"{tag}/q": q, f"{tag}/k": k, f"{tag}/v": v, f"{tag}/y": y,
                    f"{tag}/gamma": torch.tensor(0.6, dtype=torch.float64),
                    f"{tag}/scale": torch.tensor(0.8, dtype=torch.float64)})

    # ---- R5/R6: GaussianDiag with the parameter network bypassed (condition == raw parameters) ----
    GD = DISTRIBUTION["gaussian_diag"]
    ident = lambda cond, training=None: cond
    for tag, (qb, pb) in {"gd_cifar": ((-5.0, 5.0), (-1.0, 5.0)), "gd_omni": ((-5.0, 5.0), (-5.0, 5.0))}.items():
        B, H, z = 2, 8, 6
        mk = lambda b: GD(tf.TensorShape([H, H, z]), GD.get_default_hparams().update_config(
            f"log_scale_low_bound={b[0]},log_scale_upper_bound={b[1]}"), LAYER["conv2d"].get_default_hparams())
        prior, post = mk(pb), mk(qb)
        prior._params_network = post._params_network = ident
        praw, draw = 3.0 * rn(B, H, H, 2 * z), 2.0 * rn(B, H, H, 2 * z)
        tf_shim.REC.normal_draws.clear()
        prior.call(praw, training=False)                                              # vl.py:463
        xi = post.sample(draw, training=False, num_samples=1, initial_params=prior.params())   # vl.py:511
        xi = tf.reshape(xi, [-1, H, H, z])                                            # vl.py:530
        kl = tf.reduce_sum(tfp.distributions.kl_divergence(post._p[-1], prior._p[0]), axis=[-1, -2])   # vl.py:540-542
        logq = post.log_prob(xi, draw, training=False, compute_dist=False)           # vl.py:513-516
        logp = prior.log_prob(xi, praw, training=False, compute_dist=False)          # vl.py:518-521
        eps = tf_shim.REC.normal_draws[-1][1][0]
        out.update({f"{tag}/prior_raw": praw, f"{tag}/post_raw": draw, f"{tag}/eps": eps, f"{tag}/xi": xi,
                    f"{tag}/kl": kl, f"{tag}/logq": logq, f"{tag}/logp": logp,
                    f"{tag}/post_bounds": torch.tensor(qb), f"{tag}/prior_bounds": torch.tensor(pb)})
        # first layer: p = N(0, I), non-residual posterior (vl.py:535-537)
        post0 = mk(qb)
        post0._params_network = ident
        xi0 = tf.reshape(post0.sample(draw, training=False, num_samples=1), [-1, H, H, z])
        p0 = tfp.distributions.MultivariateNormalDiag(loc=tf.zeros([H, H, z]), scale_diag=1.0 + tf.zeros([H, H, z]))
        kl0 = tf.reduce_sum(tfp.distributions.kl_divergence(post0._p[-1], p0), axis=[-1, -2])
        out.update({f"{tag}/eps0": tf_shim.REC.normal_draws[-1][1][0], f"{tag}/xi0": xi0, f"{tag}/kl0": kl0})

    # ---- R7: DiscretizedLogisticMixture.log_prob (dlm.py:161-211,311-383) ----
    DLM = DISTRIBUTION["discretized_logistic_mixture"]
    dlm = DLM(tf.TensorShape([8, 8, 3]), DLM.get_default_hparams(), LAYER["conv2d"].get_default_hparams())
    dlm._params_network = ident
    B = 3
    params = 2.0 * rn(B, 8, 8, 50)
    xq = torch.randint(0, 256, (B, 8, 8, 3), generator=g).double()
    xq[0, 0, 0, :] = 0.0
    xq[0, 0, 1, :] = 255.0                                            # both edge-bin branches
    xs = f32(2.0 * xq / 255.0 - 1.0)
    out.update({"dlm/params": params, "dlm/x": xs, "dlm/logp": dlm.log_prob(xs, params, training=False)})
    params_big = f32(params * 6.0)                                    # saturates cdf_delta < 1e-5 and the -7 floor
    out.update({"dlm/params_big": params_big, "dlm/logp_big": dlm.log_prob(xs, params_big, training=False)})

    # ---- R8: Bernoulli.log_prob with the border crop (bern.py:155-179) ----
    BE = DISTRIBUTION["bernoulli"]
    be = BE(tf.TensorShape([12, 12, 1]), BE.get_default_hparams(), LAYER["conv2d"].get_default_hparams())
    be._params_network = ident
    logits = 3.0 * rn(B, 12, 12, 1)
    xb = (torch.rand(B, 12, 12, 1, generator=g) < 0.5).double()
    out.update({"bern/logits": logits, "bern/x": xb, "bern/crop": torch.tensor(2),
                "bern/logp": be.log_prob(xb, logits, training=False, crop_size=2)})

    np.savez_compressed(os.path.join(HERE, fname), **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v)
                                                       for k, v in out.items()})
    print(fname, len(out), "arrays")


if __name__ == "__main__":
    ops_case("ref_ops.npz")
    model_case(helpers.SMALL_CIFAR, "ref_small_cifar.npz", batch=2, seed=11)
    model_case(helpers.SMALL_OMNIGLOT, "ref_small_omniglot.npz", batch=2, seed=12)
    model_case(helpers.SMALL_NVAE, "ref_small_nvae.npz", batch=2, seed=13)      # no attention, no nonlocal
