"""TensorFlow side of the drop-in (STATUS: source only -- TensorFlow is not
installed in the build container or on the GPU boxes, so nothing here is
imported by the tests; the same C ABI is exercised through
deep_attentive_vi_b200/_lib.py).

Usage inside the reference tree (src/ on sys.path, TF 2.x with a CUDA build):

    import davi_tf                       # loads libdavi_tf.so, registers gradients
    davi_tf.install()                    # swaps the registries' bodies, see below

`install()` replaces, in the reference's own registries, exactly the bodies the
hot path covers and nothing else:
  layers.attention.DepthWiseAttention.apply        -> DaviDwAttn        (att.py:135-164)
  layers.attention.NonlocalResNetBlock.call (core) -> DaviNonlocalAttn  (att.py:318-389)
  variational_layer KL / sample site               -> DaviGaussKl       (vl.py:511-542, gd.py:121-240)
  DiscretizedLogisticMixture.log_prob              -> DaviDlmNll        (dlm.py:311-383)
  Bernoulli.log_prob                               -> DaviBernoulliNll  (bern.py:155-179)
hparams, constructor signatures and weights are untouched, so README command
lines and checkpoints keep working.
"""
import os

import tensorflow as tf
from tensorflow.python.framework import ops as tf_ops

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = tf.load_op_library(os.path.join(_HERE, "libdavi_tf.so"))


@tf_ops.RegisterGradient("DaviDwAttn")
def _dw_attn_grad(op, dout):
    q, k, v, s = op.inputs
    return _mod.davi_dw_attn_grad(query=q, keys=k, values=v, scale=s, dout=dout)


@tf_ops.RegisterGradient("DaviNonlocalAttn")
def _nonlocal_grad(op, dy, _dlse, _do):
    q, k, v, x, scale, gamma = op.inputs
    dq, dk, dv, dgs = _mod.davi_nonlocal_attn_grad(q=q, k=k, v=v, lse=op.outputs[1], scale=scale,
                                                   gamma=gamma, dy=dy, o=op.outputs[2])
    return dq, dk, dv, dy, dgs[1], dgs[0]          # d x == dy (residual)


@tf_ops.RegisterGradient("DaviGaussKl")
def _gauss_kl_grad(op, dxi, dkl, _dlq, _dlp):
    post, prior, eps = op.inputs
    attrs = {a: op.get_attr(a) for a in ("post_lo", "post_hi", "prior_lo", "prior_hi", "has_prior")}
    dpost, dprior = _mod.davi_gauss_kl_grad(post=post, prior=prior, eps=eps, dxi=dxi, dkl=dkl, **attrs)
    return dpost, dprior, None


@tf_ops.RegisterGradient("DaviDlmNll")
def _dlm_grad(op, dnll, _unused):
    g = op.outputs[1]                                # per-image d nll / d params from the forward launch
    return g * tf.reshape(dnll, [-1, 1, 1, 1]), None


@tf_ops.RegisterGradient("DaviBernoulliNll")
def _bern_grad(op, dnll, _unused):
    return op.outputs[1] * tf.reshape(dnll, [-1, 1, 1, 1]), None


def depthwise_attention_apply(self, query, keys, values):
    """Body for layers.attention.DepthWiseAttention.apply (att.py:135-164)."""
    scale = getattr(self._alignment_function, "scale", None)
    if scale is None:
        scale = tf.constant(1.0, tf.float32)
    return _mod.davi_dw_attn(query=query, keys=keys, values=values, scale=scale)


def nonlocal_core(q, k, v, x, scale, gamma):
    """softmax(scale q k^T) v * gamma + x for NonlocalResNetBlock.call (att.py:318-389);
    q,k,v are the outputs of the block's three 1x1 convolutions, NHWC."""
    y, _, _ = _mod.davi_nonlocal_attn(q=q, k=k, v=v, x=x, scale=scale, gamma=gamma)
    return y


def install():
    """Monkey-patch the reference registries (run once, before building DeepVAE)."""
    from layers import attention as att
    att.DepthWiseAttention.apply = depthwise_attention_apply
    att.DAVI_NONLOCAL_CORE = nonlocal_core           # consumed by the patched NonlocalResNetBlock.call
