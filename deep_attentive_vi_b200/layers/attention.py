"""Attention layers of the reference (layers/attention.py) with their bodies
routed through the sm_100a kernels.  Class names, constructor arguments and
hparams are the reference's; tensors are NHWC float32 on the GPU.
"""
import torch
import torch.nn as nn

from .. import ops
from ..util.hparams import HParams
from .util import KERAS_DEFAULT_L2
from .weight_norm import WNConv2d, to_nchw, to_nhwc


class SoftmaxAlignment(nn.Module):
    """reference layers/attention.py:60-116.  Holds the learnable logit scale
    (init 1.0, att.py:84-92); the softmax itself lives inside the fused kernels."""

    def __init__(self, hparams, key_dim=None, use_scale=False, trainable=True):
        super().__init__()
        self._hparams, self._key_dim, self._use_scale = hparams, key_dim, use_scale
        if use_scale:
            self.scale = nn.Parameter(torch.ones(()), requires_grad=trainable)
        else:
            self.scale = None
        self.lambda_reg = getattr(hparams, "lambda_reg", 1.5e-4) if hparams is not None else 1.5e-4

    def call(self, proj_query, proj_key):
        """Drop-in for att.py:98-116 on the depth-wise layout: proj_query
        [B,W,H,d,1], proj_key [B,W,H,d,n] -> softmax weights [B,W,H,1,n]."""
        q = proj_query.squeeze(-1)
        k = proj_key.permute(0, 4, 1, 2, 3).contiguous()
        return ops.dw_attn_weights(q, k, self.scale).unsqueeze(-2)

    def reg_loss(self):
        if self.scale is None or not self.scale.requires_grad:
            return 0.0
        return self.lambda_reg * self.scale.square()          # att.py:88 kernel_regularizer


ALIGNMENT_FUNCTION = {"softmax": SoftmaxAlignment}          # att.py:418


class DepthWiseAttention(nn.Module):
    """reference layers/attention.py:119-177.

    The reference class is a bare tf.Module whose SoftmaxAlignment creates
    `scale` lazily, so Keras never tracks it: it stays the constant 1.0
    (SURVEY.md 8b).  `train_scale=True` makes it a trained weight instead."""

    def __init__(self, hparams, attention_hparams, train_scale=False):
        super().__init__()
        self._key_dim = attention_hparams.key_dim
        self._query_dim = attention_hparams.query_dim
        self._use_scale = attention_hparams.use_scale
        self._hparams = hparams
        self._alignment_function = ALIGNMENT_FUNCTION[attention_hparams.alignment_function](
            hparams, self._query_dim, self._use_scale, trainable=train_scale)

    @property
    def scale(self):
        return self._alignment_function.scale

    def apply(self, query, keys, values):
        """Drop-in for att.py:135-164: query [B,W,H,d], keys [B,n,W,H,d],
        values [B,n,W,H,C] -> [B,W,H,C] (one fused kernel, no Permute copies)."""
        return ops.dw_attn(query, keys, values, self.scale)

    def apply_ctx(self, query, slots):
        """The same contraction over a slot list of ("kv", key, value) pairs and ("own"/"acc",
        ops.SharedSlot) contexts shared with other layers (lazy gradient gather)."""
        return ops.dw_attn_ctx(query, slots, self.scale)

    def apply_slots(self, query, keys, values):
        """Same contraction over python lists of per-layer tensors (the
        tf.stack of vae.py:340,350 never happens)."""
        return ops.dw_attn_slots(query, keys, values, self.scale)

    @staticmethod
    def get_default_hparams():
        return HParams(key_dim=20, query_dim=20, value_dim=20, alignment_function="softmax",
                       use_scale=True, use_layer_norm=False)               # att.py:171-177


class NonlocalResNetBlock(nn.Module):
    """reference layers/attention.py:180-416: single-head softmax attention
    over the W*H positions, y = gamma * attn(x) + x.  The three weight-normed
    1x1 projections run as ONE cuDNN GEMM; softmax(QK^T)V + the gamma residual
    run in the fused attention kernel (the [B,N,N] matrix is never in HBM)."""

    def __init__(self, hparams, nonlocop_hparams, in_channels):
        super().__init__()
        C = in_channels
        self._key_dim = nonlocop_hparams.key_dim if nonlocop_hparams.key_dim is not None else C // 8
        self._use_scale = nonlocop_hparams.use_scale
        self._subsample = nonlocop_hparams.subsample
        if self._subsample:
            raise NotImplementedError("subsample=True (att.py:323-326) is not used by any published config")
        self._hparams, self._nonlocop_hparams, self._C = hparams, nonlocop_hparams, C
        kw = dict(kernel_size=1, use_bias=hparams.use_bias, use_weight_norm=hparams.use_weight_norm,
                  lambda_reg=hparams.lambda_reg)
        self.query_conv = WNConv2d(C, self._key_dim, **kw)                  # att.py:225
        self.key_conv = WNConv2d(C, self._key_dim, **kw)                    # att.py:234
        self.value_conv = WNConv2d(C, C, **kw)                              # att.py:243
        self._alignment_function = ALIGNMENT_FUNCTION[nonlocop_hparams.alignment_function](
            hparams, self._key_dim, self._use_scale)
        self.gamma = nn.Parameter(torch.zeros(()))                          # att.py:256-261
        self._use_ln = nonlocop_hparams.use_layer_norm
        if self._use_ln:                                                    # att.py:263-282
            self.ln_a_gamma, self.ln_a_beta = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
            self.ln_b_gamma, self.ln_b_beta = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
            self._ln_conv2d = WNConv2d(C, C, **kw)
        self._pad = (-C) % 4
        # set by util/trainer.py: one leaf view [2d+C, C, 1, 1] over the three adjacent effective kernels
        self._qkv_w_override = None
        self._qkv_wt_override = None
        self._qkv_lo_override = None

    def _qkv(self, x):
        d, C, pad = self._key_dim, self._C, self._pad
        bs = [self.query_conv.bias, self.key_conv.bias, self.value_conv.bias]
        if self._qkv_w_override is not None:
            b = torch.cat(bs, 0) if bs[0] is not None else None
            return to_nhwc(ops.conv2d_bias(to_nchw(x), self._qkv_w_override, b, wt=self._qkv_wt_override,
                                           lo=self._qkv_lo_override))
        ws = [self.query_conv.kernel(), self.key_conv.kernel(), self.value_conv.kernel()]
        if pad:   # keep rows 16-byte multiples (C = 138 in the CIFAR-10 pre-processor)
            ws.append(ws[0].new_zeros(pad, C, 1, 1))
            if bs[0] is not None:
                bs.append(bs[0].new_zeros(pad))
        w = torch.cat(ws, 0).contiguous(memory_format=torch.channels_last)
        b = torch.cat(bs, 0) if bs[0] is not None else None
        return to_nhwc(ops.conv2d_bias(to_nchw(x), w, b))       # att.py:310,346,362  [B,H,W,2d+C(+pad)]

    def forward(self, inputs):
        orig = inputs
        if self._use_ln:
            inputs = ops.layer_norm(inputs, self.ln_a_gamma, self.ln_a_beta)         # att.py:293-294
        qkv = self._qkv(inputs)
        x = inputs if not self._pad else torch.nn.functional.pad(inputs, (0, self._pad))
        d = self._key_dim
        if d % 4 == 0:
            y = ops.nonlocal_attn_packed(qkv, x, d, self._alignment_function.scale, self.gamma)   # att.py:358-389
        else:   # odd key widths: slices are copied into 16-byte-aligned rows by the op
            y = ops.nonlocal_attn(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], x,
                                  self._alignment_function.scale, self.gamma)
        if self._pad:
            y = y[..., :self._C].contiguous()
        if not self._use_ln:
            return y
        # y = gamma*o + LN_a(x); the reference adds the ORIGINAL input (att.py:395)
        y = y - inputs + orig
        return orig + self._ln_conv2d(ops.layer_norm(y, self.ln_b_gamma, self.ln_b_beta))  # att.py:395-398

    def reg_loss(self):
        r = self.query_conv.reg_loss() + self.key_conv.reg_loss() + self.value_conv.reg_loss()
        r = r + self._alignment_function.reg_loss()
        r = r + KERAS_DEFAULT_L2 * self.gamma.square()                      # att.py:260: string 'l2'
        if self._use_ln:
            r = r + self._ln_conv2d.reg_loss() + KERAS_DEFAULT_L2 * (
                self.ln_a_gamma.square().sum() + self.ln_a_beta.square().sum()
                + self.ln_b_gamma.square().sum() + self.ln_b_beta.square().sum())
        return r

    @staticmethod
    def get_default_hparams():
        return HParams(key_dim=32, query_dim=32, use_scale=True, value_dim=32, num_heads=1,
                       subsample=False, alignment_function="softmax", use_layer_norm=False,
                       layer_norm_regularizer="l2")                        # att.py:407-416
