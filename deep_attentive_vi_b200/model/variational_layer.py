"""One layer of the variational hierarchy -- host mirror of the reference's
model/variational_layer.py (VariationalLayer).  The dataflow follows
SURVEY.md appendix A line by line; the depth-wise attentions, the LN/GELU glue
and the sample+KL step are single fused kernels, and the layer stacks of the
reference (tf.stack / tf.split / unstack-replace-stack) are python lists of
views into the per-layer tensors -- nothing is copied.
"""
import torch
import torch.nn as nn

from .. import ops
from ..layers import LAYER, DepthWiseAttention
from ..layers.util import KERAS_DEFAULT_L2, deep_vae_init_
from ..stat_tools import DISTRIBUTION


def _ln_params(C):
    return nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))


class VariationalLayer(nn.Module):
    def __init__(self, posterior_type, prior_type, latent_shape, posterior_hparams,
                 posterior_network_hparams, prior_hparams=None, prior_network_hparams=None,
                 encoder_hparams=None, decoder_hparams=None, merge_encoder_hparams=None,
                 merge_decoder_hparams=None, constant_decoder=False, decoder_feature_shape=None,
                 residual=False, decoder_attention_hparams=None, encoder_attention_hparams=None,
                 is_last_layer=False, encoder_channels=None, context_channels=None):
        """Extra (torch needs static shapes): encoder_channels = channels of the
        bottom-up tensors e[j]; context_channels = channels of the previous
        layer's context values (input a of the merge-decoder cell)."""
        super().__init__()
        self._latent_shape = list(latent_shape)
        self._residual = residual
        self._constant_decoder = constant_decoder
        self._is_last_layer = is_last_layer
        self._encoder_hparams, self._decoder_hparams = encoder_hparams, decoder_hparams
        self._decoder_attention_hparams = decoder_attention_hparams
        self._encoder_attention_hparams = encoder_attention_hparams
        H, Cg = decoder_feature_shape[0], decoder_feature_shape[-1]
        z = self._latent_shape[-1]

        # bottom-up residual cell (vl.py:145-148)
        self._encoder_network = (LAYER[encoder_hparams.layer_type](encoder_hparams, encoder_channels, H)
                                 if encoder_hparams is not None else None)
        # top-down cell, or the learned constant of the first layer (vl.py:151-162)
        self._decoder_network, self._decoder_const = None, None
        if decoder_hparams is not None:
            if constant_decoder:
                shape = (1, decoder_feature_shape[0], decoder_feature_shape[1], Cg)
                # VarianceScaling fans of a rank-4 shape: fan_in = prod(shape[:-1])
                self._decoder_const = nn.Parameter(deep_vae_init_(
                    torch.empty(shape), fan_in=shape[0] * shape[1] * shape[2]))
            else:
                self._decoder_network = LAYER[decoder_hparams.layer_type](decoder_hparams, Cg, H)
        # merge cells (vl.py:165-184)
        self._merge_decoder_network = None
        if merge_decoder_hparams is not None:
            self._merge_decoder_network = LAYER[merge_decoder_hparams.layer_type](
                merge_decoder_hparams, (context_channels, z), num_filters=Cg)
        self._use_decoder_attention = decoder_attention_hparams is not None
        self._use_encoder_attention = encoder_attention_hparams is not None
        dq = decoder_attention_hparams.query_dim if self._use_decoder_attention else 0
        dk = decoder_attention_hparams.key_dim if self._use_decoder_attention else 0
        # channels of the context proper after the key/query split (vl.py:375-380)
        if self._use_decoder_attention:
            self._ctx_channels = Cg - dq - (0 if is_last_layer else dk)
        else:
            self._ctx_channels = Cg
        self._merge_encoder_network = None
        enc_f = encoder_hparams["num_filters"] if encoder_hparams is not None else None
        if merge_encoder_hparams is not None:
            nf = enc_f + encoder_attention_hparams.key_dim if self._use_encoder_attention else None
            self._merge_encoder_network = LAYER[merge_encoder_hparams.layer_type](
                merge_encoder_hparams, (enc_f if self._use_encoder_attention else encoder_channels,
                                        self._ctx_channels), num_filters=nf)
        # distributions (vl.py:187-200)
        post_in = encoder_channels if self._use_encoder_attention else (
            self._merge_encoder_network.out_channels if self._merge_encoder_network is not None else encoder_channels)
        self._posterior_network = DISTRIBUTION[posterior_type](self._latent_shape, posterior_hparams,
                                                               posterior_network_hparams, post_in)
        self._prior_network = (DISTRIBUTION[prior_type](self._latent_shape, prior_hparams,
                                                        prior_network_hparams, self._ctx_channels)
                               if prior_hparams is not None else None)
        # depth-wise attention in the decoder (vl.py:204-230)
        if self._use_decoder_attention:
            self._decoder_attention = DepthWiseAttention(decoder_hparams, decoder_attention_hparams)
            self._dec_ln = decoder_attention_hparams.use_layer_norm
            if self._dec_ln:
                C = self._ctx_channels
                self.attn_ln_gamma, self.attn_ln_beta = _ln_params(C)      # _attention_layer_norm_layer
                self.attn_ln2_gamma, self.attn_ln2_beta = _ln_params(C)    # _attention_layer_norm_layer2
                self.ctx_ln_gamma, self.ctx_ln_beta = _ln_params(C)        # _ctx_layer_norm_layer
            self._gamma = nn.Parameter(torch.zeros(()))                    # vl.py:226-230
        # depth-wise attention in the encoder (vl.py:234-257)
        if self._use_encoder_attention:
            self._encoder_attention = DepthWiseAttention(encoder_hparams, encoder_attention_hparams)
            self._encoder_gamma = nn.Parameter(torch.zeros(()))            # vl.py:241-245
            self._enc_ln = encoder_attention_hparams.use_layer_norm
            if self._enc_ln:
                self.cond_ln_gamma, self.cond_ln_beta = _ln_params(enc_f)  # _cond_attention_layer_norm_layer
                self.enc_ln_gamma, self.enc_ln_beta = _ln_params(enc_f)    # _encoder_attention_layer_norm_layer
        self._enc_f = enc_f

    # ------------------------------------------------------------------
    def bottom_up(self, condition, training):
        """infer(direction='bottom-up'), vl.py:449-451."""
        return self._encoder_network(condition, training) if self._encoder_network is not None else condition

    def compute_context(self, latent_condition, context, training, num_samples):
        """vl.py:262-295."""
        if latent_condition is not None:
            if context is not None:
                prev = self._merge_decoder_network([context, latent_condition], training)   # vl.py:282
            else:
                prev = latent_condition
            return self._decoder_network(prev, training)                                    # vl.py:288
        if self._constant_decoder:
            return self._decoder_const.expand(num_samples, -1, -1, -1).contiguous()         # vl.py:291-293
        return None

    def call_decoder(self, training, latent_condition=None, context=None, num_samples=1):
        """vl.py:338-418.  `context` is the LIST c[0..i-1] of [B,H,W,values||key]
        tensors (the reference receives tf.stack of it)."""
        ctx_in = context
        slots = None
        if self._use_decoder_attention:
            # `context` = slot list: ("own"/"acc", SharedSlot) for c[0..i-2] (already layer-normed,
            # vae.py:346-348) and, last, the raw tensor c[i-1]
            qd = self._decoder_attention_hparams.query_dim
            last = context[-1]
            nv = last.shape[-1] - qd                                                        # vl.py:358
            slots = list(context[:-1])
            last_values, last_key = ops.split_channels(last, nv, qd)                        # vl.py:359 (views)
            ctx_in = last_values                                                            # vl.py:360
        new_context = self.compute_context(latent_condition, ctx_in, training, num_samples)
        key = query = None
        if self._use_decoder_attention:
            qd, kd = self._decoder_attention_hparams.query_dim, self._decoder_attention_hparams.key_dim
            full = new_context
            g = self._ctx_channels
            if not self._is_last_layer:                                                     # vl.py:375-380
                new_context, key, query = ops.split_channels(full, g, kd, full.shape[-1] - g - kd)
            else:
                new_context, query = ops.split_channels(full, g, full.shape[-1] - g)
        enc_query = None
        if self._use_encoder_attention:
            enc_query = new_context[..., :self._encoder_attention_hparams.query_dim]        # vl.py:385-386
        if self._use_decoder_attention:
            if self._dec_ln:
                new_context = ops.ln_gelu_residual(new_context, self.ctx_ln_gamma, self.ctx_ln_beta)  # vl.py:393-394
                last_values = ops.layer_norm(last_values, self.attn_ln_gamma, self.attn_ln_beta)      # vl.py:396-401
            slots.append(("kv", last_key, last_values))
            att = self._decoder_attention.apply_ctx(query, slots)                           # vl.py:403-406
            if self._dec_ln:                                                                # vl.py:408-412, fused
                prior_condition = ops.ln_gelu_res(att, self.attn_ln2_gamma, self.attn_ln2_beta, new_context,
                                                  self._gamma)
            else:
                prior_condition = new_context + self._gamma * att                           # vl.py:412
        else:
            prior_condition = new_context
        return new_context, prior_condition, key, enc_query

    def infer_top_down(self, condition, training, latent_condition=None, context=None, eps=None,
                       noise=None, compute_logp=False):
        """infer(direction='top-down'), vl.py:453-550.  `condition` is the LIST
        e[i..L] (or the single tensor e[i] without encoder attention).
        eps [B,H,W,z]; noise = (posterior_noise, prior_noise) or None.
        Returns xi, context, kl, q_logp, p_logp."""
        B = (condition[0] if isinstance(condition, (list, tuple)) else condition).shape[0]
        new_context, prior_condition, key, enc_query = self.call_decoder(
            training, latent_condition, context, B)
        not_first = latent_condition is not None
        if not_first:
            self._prior_network.call(prior_condition, training)                             # vl.py:463
        if self._use_encoder_attention:
            # `condition` = [e[i] tensor, then ("own"/"acc", SharedSlot) for e[i+1..L]]
            f = self._enc_f
            cond, key0 = ops.split_channels(condition[0], f, condition[0].shape[-1] - f)    # vl.py:472-473 (views)
        else:
            cond = condition
        if not_first:
            cond = self._merge_encoder_network([cond, prior_condition], training)           # vl.py:477
        if self._use_encoder_attention:
            if not_first:
                f = self._enc_f
                cond, key0 = ops.split_channels(cond, f, cond.shape[-1] - f)                # vl.py:483-494
            att = self._encoder_attention.apply_ctx(enc_query, [("kv", key0, cond)] + list(condition[1:]))  # vl.py:496-499
            if self._enc_ln:
                cond = ops.ln_gelu_residual(cond, self.cond_ln_gamma, self.cond_ln_beta)    # vl.py:503
                cond = ops.ln_gelu_res(att, self.enc_ln_gamma, self.enc_ln_beta, cond,
                                       self._encoder_gamma)                                 # vl.py:502,505 fused
            else:
                cond = cond + self._encoder_gamma * att                                     # vl.py:505
            cond = torch.cat([cond, key0], dim=-1)                                          # vl.py:507
        n_post, n_prior = noise if noise is not None else (None, None)
        prior = self._prior_network if (not_first and self._residual) else None
        if not_first and not self._residual:
            raise NotImplementedError("non-residual posterior with a learned prior is not used by any published config")
        out = self._posterior_network.sample_kl(cond, training, eps, prior=prior, prior_noise=n_prior,
                                                noise=n_post, compute_logp=compute_logp)   # vl.py:511-542
        xi, kl = out[0], out[1]
        q_logp, p_logp = (out[2], out[3]) if compute_logp else (None, None)
        if not self._use_decoder_attention:
            return xi, prior_condition, kl, q_logp, p_logp                                  # vl.py:545-546
        if not self._is_last_layer:
            return xi, torch.cat([new_context, key], dim=-1), kl, q_logp, p_logp            # vl.py:547-548
        return xi, new_context, kl, q_logp, p_logp

    # ------------------------------------------------------------------
    def reg_loss(self):
        r = 0.0
        for m in (self._encoder_network, self._decoder_network, self._merge_decoder_network,
                  self._merge_encoder_network, self._posterior_network, self._prior_network):
            if m is not None:
                r = r + m.reg_loss()
        k = KERAS_DEFAULT_L2   # string 'l2' regularisers (vl.py:156,215-216,229,244)
        if self._decoder_const is not None:
            r = r + k * self._decoder_const.square().sum()
        if self._use_decoder_attention:
            r = r + k * self._gamma.square()
            if self._dec_ln:
                for p in (self.attn_ln_gamma, self.attn_ln_beta, self.attn_ln2_gamma, self.attn_ln2_beta,
                          self.ctx_ln_gamma, self.ctx_ln_beta):
                    r = r + k * p.square().sum()
        if self._use_encoder_attention:
            r = r + k * self._encoder_gamma.square()
            if self._enc_ln:
                for p in (self.cond_ln_gamma, self.cond_ln_beta, self.enc_ln_gamma, self.enc_ln_beta):
                    r = r + k * p.square().sum()
        return r
