"""DeepVAE -- host mirror of the reference's model/vae.py (Deep Attentive VAE).

Same constructor keywords and hparams as the reference; `forward` returns the
reference's `[nll, sum(kl), nelbo]` (vae.py:262-265).  The attention and ELBO
hot path runs in the sm_100a kernels; the conv cells are host-side cuDNN.
Differences forced by PyTorch: shapes are static (no lazy build) and the random
draws (eps of the reparameterisation, train-time log-sigma noise) can be passed
in so that runs are reproducible against the oracle.
"""
import copy
import math

import torch
import torch.nn as nn

from .. import ops
from ..layers import LAYER
from ..layers.util import KERAS_DEFAULT_L2
from .data_layer import DataLayer
from .util import get_default_network_hparams
from .variational_layer import VariationalLayer


class DeepVAE(nn.Module):
    def __init__(self, hparams, preproc_encoder_hparams=None, encoder_hparams=None, decoder_hparams=None,
                 merge_decoder_hparams=None, merge_encoder_hparams=None, postproc_decoder_hparams=None,
                 data_dist_hparams=None, posterior_hparams=None, prior_hparams=None, image_info=None,
                 decoder_attention_hparams=None, encoder_attention_hparams=None, name="deep_autoencoder"):
        super().__init__()
        self.name = name
        hp = self._hparams = hparams
        self._image_info = image_info
        self._num_layers = L = hp.num_layers
        self._num_proc_blocks = hp.num_proc_blocks
        self._posterior_hparams, self._prior_hparams = posterior_hparams, prior_hparams
        self._decoder_attention_hparams = decoder_attention_hparams
        self._encoder_attention_hparams = encoder_attention_hparams
        net = get_default_network_hparams(hp)                                       # model/util.py:113
        img = image_info["shape"][0] + image_info["pad_size"] * 2                   # vae.py:135
        img_c = image_info["shape"][-1]
        down = 2 ** self._num_proc_blocks                                           # vae.py:95
        z = hp.layer_latent_shape[-1]

        # ---- pre-processing encoder (vae.py:104-122) ----
        num_enc = preproc_encoder_hparams["num_filters"]
        pre, c, size = [], img_c, img
        php = copy.deepcopy(preproc_encoder_hparams)
        for s in range(self._num_proc_blocks):
            if encoder_attention_hparams is not None:
                php["num_filters"] = num_enc + int(encoder_attention_hparams.key_dim / down)   # vae.py:109
            else:
                php["num_filters"] = num_enc
            layer = LAYER[preproc_encoder_hparams.layer_type](php, c, size)
            c, size = layer.out_channels, layer.out_size
            pre.append(layer)
            php = copy.deepcopy(php)
            if s + 1 < self._num_proc_blocks:
                php["use_stem"] = False
            num_enc *= 2
        self._preproc_encoder = nn.ModuleList(pre)
        enc_channels = c

        # ---- variational hierarchy (vae.py:130-207) ----
        lat = int(img / down)
        num_dec = decoder_hparams["num_filters"] * down                             # vae.py:148
        ehp = copy.deepcopy(encoder_hparams)
        ehp["num_filters"] = num_enc                                                # vae.py:154-155
        self._layer_latent_shape = [lat, lat, z]
        layers, ctx_ln = [], []
        ctx_values = None   # channels of the previous context's values
        ctx_total = None    # channels of the previous context tensor c[i-1]
        for i in range(L):
            cg = num_dec
            if decoder_attention_hparams is not None:
                if i > 0:
                    cg += decoder_attention_hparams.query_dim                       # vae.py:168-171
                if i < L - 1:
                    cg += decoder_attention_hparams.key_dim                         # vae.py:173-176
            if encoder_attention_hparams is not None:
                cg += encoder_attention_hparams.query_dim                           # vae.py:179-181
            layer = VariationalLayer(
                posterior_type=hp.posterior_type, prior_type=hp.prior_type, latent_shape=[lat, lat, z],
                encoder_hparams=ehp,
                decoder_hparams=decoder_hparams if (i > 0 or L > 1) else None,
                merge_decoder_hparams=merge_decoder_hparams if i > 0 else None,
                merge_encoder_hparams=merge_encoder_hparams if i > 0 else None,
                posterior_hparams=posterior_hparams,
                posterior_network_hparams=net.posterior if i > 0 else net.init_posterior,
                prior_hparams=prior_hparams if i > 0 else None,
                prior_network_hparams=net.prior if i > 0 else None,
                residual=hp.residual_var_layer, constant_decoder=(i == 0 and L > 1),
                decoder_feature_shape=[lat, lat, cg],
                decoder_attention_hparams=decoder_attention_hparams if i > 0 else None,
                is_last_layer=(i == L - 1), encoder_attention_hparams=encoder_attention_hparams,
                encoder_channels=enc_channels, context_channels=ctx_values)
            layers.append(layer)
            # what this layer hands to the next one (vl.py:545-550)
            if decoder_attention_hparams is not None and i > 0:
                ctx_total = layer._ctx_channels + (0 if i == L - 1 else decoder_attention_hparams.key_dim)
            else:
                ctx_total = cg
            ctx_values = (ctx_total - decoder_attention_hparams.query_dim
                          if decoder_attention_hparams is not None else ctx_total)
            if decoder_attention_hparams is not None:
                # vae.py:204-207: one LayerNormalization per layer, applied to c[i-2] for i > 1.
                # Keras creates weights lazily, so only the used ones (i >= 2) exist.
                if i >= 2:
                    cprev = layers[i - 2]._ctx_channels + decoder_attention_hparams.key_dim if i - 2 > 0 else \
                        num_dec + decoder_attention_hparams.key_dim + (
                            encoder_attention_hparams.query_dim if encoder_attention_hparams is not None else 0)
                    ctx_ln.append(nn.ParameterList([nn.Parameter(torch.ones(cprev)),
                                                    nn.Parameter(torch.zeros(cprev))]))
                else:
                    ctx_ln.append(nn.ParameterList([]))
        self._latent_layer = nn.ModuleList(layers)
        self._ctx_layer_norm = nn.ModuleList(ctx_ln)

        # ---- merge of the last context and sample, data layer (vae.py:78-89) ----
        last_ctx = ctx_total
        if L > 1 and merge_decoder_hparams is not None:
            self._merge_decoder = LAYER[merge_decoder_hparams.layer_type](merge_decoder_hparams, (last_ctx, z))
            data_in = self._merge_decoder.out_channels
        else:
            self._merge_decoder = None
            data_in = last_ctx if L > 1 else z
        self._data_layer = DataLayer(preproc_hparams=postproc_decoder_hparams, data_dist_hparams=data_dist_hparams,
                                     network_hparams=net.data, class_dist=hp.data_distribution,
                                     image_info=image_info, num_proc_blocks=self._num_proc_blocks,
                                     in_channels=data_in, in_size=lat)
        self.eval_samples, self.eval_workers = 1, 1

    # ------------------------------------------------------------------
    @property
    def image_shape(self):
        ps = self._image_info["pad_size"]
        s = self._image_info["shape"]
        return [s[0] + 2 * ps, s[1] + 2 * ps, s[-1]]

    @property
    def latent_shape(self):
        return list(self._layer_latent_shape)

    def preprocess(self, x):
        """vae.py:270-278: scale evidence to [-1, 1]."""
        if not self._image_info["binary"]:
            lo, hi = self._image_info["low"], self._image_info["high"]
            return (2.0 * (x - lo) / (hi - lo)) - 1.0
        return (2.0 * x) - 1.0

    def draw_noise(self, batch, device, generator=None, training=True):
        """The random inputs of one pass: eps[i] [B,H,W,z] per layer and, in
        training with noise_stddev > 0, the log-sigma noise (posterior, prior)."""
        H, W, z = self._layer_latent_shape
        L = self._num_layers
        rn = lambda: torch.randn(batch, H, W, z, device=device, generator=generator)
        eps = [rn() for _ in range(L)]
        noise = [None] * L
        use_q = training and self._posterior_hparams.noise_stddev > 0
        use_p = training and self._prior_hparams is not None and self._prior_hparams.noise_stddev > 0
        if use_q or use_p:
            noise = [(rn() if use_q else None, rn() if (use_p and i > 0) else None) for i in range(L)]
        return eps, noise

    def encode(self, inputs, training, num_samples=1, compute_q_logp=False, compute_p_logp=False,
               eps=None, noise=None):
        """Bi-directional inference, vae.py:282-371.  Returns c, h, kl, q_logp, p_logp."""
        L = self._num_layers
        if num_samples > 1:
            inputs = inputs.repeat(num_samples, 1, 1, 1)                                 # vae.py:305
        if eps is None:
            eps, noise = self.draw_noise(inputs.shape[0], inputs.device, training=training)
        if noise is None:
            noise = [None] * L
        e = [None] * (L + 1)
        t = inputs
        for layer in self._preproc_encoder:                                             # vae.py:310-312
            t = layer(t, training)
        e[-1] = t
        for i in range(L):                                                              # vae.py:315-320
            j = L - 1 - i
            e[j] = self._latent_layer[j].bottom_up(e[j + 1], training)
        h, c, kl = [None] * L, [None] * L, [None] * L
        q_logp, p_logp = [None] * L, [None] * L
        want_logp = compute_q_logp or compute_p_logp
        # The reference stacks e[i:] / c[:i] for every layer (vae.py:339-350).  Here every tensor that
        # several layers attend to is wrapped ONCE in a SharedSlot (values || key): the earliest consumer
        # owns its gradient, the later ones accumulate into it from their backward kernels.
        enc_att, dec_att = self._encoder_attention_hparams, self._decoder_attention_hparams
        if enc_att is not None:
            f = e[1].shape[-1] - enc_att.key_dim
            e_shared = [None] + [ops.SharedSlot(t, f) for t in e[1:]]
        c_shared = [None] * L
        for i in range(L):                                                              # vae.py:336-365
            if enc_att is not None:                                                     # vae.py:339-342
                condition = [e[i]] + [("own" if i == 0 else "acc", s) for s in e_shared[i + 1:]]
            else:
                condition = e[i]
            context = None
            if i > 0:
                if dec_att is not None:
                    if i > 1:
                        g, b = self._ctx_layer_norm[i]
                        c[i - 2] = ops.layer_norm(c[i - 2], g, b)                       # vae.py:347-348
                        c_shared[i - 2] = ops.SharedSlot(c[i - 2], c[i - 2].shape[-1] - dec_att.key_dim)
                    context = [("own" if k == i - 2 else "acc", c_shared[k]) for k in range(i - 1)] + [c[i - 1]]  # vae.py:350
                else:
                    context = c[i - 1]
            h[i], c[i], kl[i], q_logp[i], p_logp[i] = self._latent_layer[i].infer_top_down(
                condition, training, latent_condition=h[i - 1] if i > 0 else None, context=context,
                eps=eps[i], noise=noise[i], compute_logp=want_logp)
        if L == 1:
            return h, h, kl, q_logp, p_logp
        return c, h, kl, q_logp, p_logp

    def _decode_input(self, c, h, training):
        if self._num_layers > 1:
            if self._merge_decoder is not None:
                return self._merge_decoder([c[-1], h[-1]], training)                    # vae.py:250-251
            return [c[-1], h[-1]]
        return h[0]

    def forward(self, inputs, training=True, eps=None, noise=None):
        """NELBO terms [nll, sum(kl), nelbo], each [B] (vae.py:226-265)."""
        inputs = inputs.float()
        evidence = self.preprocess(inputs)
        c, h, kl, _, _ = self.encode(evidence, training, eps=eps, noise=noise)
        cc = self._decode_input(c, h, training)
        x = inputs if self._image_info["binary"] else evidence                          # vae.py:256-259
        nll = self._data_layer.decode(cc, x=x, training=training)
        kl_sum = kl[0]
        for k in kl[1:]:
            kl_sum = kl_sum + k
        return [nll, kl_sum, nll + kl_sum]                                              # vae.py:262-265

    def infer_worker_body(self, inputs, num_samples, eps=None):
        """One chunk of the importance-sampled estimate, vae.py:452-482: tiles the
        batch sample-major, returns logsumexp over the chunk's samples [B]."""
        enc_in = self.preprocess(inputs)
        rep = [num_samples] + [1] * (inputs.dim() - 1)
        inputs_t, enc_t = inputs.repeat(*rep), enc_in.repeat(*rep)                      # vae.py:458-460
        c, h, _, q_logp, p_logp = self.encode(enc_t, training=False, compute_q_logp=True,
                                              compute_p_logp=True, eps=eps)
        lw = p_logp[0] - q_logp[0]
        for j in range(1, self._num_layers):
            lw = lw + (p_logp[j] - q_logp[j])                                           # vae.py:465
        cc = self._decode_input(c, h, False)
        x = inputs_t if self._image_info["binary"] else enc_t
        lw = lw - self._data_layer.decode(cc, x=x, training=False)                      # vae.py:472-478
        return ops.is_logsumexp(lw, num_samples)                                        # vae.py:480-482

    @torch.no_grad()
    def predict_step(self, inputs, num_samples=None, workers=None):
        """Importance-sampled log-likelihood estimate [B], vae.py:502-539."""
        K = num_samples or self.eval_samples
        Wk = workers or self.eval_workers
        parts = []
        for i in range(Wk):                                                             # vae.py:524
            k = K // Wk if i != Wk - 1 else K - (K // Wk) * (Wk - 1)
            parts.append(self.infer_worker_body(inputs, k))
        stacked = torch.stack(parts, 0).reshape(-1)                                     # [Wk*B], chunk-major
        return ops.is_logsumexp(stacked, Wk) - math.log(K)                              # vae.py:539

    # ------------------------------------------------------------------
    def reg_loss(self):
        """Sum of the Keras regulariser losses (enters the training loss only)."""
        r = self._data_layer.reg_loss()
        for m in self._preproc_encoder:
            r = r + m.reg_loss()
        for m in self._latent_layer:
            r = r + m.reg_loss()
        if self._merge_decoder is not None:
            r = r + self._merge_decoder.reg_loss()
        for pl in self._ctx_layer_norm:
            for p in pl:
                r = r + KERAS_DEFAULT_L2 * p.square().sum()
        return r

    def count_params(self, verbose=False):
        """vae.py:543-570: trainable weights of the pre-processor, the merge cell, the
        latent layers and the data layer (the model-level ctx LayerNorms are NOT
        in that sum, exactly as in the reference)."""
        mods = list(self._preproc_encoder) + list(self._latent_layer) + [self._data_layer]
        if self._merge_decoder is not None:
            mods.append(self._merge_decoder)
        total = sum(p.numel() for m in mods for p in m.parameters() if p.requires_grad)
        if verbose:
            print("Total Params:")
            print(total / 1e6)
        return total
