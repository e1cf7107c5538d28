"""CPU oracle of the whole Deep Attentive VAE forward (torch CPU, any dtype).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY PINNED: the
reference's stack cannot be installed here and it publishes no golden vectors, so
the whole-model [nll, sum kl, nelbo] of three small configurations (train / eval
mode, one importance-sampled chunk) are pinned to fixtures produced by executing the
reference's own source on a float64 TensorFlow stand-in (tests/golden/make_golden.py,
tests/golden/ref_small_*.npz; tests/test_golden.py, 1e-9 relative); the published
parameter counts (README.md:32,77,105) are checked in tests/test_model_structure.py.

An independent, functional, op-for-op restatement of the reference's model
(model/vae.py, model/variational_layer.py, model/data_layer.py, layers/resnet.py,
layers/weight_norm.py, layers/attention.py): it materialises every tf.stack /
tf.split / unstack-replace-stack the reference does and evaluates the hot-path
functions through ``oracle/ops.py``.  It shares NO code with the product
package; it only consumes the product model's ``state_dict()`` (weights by
name) and a small plain-dict description of the configuration.

cfg keys: L, z, binary, low, high, pad, dist ('dlm'|'bernoulli'),
  enc_att / dec_att: None or dict(d=key_dim, ln=bool),
  preproc_scales, enc_scales, dec_scales, postproc_scales: lists of ints,
  post_bounds, prior_bounds: (lo, hi).
"""
import torch
import torch.nn.functional as F

from . import ops as R


class OracleVAE:
    def __init__(self, cfg, state_dict, dtype=torch.float64):
        self.cfg = cfg
        self.sd = {k: (v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu())
                   for k, v in state_dict.items()}
        self.dtype = dtype

    # ---- cells (layers/resnet.py, layers/weight_norm.py) -------------------
    def has(self, key):
        return key in self.sd

    def wn_conv(self, p, x, stride=1, padding="same"):
        """WeightNorm(Conv2D): kernel = l2_normalize(v) * exp(log_g)
        (weight_norm.py:126-135); TF 'same' padding pads bottom/right first."""
        v = self.sd[p + ".v"]
        w = v
        if self.has(p + ".log_g"):
            ss = (v * v).sum(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12)
            w = v * torch.rsqrt(ss) * torch.exp(self.sd[p + ".log_g"]).view(-1, 1, 1, 1)
        b = self.sd.get(p + ".bias")
        k = v.shape[-1]
        xc = x.permute(0, 3, 1, 2)
        if padding == "same":
            def pads(n):
                out = -(-n // stride)
                tot = max((out - 1) * stride + k - n, 0)
                return tot // 2, tot - tot // 2
            (pt, pb), (pl, pr) = pads(x.shape[1]), pads(x.shape[2])
            xc = F.pad(xc, (pl, pr, pt, pb))
        return F.conv2d(xc, w, b, stride=stride).permute(0, 2, 3, 1)

    def bn(self, p, x, training):
        """Keras BatchNormalization(eps=1e-5): batch statistics, biased variance."""
        g, b = self.sd[p + ".bn.weight"], self.sd[p + ".bn.bias"]
        if training:
            mean = x.mean(dim=(0, 1, 2))
            var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        else:
            mean, var = self.sd[p + ".bn.running_mean"], self.sd[p + ".bn.running_var"]
        return (x - mean) * torch.rsqrt(var + 1e-5) * g + b

    def nonlocal_block(self, p, x):
        """NonlocalResNetBlock.call (att.py:285-389), use_layer_norm=False."""
        q = self.wn_conv(p + ".query_conv", x)                       # att.py:310
        k = self.wn_conv(p + ".key_conv", x)                         # att.py:346
        v = self.wn_conv(p + ".value_conv", x)                       # att.py:362
        scale = self.sd.get(p + "._alignment_function.scale")
        return R.nonlocal_core(q, k, v, x, self.sd[p + ".gamma"], scale)

    @staticmethod
    def act(name, x):
        if name == "swish":
            return x * torch.sigmoid(x)
        if name == "elu":
            return F.elu(x)
        if name == "":
            return x
        raise ValueError(name)

    def conv2d_cell(self, p, x, training, activation, stride=1):
        """Conv2D.call (res.py:613-633): BN -> [nonlocal] -> (act -> conv) x num_times."""
        if self.has(p + "._batch_norm_layer.bn.weight"):
            x = self.bn(p + "._batch_norm_layer", x, training)
        if self.has(p + "._nonlocal_layer.gamma"):
            x = self.nonlocal_block(p + "._nonlocal_layer", x)
        i = 0
        while self.has(f"{p}._conv2d_layers.{i}.v"):
            x = self.wn_conv(f"{p}._conv2d_layers.{i}", self.act(activation, x), stride=stride)
            i += 1
        return x

    def factorized_reduce(self, p, x):
        """FactorizedReduce.call (res.py:125-135)."""
        o = x * torch.sigmoid(x)
        views = [o, o[:, 1:, 1:, :], o[:, :, 1:, :], o[:, 1:, :, :]]
        return torch.cat([self.wn_conv(f"{p}.convs.{i}", v, stride=2, padding="valid")
                          for i, v in enumerate(views)], dim=-1)

    def resnet_block(self, p, x, training, scale, activation):
        """ResNetBlock.call (res.py:273-286)."""
        if scale == 0:
            shortcut = x
        elif scale < 0:
            shortcut = self.factorized_reduce(p + "._shortcut", x)
        else:
            size = (int(x.shape[1] * scale), int(x.shape[2] * scale))
            up = F.interpolate(x.permute(0, 3, 1, 2), size=size, mode="bilinear",
                               align_corners=True).permute(0, 2, 3, 1)              # res.py:232
            shortcut = self.conv2d_cell(p + "._shortcut", up, training, "")
        t = x
        i = 0
        while self.has(f"{p}._nodes.{i}._conv2d_layers.0.v"):
            if i == 0 and scale > 0:
                size = (int(x.shape[1] * scale), int(x.shape[2] * scale))
                t = F.interpolate(t.permute(0, 3, 1, 2), size=size, mode="nearest").permute(0, 2, 3, 1)  # res.py:252
            stride = abs(scale) if (i == 0 and scale < 0) else 1
            t = self.conv2d_cell(f"{p}._nodes.{i}", t, training, activation, stride=stride)
            i += 1
        return 0.1 * t + shortcut                                                   # res.py:284-286

    def resnet(self, p, x, training, scales, activation="swish"):
        """ResNet.call (res.py:739-753)."""
        if self.has(p + "._stem_layer.v"):
            x = self.wn_conv(p + "._stem_layer", x)
        for k, s in enumerate(scales):
            x = self.resnet_block(f"{p}._residual_blocks.{k}", x, training, s, activation)
        return x

    def concat_conv(self, p, a, b):
        """ConcatConv2D.call (res.py:404-419)."""
        r = self.wn_conv(p + "._conv2d_layer", torch.cat([a, b], dim=-1))
        if self.has(p + "._nonlocal_layer.gamma"):
            r = self.nonlocal_block(p + "._nonlocal_layer", r)
        return r

    def conv_sum(self, p, a, b):
        """Conv2DSum.call (res.py:496-517)."""
        b = self.wn_conv(p + "._conv2d_layer", b)
        if self.has(p + "._conv2d_layer_a.v"):
            a = self.wn_conv(p + "._conv2d_layer_a", a)
        r = a + b
        if self.has(p + "._nonlocal_layer.gamma"):
            r = self.nonlocal_block(p + "._nonlocal_layer", r)
        return r

    def ln(self, p_gamma, p_beta):
        return (self.sd[p_gamma], self.sd[p_beta])

    # ---- one variational layer, top-down (vl.py:338-550) --------------------
    def layer_top_down(self, i, e_stack, c_list, h_prev, eps, training, noise, compute_logp):
        cfg, sd = self.cfg, self.sd
        p = f"_latent_layer.{i}"
        L = cfg["L"]
        B = e_stack.shape[0] if e_stack.dim() == 5 else e_stack.shape[0]
        dec_att = cfg["dec_att"] if i > 0 else None
        enc_att = cfg["enc_att"]
        key = None
        if i == 0:
            new_context = sd[p + "._decoder_const"].repeat(B, 1, 1, 1)              # vl.py:291-293
        else:
            if dec_att is not None:
                context = torch.stack(c_list, dim=1)                                 # vae.py:350
                nv = context.shape[-1] - dec_att["d"]                                # vl.py:358
                values, keys = context[..., :nv], context[..., nv:]                  # vl.py:359
                ctx_in = values[:, -1]                                               # vl.py:360
            else:
                ctx_in = c_list[-1]
            prev = self.concat_conv(p + "._merge_decoder_network", ctx_in, h_prev)   # vl.py:282
            new_context = self.resnet(p + "._decoder_network", prev, training, cfg["dec_scales"])  # vl.py:288
        if dec_att is not None:
            d = dec_att["d"]
            if i < L - 1:                                                            # vl.py:375-380
                g = new_context.shape[-1] - 2 * d
                new_context, key, query = new_context[..., :g], new_context[..., g:g + d], new_context[..., g + d:]
            else:
                g = new_context.shape[-1] - d
                new_context, query = new_context[..., :g], new_context[..., g:]
        enc_query = new_context[..., :enc_att["d"]] if enc_att is not None else None  # vl.py:385-386
        if dec_att is not None:
            new_context, prior_condition = R.decoder_attention_glue(
                new_context, query, keys, values,
                self.ln(p + ".ctx_ln_gamma", p + ".ctx_ln_beta") if dec_att["ln"] else None,
                self.ln(p + ".attn_ln_gamma", p + ".attn_ln_beta") if dec_att["ln"] else None,
                self.ln(p + ".attn_ln2_gamma", p + ".attn_ln2_beta") if dec_att["ln"] else None,
                sd[p + "._gamma"], use_layer_norm=dec_att["ln"],
                scale=sd.get(p + "._decoder_attention._alignment_function.scale"))   # vl.py:390-412
        else:
            prior_condition = new_context
        n_post, n_prior = noise if noise is not None else (None, None)
        prior_params = None
        if i > 0:
            prior_params = self.conv2d_cell(p + "._prior_network._params_network", prior_condition,
                                            training, "elu")                         # vl.py:463
            if n_prior is not None:                                                  # gd.py:165-166
                zc = prior_params.shape[-1] // 2
                prior_params = torch.cat([prior_params[..., :zc],
                                          prior_params[..., zc:] + cfg["prior_noise_std"] * n_prior], dim=-1)
        if enc_att is not None:
            f = e_stack.shape[-1] - enc_att["d"]
            values_e, keys_e = e_stack[..., :f], e_stack[..., f:]                    # vl.py:472
            condition = values_e[:, 0]                                               # vl.py:473
        else:
            condition = e_stack
        if i > 0:
            condition = self.conv_sum(p + "._merge_encoder_network", condition, prior_condition)  # vl.py:477
        if enc_att is not None:
            new_key = None
            if i > 0:
                condition, new_key = condition[..., :f], condition[..., f:]          # vl.py:483
            condition = R.encoder_attention_glue(
                condition, new_key, enc_query, keys_e, values_e,
                self.ln(p + ".enc_ln_gamma", p + ".enc_ln_beta") if enc_att["ln"] else None,
                self.ln(p + ".cond_ln_gamma", p + ".cond_ln_beta") if enc_att["ln"] else None,
                sd[p + "._encoder_gamma"], use_layer_norm=enc_att["ln"], replace_first=(i > 0),
                scale=sd.get(p + "._encoder_attention._alignment_function.scale"))   # vl.py:480-507
        post_act = "elu" if i == 0 else ""                                           # model/util.py:183,185
        post = self.conv2d_cell(p + "._posterior_network._params_network", condition, training, post_act)
        if n_post is not None:
            zc = post.shape[-1] // 2
            post = torch.cat([post[..., :zc], post[..., zc:] + cfg["post_noise_std"] * n_post], dim=-1)
        xi, kl, q_logp, p_logp = R.variational_sample_kl(
            post, eps, prior_params, cfg["post_bounds"], cfg["prior_bounds"], compute_logp)  # vl.py:511-542
        if dec_att is None:
            ctx_out = prior_condition                                                # vl.py:545-546
        elif i < L - 1:
            ctx_out = torch.cat([new_context, key], dim=-1)                          # vl.py:547-548
        else:
            ctx_out = new_context
        return xi, ctx_out, kl, q_logp, p_logp

    # ---- whole model (vae.py:226-371) --------------------------------------
    def preprocess(self, x):
        c = self.cfg
        if not c["binary"]:
            return (2.0 * (x - c["low"]) / (c["high"] - c["low"])) - 1.0             # vae.py:277
        return (2.0 * x) - 1.0

    def encode(self, evidence, eps, training=True, noise=None, compute_logp=False):
        cfg = self.cfg
        L = cfg["L"]
        e = [None] * (L + 1)
        e[-1] = self.resnet("_preproc_encoder.0", evidence, training, cfg["preproc_scales"])  # vae.py:310-312
        for i in range(L):                                                           # vae.py:315-320
            j = L - 1 - i
            e[j] = self.resnet(f"_latent_layer.{j}._encoder_network", e[j + 1], training, cfg["enc_scales"])
        h, c, kl, ql, pl = [None] * L, [None] * L, [None] * L, [None] * L, [None] * L
        for i in range(L):                                                           # vae.py:336-365
            cond = torch.stack(e[i:], dim=1) if cfg["enc_att"] is not None else e[i]  # vae.py:339-342
            if i > 1 and cfg["dec_att"] is not None:
                c[i - 2] = R.layer_norm(c[i - 2], self.sd[f"_ctx_layer_norm.{i}.0"],
                                        self.sd[f"_ctx_layer_norm.{i}.1"])            # vae.py:347-348
            h[i], c[i], kl[i], ql[i], pl[i] = self.layer_top_down(
                i, cond, c[:i], h[i - 1] if i > 0 else None, eps[i], training,
                noise[i] if noise is not None else None, compute_logp)
        return c, h, kl, ql, pl

    def decode_nll(self, c, h, inputs, evidence, training):
        cfg = self.cfg
        cc = self.concat_conv("_merge_decoder", c[-1], h[-1])                        # vae.py:250-251
        cc = self.resnet("_data_layer._preproc_layer.0", cc, training, cfg["postproc_scales"])  # data_layer.py:173-175
        params = self.conv2d_cell("_data_layer._data_dist._params_network", cc, training, "elu")
        if cfg["dist"] == "dlm":
            return -R.dlm_log_prob(params, evidence)                                 # data_layer.py:179-181
        return -R.bernoulli_log_prob(params, inputs, cfg["pad"])

    def forward(self, inputs, eps, training=True, noise=None):
        """[nll, sum(kl), nelbo] (vae.py:226-265)."""
        inputs = inputs.to(self.dtype)
        eps = [t.to(self.dtype) for t in eps]
        if noise is not None:
            noise = [None if n is None else tuple(None if t is None else t.to(self.dtype) for t in n)
                     for n in noise]
        evidence = self.preprocess(inputs)
        c, h, kl, _, _ = self.encode(evidence, eps, training, noise)
        nll = self.decode_nll(c, h, inputs, evidence, training)
        return R.nelbo(nll, kl)

    def is_chunk(self, inputs, eps, num_samples):
        """infer_worker_body (vae.py:452-482) for one chunk of `num_samples`."""
        inputs = inputs.to(self.dtype)
        eps = [t.to(self.dtype) for t in eps]
        rep = [num_samples] + [1] * (inputs.dim() - 1)
        inputs_t = inputs.repeat(*rep)
        ev_t = self.preprocess(inputs).repeat(*rep)
        c, h, _, ql, pl = self.encode(ev_t, eps, training=False, compute_logp=True)
        nll = self.decode_nll(c, h, inputs_t, ev_t, False)
        return R.is_log_likelihood(pl, ql, nll, num_samples)
