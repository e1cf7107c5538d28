"""Shared test helpers: small model configurations and the plain-dict
description the oracle model consumes."""
import torch

SMALL_CIFAR = dict(
    dataset="cifar10",
    hparams="layer_latent_shape=[16,16,4],num_layers=4,data_distribution=discretized_logistic_mixture",
    encoder_attention_hparams="key_dim=4,query_dim=4,use_layer_norm=True",
    decoder_attention_hparams="key_dim=4,query_dim=4,use_layer_norm=True",
    decoder_hparams="scale=[0,0],use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8],num_blocks=2,num_filters=8",
    encoder_hparams="scale=[0,0],use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8],num_blocks=2,num_filters=8",
    preproc_encoder_hparams="scale=[0,-2],use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8],num_blocks=2,num_filters=8",
    postproc_decoder_hparams="scale=[2,0],num_blocks=2,num_filters=8",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5,noise_stddev=0.001",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-1.0,noise_stddev=0.001",
)

SMALL_OMNIGLOT = dict(
    dataset="omniglot",
    hparams="layer_latent_shape=[16,16,4],num_layers=3,num_proc_blocks=1,data_distribution=bernoulli",
    decoder_attention_hparams="key_dim=8,query_dim=8,use_layer_norm=True",
    encoder_attention_hparams="key_dim=8,query_dim=8,use_layer_norm=False",
    merge_encoder_hparams="use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8]",
    merge_decoder_hparams="use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8]",
    decoder_hparams="scale=[0,0],num_blocks=2,num_filters=8",
    encoder_hparams="scale=[0,0],num_blocks=2,num_filters=8",
    preproc_encoder_hparams="scale=[0,0,-2],num_nodes=2,num_blocks=3,num_filters=8",
    postproc_decoder_hparams="scale=[2,0,0],num_nodes=2,use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8],num_blocks=3,num_filters=8",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
)


# NVAE-like baseline of the reference README (no depth-wise attention, no nonlocal blocks)
SMALL_NVAE = dict(
    dataset="cifar10",
    hparams="layer_latent_shape=[16,16,4],num_layers=3,data_distribution=discretized_logistic_mixture",
    decoder_hparams="scale=[0,0],num_blocks=2,num_filters=8",
    encoder_hparams="scale=[0,0],num_blocks=2,num_filters=8",
    preproc_encoder_hparams="scale=[0,-2],num_blocks=2,num_filters=8",
    postproc_decoder_hparams="scale=[2,0],num_blocks=2,num_filters=8",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5.0",
)


def oracle_cfg(kw):
    """DeepVAE constructor kwargs (HParams objects) -> the oracle's plain dict."""
    hp, info = kw["hparams"], kw["image_info"]
    att = lambda a: None if a is None else dict(d=a.key_dim, ln=a.use_layer_norm)
    sc = lambda h: list(h.scale) if isinstance(h.scale, list) else [h.scale] * h.num_blocks
    return dict(
        L=hp.num_layers, z=hp.layer_latent_shape[-1], binary=info["binary"], low=info["low"],
        high=info["high"], pad=info["pad_size"],
        dist="dlm" if hp.data_distribution == "discretized_logistic_mixture" else "bernoulli",
        enc_att=att(kw["encoder_attention_hparams"]), dec_att=att(kw["decoder_attention_hparams"]),
        preproc_scales=sc(kw["preproc_encoder_hparams"]), enc_scales=sc(kw["encoder_hparams"]),
        dec_scales=sc(kw["decoder_hparams"]), postproc_scales=sc(kw["postproc_decoder_hparams"]),
        post_bounds=(kw["posterior_hparams"].log_scale_low_bound, kw["posterior_hparams"].log_scale_upper_bound),
        prior_bounds=(kw["prior_hparams"].log_scale_low_bound, kw["prior_hparams"].log_scale_upper_bound),
        post_noise_std=max(kw["posterior_hparams"].noise_stddev, 0.0),
        prior_noise_std=max(kw["prior_hparams"].noise_stddev, 0.0),
    )


def randomize_zero_init_scalars(model, seed=0):
    """gamma / enc_gamma / nonlocal gamma are zero-initialised in the reference
    (vl.py:226-245, att.py:256-261), which would switch every attention path off;
    parity runs give them N(0, 0.5^2) values (and non-trivial LN weights, scales)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            leaf = name.split(".")[-1]
            if leaf in ("gamma", "_gamma", "_encoder_gamma") and p.dim() == 0:
                p.copy_(torch.randn((), generator=g) * 0.5)
            elif "ln_gamma" in leaf or "ln2_gamma" in leaf or (name.startswith("_ctx_layer_norm") and leaf == "0"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif "ln_beta" in leaf or "ln2_beta" in leaf or (name.startswith("_ctx_layer_norm") and leaf == "1"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            elif leaf == "scale" and p.dim() == 0:
                p.copy_(1.0 + 0.3 * torch.randn((), generator=g))


def synthetic_images(kw, batch, seed=0):
    """SURVEY.md 8d synthetic inputs: CIFAR randint(0,256); OMNIGLOT Bernoulli(0.5)
    28x28 zero-padded to 32x32."""
    g = torch.Generator().manual_seed(seed)
    info = kw["image_info"]
    h, w, c = info["shape"]
    if info["binary"]:
        x = (torch.rand(batch, h, w, c, generator=g) < 0.5).float()
        p = info["pad_size"]
        return torch.nn.functional.pad(x, (0, 0, p, p, p, p))
    return torch.randint(0, 256, (batch, h, w, c), generator=g).float()


# ---------------------------------------------------------------------------
# golden fixtures produced by the reference's own code (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------
import os as _os
import re as _re

GOLDEN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")

_LN_NAMES = {"_cond_attention_layer_norm_layer": "cond_ln", "_encoder_attention_layer_norm_layer": "enc_ln",
             "_attention_layer_norm_layer": "attn_ln", "_attention_layer_norm_layer2": "attn_ln2",
             "_ctx_layer_norm_layer": "ctx_ln"}


def reference_key_to_host(path, name):
    """Attribute path + variable name of the reference (Keras) model -> (host state_dict key, is_conv_kernel)."""
    path = _re.sub(r"_shortcut\.conv_(\d)", lambda m: f"_shortcut.convs.{int(m.group(1)) - 1}", path)
    path = path.replace(".layers.1.", ".")        # Sequential([Lambda(resize), Conv2D]) of up-sampling blocks
    leaf = path.rsplit(".", 1)[-1]
    if path.endswith(".layer"):                                   # the wrapped Keras Conv2D of a WeightNorm
        base = path[:-len(".layer")]
        return (f"{base}.v", True) if name == "kernel" else (f"{base}.bias", False)
    if name == "g":
        return f"{path}.log_g", False
    if leaf == "_batch_norm_layer":
        return f"{path}.bn." + {"gamma": "weight", "beta": "bias", "moving_mean": "running_mean",
                                "moving_variance": "running_var"}[name], False
    if leaf in _LN_NAMES:
        return f"{path.rsplit('.', 1)[0]}.{_LN_NAMES[leaf]}_{name}", False
    if _re.fullmatch(r"_ctx_layer_norm\.\d+", path):
        return f"{path}.{0 if name == 'gamma' else 1}", False
    if _re.fullmatch(r"_latent_layer\.\d+", path):
        return f"{path}." + {"gamma": "_gamma", "enc_gamma": "_encoder_gamma"}[name], False
    if leaf == "_decoder_network" and name == "constant":
        return f"{path.rsplit('.', 1)[0]}._decoder_const", False
    return f"{path}.{name}", False                                # nonlocal gamma, alignment scale ...


def load_reference_golden(fname, model):
    """Loads tests/golden/<fname>: returns (fixture dict, host state_dict built from the reference's
    weights).  Every floating-point entry of the host model must be covered."""
    import numpy as np
    z = np.load(_os.path.join(GOLDEN_DIR, fname))
    want = {k: v for k, v in model.state_dict().items()}
    sd = {}
    for key in z.files:
        if not key.startswith("w/"):
            continue
        _, path, name = key.split("/")
        hk, is_kernel = reference_key_to_host(path, name)
        w = torch.from_numpy(z[key])
        if is_kernel:
            w = w.permute(3, 2, 0, 1).contiguous()                # TF [kh,kw,in,out] -> torch [out,in,kh,kw]
        assert hk in want, f"reference weight {path}/{name} -> {hk} has no counterpart in the host model"
        assert tuple(want[hk].shape) == tuple(w.shape), (hk, tuple(want[hk].shape), tuple(w.shape))
        sd[hk] = w
    missing = [k for k, v in want.items() if k not in sd and v.is_floating_point()]
    assert not missing, f"host parameters without a reference weight: {missing[:8]}"
    for k, v in want.items():
        sd.setdefault(k, v)
    return {k: z[k] for k in z.files if not k.startswith("w/")}, sd


def golden_noise(fix, mode, L, training_noise):
    """The recorded normal draws in call order -> (eps list, noise list) of DeepVAE.forward:
    layer 0 draws [posterior noise], eps; layer i > 0 draws [prior noise, posterior noise], eps
    (vl.py:463 before vl.py:511)."""
    draws = [(k.rsplit("/", 1)[-1], torch.from_numpy(fix[k])) for k in sorted(fix) if k.startswith(f"{mode}/draw/")]
    it = iter(draws)
    eps, noise = [], []
    for i in range(L):
        pn = qn = None
        if training_noise:
            if i > 0:
                tag, pn = next(it); assert tag == "gaussian_noise"
            tag, qn = next(it); assert tag == "gaussian_noise"
        tag, e = next(it); assert tag == "mvn_sample"
        eps.append(e[0])
        noise.append((qn, pn))
    assert next(it, None) is None
    return eps, (noise if training_noise else None)


def golden_is_draws(fix):
    """The recorded eps draws of the importance-sampled chunk (one [S*B,H,W,z] tensor per layer, in call order)."""
    keys = sorted(k for k in fix if k.startswith("is/draw/"))
    assert all(k.endswith("mvn_sample") for k in keys)
    return [torch.from_numpy(fix[k])[0] for k in keys]
