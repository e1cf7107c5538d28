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
