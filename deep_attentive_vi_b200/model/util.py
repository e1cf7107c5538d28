"""Hyper-parameter plumbing of the model -- mirror of the reference's
model/util.py (_DeepVAE) and of the flag parsing in run.py:59-85."""
from ..layers import LAYER, DepthWiseAttention
from ..stat_tools import DISTRIBUTION
from ..util.hparams import HParams

# run.py:45-50
image_info = {
    "cifar10": {"shape": (32, 32, 3), "high": 255, "low": 0, "binary": False, "pad_size": 0},
    "mnist": {"shape": (28, 28, 1), "high": 255, "low": 0, "binary": True, "pad_size": 2},
    "omniglot": {"shape": (28, 28, 1), "high": 1, "low": 0, "binary": True, "pad_size": 2},
}


def get_default_hparams():
    """model/util.py:158-174."""
    return HParams(posterior_type="gaussian_diag", prior_type="gaussian_diag",
                   data_distribution="discretized_logistic_mixture", num_layers=1,
                   layer_latent_shape=100, data_layer_type="conv2d", posterior_layer_type="conv2d",
                   prior_layer_type="conv2d", residual_var_layer=True, preproc_encoder_type="resnet",
                   encoder_type="resnet", postproc_decoder_type="resnet", decoder_type="resnet",
                   merge_encoder_type="conv2d_sum", merge_decoder_type="concat_conv2d", num_proc_blocks=1)


def get_default_network_hparams(h):
    """model/util.py:176-215."""
    d = lambda kind, s: LAYER[kind].get_default_hparams().update_config(s)
    res = "use_stem=False,activation=swish,use_batch_norm=True"
    return HParams(
        prior=d(h.prior_layer_type, "kernel_size=1,num_times=1,activation=elu"),
        posterior=d(h.posterior_layer_type, "kernel_size=3,num_times=1,activation="),
        init_posterior=d(h.posterior_layer_type, "kernel_size=3,num_times=2,activation=elu"),
        data=d(h.data_layer_type, "kernel_size=3,num_times=1,activation=elu"),
        preproc_encoder=d(h.preproc_encoder_type, "use_stem=True,activation=swish,use_batch_norm=True"),
        encoder=d(h.encoder_type, res),
        postproc_decoder=d(h.postproc_decoder_type, res),
        decoder=d(h.decoder_type, res),
    )


def parse_flags(dataset="cifar10", hparams="", prior_hparams="", posterior_hparams="",
                data_dist_hparams="", preproc_encoder_hparams="", encoder_hparams="",
                merge_encoder_hparams="", decoder_hparams="", postproc_decoder_hparams="",
                merge_decoder_hparams="", decoder_attention_hparams="", encoder_attention_hparams=""):
    """The flag strings of run.py:156-179 -> DeepVAE constructor kwargs (run.py:59-109)."""
    hp = get_default_hparams().update_config(hparams)
    net = lambda: get_default_network_hparams(hp)
    return dict(
        hparams=hp,
        data_dist_hparams=DISTRIBUTION[hp.data_distribution].get_default_hparams().update_config(data_dist_hparams),
        preproc_encoder_hparams=net().preproc_encoder.update_config(preproc_encoder_hparams),
        postproc_decoder_hparams=net().postproc_decoder.update_config(postproc_decoder_hparams),
        encoder_hparams=net().encoder.update_config(encoder_hparams),
        merge_encoder_hparams=LAYER[hp.merge_encoder_type].get_default_hparams().update_config(merge_encoder_hparams),
        decoder_hparams=net().decoder.update_config(decoder_hparams),
        merge_decoder_hparams=LAYER[hp.merge_decoder_type].get_default_hparams().update_config(merge_decoder_hparams),
        posterior_hparams=DISTRIBUTION[hp.posterior_type].get_default_hparams().update_config(posterior_hparams),
        prior_hparams=DISTRIBUTION[hp.prior_type].get_default_hparams().update_config(prior_hparams),
        image_info=dict(image_info[dataset]),
        decoder_attention_hparams=(DepthWiseAttention.get_default_hparams().update_config(decoder_attention_hparams)
                                   if decoder_attention_hparams else None),
        encoder_attention_hparams=(DepthWiseAttention.get_default_hparams().update_config(encoder_attention_hparams)
                                   if encoder_attention_hparams else None),
    )


# the two published Attentive-VAE configurations (reference README.md:38-48, 84-96)
CIFAR10_16L = dict(
    dataset="cifar10",
    hparams="layer_latent_shape=[16,16,20],num_layers=16,data_distribution=discretized_logistic_mixture",
    encoder_attention_hparams="key_dim=20,query_dim=20,use_layer_norm=True",
    decoder_attention_hparams="key_dim=20,query_dim=20,use_layer_norm=True",
    decoder_hparams="scale=[0,0],use_nonlocal=True,nonlocop_hparams=[key_dim=32,query_dim=32],num_blocks=2,num_filters=128",
    encoder_hparams="scale=[0,0],use_nonlocal=True,nonlocop_hparams=[key_dim=32,query_dim=32],num_blocks=2,num_filters=128",
    preproc_encoder_hparams="scale=[0,-2],use_nonlocal=True,nonlocop_hparams=[key_dim=32,query_dim=32],num_blocks=2,num_filters=128",
    postproc_decoder_hparams="scale=[2,0],num_blocks=2,num_filters=128",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5,noise_stddev=0.001",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-1.0,noise_stddev=0.001",
)

OMNIGLOT_15L = dict(
    dataset="omniglot",
    hparams="layer_latent_shape=[16,16,20],num_layers=15,num_proc_blocks=1,data_distribution=bernoulli",
    decoder_attention_hparams="key_dim=8,query_dim=8,use_layer_norm=True",
    encoder_attention_hparams="key_dim=8,query_dim=8,use_layer_norm=False",
    merge_encoder_hparams="use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8]",
    merge_decoder_hparams="use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8]",
    decoder_hparams="scale=[0,0],num_blocks=2,num_filters=32",
    encoder_hparams="scale=[0,0],num_blocks=2,num_filters=32",
    preproc_encoder_hparams="scale=[0,0,-2],num_nodes=2,num_blocks=3,num_filters=32",
    postproc_decoder_hparams="scale=[2,0,0],num_nodes=2,use_nonlocal=True,nonlocop_hparams=[key_dim=8,query_dim=8],num_blocks=3,num_filters=32",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
)
