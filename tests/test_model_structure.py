"""Known-answer test: the trainable-parameter counts the reference publishes
(README.md:32,54,77,105), reproduced by the host model's count_params()
(reference vae.py:543-570).  These are the only golden numbers the reference has."""
import pytest

from deep_attentive_vi_b200.model import CIFAR10_16L, OMNIGLOT_15L, DeepVAE, parse_flags

NVAE_16L = dict(
    dataset="cifar10",
    hparams="layer_latent_shape=[16,16,20],num_layers=16,data_distribution=discretized_logistic_mixture",
    decoder_hparams="scale=[0,0],num_blocks=2,num_filters=128",
    encoder_hparams="scale=[0,0],num_blocks=2,num_filters=128",
    preproc_encoder_hparams="scale=[0,-2],num_blocks=2,num_filters=128",
    postproc_decoder_hparams="scale=[2,0],num_blocks=2,num_filters=128",
    posterior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5",
    prior_hparams="log_scale_upper_bound=5,log_scale_low_bound=-5.0",
)


@pytest.mark.parametrize("flags,millions,exact", [
    (CIFAR10_16L, "118.97", 118966391),      # README.md:32,54 ; SURVEY.md section 0 analytic count
    (OMNIGLOT_15L, "7.87", None),            # README.md:77
    (NVAE_16L, "79.21", None),               # README.md:105
])
def test_parameter_counts_match_reference_readme(flags, millions, exact):
    n = DeepVAE(**parse_flags(**flags)).count_params()
    assert f"{n / 1e6:.2f}" == millions
    if exact is not None:
        assert n == exact


def test_hparams_string_syntax():
    from deep_attentive_vi_b200.util.hparams import HParams
    h = HParams(a=1, b=2.0, c=True, d="x", e=HParams(k=3, q=4), f=[0])
    h.update_config("a=5,b=-1.0,c=False,d=swish,e=[k=8,q=9],f=[0,-2]")
    assert (h.a, h.b, h.c, h.d, h.e.k, h.e.q, h.f) == (5, -1.0, False, "swish", 8, 9, [0, -2])
    kw = parse_flags(**CIFAR10_16L)
    assert kw["prior_hparams"].log_scale_low_bound == -1.0
    assert kw["decoder_hparams"].nonlocop_hparams.key_dim == 32
    assert kw["hparams"].layer_latent_shape == [16, 16, 20]


def test_channel_plan_matches_survey():
    m = DeepVAE(**parse_flags(**CIFAR10_16L))
    sd = m.state_dict()
    assert sd["_latent_layer.0._decoder_const"].shape == (1, 16, 16, 296)
    assert sd["_latent_layer.1._merge_decoder_network._conv2d_layer.v"].shape == (316, 296, 1, 1)
    assert sd["_latent_layer.15._merge_decoder_network._conv2d_layer.v"].shape == (296, 296, 1, 1)
    assert sd["_latent_layer.1._merge_encoder_network._conv2d_layer_a.v"].shape == (276, 256, 1, 1)
    assert sd["_data_layer._data_dist._params_network._conv2d_layers.0.v"].shape == (50, 138, 3, 3)
    assert sd["_ctx_layer_norm.2.0"].shape == (296,)
