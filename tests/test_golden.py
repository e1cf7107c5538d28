"""Pins the oracle (and through it every parity test) to vectors produced by the reference's OWN
source code, executed in float64 on the TensorFlow stand-in of tests/golden/tf_shim.py
(generator: tests/golden/make_golden.py, run where /root/reference exists).  CPU only."""
import numpy as np
import pytest
import torch

import helpers
from oracle import ops as R
from oracle.model import OracleVAE


@pytest.fixture(scope="module")
def ops_fix():
    z = np.load(helpers.GOLDEN_DIR + "/ref_ops.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def close(got, want, tol=1e-10):
    err = (got.double() - want).abs().max().item()
    assert err <= tol * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("tag", ["dw_a", "dw_b", "dw_c"])
def test_depthwise_attention_matches_reference_code(ops_fix, tag):
    q, k, v = (ops_fix[f"{tag}/{n}"] for n in "qkv")
    close(R.depthwise_attention_apply(q, k, v), ops_fix[f"{tag}/out"])
    close(R.depthwise_attention_apply(q, k, v, scale=0.7), ops_fix[f"{tag}/out_scale0.7"])


@pytest.mark.parametrize("tag", ["nl_a", "nl_b"])
def test_nonlocal_core_matches_reference_code(ops_fix, tag):
    f = lambda n: ops_fix[f"{tag}/{n}"]
    close(R.nonlocal_core(f("q"), f("k"), f("v"), f("x"), f("gamma"), f("scale")), f("y"))


@pytest.mark.parametrize("tag", ["gd_cifar", "gd_omni"])
def test_gaussian_sample_kl_logp_match_reference_code(ops_fix, tag):
    f = lambda n: ops_fix[f"{tag}/{n}"]
    qb, pb = tuple(f("post_bounds").tolist()), tuple(f("prior_bounds").tolist())
    xi, kl, lq, lp = R.variational_sample_kl(f("post_raw"), f("eps"), f("prior_raw"), qb, pb, compute_logp=True)
    close(xi, f("xi")); close(kl, f("kl")); close(lq, f("logq")); close(lp, f("logp"))
    xi0, kl0, _, _ = R.variational_sample_kl(f("post_raw"), f("eps0"), None, qb, pb)
    close(xi0, f("xi0")); close(kl0, f("kl0"))


def test_dlm_log_prob_matches_reference_code(ops_fix):
    close(R.dlm_log_prob(ops_fix["dlm/params"], ops_fix["dlm/x"]), ops_fix["dlm/logp"])
    close(R.dlm_log_prob(ops_fix["dlm/params_big"], ops_fix["dlm/x"]), ops_fix["dlm/logp_big"])


def test_bernoulli_log_prob_matches_reference_code(ops_fix):
    close(R.bernoulli_log_prob(ops_fix["bern/logits"], ops_fix["bern/x"], int(ops_fix["bern/crop"])),
          ops_fix["bern/logp"])


@pytest.mark.parametrize("flags,fname", [(helpers.SMALL_CIFAR, "ref_small_cifar.npz"),
                                         (helpers.SMALL_OMNIGLOT, "ref_small_omniglot.npz"),
                                         (helpers.SMALL_NVAE, "ref_small_nvae.npz")])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_whole_model_elbo_matches_reference_code(flags, fname, mode):
    """[nll, sum kl, nelbo] of the reference's DeepVAE.call on the reference's weights and draws."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    kw = parse_flags(**flags)
    model = DeepVAE(**kw)
    fix, sd = helpers.load_reference_golden(fname, model)
    training = mode == "train"
    has_noise = training and kw["posterior_hparams"].noise_stddev > 0
    eps, noise = helpers.golden_noise(fix, mode, kw["hparams"].num_layers, has_noise)
    o = OracleVAE(helpers.oracle_cfg(kw), sd, dtype=torch.float64)
    nll, kl, nelbo = o.forward(torch.from_numpy(fix["x"]), eps, training, noise)
    for got, name in ((nll, "nll"), (kl, "kl"), (nelbo, "nelbo")):
        want = torch.from_numpy(fix[f"{mode}/{name}"])
        rel = ((got - want).abs() / want.abs()).max().item()
        assert rel < 1e-9, (name, rel, got, want)


@pytest.mark.parametrize("flags,fname", [(helpers.SMALL_CIFAR, "ref_small_cifar.npz"),
                                         (helpers.SMALL_OMNIGLOT, "ref_small_omniglot.npz"),
                                         (helpers.SMALL_NVAE, "ref_small_nvae.npz")])
def test_importance_sampled_chunk_matches_reference_code(flags, fname):
    """One worker chunk of the importance-sampled estimator (DeepVAE.infer_worker_body, vae.py:452-482:
    sample-major tiling, per-layer log p - log q, minus nll, logsumexp over the samples) executed by the
    reference's own code on its weights and recorded draws."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    kw = parse_flags(**flags)
    model = DeepVAE(**kw)
    fix, sd = helpers.load_reference_golden(fname, model)
    eps = helpers.golden_is_draws(fix)
    assert len(eps) == kw["hparams"].num_layers
    o = OracleVAE(helpers.oracle_cfg(kw), sd, dtype=torch.float64)
    got = o.is_chunk(torch.from_numpy(fix["x"]), eps, int(fix["is/num_samples"]))
    want = torch.from_numpy(fix["is/logsumexp"])
    assert ((got - want).abs() / want.abs()).max().item() < 1e-9, (got, want)
