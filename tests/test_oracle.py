"""The oracle checked against closed-form invariants and against an independent
numpy restatement of the DLM channel map.  (The pins proper -- fixtures produced by
running the reference's own source -- are in tests/test_golden.py, see
oracle/__init__.py.)"""
import math

import numpy as np
import torch

from oracle import ops as ref

torch.manual_seed(0)
D = torch.float64


def test_softmax_rows_sum_to_one_and_n1_identity():
    q = torch.randn(2, 3, 4, 8, dtype=D)
    k = torch.randn(2, 5, 3, 4, 8, dtype=D)
    a = ref.softmax_alignment(q.unsqueeze(-1), k.permute(0, 2, 3, 4, 1), 0.7)
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)))
    v = torch.randn(2, 1, 3, 4, 12, dtype=D)
    out = ref.depthwise_attention_apply(q, k[:, :1], v)
    assert torch.allclose(out, v[:, 0])


def test_depthwise_matches_einsum():
    q = torch.randn(2, 3, 4, 8, dtype=D)
    k = torch.randn(2, 5, 3, 4, 8, dtype=D)
    v = torch.randn(2, 5, 3, 4, 12, dtype=D)
    s = torch.einsum("bhwd,bnhwd->bhwn", q, k) * 1.3
    want = torch.einsum("bhwn,bnhwc->bhwc", torch.softmax(s, -1), v)
    assert torch.allclose(ref.depthwise_attention_apply(q, k, v, 1.3), want)


def test_nonlocal_gamma_zero_is_identity_and_matches_sdpa():
    B, W, H, d, C = 2, 4, 4, 8, 12
    q, k = torch.randn(B, W, H, d, dtype=D), torch.randn(B, W, H, d, dtype=D)
    v, x = torch.randn(B, W, H, C, dtype=D), torch.randn(B, W, H, C, dtype=D)
    assert torch.equal(ref.nonlocal_core(q, k, v, x, 0.0, 1.0), x)
    y = ref.nonlocal_core(q, k, v, x, 0.5, 0.9)
    att = torch.softmax(q.reshape(B, -1, d) @ k.reshape(B, -1, d).transpose(1, 2) * 0.9, -1)
    want = 0.5 * (att @ v.reshape(B, -1, C)).reshape(B, W, H, C) + x
    assert torch.allclose(y, want)


def test_layer_norm_and_gelu_match_torch():
    x = torch.randn(5, 7, 16, dtype=D)
    g, b = torch.randn(16, dtype=D), torch.randn(16, dtype=D)
    assert torch.allclose(ref.layer_norm(x, g, b), torch.nn.functional.layer_norm(x, (16,), g, b, 1e-5))
    assert torch.allclose(ref.gelu(x), torch.nn.functional.gelu(x, approximate="tanh"))


def test_bound_param_forms():
    x = torch.linspace(-20, 20, 101, dtype=D)
    t = ref.bound_param(x, -5.0, 5.0)
    assert torch.allclose(t, 5 * torch.tanh(x / 10)) and t.abs().max() < 5
    s = ref.bound_param(x, -1.0, 5.0)
    assert s.min() == -1.0 and s.max() == 5.0
    assert torch.all(s[1:] >= s[:-1])
    mid = ref.bound_param(torch.tensor([2.0], dtype=D), -1.0, 5.0)   # t = .5 -> S = .5
    assert torch.allclose(mid, torch.tensor([2.0], dtype=D))


def test_kl_properties_and_torch_distributions():
    B, W, H, z = 2, 3, 3, 4
    post, prior = torch.randn(B, W, H, 2 * z, dtype=D), torch.randn(B, W, H, 2 * z, dtype=D)
    eps = torch.randn(B, W, H, z, dtype=D)
    xi, kl, lq, lp = ref.variational_sample_kl(post, eps, prior, (-5, 5), (-1, 5), compute_logp=True)
    assert torch.all(kl >= 0)
    # zero residual with identical bounds: q == p, KL == 0
    _, kl0, _, _ = ref.variational_sample_kl(torch.zeros_like(post), eps, prior, (-5, 5), (-5, 5))
    assert torch.allclose(kl0, torch.zeros_like(kl0), atol=1e-12)
    # against torch.distributions
    mu_p, ls_p = prior[..., :z], prior[..., z:]
    loc_p, sc_p = ref.gaussian_dist(mu_p, ls_p, -1, 5)
    loc_q, sc_q = ref.gaussian_dist(mu_p + post[..., :z], ls_p + post[..., z:], -5, 5)
    qd, pd = torch.distributions.Normal(loc_q, sc_q), torch.distributions.Normal(loc_p, sc_p)
    assert torch.allclose(kl, torch.distributions.kl_divergence(qd, pd).sum((1, 2, 3)))
    assert torch.allclose(lq, qd.log_prob(xi).sum((1, 2, 3)))
    assert torch.allclose(lp, pd.log_prob(xi).sum((1, 2, 3)))
    # layer 0: standard normal prior
    _, kl_n, _, lp_n = ref.variational_sample_kl(post, eps, None, (-5, 5), (-5, 5), compute_logp=True)
    loc, sc = ref.gaussian_dist(post[..., :z], post[..., z:], -5, 5)
    want = torch.distributions.kl_divergence(torch.distributions.Normal(loc, sc),
                                             torch.distributions.Normal(0.0, 1.0)).sum((1, 2, 3))
    assert torch.allclose(kl_n, want)


def _dlm_by_channel_map(params, x):
    """Independent numpy restatement using the channel table of SURVEY.md R7."""
    p = params.numpy(); xv = x.numpy()
    B, W, H, _ = p.shape
    loc_ch = {0: [8, 9, 10, 11, 12], 1: [21, 22, 26, 27, 28], 2: [37, 38, 39, 40, 44]}
    ls_ch = {0: [13, 17, 18, 19, 20], 1: [29, 30, 31, 35, 36], 2: [45, 46, 47, 48, 49]}
    out = np.zeros(B)
    sp = lambda t: np.logaddexp(0, t)
    sg = lambda t: 1 / (1 + np.exp(-t))
    for b in range(B):
        for w in range(W):
            for h in range(H):
                r = p[b, w, h]; xx = xv[b, w, h]
                logits = r[:5] - (r[:5].max() + np.log(np.exp(r[:5] - r[:5].max()).sum()))
                S = np.zeros(5)
                for k in range(5):
                    c0, c1, c2 = np.tanh(r[5 + 9 * k: 8 + 9 * k])
                    means = [r[loc_ch[0][k]], r[loc_ch[1][k]] + c0 * xx[0],
                             r[loc_ch[2][k]] + c1 * xx[0] + c2 * xx[1]]
                    tot = 0.0
                    for c in range(3):
                        ls = max(r[ls_ch[c][k]], -7.0); inv = np.exp(-ls); u = xx[c] - means[c]
                        pl, mn, mid = inv * (u + 1 / 255), inv * (u - 1 / 255), inv * u
                        if xx[c] < -0.999: lp = pl - sp(pl)
                        elif xx[c] > 0.99: lp = -sp(mn)
                        elif sg(pl) - sg(mn) > 1e-5: lp = np.log(max(sg(pl) - sg(mn), 1e-12))
                        else: lp = mid - ls - 2 * sp(mid) - np.log(127.5)
                        tot += lp
                    S[k] = tot + logits[k]
                out[b] += S.max() + np.log(np.exp(S - S.max()).sum())
    return out


def test_dlm_channel_map_matches_literal_reshape_replay():
    params = torch.randn(2, 3, 3, 50, dtype=D) * 1.5
    x = torch.randint(0, 256, (2, 3, 3, 3)).double() / 127.5 - 1
    x[0, 0, 0] = torch.tensor([-1.0, 1.0, 254 / 127.5 - 1])   # both edge branches + the 0.99 asymmetry
    got = ref.dlm_log_prob(params, x).numpy()
    assert np.allclose(got, _dlm_by_channel_map(params, x), rtol=1e-10, atol=1e-10)


def test_dlm_bins_sum_to_one():
    """For 1 mixture-equivalent logits the 256 bin probabilities of the R channel sum to 1."""
    params = torch.randn(1, 1, 1, 50, dtype=D)
    total = 0.0
    for v in range(256):
        x = torch.full((1, 1, 1, 3), 0.0, dtype=D)
        x[..., 0] = v / 127.5 - 1
        # isolate the R channel: log p(R) = log p(R,G,B) summed over G,B is costly; instead
        # check each mixture's R-channel bins through the per-component formula
        logits, locs, ls = ref.dlm_create_params(params, x)
        inv = torch.exp(-ls[..., 0, :]); u = x[..., 0:1] - locs[..., 0, :]
        pl, mn = inv * (u + 1 / 255), inv * (u - 1 / 255)
        if v == 0: pr = torch.sigmoid(pl)
        elif v == 255: pr = 1 - torch.sigmoid(mn)
        else: pr = torch.sigmoid(pl) - torch.sigmoid(mn)
        total = total + pr
    assert torch.allclose(total, torch.ones_like(total), atol=1e-9)


def test_bernoulli_is_negative_bce_with_crop():
    l = torch.randn(2, 8, 8, 1, dtype=D)
    x = torch.randint(0, 2, (2, 8, 8, 1)).double()
    want = -torch.nn.functional.binary_cross_entropy_with_logits(
        l[:, 2:-2, 2:-2], x[:, 2:-2, 2:-2], reduction="none").sum((1, 2, 3))
    assert torch.allclose(ref.bernoulli_log_prob(l, x, 2), want)


def test_is_estimator_is_associative_over_chunks():
    S, B, L = 8, 3, 4
    p = [torch.randn(S * B, dtype=D) for _ in range(L)]
    q = [torch.randn(S * B, dtype=D) for _ in range(L)]
    nll = torch.randn(S * B, dtype=D)
    full = ref.is_log_likelihood(p, q, nll, S) - math.log(S)
    sl = lambda t, a, b: t.reshape(S, B)[a:b].reshape(-1)
    c1 = ref.is_log_likelihood([sl(t, 0, 5) for t in p], [sl(t, 0, 5) for t in q], sl(nll, 0, 5), 5)
    c2 = ref.is_log_likelihood([sl(t, 5, 8) for t in p], [sl(t, 5, 8) for t in q], sl(nll, 5, 8), 3)
    assert torch.allclose(ref.is_combine_chunks([c1, c2], S), full)


def test_glue_matches_manual_composition():
    B, W, H, d, C, n = 2, 3, 3, 4, 8, 3
    ctx = torch.randn(B, W, H, C, dtype=D); q = torch.randn(B, W, H, d, dtype=D)
    keys = torch.randn(B, n, W, H, d, dtype=D); vals = torch.randn(B, n, W, H, C, dtype=D)
    ln = lambda: (torch.randn(C, dtype=D), torch.randn(C, dtype=D))
    a, b_, c = ln(), ln(), ln()
    nc, pc = ref.decoder_attention_glue(ctx, q, keys, vals, a, b_, c, 0.3)
    v2 = vals.clone(); v2[:, -1] = ref.layer_norm(vals[:, -1], *b_)
    att = ref.depthwise_attention_apply(q, keys, v2)
    att = att + ref.gelu(ref.layer_norm(att, *c))
    nc2 = ctx + ref.gelu(ref.layer_norm(ctx, *a))
    assert torch.allclose(nc, nc2) and torch.allclose(pc, nc2 + 0.3 * att)


def test_conv2d_same_restatement_matches_library_conv():
    """oracle.ops.conv2d_same (tap-by-tap cross-correlation with 'same' zero padding, res.py:589-601) against
    torch's convolution in float64, value and all gradients -- the checker of the tcgen05 convolution kernels."""
    torch.manual_seed(0)
    for k in (1, 3):
        x = torch.randn(2, 8, 8, 5, dtype=torch.float64, requires_grad=True)
        w = torch.randn(k, k, 5, 7, dtype=torch.float64, requires_grad=True)
        b = torch.randn(7, dtype=torch.float64, requires_grad=True)
        y = ref.conv2d_same(x, w, b)
        yl = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=k // 2).permute(0, 2, 3, 1)
        assert torch.allclose(y, yl, atol=1e-12)
        go = torch.randn_like(y)
        for a, c in zip(torch.autograd.grad(y, (x, w, b), go), torch.autograd.grad(yl, (x, w, b), go)):
            assert torch.allclose(a, c, atol=1e-12)


def test_tf32_split_restatement_and_the_gpu_tests_integer_helper_agree():
    """oracle.ops.tf32_split (frexp / ldexp arithmetic) against the bit-mask helper the GPU tests hold davi_tf32_lo to
    (tests/test_gpu_kernels._tf32_lo_restatement): identical lo parts, including exact values, rounding ties (away
    from zero) and both signs; hi + lo reproduces v to 2^-21 relative, lo has at most 11 significant bits."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "gpu_kernel_tests", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_kernels.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(3)
    v = torch.cat([torch.randn(4096, generator=g), torch.randn(4096, generator=g) * 1e-3, torch.randn(4096, generator=g) * 300,
                   torch.tensor([0.0, 1.0, -1.0, 2.0 ** -20, 1.0 + 2.0 ** -12, 1.0 + 2.0 ** -12 + 2.0 ** -23,
                                 -(1.0 + 2.0 ** -12 + 2.0 ** -23), 1.0 + 2.0 ** -11 + 2.0 ** -12, 3.0 - 2.0 ** -22])]).float()
    hi, lo = ref.tf32_split(v)
    assert torch.equal(lo, mod._tf32_lo_restatement(v))
    assert torch.equal(hi, (v.view(torch.int32) & -8192).view(torch.float32))
    assert bool(((lo.view(torch.int32) & 8191) == 0).all())
    err = (hi.double() + lo.double() - v.double()).abs()
    assert bool((err <= v.double().abs() * 2.0 ** -21).all())
    i = int((v == 1.0 + 2.0 ** -12 + 2.0 ** -23).nonzero()[0])            # a tie: 2^-12 (1 + 2^-11) -> away from zero
    assert lo[i].item() == 2.0 ** -12 * (1 + 2.0 ** -10) and lo[i + 1].item() == -lo[i].item()

