"""GPU tests of the training step: fused Adamax against the Keras update rule
(reference util/trainer.py:226-231), and the CUDA-graph replay of the whole step
against the eager step on identical inputs and noise."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def test_adamax_matches_keras_rule(cuda):
    from deep_attentive_vi_b200 import _lib
    torch.manual_seed(0)
    n = 1003                                     # ragged tail (n % 4 != 0)
    p = torch.randn(n, device=cuda)
    m = torch.zeros(n + 1, device=cuda)[:n]
    u = torch.zeros(n + 1, device=cuda)[:n]
    pr, mr, ur = p.double().cpu(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    lr, b1, b2, eps, l2 = 0.01, 0.9, 0.999, 1e-3, 1.5e-4
    hyper = torch.zeros(2, device=cuda)
    stream = torch.cuda.current_stream().cuda_stream
    for t in range(1, 5):
        g = torch.randn(n, device=cuda)
        if t % 2:                                # host-scalar entry point
            _lib.call("davi_adamax_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), u.data_ptr(), n,
                      lr, b1, b2, eps, t, l2, 1.0, stream)
        else:                                    # device-scalar entry point (CUDA-graph safe)
            hyper.copy_(torch.tensor([lr / (1 - b1 ** t), 1.0]))
            _lib.call("davi_adamax_step_dev", p.data_ptr(), g.data_ptr(), m.data_ptr(), u.data_ptr(), n,
                      hyper.data_ptr(), b1, b2, eps, l2, stream)
        gg = g.double().cpu() + 2 * l2 * pr      # Keras l2 regulariser gradient
        mr = b1 * mr + (1 - b1) * gg
        ur = torch.maximum(b2 * ur, gg.abs())
        pr = pr - lr / (1 - b1 ** t) * mr / (ur + eps)
    assert (p.double().cpu() - pr).abs().max() < 1e-5


def test_graph_replay_equals_eager_step(cuda, convs):
    """Same weights, images and (fixed) noise: 3 eager warm-up steps + 3 graph
    replays must equal 6 eager steps."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 4
    xs = [helpers.synthetic_images(kw, B, seed=s).to(cuda) for s in range(6)]
    results = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        m = DeepVAE(**kw)
        helpers.randomize_zero_init_scalars(m)
        eps, noise = m.draw_noise(B, "cpu", generator=torch.Generator().manual_seed(1), training=True)
        eps = [e.to(cuda) for e in eps]
        noise = [None if nz is None else tuple(None if t is None else t.to(cuda) for t in nz) for nz in noise]
        m = m.to(cuda)
        m.draw_noise = lambda *a, **k: (eps, noise)      # fixed noise for both runs
        tr = Trainer(m, device=cuda, use_graph=use_graph, graph_warmup=3)
        losses = []
        for x in xs:
            nll, kl = tr.train_step(x)
            losses.append((nll.mean() + kl.mean()).item())
        assert (tr._graph is not None) == use_graph
        results.append((losses, tr.state.P.clone(), tr.step_dev.item()))
    (l0, p0, s0), (l1, p1, s1) = results
    assert s0 == s1 == 6.0
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-4 * abs(a), (l0, l1)
    # Adamax moves a weight by up to lr = 1e-2 per step: allow 10 % of ONE step after six
    assert (p0 - p1).abs().max().item() < 1e-3


def test_fused_weight_norm_matches_per_layer_path(cuda, convs):
    """davi_weight_norm_fwd/bwd (all conv kernels in one launch over the flat parameter array)
    against the per-layer torch restatement of layers/weight_norm.py:126-135: same loss, same
    gradients of every v and log_g."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 3
    x = helpers.synthetic_images(kw, B, seed=3).to(cuda)
    torch.manual_seed(0)
    m = DeepVAE(**kw)
    helpers.randomize_zero_init_scalars(m)
    eps, noise = m.draw_noise(B, "cpu", generator=torch.Generator().manual_seed(1), training=True)
    eps = [e.to(cuda) for e in eps]
    noise = [None if nz is None else tuple(None if t is None else t.to(cuda) for t in nz) for nz in noise]
    m = m.to(cuda)
    tr = Trainer(m, device=cuda)
    assert tr.state.wn_rows > 0
    nll1, kl1 = tr.forward_backward(x, eps, noise)
    g1 = tr.state.G.clone()
    tr.state.release()
    nll2, kl2 = tr.forward_backward(x, eps, noise)        # per-layer path; wn kernels find no override
    g2 = tr.state.G.clone()
    assert torch.allclose(nll1, nll2, rtol=1e-5) and torch.allclose(kl1, kl2, rtol=1e-5)
    err = (g1 - g2).norm() / g2.norm()
    assert err < 1e-4, err.item()


def test_batched_kernel_transpose_matches_the_per_kernel_entry_point(cuda, convs):
    """davi_conv2d_wt_batched (one launch per training step for every convolution's input-gradient weights) against
    davi_conv2d_wt per kernel and against the index rule wt[ci][taps-1-tap][co] = w[co][tap][ci] itself: bit-exact
    (pure data movement), for every weight-normed convolution and every packed q|k|v projection of the model."""
    from deep_attentive_vi_b200 import _lib
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 2
    x = helpers.synthetic_images(kw, B, seed=5).to(cuda)
    torch.manual_seed(0)
    m = DeepVAE(**kw)
    helpers.randomize_zero_init_scalars(m)
    m = m.to(cuda)
    tr = Trainer(m, device=cuda)
    st = tr.state
    assert st.wt_tables[5] > 0 and len(st.qkv_blocks) > 0
    st.Wt.fill_(float("nan"))
    tr.forward_backward(x)                                        # fills W, then Wt, before the forward pass
    stream = torch.cuda.current_stream().cuda_stream
    views = [(mod._w_override, mod._wt_override) for mod in st.wn_modules if mod._w_override is not None]
    views += [(blk._qkv_w_override, blk._qkv_wt_override) for blk in st.qkv_blocks]
    assert len(views) == st.wt_tables[5]
    for w, wt in views:
        wh = w.detach().permute(0, 2, 3, 1)                        # [Cout,kh,kw,Cin] memory
        assert wh.is_contiguous() and wt.is_contiguous()
        co, kh, kw_, ci = wh.shape
        assert tuple(wt.shape) == (ci, kh, kw_, co)
        want = wh.flip(1, 2).permute(3, 1, 2, 0).contiguous()
        assert torch.equal(wt, want)
        single = torch.empty_like(want)
        _lib.call("davi_conv2d_wt", wh.data_ptr(), single.data_ptr(), co, kh * kw_, ci, stream)
        assert torch.equal(single, want)
    # the lo images of the step (one davi_tf32_lo launch over W | Wt) and the views the layers read them through
    hi = (st.WW.view(torch.int32) & -8192).view(torch.float32)
    want_lo = (((st.WW - hi).view(torch.int32) + 4096) & -8192).view(torch.float32)
    ok = ~torch.isnan(st.WW)
    assert torch.equal(st.LL[ok], want_lo[ok])
    los = [mod._lo_override for mod in st.wn_modules if mod._w_override is not None] + [blk._qkv_lo_override for blk in st.qkv_blocks]
    for (w, wt), (w_lo, wt_lo) in zip(views, los):
        assert w_lo.data_ptr() - st.Wlo.data_ptr() == w.data_ptr() - st.W.data_ptr() and w_lo.shape == w.permute(0, 2, 3, 1).shape
        assert wt_lo.data_ptr() - st.Wtlo.data_ptr() == wt.data_ptr() - st.Wt.data_ptr() and wt_lo.shape == wt.shape
    # padding between kernels in the flat array is never written
    covered = torch.zeros_like(st.Wt, dtype=torch.bool)
    for _, wt in views:
        off = (wt.data_ptr() - st.Wt.data_ptr()) // 4
        covered[off:off + wt.numel()] = True
    assert torch.isnan(st.Wt[~covered]).all()


def test_step_with_precomputed_lo_images_equals_the_on_chip_path(cuda):
    """The training step with the per-step lo images of all kernels (davi_conv2d_fwd_pre) against the same step with the
    lo images derived on chip (ops.conv_pre_lo = False): identical loss terms and an identical flat gradient."""
    from deep_attentive_vi_b200 import ops
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer
    ops.set_conv_precision("3xtf32")
    ops.set_conv_impl("tcgen05")
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 4
    x = helpers.synthetic_images(kw, B, seed=2).to(cuda)
    torch.manual_seed(0)
    m = DeepVAE(**kw)
    helpers.randomize_zero_init_scalars(m)
    eps, noise = m.draw_noise(B, "cpu", generator=torch.Generator().manual_seed(1), training=True)
    eps = [e.to(cuda) for e in eps]
    noise = [None if nz is None else tuple(None if t is None else t.to(cuda) for t in nz) for nz in noise]
    m = m.to(cuda)
    tr = Trainer(m, device=cuda)
    out = []
    try:
        for pre in (True, False):
            ops.conv_pre_lo = pre
            n0 = ops._lib.launch_count()
            nll, kl = tr.forward_backward(x, eps, noise)
            out.append((nll.clone(), kl.clone(), tr.state.G.clone(), ops._lib.launch_count() - n0))
    finally:
        ops.conv_pre_lo = True
    (n1, k1, g1, c1), (n2, k2, g2, c2) = out
    assert c1 == c2                                               # same launches, other entry point
    assert torch.allclose(n1, n2, rtol=1e-6) and torch.allclose(k1, k2, rtol=1e-6)
    err = ((g1 - g2).norm() / g2.norm()).item()
    assert err < 1e-6, err


@pytest.mark.parametrize("use_graph", [False, True])
def test_bucketed_gradient_exchange_equals_the_single_exchange(cuda, use_graph):
    """The bucketed path (gradients gathered, weight-norm chain rule finished and the bucket's all-reduce started
    from gradient hooks while the backward runs; run.py:94-95, tr.py:160-174) must leave the same flat gradient
    array and parameters as the gather-everything-after-the-backward path -- here on one rank, i.e. without the
    collective itself (tests/test_host_logic.py runs it across two gloo ranks)."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 4
    xs = [helpers.synthetic_images(kw, B, seed=s).to(cuda) for s in range(4)]
    results = []
    for bucket_mb in (0, 0.1):
        torch.manual_seed(0)
        m = DeepVAE(**kw)
        helpers.randomize_zero_init_scalars(m)
        eps, noise = m.draw_noise(B, "cpu", generator=torch.Generator().manual_seed(1), training=True)
        eps = [e.to(cuda) for e in eps]
        noise = [None if nz is None else tuple(None if t is None else t.to(cuda) for t in nz) for nz in noise]
        m = m.to(cuda)
        m.draw_noise = lambda *a, **k: (eps, noise)
        tr = Trainer(m, device=cuda, use_graph=use_graph, graph_warmup=2, bucket_mb=bucket_mb)
        for x in xs:
            tr.train_step(x)
        if bucket_mb:
            assert len(tr._plan.bounds) >= 3                      # several buckets at this size
            if not use_graph:
                # overlap: all but the last flushes happened while other buckets were still waiting for gradients
                assert sum(1 for _, outstanding in tr.bucket_log if outstanding > 0) >= len(tr._plan.bounds) - 2
                assert sorted(b for b, _ in tr.bucket_log) == list(range(len(tr._plan.bounds)))
        results.append((tr.state.G.clone(), tr.state.P.clone()))
    (g0, p0), (g1, p1) = results
    # two runs of the SAME path differ in the last bits (atomic float sums in the scalar-gradient reductions)
    assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item(), ((g0 - g1).abs().max().item(), g0.abs().max().item())
    assert (p0 - p1).abs().max().item() < 1e-3
