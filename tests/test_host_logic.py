"""CPU tests of the host-side logic around the kernels: regulariser classes of
the fused optimizer, schedules, sample sharding, and the N>1 paths with gloo."""
import math
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def test_regulariser_coefficients_come_from_reg_loss():
    """The fused optimizer's per-tensor l2 strength is read off the modules' own reg_loss():
    d reg_loss / d p == 2 * c * p for every trainable parameter, per-layer lambda_reg included, and
    the model's weights are left untouched by the probe."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import regulariser_coefficients
    for flags in (helpers.SMALL_CIFAR, helpers.SMALL_OMNIGLOT):
        flags = dict(flags)
        flags["decoder_hparams"] = flags["decoder_hparams"] + ",lambda_reg=0.0003"    # a non-default strength
        m = DeepVAE(**parse_flags(**flags))
        helpers.randomize_zero_init_scalars(m)
        before = {n: p.detach().clone() for n, p in m.named_parameters()}
        coef = regulariser_coefficients(m)
        assert all(torch.equal(before[n], p.detach()) for n, p in m.named_parameters())
        m.reg_loss().backward()
        seen = set()
        for n, p in m.named_parameters():
            if not p.requires_grad:
                assert p.grad is None
                continue
            want = 2 * coef[p] * p.detach()
            assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-12), n
            seen.add(round(coef[p], 9))
        assert {round(1.5e-4, 9), round(3e-4, 9), 0.01} <= seen, seen


def test_schedules():
    from deep_attentive_vi_b200.util.trainer import cosine_warmup_lr, kl_beta
    assert kl_beta(0) == pytest.approx(0.1) and kl_beta(16) == 1.0 and kl_beta(100) == 1.0
    assert kl_beta(8) == pytest.approx(0.55)
    lr = lambda s: cosine_warmup_lr(s, 100, 1e-2, 1e-4, 5, 395)
    assert lr(1) == pytest.approx(1e-4 + (1e-2 - 1e-4) / 500) and lr(500) == pytest.approx(1e-2)
    assert lr(500 + 395 * 100) == pytest.approx(1e-4) and lr(10 ** 6) == pytest.approx(1e-4)
    assert lr(250) == pytest.approx(1e-4 + (1e-2 - 1e-4) * 0.5)


def test_cosine_warmup_matches_the_reference_schedule_golden():
    """(step -> lr) points produced by the reference's own CosineWarmup, driven like its Keras callback
    (tests/golden/make_golden_lr.py): the host schedule reproduces every point, with the schedule lengths
    derived like keras_train (tr.py:139-140: steps_per_epoch = int(train_size / (num_gpus * batch)),
    decay_epochs = epochs - lr_warmup_epochs)."""
    import json
    from deep_attentive_vi_b200.util.trainer import cosine_warmup_lr
    fix = json.load(open(os.path.join(helpers.GOLDEN_DIR, "ref_lr_schedule.json")))
    assert len(fix["cases"]) >= 3
    for c in fix["cases"]:
        spe = int(c["train_size"] / (c["num_gpus"] * c["batch_size"]))
        decay = c["epochs"] - c["lr_warmup_epochs"]
        assert (spe, decay) == (c["steps_per_epoch"], c["lr_decay_epochs"])
        for step, want in c["points"]:
            got = cosine_warmup_lr(step, spe, c["learning_rate"], c["lr_min_learning_rate"],
                                   c["lr_warmup_epochs"], decay)
            assert got == pytest.approx(want, rel=1e-6), (step, got, want)


def test_empty_sample_shards_and_rank_streams():
    """K < world: ranks without samples contribute -inf and the combination is still the K-sample estimate;
    ranks that share a seed draw different samples."""
    from deep_attentive_vi_b200.util.eval import combine_partial_logsumexp, decorrelate_rank_rng, shard_samples
    K, world = 3, 8
    lw = torch.randn(K, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    partials = []
    for r in range(world):
        lo, hi = shard_samples(K, world, r)
        partials.append(torch.logsumexp(lw[lo:hi], 0) if hi > lo else torch.full((5,), float("-inf"), dtype=torch.float64))
    got = combine_partial_logsumexp(torch.stack(partials), K)
    assert torch.allclose(got, torch.logsumexp(lw, 0) - math.log(K))
    draws = []
    for r in range(2):
        torch.manual_seed(0)
        decorrelate_rank_rng("cpu", r, 2)
        draws.append(torch.randn(4))
    torch.manual_seed(0)
    assert not torch.equal(draws[0], draws[1])


def test_sample_sharding_covers_all_samples():
    from deep_attentive_vi_b200.util.eval import shard_samples
    for K, W in ((1000, 8), (100, 3), (7, 2)):
        spans = [shard_samples(K, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == K
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.eval import combine_partial_logsumexp, shard_samples
    from oracle.model import OracleVAE
    torch.set_num_threads(2)
    kw = parse_flags(**helpers.SMALL_OMNIGLOT)
    torch.manual_seed(0)
    m = DeepVAE(**kw)
    helpers.randomize_zero_init_scalars(m)
    B = 2
    x_all = helpers.synthetic_images(kw, B * world)
    eps_all, _ = m.draw_noise(B * world, "cpu", generator=torch.Generator().manual_seed(3))
    probe = "_latent_layer.1._gamma"

    def shard_grad(r):
        o = OracleVAE(helpers.oracle_cfg(kw), m.state_dict())
        o.sd[probe].requires_grad_()
        sl = slice(r * B, (r + 1) * B)
        nll, kl, _ = o.forward(x_all[sl], [e[sl] for e in eps_all], True, None)
        # trainer.forward_backward: local sums scaled by the GLOBAL batch (tr.py:174)
        ((nll.sum() + 0.5 * kl.sum()) / (B * world)).backward()
        return o.sd[probe].grad.clone()

    g = shard_grad(rank)
    dist.all_reduce(g, op=dist.ReduceOp.SUM)             # trainer.all_reduce
    if rank == 0:
        want = sum(shard_grad(r) for r in range(world))  # single process, per-shard BN statistics
        ret["dp_err"] = float((g - want).abs() / want.abs())
    # sample-sharded importance-weighted estimate == unsharded
    K = 6
    lw = torch.randn(K, 3, generator=torch.Generator().manual_seed(4), dtype=torch.float64) * 20
    lo, hi = shard_samples(K, world, rank)
    local = torch.logsumexp(lw[lo:hi], 0)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    got = combine_partial_logsumexp(torch.stack(gathered), K)
    if rank == 0:
        ret["is_err"] = float((got - (torch.logsumexp(lw, 0) - math.log(K))).abs().max())
    dist.destroy_process_group()


def test_data_parallel_and_sample_sharding_with_gloo():
    world = 2
    port = 29500 + os.getpid() % 1000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret["dp_err"] < 1e-10
        assert ret["is_err"] < 1e-10


def test_conv_precision_modes_are_validated():
    """ops.set_conv_precision: three named modes (default = fp32-grade 3xTF32); anything else is an error,
    and the mode sets torch's cuDNN TF32 switch the way the mode needs it."""
    import pytest
    import torch
    from deep_attentive_vi_b200 import ops
    assert ops.get_conv_precision() == "3xtf32"
    try:
        ops.set_conv_precision("fp32")
        assert ops.get_conv_precision() == "fp32" and not torch.backends.cudnn.allow_tf32
        ops.set_conv_precision("tf32")
        assert torch.backends.cudnn.allow_tf32 and torch.backends.cuda.matmul.allow_tf32
        with pytest.raises(ValueError):
            ops.set_conv_precision("bf16")
    finally:
        ops.set_conv_precision("3xtf32")
    assert torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32


def test_gradient_bucket_plan():
    """GradBuckets: buckets are contiguous, cover the flat array, never split a source's hull, and report
    completion exactly once per bucket."""
    from deep_attentive_vi_b200.util.trainer import GradBuckets
    # a: plain parameter; b: conv kernel (v 100..400, log_g 400..410); c, d: parameters; e: conv with v / log_g apart
    sources = {"a": [(0, 64)], "b": [(100, 400), (400, 410)], "c": [(448, 500)], "d": [(512, 900)],
               "e": [(960, 1500), (1600, 1610)], "f": [(1520, 1580)]}
    plan = GradBuckets(sources, 2048, 400)
    assert plan.bounds[0][0] == 0 and plan.bounds[-1][1] == 2048
    assert all(a[1] == b[0] for a, b in zip(plan.bounds, plan.bounds[1:]))
    for k, ranges in sources.items():                              # every range of a source inside ITS bucket
        lo, hi = plan.bounds[plan.bucket_of[k]]
        assert all(lo <= r0 and r1 <= hi for r0, r1 in ranges), (k, plan.bounds)
    assert plan.bucket_of["e"] == plan.bucket_of["f"]               # f lies inside e's hull: one atom
    assert len(plan.bounds) >= 3
    done = []
    for k in ["f", "e", "d", "d", "c", "b"]:                        # reverse order of the flat layout, one repeat
        b = plan.arrived(k)
        if b is not None:
            done.append(b)
    assert done == sorted(set(done), reverse=True) and len(done) == len(set(done))
    rest = plan.remaining()                                         # "a" never arrived: its bucket is flushed at the end
    assert plan.bucket_of["a"] in rest and not (set(rest) & set(done))
    assert sorted(rest + done) == list(range(len(plan.bounds)))
    assert plan.remaining() == []
    plan.reset()
    assert plan.arrived("a") is not None or len(plan.members[plan.bucket_of["a"]]) > 1


def _bucket_worker(rank, world, port, ret):
    """The bucketed exchange across two gloo ranks: per-bucket async all-reduces started from gradient hooks
    during the backward, joined afterwards, against ONE all-reduce of the flat gradient array."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deep_attentive_vi_b200.util.trainer import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(32, 32) for _ in range(6)]).double()
    params = list(net.parameters())
    offs, off = {}, 0
    for p in params:
        offs[id(p)] = (off, off + p.numel()); off += p.numel() + 8          # gaps, like the padded flat layout
    flat = torch.zeros(off, dtype=torch.float64)
    plan = GradBuckets({k: [r] for k, r in offs.items()}, off, 1500)
    works, order = [], []

    def arrived(p):
        b = plan.arrived(id(p))
        if b is None:
            return
        for k in plan.members[b]:
            q = next(t for t in params if id(t) == k)
            flat[offs[k][0]:offs[k][1]].copy_(q.grad.reshape(-1))
        lo, hi = plan.bounds[b]
        order.append(b)
        works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    for p in params:
        p.register_post_accumulate_grad_hook(arrived)
    x = torch.randn(5, 32, dtype=torch.float64, generator=torch.Generator().manual_seed(10 + rank))
    net(x).square().sum().backward()
    assert plan.remaining() == []                                   # every bucket was flushed from a hook
    for w in works:
        w.wait()
    want = torch.zeros_like(flat)
    for p in params:
        want[offs[id(p)][0]:offs[id(p)][1]].copy_(p.grad.reshape(-1))
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["err"] = float((flat - want).abs().max())
        ret["order"] = list(order)
        ret["nb"] = len(plan.bounds)
    dist.destroy_process_group()


def test_bucketed_all_reduce_with_gloo():
    world = 2
    port = 29500 + (os.getpid() + 17) % 1000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_bucket_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret["err"] == 0.0
        assert ret["nb"] >= 3 and ret["order"] == sorted(ret["order"], reverse=True)   # last layers' buckets first


def test_convolution_launch_plans_of_the_model_shapes(libdavi):
    """davi_conv2d_plan (host-only): every convolution shape of the CIFAR-10 / OMNIGLOT models, forward, input gradient
    and weight gradient on 148 SMs, fits shared memory (227 KB) and TMEM (512 columns), keeps at least two ring stages,
    stays inside the workspace bound, and cuts its (tile, k-chunk) units so that no CTA pair idles."""
    import ctypes
    shapes = [(40, 16, 16, 276, 276, 3), (40, 16, 16, 316, 316, 3), (40, 16, 16, 296, 296, 3), (40, 16, 16, 276, 40, 3),
              (40, 16, 16, 276, 340, 1), (40, 16, 16, 316, 380, 1), (40, 16, 16, 296, 360, 1), (40, 16, 16, 276, 256, 1),
              (40, 32, 32, 140, 140, 3), (40, 32, 32, 4, 140, 3), (40, 32, 32, 140, 52, 3), (40, 32, 32, 140, 204, 1),
              (16, 16, 16, 72, 72, 3), (16, 16, 16, 88, 88, 3), (16, 32, 32, 36, 36, 3), (400, 16, 16, 316, 316, 3),
              (1, 16, 16, 8, 8, 1), (2, 32, 8, 36, 40, 3)]
    wsb = libdavi.davi_conv2d_workspace_bytes()
    for (B, H, W, ci, co, k) in shapes:
        for mode, cin, cout in ((0, ci, co), (0, co, ci), (1, ci, co)):        # forward, input gradient, weight gradient
            plan = (ctypes.c_int64 * 12)()
            assert libdavi.davi_conv2d_supported(B, H, W, cin, cout, k, k) == 1
            assert libdavi.davi_conv2d_plan(mode, B, H, W, cin, cout, k, k, 148, plan) == 0, (mode, B, H, W, cin, cout, k)
            tiles, kch, pairs, n_total, np0, np1, stages, slots, reduce, smem, ws, tile_ok = list(plan)
            assert 1 <= pairs <= 74 and pairs <= tiles * kch
            assert n_total == np0 + np1 and np0 % 32 == 0 and np1 % 32 == 0 and np0 <= 256 and n_total >= cout
            assert n_total + 64 * slots <= 512 and slots >= 2 and stages >= 2
            assert smem <= 227 * 1024 and ws <= wsb
            if reduce:
                assert -(-tiles * kch // pairs) <= kch             # at most two segments per pair
    plan = (ctypes.c_int64 * 12)()
    assert libdavi.davi_conv2d_supported(40, 16, 16, 276, 584, 1, 1) == 0       # more than 384 channels
    assert libdavi.davi_conv2d_plan(0, 40, 16, 16, 276, 584, 1, 1, 148, plan) != 0
    assert libdavi.davi_conv2d_supported(40, 32, 32, 138, 138, 3, 3) == 0       # rows that are not 16-byte multiples


def test_split_channels_gradient_is_the_concatenation_of_the_piece_gradients():
    """ops.split_channels (views out, ONE concatenation on the way back; the channel splits of vl.py:359,375-380,
    472-473,483-494) against plain slices: same values, same gradient, including an unused piece, a slice of a piece
    (the encoder query, vl.py:385-386) and a piece that is concatenated again (vl.py:507,547-548)."""
    import torch
    from deep_attentive_vi_b200 import ops
    torch.manual_seed(0)
    t = torch.randn(2, 3, 3, 10, dtype=torch.float64, requires_grad=True)

    def loss(parts):
        a, b, c = parts
        return a.sin().sum() + 2 * (b * b).sum() + (a[..., :2] * c[..., :2]).sum() + torch.cat([a, c * 3], -1).sum()

    keep = ops.split_channels_fused
    try:
        grads = []
        for fused in (True, False):
            ops.split_channels_fused = fused
            parts = ops.split_channels(t * 1.0, 5, 2, 3)
            assert [p.shape[-1] for p in parts] == [5, 2, 3] and all(p._base is not None for p in parts)
            grads.append(torch.autograd.grad(loss(parts), t)[0])
        assert torch.equal(grads[0], grads[1])
        ops.split_channels_fused = True
        a, b = ops.split_channels(t * 1.0, 4, 6)
        g, = torch.autograd.grad(a.sum(), t)                         # b unused: its gradient is a zero block
        assert bool((g[..., :4] == 1).all()) and bool((g[..., 4:] == 0).all())
        with torch.no_grad():                                        # evaluation: plain views, no autograd node
            a, b = ops.split_channels(t, 4, 6)
            assert a.data_ptr() == t.data_ptr() and b.shape[-1] == 6
    finally:
        ops.split_channels_fused = keep
