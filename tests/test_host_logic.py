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


def test_l2_classes_match_reg_loss_gradient():
    """d reg_loss / d p == 2 * l2_coefficient(name) * p for every trainable parameter."""
    from deep_attentive_vi_b200.model import DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import l2_coefficient
    for flags in (helpers.SMALL_CIFAR, helpers.SMALL_OMNIGLOT):
        m = DeepVAE(**parse_flags(**flags))
        helpers.randomize_zero_init_scalars(m)
        m.reg_loss().backward()
        for n, p in m.named_parameters():
            if not p.requires_grad:
                assert p.grad is None
                continue
            want = 2 * l2_coefficient(n) * p.detach()
            assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-12), n


def test_schedules():
    from deep_attentive_vi_b200.util.trainer import cosine_warmup_lr, kl_beta
    assert kl_beta(0) == pytest.approx(0.1) and kl_beta(16) == 1.0 and kl_beta(100) == 1.0
    assert kl_beta(8) == pytest.approx(0.55)
    lr = lambda s: cosine_warmup_lr(s, 100, 1e-2, 1e-4, 5, 395)
    assert lr(0) == pytest.approx(1e-4) and lr(500) == pytest.approx(1e-2)
    assert lr(500 + 395 * 100) == pytest.approx(1e-4)
    assert lr(250) == pytest.approx(1e-4 + (1e-2 - 1e-4) * 0.5)


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
