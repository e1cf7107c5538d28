"""Importance-sampled test log-likelihood, sharded by SAMPLE across ranks --
replaces the reference's util/eval.py:140-146 + model/vae.py:485-539 (there:
data-parallel over images with a serial while_loop over sample chunks).

Rank r draws K/world of the K importance samples for the SAME image batch,
reduces them with a local logsumexp and the ranks' partial results are combined
with one tiny all-gather: logsumexp is associative, exactly like the reference's
own nesting of two logsumexps over chunks (vae.py:482,539)."""
import math

import torch
import torch.distributed as dist

from ..util.hparams import HParams


def get_default_eval_hparams():
    """ev.py:189-201."""
    return HParams(num_gpus=2, generate=True, reconstruct=True, compute_logp=True,
                   num_importance_samples=100, num_sample_images=32, num_test_images=32, workers=25,
                   batch_size=16, shuffle=True, verbose=True)


def shard_samples(num_samples, world, rank):
    """Samples [lo, hi) owned by `rank` (remainder to the last rank, as the
    reference gives the remainder to the last chunk, vae.py:524)."""
    per = num_samples // world
    lo = rank * per
    hi = num_samples if rank == world - 1 else lo + per
    return lo, hi


def combine_partial_logsumexp(partials, num_samples):
    """partials [world, B] -> log p_hat(x) [B] (vae.py:539)."""
    return torch.logsumexp(partials, dim=0) - math.log(num_samples)


class GraphedChunks:
    """CUDA graphs of DeepVAE.infer_worker_body, one per chunk size: a chunk is ~5 000 kernel launches,
    so the eager loop of the reference (vae.py:515-535, a serial tf.while_loop) is host-bound; the graph
    replays it with one host call.  The random draws come from the default CUDA generator, which
    torch.cuda.graph registers, so every replay uses fresh importance samples."""

    def __init__(self, model, x):
        self.model = model
        self.x = x.clone()
        self.graphs = {}
        self.side = torch.cuda.Stream(device=x.device)

    @torch.no_grad()
    def __call__(self, x, n):
        self.x.copy_(x)
        if n not in self.graphs:
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                self.model.infer_worker_body(self.x, n)               # warm-up (cuDNN algorithm selection)
            cur.wait_stream(self.side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.side):
                out = self.model.infer_worker_body(self.x, n)
            self.graphs[n] = (g, out)
        g, out = self.graphs[n]
        g.replay()
        return out.clone()


_rank_streams = set()


def decorrelate_rank_rng(device, rank, world):
    """Sample sharding is only an estimator with K samples if every rank draws DIFFERENT ones; ranks
    usually share one seed (identical weights).  Offsets this rank's CUDA generator by a rank-dependent
    amount once per (device, rank): same seed, disjoint Philox subsequences."""
    key = (str(device), rank, world)
    if world <= 1 or key in _rank_streams:
        return
    _rank_streams.add(key)
    gen = torch.cuda.default_generators[torch.device(device).index or 0] if torch.device(device).type == "cuda" \
        else torch.default_generator
    if torch.device(device).type == "cuda":
        gen.set_offset(gen.get_offset() + (rank + 1) * (1 << 40))          # multiples of 4 (Philox requirement)
    else:
        gen.manual_seed(gen.initial_seed() + 7919 * (rank + 1))


@torch.no_grad()
def importance_log_likelihood(model, x, num_samples, max_chunk=25, world=1, rank=0, group=None, chunks=None):
    """log p_hat(x) [B] with K = num_samples importance samples, this rank
    computing its shard in chunks of at most `max_chunk` samples.  `chunks`: a GraphedChunks
    instance to replay CUDA graphs instead of launching every chunk eagerly.  Ranks draw from
    decorrelated generator streams; a rank whose shard is empty (K < world) contributes -inf."""
    lo, hi = shard_samples(num_samples, world, rank)
    decorrelate_rank_rng(x.device, rank, world)
    parts = []
    k = lo
    while k < hi:
        n = min(max_chunk, hi - k)
        parts.append(chunks(x, n) if chunks is not None else model.infer_worker_body(x, n))
        k += n
    if parts:
        local = torch.logsumexp(torch.stack(parts, 0), dim=0)        # [B]
    else:
        local = torch.full((x.shape[0],), float("-inf"), device=x.device, dtype=torch.float32)
    if world > 1:
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local, group=group)
        partials = torch.stack(gathered, 0)
    else:
        partials = local.unsqueeze(0)
    return combine_partial_logsumexp(partials, num_samples)
