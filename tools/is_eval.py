#!/usr/bin/env python
"""BASELINE config 5: importance-sampled test NLL of the 16-layer CIFAR-10 model, K = 1000 samples,
sharded by SAMPLE across the ranks (reference: util/eval.py:140-146 + model/vae.py:452-539, which is
data-parallel over images with a serial while_loop over sample chunks).

  python tools/is_eval.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/is_eval.py

Prints one JSON line (rank 0): importance samples x images per second over all ranks, device-timed,
max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import ops  # noqa: E402
from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags  # noqa: E402
from deep_attentive_vi_b200.util.eval import GraphedChunks, importance_log_likelihood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=16)         # ev.py:198
    ap.add_argument("--chunk", type=int, default=25)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--eager", action="store_true", help="launch every chunk eagerly (no CUDA graph)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_conv_precision(os.environ.get("CONVS", "3xtf32"))
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)                                      # same weights on every rank
    model = DeepVAE(**parse_flags(**CIFAR10_16L)).to(dev).eval()
    x = torch.randint(0, 256, (args.batch, 32, 32, 3), generator=torch.Generator().manual_seed(7)).float().to(dev)
    torch.manual_seed(1000 + rank)                            # different importance samples per rank
    chunks = None if args.eager else GraphedChunks(model, x)
    importance_log_likelihood(model, x, 2 * args.chunk * world, args.chunk, world, rank, chunks=chunks)   # warm-up
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        logp = importance_log_likelihood(model, x, args.samples, args.chunk, world, rank, chunks=chunks)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(ms.item()) / 1e3 / args.reps
        print(json.dumps({"metric": "cifar10_16layer_is_eval_sample_images_per_sec",
                          "value": args.samples * args.batch / sec, "unit": "sample-images/s", "n_gpus": world,
                          "seconds_per_batch": sec, "num_importance_samples": args.samples, "batch": args.batch,
                          "chunk": args.chunk, "sharding": "by sample", "cuda_graph": not args.eager, "data": "synthetic",
                          "convs": ops.get_conv_precision(),
                          "mean_log_likelihood_nats": float(logp.mean().item())}), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
