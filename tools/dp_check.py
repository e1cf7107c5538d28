"""Dev tool (GPU box, N >= 2 GPUs): the bucketed, overlapped gradient exchange against the single all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Each rank trains the small CIFAR configuration for a few steps on its own shard, once with ONE all-reduce after the
backward (bucket_mb = 0) and once with 0.1 MB buckets started from the gradient hooks (eager and CUDA graph); the
parameters must agree between the modes and between the ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from deep_attentive_vi_b200.model import DeepVAE, parse_flags  # noqa: E402
from deep_attentive_vi_b200.util.trainer import Trainer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    kw = parse_flags(**helpers.SMALL_CIFAR)
    B = 4
    xs = [helpers.synthetic_images(kw, B, seed=100 * rank + s).to(dev) for s in range(5)]
    out = {}
    for name, bucket_mb, use_graph in (("single", 0, False), ("bucketed", 0.1, False), ("bucketed_graph", 0.1, True)):
        torch.manual_seed(0)
        m = DeepVAE(**kw)
        helpers.randomize_zero_init_scalars(m)
        eps, noise = m.draw_noise(B, "cpu", generator=torch.Generator().manual_seed(1 + rank), training=True)
        eps = [e.to(dev) for e in eps]
        noise = [None if nz is None else tuple(None if t is None else t.to(dev) for t in nz) for nz in noise]
        m = m.to(dev)
        m.draw_noise = lambda *a, **k: (eps, noise)
        tr = Trainer(m, device=dev, world_size=world, rank=rank, use_graph=use_graph, graph_warmup=2, bucket_mb=bucket_mb)
        for x in xs:
            tr.train_step(x)
        torch.cuda.synchronize()
        out[name] = tr.state.P.clone()
        ref = out[name].clone()
        dist.broadcast(ref, 0)
        out[name + "_rank_diff"] = float((ref - out[name]).abs().max())
        if bucket_mb:
            out[name + "_buckets"] = len(tr._plan.bounds)
    res = {"rank": rank, "world": world,
           "bucketed_vs_single": float((out["bucketed"] - out["single"]).abs().max()),
           "graph_vs_single": float((out["bucketed_graph"] - out["single"]).abs().max()),
           "rank_diff": [out[k] for k in ("single_rank_diff", "bucketed_rank_diff", "bucketed_graph_rank_diff")],
           "buckets": out["bucketed_buckets"]}
    print(json.dumps(res), flush=True)
    # a captured CUDA graph holds NCCL kernels: destroy_process_group() then blocks (same tear-down as bench.py)
    dist.barrier(device_ids=[torch.cuda.current_device()])
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
