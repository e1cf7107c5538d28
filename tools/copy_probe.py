"""Dev tool: WHERE do the element-wise / copy / fill kernels of a training step come from?  One eager step of the
16-layer CIFAR-10 model at B = 40 under the torch profiler with shapes and Python stacks; every aten op outside the
repo's own kernels is grouped by (op, input shapes, autograd node, innermost repo frames) with its launch count and
CUDA time.  Usage: python tools/copy_probe.py [out.txt]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import ops  # noqa: E402
from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags  # noqa: E402
from deep_attentive_vi_b200.util.trainer import Trainer  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
ops.set_conv_precision("3xtf32")
dev = torch.device("cuda")
torch.manual_seed(0)
m = DeepVAE(**parse_flags(**CIFAR10_16L)).to(dev)
B = int(os.environ.get("B", 40))
x = torch.randint(0, 256, (B, 32, 32, 3), device=dev).float()
tr = Trainer(m, device=dev, use_graph=False) if "use_graph" in Trainer.__init__.__code__.co_varnames else Trainer(m, device=dev)
for _ in range(2):
    tr.train_step(x)
torch.cuda.synchronize()

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    tr.train_step(x)
    torch.cuda.synchronize()


def dev_time(e):
    for name in ("self_device_time_total", "self_cuda_time_total"):
        v = getattr(e, name, None)
        if v is not None:
            return float(v)
    return 0.0


groups = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for e in prof.events():
    t = dev_time(e)
    if t <= 0 or not e.name.startswith("aten::"):
        continue
    node, p = "", e.cpu_parent
    while p is not None:
        if p.name.startswith("autograd::engine::evaluate_function"):
            node = p.name.split(": ")[-1]
            break
        p = p.cpu_parent
    frames = [f for f in (e.stack or []) if "deep_attentive_vi_b200" in f or "bench.py" in f][:3]
    frames = tuple(f.split("deep_attentive_vi_b200/")[-1] for f in frames)
    shapes = str([s for s in (e.input_shapes or []) if s])[:90]
    g = groups[(e.name, shapes, node, frames)]
    g[0] += 1
    g[1] += t
    total += t
print(f"aten ops with device time: {total / 1e3:.2f} ms in this step", file=out)
for (name, shapes, node, frames), (n, t) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{t / 1e3:7.3f} ms {n:5d} x {name:28s} {shapes}  [{node}]", file=out)
    for f in frames:
        print(f"{'':24s}{f}", file=out)
