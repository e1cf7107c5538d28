"""Dev tool (GPU box): per-step loss of the full CIFAR-10 model, eager vs CUDA-graph trainer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags
from deep_attentive_vi_b200.util.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
from deep_attentive_vi_b200 import ops as _ops; _ops.set_conv_precision(os.environ.get("CONVS", "3xtf32"))
torch.backends.cudnn.benchmark = True
for use_graph in (False, True):
    torch.manual_seed(0)
    m = DeepVAE(**parse_flags(**CIFAR10_16L)).to(dev)
    tr = Trainer(m, device=dev, use_graph=use_graph, graph_warmup=2)
    g = torch.Generator().manual_seed(1000)
    xs = [torch.randint(0, 256, (40, 32, 32, 3), generator=g).float().to(dev) for _ in range(4)]
    out = []
    for s in range(steps):
        nll, kl = tr.train_step(xs[s % 4])
        out.append((float(nll.mean()), float(kl.mean())))
    print("graph" if use_graph else "eager", " ".join(f"{a:.4g}/{b:.4g}" for a, b in out), flush=True)
    del tr, m
    torch.cuda.empty_cache()
