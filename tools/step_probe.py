"""Quick probe: one fwd+bwd of the full 16-layer CIFAR-10 model at B=40 on the GPU,
with a torch profiler table of the top CUDA kernels (dev tool, not the bench)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags  # noqa: E402

from deep_attentive_vi_b200 import ops  # noqa: E402
tf32 = os.environ.get("CONVS", "3xtf32")        # 3xtf32 (default) | fp32 | tf32
ops.set_conv_precision(tf32)
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
torch.manual_seed(0)
m = DeepVAE(**parse_flags(**CIFAR10_16L)).to(dev)
B = int(os.environ.get("B", 40))
x = torch.randint(0, 256, (B, 32, 32, 3), device=dev).float()


from deep_attentive_vi_b200.util.trainer import Trainer  # noqa: E402
tr = Trainer(m, device=dev)


def step():
    return tr.train_step(x)


for _ in range(3):
    step()
torch.cuda.synchronize()
t = time.time()
n = 5
for _ in range(n):
    step()
torch.cuda.synchronize()
dt = (time.time() - t) / n
print(f"step {dt*1e3:.1f} ms  -> {B/dt:.1f} img/s (convs={tf32}), mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70))
