"""Dev tool (GPU box): what TF32 cuDNN convolutions change.  One forward+backward of the full
16-layer CIFAR-10 model on the same batch and the same random draws with fp32, 3xTF32 (default) and plain TF32
convolutions: relative difference of the ELBO terms and of the flat gradient."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags  # noqa: E402
from deep_attentive_vi_b200.util.trainer import Trainer  # noqa: E402

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
m = DeepVAE(**parse_flags(**CIFAR10_16L)).to(dev)
B = int(os.environ.get("B", 40))
x = torch.randint(0, 256, (B, 32, 32, 3), device=dev).float()
tr = Trainer(m, device=dev)


from deep_attentive_vi_b200 import ops  # noqa: E402


def run(convs):
    ops.set_conv_precision(convs)
    torch.manual_seed(123)
    nll, kl = tr.forward_backward(x)
    torch.cuda.synchronize()
    return nll.double().sum().item(), kl.double().sum().item(), tr.state.G.double().clone()


n0, k0, g0 = run("fp32")
rel = lambda a, b: abs(a - b) / abs(b)
out = {"batch": B, "nll_fp32": n0, "kl_fp32": k0}
for mode in ("fp32", "3xtf32", "tf32"):      # the second fp32 run is the run-to-run noise floor
    n1, k1, g1 = run(mode)
    out[f"{mode}_vs_fp32"] = {"nll": rel(n1, n0), "kl": rel(k1, k0), "nelbo": rel(n1 + k1, n0 + k0),
                              "grad_rel_to_norm": ((g1 - g0).norm() / g0.norm()).item(),
                              "grad_max_rel_to_max": ((g1 - g0).abs().max() / g0.abs().max()).item()}
print(json.dumps(out))
