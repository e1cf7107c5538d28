"""Dev tool (GPU box): launch the nonlocal kernels a few times for ONE shape (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import _lib

B, HW, d, C = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (40, 16, 32, 276)))
bwd = "--bwd" in sys.argv
dev = "cuda"
N = HW * HW
torch.manual_seed(0)
rs = 2 * d + C
qkv = torch.randn(B, N, rs, device=dev) * 0.5
x = torch.randn(B, N, C, device=dev)
go = torch.randn(B, N, C, device=dev)
y = torch.empty_like(x)
osave = torch.empty_like(x)
lse = torch.empty(B, N, device=dev)
dqkv = torch.empty_like(qkv)
dgs = torch.empty(2, device=dev)
scale = torch.tensor(0.8, device=dev); gamma = torch.tensor(0.6, device=dev)
wsb = _lib.load().davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
ws = torch.empty(max(wsb // 4, 1), device=dev)
st = torch.cuda.current_stream().cuda_stream
qp = qkv.data_ptr()
for _ in range(3):
    _lib.call("davi_nonlocal_attn_fwd", qp, qp + 4 * d, qp + 8 * d, x.data_ptr(), y.data_ptr(), lse.data_ptr(),
              osave.data_ptr(), B, N, d, C, rs, rs, rs, C, C, C, scale.data_ptr(), gamma.data_ptr(), st)
    if bwd:
        dp = dqkv.data_ptr()
        _lib.call("davi_nonlocal_attn_bwd", qp, qp + 4 * d, qp + 8 * d, lse.data_ptr(), osave.data_ptr(),
                  go.data_ptr(), dp, dp + 4 * d,
                  dp + 8 * d, dgs.data_ptr(), B, N, d, C, rs, rs, rs, C, C, scale.data_ptr(), gamma.data_ptr(),
                  ws.data_ptr(), wsb, st)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
