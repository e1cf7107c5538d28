"""Dev tool (GPU box): launch the depth-wise attention kernels a few times for ONE shape (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import _lib

B, n, HW, d, C = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (40, 15, 16, 20, 276)))
dev = "cuda"
P = HW * HW
torch.manual_seed(0)
q = torch.randn(B, P, d, device=dev)
K = torch.randn(B, n, P, d, device=dev)
V = torch.randn(B, n, P, C, device=dev)
go = torch.randn(B, P, C, device=dev)
out = torch.empty(B, P, C, device=dev)
dq, dK, dV = torch.empty_like(q), torch.empty_like(K), torch.empty_like(V)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    flush.fill_(0.0)
    _lib.call("davi_dw_attn_fwd", q.data_ptr(), K.data_ptr(), V.data_ptr(), out.data_ptr(), 0, B, n, P, d, C,
              d, n * P * d, P * d, d, n * P * C, P * C, C, C, 0, st)
    flush.fill_(0.0)
    _lib.call("davi_dw_attn_bwd", q.data_ptr(), K.data_ptr(), V.data_ptr(), go.data_ptr(), dq.data_ptr(), dK.data_ptr(),
              dV.data_ptr(), 0, B, n, P, d, C, d, n * P * d, P * d, d, n * P * C, P * C, C, C, 0, st)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
