"""Dev tool (GPU box): colsum micro-benchmark (CUDA events) for the model's bias-gradient shapes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import ops
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda"); flush_rd = torch.zeros_like(flush)
for rows, C in ((10240, 316), (10240, 276), (40960, 140), (10240, 40)):
    g = torch.randn(rows, C, device="cuda")
    for warm in (True, False):
        ts = []
        for it in range(6):
            if not warm:
                flush.fill_(0.0); flush_rd.sum()
            else:
                g.add_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ops.colsum_nhwc(g); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); ref = g.sum(0); t1.record(); t1.synchronize()
        print(json.dumps({"rows": rows, "C": C, "l2_warm": warm, "colsum_us": round(sorted(ts)[len(ts) // 2], 1),
                          "torch_sum_us": round(t0.elapsed_time(t1) * 1e3, 1)}), flush=True)
