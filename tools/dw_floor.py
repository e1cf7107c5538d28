"""Dev tool (GPU box): what a plain streaming copy achieves at the depth-wise calls' sizes.

torch `b.copy_(a)` launched back to back on rotating operand sets (footprint > 3x L2), one CUDA-event pair
around the lot -- the same protocol as tools/dw_replay.py -- so that the per-launch fixed cost of a
24-230 MB launch on this machine is known next to the kernels' numbers."""
import json
import sys
import torch

dev = torch.device("cuda")
L2 = 126e6
out = []
for mb in (24, 36, 48, 73, 97, 121, 145, 194, 230, 1700):
    n = int(mb * 1e6 / 8)                                   # floats per side: read n*4 + write n*4 = mb
    sets = max(2, min(24, int(3 * L2 // (mb * 1e6)) + 1))
    a = [torch.randn(n, device=dev) for _ in range(sets)]
    b = [torch.empty(n, device=dev) for _ in range(sets)]
    for x, y in zip(a, b):
        y.copy_(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()                              # replayed from a graph: no host launch cost
    with torch.cuda.graph(g):
        for x, y in zip(a, b):
            y.copy_(x)
    g.replay()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) / sets)
    t = sorted(ts)[2]
    out.append({"MB": mb, "sets": sets, "us": round(t * 1e3, 2), "GBs": round(mb / t / 1e3 * 1e3)})
    del a, b
print(json.dumps(out))
