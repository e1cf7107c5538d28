"""Kernel micro-benchmarks (run on the GPU box): depth-wise attention sweep
(BASELINE config 4), nonlocal blocks, ELBO reductions.  Times each launch with
CUDA events on the launching stream; L2 is flushed before every timed launch by
writing a 512 MB buffer.  Prints one JSON line per case."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import ops  # noqa: E402

PEAK = 6553.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            # L2 flush: write a buffer larger than L2, then READ a second one so that the cache is
            # left full of CLEAN lines (otherwise the timed kernel pays the write-back of the flush)
            flush.fill_(1.0)
            FLUSH_READ.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


FLUSH_READ = None


def dw_case(B, n, HW, d, C, flush, bwd=True):
    dev = "cuda"
    q = torch.randn(B, HW, HW, d, device=dev)
    k = torch.randn(B, n, HW, HW, d, device=dev)
    v = torch.randn(B, n, HW, HW, C, device=dev)
    go = torch.randn(B, HW, HW, C, device=dev)
    P = HW * HW
    fb = 4 * B * P * (n * (d + C) + d + C)
    bb = 4 * B * P * (2 * n * (d + C) + 2 * d + C)
    f = ops._DwAttnStacked
    class Ctx:  # minimal stand-in so the raw forward/backward can be timed without autograd overhead
        needs_input_grad = (True, True, True, False)
        def save_for_backward(self, *a): self.saved_tensors = a
    ctx = Ctx()
    med, best = timeit(lambda: f.forward(ctx, q, k, v, None), flush=flush)
    res = {"op": "dw_attn_fwd", "B": B, "n": n, "HW": HW, "d": d, "C": C, "MB": fb / 1e6,
           "ms": med, "GBs": fb / med / 1e6, "best_GBs": fb / best / 1e6, "frac_measured": fb / med / 1e6 / PEAK}
    print(json.dumps(res), flush=True)
    if bwd:
        med, best = timeit(lambda: f.backward(ctx, go), flush=flush)
        res = {"op": "dw_attn_bwd", "B": B, "n": n, "HW": HW, "d": d, "C": C, "MB": bb / 1e6,
               "ms": med, "GBs": bb / med / 1e6, "best_GBs": bb / best / 1e6, "frac_measured": bb / med / 1e6 / PEAK}
        print(json.dumps(res), flush=True)


def copy_case(flush):
    a = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a), flush=flush)
    print(json.dumps({"op": "torch_copy_256MB", "ms": med, "GBs": 2 * a.numel() * 4 / med / 1e6}), flush=True)


def nonlocal_case(B, HW, d, C, flush):
    dev = "cuda"
    N = HW * HW
    qkv = torch.randn(B, HW, HW, 2 * d + C, device=dev) * 0.3
    x = torch.randn(B, HW, HW, C, device=dev)
    go = torch.randn(B, HW, HW, C, device=dev)
    scale = torch.tensor(1.0, device=dev); gamma = torch.tensor(0.5, device=dev)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    fl = 2 * B * N * N * (d + C)
    f = ops._NonlocalAttn
    class Ctx:
        needs_input_grad = (True,) * 6
        def save_for_backward(self, *a): self.saved_tensors = a
    ctx = Ctx()
    med, _ = timeit(lambda: f.forward(ctx, q, k, v, x, scale, gamma), flush=flush, iters=5)
    print(json.dumps({"op": "nonlocal_fwd", "B": B, "N": N, "d": d, "C": C, "ms": med,
                      "TFLOPs": fl / med / 1e9}), flush=True)
    med, _ = timeit(lambda: f.backward(ctx, go), flush=flush, iters=5)
    print(json.dumps({"op": "nonlocal_bwd", "B": B, "N": N, "d": d, "C": C, "ms": med,
                      "TFLOPs": 2.5 * fl / med / 1e9}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true", help="full BASELINE config-4 sweep")
    ap.add_argument("--nonlocal", dest="nl", action="store_true")
    args = ap.parse_args()
    global FLUSH_READ
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    FLUSH_READ = torch.zeros(256 * 1024 * 1024 // 4, device="cuda")
    copy_case(flush)
    # the 31 calls of the CIFAR-16 model at B=40
    for n in (1, 2, 4, 8, 12, 15):
        dw_case(40, n, 16, 20, 276, flush)
    for n in (2, 9, 17):
        dw_case(40, n, 16, 20, 256, flush)
    dw_case(16, 14, 16, 8, 72, flush)
    if args.sweep:
        for HW in (16, 32):
            for d, C in ((8, 72), (20, 276), (32, 288)):
                for n in (4, 8, 16, 17, 24, 32):
                    for B in (40, 80, 160, 320):
                        if 4 * B * HW * HW * n * (d + C) * 2 > 60e9:
                            continue
                        dw_case(B, n, HW, d, C, flush)
    if args.nl:
        nonlocal_case(40, 16, 32, 276, flush)
        nonlocal_case(40, 16, 32, 316, flush)
        nonlocal_case(40, 32, 32, 138, flush)
        nonlocal_case(16, 16, 8, 72, flush)


if __name__ == "__main__":
    main()
