"""Replay of the depth-wise attention launches of ONE training step on synthetic operands.

`ops.dw_trace` (deep_attentive_vi_b200/ops.py) records the structure of every launch a step issues --
davi_dw_attn_slots_fwd, davi_dw_attn_slots_bwd (which gradients it writes, which are left to the gather,
whether it saves its weights) and davi_dw_attn_gather (contexts and their consumers).  This module rebuilds
operands of exactly those shapes / strides / slice layouts and launches the same entry points through the
C ABI, so that bench.py can time "what the step runs" stand-alone with cold inputs, and ncu can count the
DRAM bytes of the same launches:

    python tools/dw_replay.py --trace profiles/r02_dw_trace_cifar16_b40.json [--only fwd:15 bwd:15 gather]

Algorithmic bytes: fwd 4 B P (n (d + C) + d + C), bwd
4 B P (2 n (d + C) + 2 d + C) per attention call; the gather launches move part of the backward's bytes
(every dK / dV row of a shared context is written once, there) and add none of their own.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from deep_attentive_vi_b200 import _lib  # noqa: E402

L2_BYTES = 126e6


def algorithmic_bytes(e):
    if e["op"] == "gather":
        return 0
    B, n, P, d, C = e["B"], e["n"], e["P"], e["d"], e["C"]
    if e["op"] == "fwd":
        return 4 * B * P * (n * (d + C) + d + C)
    return 4 * B * P * (2 * n * (d + C) + 2 * d + C)


def moved_bytes(e):
    """Bytes the launch itself has to move (compulsory traffic of THIS implementation), for comparison with
    ncu's dram__bytes_read.sum + dram__bytes_write.sum."""
    if e["op"] == "gather":
        tot = 0
        for p in e["problems"]:
            tot += 4 * e["B"] * e["P"] * (len(p["consumers"]) * (e["C"] + e["d"] + 2) + p["ct"])
        return tot
    B, n, P, d, C = e["B"], e["n"], e["P"], e["d"], e["C"]
    rd = n * (d + C) + d + C
    if e["op"] == "fwd":
        return 4 * B * P * rd
    wr = d + sum((d + C) for g in e["grads"] if g is not None) + (2 * n if e["save"] else 0)
    return 4 * B * P * (rd + wr)


def _slot_operands(slots, rows, dev, fill=None):
    """tensors + (k_ptrs, v_ptrs, k_sps, v_sps) of a slot list; a slot with koff keeps key and values in ONE
    tensor (values || key channel slices of a per-layer tensor), as in the model."""
    keep, kp, vp, ks, vs = [], [], [], [], []
    mk = (lambda n: torch.randn(n, device=dev)) if fill is None else (lambda n: torch.full((n,), fill, device=dev))
    for s in slots:
        if s is None:
            kp.append(0); vp.append(0); ks.append(4); vs.append(4)
            continue
        v = mk(rows * s["v_sp"])
        keep.append(v)
        vp.append(v.data_ptr()); vs.append(s["v_sp"])
        if s["koff"] is not None:
            kp.append(v.data_ptr() + 4 * s["koff"]); ks.append(s["v_sp"])
        else:
            k = mk(rows * s["k_sp"])
            keep.append(k)
            kp.append(k.data_ptr()); ks.append(s["k_sp"])
    return keep, kp, vp, ks, vs


def build(e, dev):
    """-> (launch(), keep-alive list) for one trace entry on fresh synthetic operands."""
    T, I = _lib.ptr_table, _lib.int_table
    stream = lambda: torch.cuda.current_stream().cuda_stream
    if e["op"] in ("fwd", "bwd"):
        B, n, P, d, C = e["B"], e["n"], e["P"], e["d"], e["C"]
        rows = B * P
        q = torch.randn(rows * e["q_sp"], device=dev)
        keep, kp, vp, ks, vs = _slot_operands(e["slots"], rows, dev)
        tabs = [T(kp), T(vp), I(ks), I(vs)]
        if e["op"] == "fwd":
            out = torch.empty(rows * C, device=dev)

            def launch():
                _lib.call("davi_dw_attn_slots_fwd", q.data_ptr(), tabs[0][0], tabs[1][0], tabs[2][0], tabs[3][0],
                          out.data_ptr(), 0, B, n, P, d, C, e["q_sp"], C, 0, stream())
            return launch, [q, out, keep, tabs]
        dout = torch.randn(rows * C, device=dev)
        dq = torch.empty(rows * e["q_sp"], device=dev)
        gkeep, dkp, dvp, dks, dvs = _slot_operands(e["grads"], rows, dev, fill=0.0)
        gt = [T(dkp), T(dvp), I(dks), I(dvs)]
        a = torch.empty(rows * n, device=dev) if e["save"] else None
        ds = torch.empty(rows * n, device=dev) if e["save"] else None

        def launch():
            _lib.call("davi_dw_attn_slots_bwd", q.data_ptr(), tabs[0][0], tabs[1][0], tabs[2][0], tabs[3][0],
                      dout.data_ptr(), dq.data_ptr(), gt[0][0], gt[1][0], gt[2][0], gt[3][0], 0, B, n, P, d, C,
                      e["q_sp"], C, 0, 0, a.data_ptr() if a is not None else 0, ds.data_ptr() if ds is not None else 0,
                      stream())
        return launch, [q, dout, dq, keep, gkeep, tabs, gt, a, ds]
    B, P, d, C = e["B"], e["P"], e["d"], e["C"]
    rows = B * P
    ms, ov, ok, ovs, oks, vp, vs, kp, ks, wa, wd, ws, keep = ([] for _ in range(13))
    for p in e["problems"]:
        out = torch.empty(rows * p["ct"], device=dev)
        keep.append(out)
        ms.append(len(p["consumers"]))
        ov.append(out.data_ptr()); ok.append(out.data_ptr() + 4 * p["nv"]); ovs.append(p["ct"]); oks.append(p["ct"])
        for c in p["consumers"]:
            dout, q = torch.randn(rows * C, device=dev), torch.randn(rows * c["q_sp"], device=dev)
            a, ds = torch.rand(rows * c["n"], device=dev), torch.randn(rows * c["n"], device=dev)
            keep += [dout, q, a, ds]
            vp.append(dout.data_ptr()); vs.append(C); kp.append(q.data_ptr()); ks.append(c["q_sp"])
            wa.append(a.data_ptr() + 4 * c["j"]); wd.append(ds.data_ptr() + 4 * c["j"]); ws.append(c["n"])
    mt = (ctypes.c_int * len(ms))(*ms)
    tabs = [T(ov), T(ok), I(ovs), I(oks), T(vp), I(vs), T(kp), I(ks), T(wa), T(wd), I(ws)]

    def launch():
        _lib.call("davi_dw_attn_gather", len(ms), mt, *[t[0] for t in tabs], B, P, d, C, stream())
    return launch, [keep, tabs, mt]


def time_entry(e, dev, reps=3, footprint=3 * L2_BYTES, graph=True):
    """Mean time (ms) of one launch of entry e with cold inputs: the launch is issued back to back on ROTATING
    operand sets whose total footprint exceeds 3x L2, all between ONE pair of CUDA events on the launching
    stream; median of `reps` rounds.  The launches are replayed from a captured CUDA graph, the way the training
    step issues them (a ctypes call costs the host ~10 us, more than the small kernels run)."""
    per_set = max(moved_bytes(e), 1)
    sets = max(2, min(24, int(footprint // per_set) + 1))
    built = [build(e, dev) for _ in range(sets)]
    for fn, _ in built:
        fn()
    torch.cuda.synchronize()
    run = None
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for fn, _ in built:
                fn()
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for fn, _ in built:
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        run = g.replay
        run()
    else:
        def run():
            for fn, _ in built:
                fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / sets)
    del built
    return sorted(ts)[len(ts) // 2], sets


def replay_roofline(trace, dev):
    """-> (algorithmic bytes, total ms, per-launch list) over every launch of the trace."""
    tot_b = tot_t = 0.0
    per = []
    for e in trace:
        ms, sets = time_entry(e, dev)
        tot_b += algorithmic_bytes(e)
        tot_t += ms
        row = {"op": e["op"], "us": round(ms * 1e3, 1), "sets": sets, "moved_MB": round(moved_bytes(e) / 1e6, 1),
               "moved_GBs": round(moved_bytes(e) / ms / 1e6)}
        if e["op"] == "gather":
            row["problems"] = len(e["problems"]); row["consumers"] = sum(len(p["consumers"]) for p in e["problems"])
        else:
            row.update(n=e["n"], C=e["C"], alg_MB=round(algorithmic_bytes(e) / 1e6, 1))
        per.append(row)
    return tot_b, tot_t, per


def select(trace, only):
    """`only`: list like ["fwd:15", "bwd:15", "gather", "gather:16"] (op[:n or problem count]) -> FIRST match each."""
    out = []
    for pat in only:
        op, _, num = pat.partition(":")
        for e in trace:
            if e["op"] != op:
                continue
            size = len(e["problems"]) if op == "gather" else e["n"]
            if not num or int(num) == size:
                out.append(e)
                break
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trace", required=True)
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--once", action="store_true", help="launch every selected entry ONCE after a warm-up launch "
                    "(the mode to run under ncu: -k regex:dw_ -s <number of entries> skips the warm-up launches)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    trace = json.load(open(args.trace))
    if args.only:
        trace = select(trace, args.only)
    if args.once:
        built = [build(e, dev) for e in trace]
        for fn, _ in built:
            fn()
        torch.cuda.synchronize()
        flush = torch.empty(int(2 * L2_BYTES) // 4, device=dev)
        flush.zero_(); flush.sum().item()                       # clean L2: written, then read
        for fn, _ in built:
            fn()
        torch.cuda.synchronize()
        print(json.dumps({"launched": [e["op"] + ":" + str(len(e["problems"]) if e["op"] == "gather" else e["n"])
                                       for e in trace],
                          "algorithmic_MB": [round(algorithmic_bytes(e) / 1e6, 1) for e in trace],
                          "moved_MB": [round(moved_bytes(e) / 1e6, 1) for e in trace]}))
        return
    tb, tt, per = replay_roofline(trace, dev)
    print(json.dumps({"algorithmic_MB": tb / 1e6, "ms": tt, "GBs": tb / tt / 1e6, "per_launch": per}))


if __name__ == "__main__":
    main()
