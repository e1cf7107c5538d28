"""Dev tool (GPU box): the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against an fp64 convolution,
with timings beside the cuDNN 3xTF32 path it replaces.

    python tools/conv_check.py fwd 40 16 16 276 276 3        # kind B H W Cin Cout k
    python tools/conv_check.py all                           # the model's shapes, one subprocess per case

Prints one JSON line per case: max error relative to max |reference| and microseconds per call."""
import ctypes
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deep_attentive_vi_b200 import _lib, ops  # noqa: E402

CASES = [
    ("fwd", 2, 16, 16, 32, 32, 1), ("fwd", 2, 16, 16, 64, 48, 3), ("fwd", 3, 16, 16, 276, 276, 3),
    ("fwd", 40, 16, 16, 276, 276, 3), ("fwd", 40, 16, 16, 316, 316, 3), ("fwd", 40, 16, 16, 316, 316, 1),
    ("fwd", 40, 16, 16, 276, 340, 1), ("fwd", 40, 16, 16, 276, 40, 3), ("fwd", 8, 32, 32, 64, 64, 3),
    ("wgrad", 2, 16, 16, 32, 64, 1), ("wgrad", 2, 16, 16, 64, 48, 3), ("wgrad", 40, 16, 16, 276, 276, 3),
    ("wgrad", 40, 16, 16, 316, 316, 3), ("wgrad", 40, 16, 16, 316, 316, 1), ("wgrad", 40, 16, 16, 276, 40, 3),
]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def one(kind, B, H, W, Cin, Cout, k):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev)      # OHWI
    bias = torch.randn(Cout, generator=g).to(dev)
    gy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    wsb = _lib.load().davi_conv2d_workspace_bytes()
    ws = torch.zeros(wsb // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    xn, wn = x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2)                                # NCHW views
    res = dict(kind=kind, B=B, H=H, W=W, Cin=Cin, Cout=Cout, k=k)
    flops = 2.0 * B * H * W * Cin * Cout * k * k
    if kind == "fwd":
        y = torch.empty(B, H, W, Cout, device=dev)
        run = lambda: _lib.call("davi_conv2d_fwd", p(x), p(w), p(bias), p(y), B, H, W, Cin, Cout, k, k, Cin, Cout,
                                p(ws), wsb, st)
        run()
        torch.cuda.synchronize()
        ref = torch.nn.functional.conv2d(xn.double(), wn.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
        res["err"] = float((y.double() - ref).abs().max() / ref.abs().max())
        run()                                                                             # flags left clean?
        torch.cuda.synchronize()
        res["err_2nd_call"] = float((y.double() - ref).abs().max() / ref.abs().max())
        # input gradient through the same kernel
        wt = torch.empty(Cin, k, k, Cout, device=dev)
        gx = torch.empty(B, H, W, Cin, device=dev)
        _lib.call("davi_conv2d_wt", p(w), p(wt), Cout, k * k, Cin, st)
        rung = lambda: _lib.call("davi_conv2d_fwd", p(gy), p(wt), None, p(gx), B, H, W, Cout, Cin, k, k, Cout, Cin,
                                 p(ws), wsb, st)
        rung()
        torch.cuda.synchronize()
        refg = torch.nn.grad.conv2d_input(xn.shape, wn.double(), gy.permute(0, 3, 1, 2).double(), padding=k // 2)
        res["err_dgrad"] = float((gx.double() - refg.permute(0, 2, 3, 1)).abs().max() / refg.abs().max())
        res["us"] = round(timed(run), 1)
        res["us_dgrad"] = round(timed(rung), 1)
        res["TFLOPs"] = round(flops / res["us"] * 1e-6, 1)
        xc = xn.contiguous(memory_format=torch.channels_last)
        wc = wn.contiguous(memory_format=torch.channels_last)
        ops.set_conv_precision("3xtf32")
        res["us_cudnn_3x"] = round(timed(lambda: ops.conv2d_bias(xc, wc, bias, padding=k // 2)), 1)
    elif kind == "fwd_pre":
        # the trainer's path: lo image of the weights computed once (davi_tf32_lo), loaded by TMA (davi_conv2d_fwd_pre);
        # must equal davi_conv2d_fwd_ex bit for bit
        y, y2 = torch.empty(B, H, W, Cout, device=dev), torch.empty(B, H, W, Cout, device=dev)
        w_lo = torch.empty_like(w)
        _lib.call("davi_tf32_lo", p(w), p(w_lo), w.numel(), st)
        run = lambda: _lib.call("davi_conv2d_fwd_ex", p(x), p(w), p(bias), p(y), B, H, W, Cin, Cout, k, k, Cin, Cout, 0,
                                p(ws), wsb, st)
        runp = lambda: _lib.call("davi_conv2d_fwd_pre", p(x), p(w), p(w_lo), p(bias), p(y2), B, H, W, Cin, Cout, k, k, Cin,
                                 Cout, 0, p(ws), wsb, st)
        run()
        runp()
        torch.cuda.synchronize()
        res["bit_identical"] = bool(torch.equal(y, y2))
        res["us"] = round(timed(run), 1)
        res["us_pre"] = round(timed(runp), 1)
        res["TFLOPs_pre"] = round(flops / res["us_pre"] * 1e-6, 1)
    else:
        gw = torch.full((Cout, k, k, Cin), float("nan"), device=dev)
        run = lambda: _lib.call("davi_conv2d_wgrad", p(x), p(gy), p(gw), B, H, W, Cin, Cout, k, k, Cin, Cout, p(ws),
                                wsb, st)
        run()
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_weight(xn.double(), wn.shape, gy.permute(0, 3, 1, 2).double(), padding=k // 2)
        ref = ref.permute(0, 2, 3, 1)
        res["err"] = float((gw.double() - ref).abs().max() / ref.abs().max())
        run()
        torch.cuda.synchronize()
        res["err_2nd_call"] = float((gw.double() - ref).abs().max() / ref.abs().max())
        res["us"] = round(timed(run), 1)
        res["TFLOPs"] = round(flops / res["us"] * 1e-6, 1)
    res["flags_clean"] = bool((ws[:256] == 0).all())
    print(json.dumps(res), flush=True)


def main():
    if sys.argv[1] == "all":                       # one process per kind: a trapped kernel ends only its own list
        for kind in ("fwd", "wgrad"):
            r = subprocess.run([sys.executable, __file__, "list", kind], capture_output=True, text=True, timeout=600)
            print(r.stdout, end="", flush=True)
            if r.returncode:
                print(json.dumps(dict(kind=kind, rc=r.returncode, err=r.stderr[-600:])), flush=True)
        return
    if sys.argv[1] == "dbg":                       # where the time goes: DAVI_OPT_CONV_DEBUG bits (results are wrong)
        B, H, W, Cin, Cout, k = 40, 16, 16, 316, 316, 3
        dev = torch.device("cuda")
        x = torch.randn(B, H, W, Cin, device=dev); w = torch.randn(Cout, k, k, Cin, device=dev)
        gy = torch.randn(B, H, W, Cout, device=dev)
        y = torch.empty(B, H, W, Cout, device=dev); gw = torch.empty(Cout, k, k, Cin, device=dev)
        wsb = _lib.load().davi_conv2d_workspace_bytes()
        ws = torch.zeros(wsb // 4, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        for dbg in (0, 64, 128, 192, 4, 0):
            _lib.set_option(4, dbg)
            f = timed(lambda: _lib.call("davi_conv2d_fwd", p(x), p(w), None, p(y), B, H, W, Cin, Cout, k, k, Cin, Cout, p(ws), wsb, st))
            g = timed(lambda: _lib.call("davi_conv2d_wgrad", p(x), p(gy), p(gw), B, H, W, Cin, Cout, k, k, Cin, Cout, p(ws), wsb, st))
            print(json.dumps(dict(debug=dbg, fwd_us=round(f, 1), wgrad_us=round(g, 1))), flush=True)
        _lib.set_option(4, 0)
        return
    if sys.argv[1] == "fused":                     # cost of the fused activation (swish in the operand path)
        dev = torch.device("cuda")
        lib = _lib.load()
        wsb = lib.davi_conv2d_workspace_bytes()
        ws = torch.zeros(wsb // 4, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        for (B, H, W, Cin, Cout, k) in ((40, 16, 16, 316, 316, 3), (40, 16, 16, 316, 316, 1)):
            x = torch.randn(B, H, W, Cin, device=dev); w = torch.randn(Cout, k, k, Cin, device=dev) * 0.05
            gy = torch.randn(B, H, W, Cout, device=dev)
            y = torch.empty(B, H, W, Cout, device=dev); gw = torch.empty(Cout, k, k, Cin, device=dev)
            f = lambda act: timed(lambda: _lib.call("davi_conv2d_fwd_ex", p(x), p(w), None, p(y), B, H, W, Cin, Cout, k, k, Cin,
                                                    Cout, act, p(ws), wsb, st))
            g = lambda act: timed(lambda: _lib.call("davi_conv2d_wgrad_ex", p(x), p(gy), p(gw), B, H, W, Cin, Cout, k, k, Cin,
                                                    Cout, act, p(ws), wsb, st))
            silu = timed(lambda: torch.nn.functional.silu(x))
            print(json.dumps(dict(k=k, fwd=round(f(0), 1), fwd_act=round(f(1), 1), wgrad=round(g(0), 1),
                                  wgrad_act=round(g(1), 1), torch_silu=round(silu, 1))), flush=True)
        return
    if sys.argv[1] == "list":
        for c in CASES:
            if c[0] == sys.argv[2]:
                one(*c)
        return
    kind = sys.argv[1]
    one(kind, *[int(v) for v in sys.argv[2:8]])


if __name__ == "__main__":
    main()
