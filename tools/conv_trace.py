"""Dev tool (GPU box): event timeline of CTA pair 0 of the convolution kernel (csrc/conv_tc.cu).

Builds a second library with -DDAVI_CV_TRACE (tools/_bin/libdavi_cvtrace.so; the product libdavi.so has none of the
stamps), launches one convolution and prints, per k-chunk, the %globaltimer stamps (ns since the first TMA issue):
TMA issued | stage landed (lo warps) | lo slot free | lo done | MMA thread saw `ready` | MMAs + commits issued.

    python tools/conv_trace.py --build-only        # here: nvcc
    python tools/conv_trace.py fwd 40 16 16 316 316 3
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_bin", "libdavi_cvtrace.so")


def build(pair=0):
    from deep_attentive_vi_b200 import build as B
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [B.nvcc_path()] + B.NVCC_FLAGS + ["-DDAVI_HAVE_NONLOCAL_TC", "-DDAVI_CV_TRACE", f"-DDAVI_CV_TRACE_PAIR={pair}", "-I", os.path.join(ROOT, "include"),
                                            "-I", B.CSRC, "-o", OUT] + B._sources()
    subprocess.run(cmd, check=True)
    return OUT


def main():
    if "--build-only" in sys.argv:                 # --build-only [PAIR]: the CTA pair whose events are recorded
        rest = [a for a in sys.argv[1:] if a.isdigit()]
        print(build(int(rest[0]) if rest else 0))
        return
    if not os.path.exists(OUT):
        build()
    import torch
    from deep_attentive_vi_b200 import _lib
    _lib.LIB_PATH = OUT
    lib = _lib.load()
    lib.davi_cv_trace_read.restype = ctypes.c_int
    lib.davi_cv_trace_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    B, H, W, Cin, Cout, k = [int(a) for a in sys.argv[2:8]] if len(sys.argv) >= 8 else (40, 16, 16, 316, 316, 3)
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev); w = torch.randn(Cout, k, k, Cin, device=dev)
    gy = torch.randn(B, H, W, Cout, device=dev)
    y = torch.empty(B, H, W, Cout, device=dev); gw = torch.empty(Cout, k, k, Cin, device=dev)
    wsb = lib.davi_conv2d_workspace_bytes()
    ws = torch.zeros(wsb // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(3):
        if kind == "fwd":
            _lib.call("davi_conv2d_fwd", p(x), p(w), None, p(y), B, H, W, Cin, Cout, k, k, Cin, Cout, p(ws), wsb, st)
        else:
            _lib.call("davi_conv2d_wgrad", p(x), p(gy), p(gw), B, H, W, Cin, Cout, k, k, Cin, Cout, p(ws), wsb, st)
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (2 * 12 * 64))()
    assert lib.davi_cv_trace_read(buf) == 0
    t = lambda rank, role, i: buf[(rank * 12 + role) * 64 + i]
    t0 = t(0, 0, 0)
    print(f"{kind} B={B} {H}x{W} Cin={Cin} Cout={Cout} k={k}; ns since the leader's first TMA issue")
    print("chunk | r0: tma  landed lofree lodone | r1: tma  landed lofree lodone | mma: ready issued")
    for i in range(56):
        if not t(0, 0, i):
            break
        f = lambda v: f"{v - t0:7d}" if v else "      -"
        print(f"{i:5d} | " + " ".join(f(t(0, r, i)) for r in (0, 1, 7, 2)) + " | " + " ".join(f(t(1, r, i)) for r in (0, 1, 7, 2))
              + " | " + f(t(0, 3, i)) + " " + f(t(0, 4, i)))
    print("lo pass of the leader, ns after the stage landed: slot free | A in flight to TMEM | B lo written | TMEM stores done | fences | arrived")
    for i in range(20, 28):
        if t(0, 1, i):
            print(f"{i:5d} | " + " ".join(f"{t(0, r, i) - t(0, 1, i):6d}" for r in (7, 8, 9, 10, 11, 2)))
    for si in range(2):
        if t(0, 6, 8 + 8 * si):
            print(f"epilogue of segment {si}, warp 0 of the leader: accumulator complete {t(0, 6, 2 * si) - t0} | partials available "
                  f"{t(0, 6, 8 + 8 * si) - t0} | column groups done " + " ".join(str(t(0, 6, 9 + 8 * si + j) - t0) for j in range(6) if t(0, 6, 9 + 8 * si + j)))
    print("epilogue (acc_full seen, TMEM drained) per segment: r0 " + " ".join(str(t(0, 6, i) - t0) for i in range(4) if t(0, 6, i))
          + " | r1 " + " ".join(str(t(1, 6, i) - t0) for i in range(4) if t(1, 6, i)))


if __name__ == "__main__":
    main()
