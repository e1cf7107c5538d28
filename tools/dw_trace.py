"""Dev tool (GPU box): phase timeline of the first and the last CTA of the persistent depth-wise backward kernel.

Builds a second library with -DDAVI_DW_TRACE (tools/_bin/libdavi_dwtrace.so; the product libdavi.so has none of the
stamps), replays ONE recorded launch of the training step (tools/dw_replay.py operands) on cold inputs and prints
the clock64 stamps relative to the CTA's entry:

    python tools/dw_trace.py --build-only                 # here: nvcc
    python tools/dw_trace.py bwd:10 [bwd:15 ...]          # GPU box

Legend: 1 setup done | 2 producer: ring filled once | 3 producer: all copies issued | 8+t score: weights of tile t
published | 20+t score: tile t finished (dq, dK written) | 33 score: done | 36+t stream: dOut tile of tile t
received | 50+t stream: tile t consumed.  Slots 4/5/6/60/61/62 accumulate the time a role spent blocked.
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
OUT = os.path.join(ROOT, "tools", "_bin", "libdavi_dwtrace.so")


def build():
    from deep_attentive_vi_b200 import build as B
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [B.nvcc_path()] + B.NVCC_FLAGS + ["-DDAVI_HAVE_NONLOCAL_TC", "-DDAVI_DW_TRACE", "-I", os.path.join(ROOT, "include"),
                                            "-I", B.CSRC, "-o", OUT] + B._sources()
    subprocess.run(cmd, check=True)
    return OUT


def main():
    if "--build-only" in sys.argv:
        print(build())
        return
    if not os.path.exists(OUT):
        build()
    import torch
    from deep_attentive_vi_b200 import _lib
    _lib.LIB_PATH = OUT
    lib = _lib.load()
    lib.davi_dw_trace_read.restype = ctypes.c_int
    lib.davi_dw_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    import dw_replay
    trace = json.load(open(os.path.join(ROOT, "profiles", "r02_dw_trace_cifar16_b40.json")))
    pats = [a for a in sys.argv[1:] if not a.startswith("-")] or ["bwd:10"]
    dev = torch.device("cuda")
    mhz = 1965.0
    for e in dw_replay.select(trace, pats):
        built = [dw_replay.build(e, dev) for _ in range(4)]
        for fn, _ in built:
            fn()
            torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 128)()
        assert lib.davi_dw_trace_read(buf) == 0
        print(f"{e['op']} n={e['n']} C={e['C']}  (us since CTA entry)")
        for c, name in ((0, "first CTA"), (1, "last CTA ")):
            row = list(buf[64 * c:64 * c + 64])
            t0 = row[0]
            acc = {4: "producer blocked on free stages", 5: "score blocked on da", 6: "score blocked on a_empty",
                   60: "stream blocked on dOut tile", 61: "stream blocked on weights", 62: "stream blocked on value tiles"}
            pts = sorted(((i, (v - t0) / mhz) for i, v in enumerate(row) if v and i and i not in acc), key=lambda p: p[1])
            print("  " + name + "  " + "  ".join(f"[{i}]{t:6.2f}" for i, t in pts))
            print("            " + "; ".join(f"{txt} {row[i] / mhz:.2f} us" for i, txt in acc.items()))


if __name__ == "__main__":
    main()
