"""Dev tool (GPU box): phase timeline of ONE CTA of the nonlocal attention kernels.

Builds a second library with -DDAVI_NL_TRACE (tools/_bin/libdavi_trace.so; the product libdavi.so has none
of the stamps: NL_STAMP expands to nothing there), launches forward and backward once for one shape and prints
the clock64 stamps of CTA (0, 0, z) relative to its entry.  Build here (no GPU needed), run on the GPU box:

    python tools/nl_trace.py --build-only          # here: nvcc
    python tools/nl_trace.py 40 16 32 276          # GPU box: B, H = W, d, C

Stamp legend -- forward / dV (nl_pv_body):  0 entry | 1 setup done | 2 MMA: X ready | 3 MMA: pass A + first
pass-B S issued | 4 MMA: first P/W chunk available | 5 MMA: all issued | 6 softmax: pass A done | 7 softmax:
first P chunk out | 8 softmax: last P chunk out | 9 accumulator complete | 10 tile staged | 11 stores issued.
dQ / dK for N = 256 (nl_ds256_body): 0 entry | 2 MMA: first U/W chunk ready | 3 MMA: dp issued | 4/6 MMA: S(0/1)
issued | 5/7 MMA: dS(0/1) received | 8 MMA: all issued | 10/12 dS warps: S(0/1) complete | 11/13 dS warps:
dS(0/1) handed over | 14 accumulator complete.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_bin", "libdavi_trace.so")


def build():
    from deep_attentive_vi_b200 import build as B
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [B.nvcc_path()] + B.NVCC_FLAGS + ["-DDAVI_HAVE_NONLOCAL_TC", "-DDAVI_NL_TRACE", "-I", os.path.join(ROOT, "include"),
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
    _lib.LIB_PATH = OUT                       # every entry point from the trace build
    lib = _lib.load()
    lib.davi_nl_trace_read.restype = ctypes.c_int
    lib.davi_nl_trace_read.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    args = [int(a) for a in sys.argv[1:] if not a.startswith("-")]
    B, HW, d, C = args if len(args) == 4 else (40, 16, 32, 276)
    N, rs, dev = HW * HW, 2 * d + C, "cuda"
    torch.manual_seed(0)
    qkv = torch.randn(B, N, rs, device=dev) * 0.5
    x, go = torch.randn(B, N, C, device=dev), torch.randn(B, N, C, device=dev)
    y, o, lse = torch.empty_like(x), torch.empty_like(x), torch.empty(B, N, device=dev)
    dqkv, dgs = torch.empty_like(qkv), torch.empty(2, device=dev)
    scale, gamma = torch.tensor(0.8, device=dev), torch.tensor(0.6, device=dev)
    wsb = lib.davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
    ws = torch.empty(max(wsb // 4, 1), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    q, dq = qkv.data_ptr(), dqkv.data_ptr()
    mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965

    def read():
        buf = (ctypes.c_longlong * 96)()
        assert lib.davi_nl_trace_read(buf) == 0
        return [list(buf[32 * z:32 * z + 32]) for z in range(3)]

    def show(title, row):
        t0 = row[0]
        pts = [(i, v - t0) for i, v in enumerate(row) if v and i]
        print(title + "  " + "  ".join(f"[{i}] {c / mhz:6.2f}us" for i, c in sorted(pts, key=lambda p: p[1])))

    for rep in range(3):                      # the last repetition (warm instruction cache, L2) is printed
        _lib.call("davi_nonlocal_attn_fwd", q, q + 4 * d, q + 8 * d, x.data_ptr(), y.data_ptr(), lse.data_ptr(),
                  o.data_ptr(), B, N, d, C, rs, rs, rs, C, C, C, scale.data_ptr(), gamma.data_ptr(), st)
        torch.cuda.synchronize()
        fwd = read()
        _lib.call("davi_nonlocal_attn_bwd", q, q + 4 * d, q + 8 * d, lse.data_ptr(), o.data_ptr(), go.data_ptr(),
                  dq, dq + 4 * d, dq + 8 * d, dgs.data_ptr(), B, N, d, C, rs, rs, rs, C, C, scale.data_ptr(),
                  gamma.data_ptr(), ws.data_ptr(), wsb, st)
        torch.cuda.synchronize()
        bwd = read()
    print(f"B={B} N={N} d={d} C={C}; SM clock assumed {mhz} MHz; times since the CTA's entry")
    show("forward          ", fwd[0])
    show("backward dV  (z=0)", bwd[0])
    show("backward dQ  (z=1)", bwd[1])
    show("backward dK  (z=2)", bwd[2])


if __name__ == "__main__":
    main()
