"""Dev tool (GPU box): nonlocal attention forward/backward vs the fp64 oracle for
the shapes of both published configs, with timings.  --simt (davi_set_option)
selects the exact CUDA-core kernels for comparison."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_attentive_vi_b200 import ops  # noqa: E402
from oracle import ops as ref  # noqa: E402


def rel(got, want):
    want = want.detach().double().cpu()
    got = got.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def case(B, HW, d, C, bwd, timing=True):
    torch.manual_seed(HW * 1000 + C)
    dev = "cuda"
    qkv = torch.randn(B, HW, HW, 2 * d + C)
    qkv[..., :2 * d] *= 0.6
    x = torch.randn(B, HW, HW, C)
    go = torch.randn(B, HW, HW, C)
    scale, gamma = torch.tensor(0.8), torch.tensor(0.6)
    qkv_g = qkv.to(dev).requires_grad_()
    x_g = x.to(dev).requires_grad_()
    s_g, g_g = scale.to(dev).requires_grad_(), gamma.to(dev).requires_grad_()
    y = ops.nonlocal_attn(qkv_g[..., :d], qkv_g[..., d:2 * d], qkv_g[..., 2 * d:], x_g, s_g, g_g)
    torch.cuda.synchronize()
    nb = min(B, 2)   # oracle on the first images only (fp64 CPU)
    qkv_r = qkv[:nb].double().requires_grad_()
    x_r = x[:nb].double().requires_grad_()
    s_r, g_r = scale.double().requires_grad_(), gamma.double().requires_grad_()
    want = ref.nonlocal_core(qkv_r[..., :d], qkv_r[..., d:2 * d], qkv_r[..., 2 * d:], x_r, g_r, s_r)
    out = {"B": B, "N": HW * HW, "d": d, "C": C, "fwd_err": rel(y[:nb], want)}
    if bwd:
        y.backward(go.to(dev))
        want.backward(go[:nb].double())
        out["dq_err"] = rel(qkv_g.grad[:nb, ..., :d], qkv_r.grad[..., :d])
        out["dk_err"] = rel(qkv_g.grad[:nb, ..., d:2 * d], qkv_r.grad[..., d:2 * d])
        out["dv_err"] = rel(qkv_g.grad[:nb, ..., 2 * d:], qkv_r.grad[..., 2 * d:])
        if nb == B:
            out["dgamma_err"] = abs(g_g.grad.item() - g_r.grad.item()) / max(abs(g_r.grad.item()), 1e-30)
            out["dscale_err"] = abs(s_g.grad.item() - s_r.grad.item()) / max(abs(s_r.grad.item()), 1e-30)
    if timing:
        from deep_attentive_vi_b200 import _lib
        N = HW * HW
        qd = qkv_g.detach()
        xx, god = x_g.detach(), go.to(dev)
        yb = torch.empty_like(xx)
        lse = torch.empty(B, N, device=dev)
        dqkv = torch.empty_like(qd)
        dgs = torch.empty(2, device=dev)
        wsb = _lib.load().davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
        ws = torch.empty(max(wsb // 4, 1), device=dev)
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        rs = 2 * d + C
        qp, kp, vp = qd.data_ptr(), qd.data_ptr() + 4 * d, qd.data_ptr() + 8 * d
        dqp, dkp, dvp = dqkv.data_ptr(), dqkv.data_ptr() + 4 * d, dqkv.data_ptr() + 8 * d
        sp, gp = s_g.detach().data_ptr(), g_g.detach().data_ptr()
        osave = torch.empty_like(xx)
        fwd = lambda: _lib.call("davi_nonlocal_attn_fwd", qp, kp, vp, xx.data_ptr(), yb.data_ptr(), lse.data_ptr(),
                                osave.data_ptr(), B, N, d, C, rs, rs, rs, C, C, C, sp, gp, st)
        bwdf = lambda: _lib.call("davi_nonlocal_attn_bwd", qp, kp, vp, lse.data_ptr(), osave.data_ptr(),
                                 god.data_ptr(), dqp, dkp, dvp, dgs.data_ptr(), B, N, d, C, rs, rs, rs, C, C, sp, gp,
                                 ws.data_ptr(), wsb, st)

        def timeit(fn):
            ts = []
            for it in range(6):
                for _ in range(3):            # keep the GPU busy while the host enqueues the timed launch
                    flush.fill_(0.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                if it:
                    ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]
        fl = 2 * B * N * N * (d + C)
        ms = timeit(fwd)
        out["fwd_us"] = round(ms * 1e3, 1)
        out["fwd_TFLOPs"] = round(fl / ms / 1e9, 1)
        if bwd:
            ms = timeit(bwdf)
            out["bwd_us"] = round(ms * 1e3, 1)
            out["bwd_TFLOPs"] = round(2.5 * fl / ms / 1e9, 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    bwd = "--bwd" in sys.argv
    small = "--small" in sys.argv
    simt = "--simt" in sys.argv
    if simt:
        from deep_attentive_vi_b200 import _lib
        _lib.set_option(_lib.OPT_NONLOCAL_IMPL, 1)
    print(json.dumps({"impl": "simt" if simt else "tc"}))
    case(2, 16, 32, 64, bwd)
    if not small:
        case(40, 16, 32, 276, bwd)
        case(40, 16, 32, 316, bwd)
        case(40, 16, 32, 296, bwd)
        case(40, 32, 32, 140, bwd)
        case(16, 16, 8, 72, bwd)
        case(16, 16, 8, 88, bwd)
        case(16, 32, 8, 36, bwd)
