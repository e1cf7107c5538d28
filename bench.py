#!/usr/bin/env python
"""Benchmark of BASELINE.json's metric: CIFAR-10 16-layer attentive VAE train
images/sec (batch 40 per GPU, synthetic data, random-init weights).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, restated
                                                            # (oracle/, torch CPU, all host threads)
  python bench.py --workload omniglot15|sweep4|is_eval     # BASELINE configs 0, 3, 4 (one JSON line each)

One JSON line on stdout (rank 0).  A "step" is one full optimisation step of the
model (forward, backward, gradient all-reduce, fused Adamax) on one batch of 40
synthetic images per GPU.  `value` is measured with the batches already resident
in HBM; `e2e` repeats the measurement through the public API with pinned-host
inputs copied H2D every step and the step's loss read back D2H.
`roofline` reports the depth-wise attention kernels (the memory-bound hot kernel
north_star names) as THE STEP RUNS THEM: the structure of every launch of one
step (slot tables, strides, which gradients are written directly and which are
gathered) is recorded and replayed stand-alone on cold operands (`frac`), and
the same launches are timed inside an eager step (`in_step`); `traffic` is read
from the committed ncu capture of those launches (profiles/).
`roofline_nonlocal` is the tensor-bound view of the nonlocal attention calls,
`roofline_conv` the same view of the convolution launches of one step (the
repo's tcgen05 kernels, recorded and replayed per shape), `roofline_elbo` the
HBM view of the ELBO reductions.  Precision: every kernel computes at fp32
grade (3xTF32 on the tensor cores); `--convs` selects other convolution paths
(`3xtf32-cudnn` = the round-1 path: cuDNN over split operands; `fp32`; `tf32`,
below the reference's precision), reported as `conv_dtype`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cifar10_16layer_attentive_vae_train_images_per_sec"
UNIT = "images/s"
WORKLOAD = "cifar10-16layer-attentive-vae-118.97M, latent [16,16,20], key/query 20, nonlocal 32, batch 40/GPU"
# BASELINE.json configs beside the headline one (configs[1]); each is one driver-runnable JSON line:
#   --workload omniglot15   configs[0]: 15-layer OMNIGLOT model (7.87 M), batch 16, one ELBO train step
#   --workload sweep4       configs[3]: depth-wise attention kernel sweep
#   --workload is_eval      configs[4]: K = 1000 importance samples, sharded by sample over the ranks
TRAIN_WORKLOADS = {
    "cifar16": dict(flags="CIFAR10_16L", batch=40, metric=METRIC, workload=WORKLOAD, shape=(32, 32, 3), binary=False),
    "omniglot15": dict(flags="OMNIGLOT_15L", batch=16,
                       metric="omniglot_15layer_attentive_vae_train_images_per_sec",
                       workload="omniglot-15layer-attentive-vae-7.87M, latent [16,16,20], key/query 8, nonlocal 8, "
                                "batch 16/GPU, Bernoulli(0.5) 28x28 images padded to 32x32",
                       shape=(32, 32, 1), binary=True),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 7 for n, v in zip(names, s[3:7]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# reference arm: the reference's CPU path restated op for op (oracle/)
# ---------------------------------------------------------------------------

def oracle_setup(batch, seed=0, flags="CIFAR10_16L"):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from deep_attentive_vi_b200 import model as M
    from oracle.model import OracleVAE
    kw = M.parse_flags(**getattr(M, flags))
    torch.manual_seed(seed)
    m = M.DeepVAE(**kw)                     # CPU construction only: supplies shapes + random-init weights
    o = OracleVAE(helpers.oracle_cfg(kw), m.state_dict(), dtype=torch.float32)
    params = [v for k, v in o.sd.items() if v.is_floating_point() and "running" not in k]
    for v in params:
        v.requires_grad_()
    x = helpers.synthetic_images(kw, batch, seed=1000)
    return m, o, params, x


def oracle_step(m, o, params, x, state, lr=0.01):
    """fwd + bwd (autograd = what TF autodiff does) + Adamax, all on the host cores."""
    import torch
    eps, noise = m.draw_noise(x.shape[0], "cpu", training=True)
    nll, kl, _ = o.forward(x, eps, True, noise)
    loss = nll.mean() + kl.mean()
    for p in params:
        p.grad = None
    loss.backward()
    state["t"] += 1
    with torch.no_grad():
        for i, p in enumerate(params):
            if p.grad is None:
                continue
            mom, u = state["m"].setdefault(i, torch.zeros_like(p)), state["u"].setdefault(i, torch.zeros_like(p))
            mom.mul_(0.9).add_(p.grad, alpha=0.1)
            torch.maximum(u.mul_(0.999), p.grad.abs(), out=u)
            p.addcdiv_(mom, u + 1e-3, value=-lr / (1 - 0.9 ** state["t"]))
    return float(loss.detach())


def cpu_steps(flags, batch, steps, warmup):
    """`steps` timed oracle train steps at `batch` after `warmup` untimed ones -> list of seconds per step."""
    m, o, params, x = oracle_setup(batch, flags=flags)
    state = {"t": 0, "m": {}, "u": {}}
    for _ in range(warmup):
        oracle_step(m, o, params, x, state)
    ts = []
    for _ in range(steps):
        t0 = time.time()
        oracle_step(m, o, params, x, state)
        ts.append(time.time() - t0)
    return ts


def train_config(wl, batch, world, cuda_graph=None, convs=None):
    """The `config` object of a training workload: identical for this repo's arm and the reference arm (the
    CPU arm times a bounded sample of THIS workload; what the sample was is said in cpu_baseline.sample)."""
    cfg = {"workload": wl["workload"], "per_gpu_batch": batch, "global_batch": batch * world,
           "parallelism": f"dp{world}"}
    if cuda_graph is not None:
        cfg["cuda_graph"] = cuda_graph
    if convs is not None:
        cfg["convs"] = convs
    return cfg


def run_reference(args):
    """The reference's CPU path (oracle port: TensorFlow is not installable here) on the box's host cores.
    Each of the K timed steps is a bounded sample of the workload (a full train step at --ref-batch images);
    one more step at the workload's own batch size is timed beside them (`full_batch_step`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    if args.workload not in TRAIN_WORKLOADS:
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU arm covers the training workloads, not "
                                                              f"--workload {args.workload}"}), flush=True)
        return 0
    wl = TRAIN_WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = min(args.ref_batch, wl["batch"])
    ts = cpu_steps(wl["flags"], batch, args.steps, max(1, min(args.warmup, 2)))
    dt = sum(ts) / len(ts)
    med = sorted(ts)[len(ts) // 2]
    val = batch / dt
    full = None
    if batch != wl["batch"] and not args.no_full_batch_step:
        tf = cpu_steps(wl["flags"], wl["batch"], 1, 0)[0]
        full = {"batch": wl["batch"], "seconds": tf, "value": wl["batch"] / tf, "unit": UNIT}
    sample = (f"{args.steps} full train steps (fwd+bwd+Adamax) of the same model at batch {batch} per step "
              f"(oracle/ port of the reference's ops, torch CPU fp32, {cores} threads); median step {med:.3f} s")
    line = {
        "impl": "reference", "metric": wl["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(wl, wl["batch"], max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "median_value": batch / med, "full_batch_step": full},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# roofline of the depth-wise attention kernels: the launches ONE training step issues
# ---------------------------------------------------------------------------

def dw_roofline_of_step(trainer, x, device):
    """Runs one EAGER step of the trainer with (a) ops.dw_trace recording the structure of every depth-wise
    attention launch (slot tables, strides, skipped / direct / gathered gradients) and (b) a CUDA-event pair
    around each of those launches (`in_step`: operands as the neighbouring kernels left them), then replays
    exactly the recorded launches on synthetic operands with cold inputs (`stand_alone`, tools/dw_replay.py)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dw_replay
    from deep_attentive_vi_b200 import _lib, ops
    # Rank 0 runs these eager steps ALONE: no collective may be issued (the bucketed gradient exchange starts its
    # all-reduces from gradient hooks inside _step_body when world > 1 -- the other ranks are not here to answer)
    world_saved, trainer.world = trainer.world, 1
    try:
        return _dw_roofline_of_step(trainer, x, device, torch, dw_replay, _lib, ops)
    finally:
        trainer.world = world_saved


def _dw_roofline_of_step(trainer, x, device, torch, dw_replay, _lib, ops):
    trainer._step_body(x)                                   # eager warm-up of this code path
    torch.cuda.synchronize()
    ops.dw_trace, ops.conv_trace, _lib.timed = [], [], {"prefix": "davi_dw_attn", "events": []}
    try:
        trainer._step_body(x)
        torch.cuda.synchronize()
        trace, conv_trace, events = ops.dw_trace, ops.conv_trace, _lib.timed["events"]
    finally:
        ops.dw_trace, ops.conv_trace, _lib.timed = None, None, None
    in_step_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in events)
    by_name = {}
    for name, e0, e1 in events:
        by_name[name] = by_name.get(name, 0.0) + e0.elapsed_time(e1)
    # The event pairs also contain each launch's latency (~10 us per pair on these 13-60 us kernels), so the
    # kernels' own durations inside a step are read from a CUPTI activity trace of one more eager step
    # (torch.profiler, kernel records only) -- the per-kernel table VERDICT r01 computed its 0.44 from.
    kern_ms, kern_by_name = None, {}
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            trainer._step_body(x)
            torch.cuda.synchronize()
        for ev in prof.key_averages():
            if "davi::dw_" in ev.key:
                us = getattr(ev, "self_device_time_total", None)
                if us is None:
                    us = getattr(ev, "self_cuda_time_total", 0.0)
                short = ev.key.split("davi::")[1].split("(")[0]
                kern_by_name[short] = {"ms": us / 1e3, "launches": ev.count}
        kern_ms = sum(v["ms"] for v in kern_by_name.values())
    except Exception as exc:                                  # CUPTI unavailable: the event numbers remain
        kern_by_name = {"error": repr(exc)}
    alg = sum(dw_replay.algorithmic_bytes(e) for e in trace)
    moved = sum(dw_replay.moved_bytes(e) for e in trace)
    tb, tt, per = dw_replay.replay_roofline(trace, device)
    assert tb == alg
    return {"trace": trace, "conv_trace": conv_trace, "launches": len(trace), "algorithmic_bytes": alg, "moved_bytes": moved,
            "in_step_ms": in_step_ms, "in_step_ms_by_entry_point": by_name, "in_step_kernel_ms": kern_ms,
            "in_step_kernels": kern_by_name, "stand_alone_ms": tt, "per_launch": per}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernels from the committed ncu capture
    (profiles/r02_ncu_dw_attn_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_dw_attn_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def elbo_roofline(batch, device):
    """HBM view of the ELBO reductions at the model's shapes: 16 gauss_kl layers ([B,16,16,2*20]) forward +
    backward and the DLM NLL ([B,32,32,50], gradient fused into the forward pass), rotating operand sets."""
    import torch
    from deep_attentive_vi_b200 import _lib
    stream = torch.cuda.current_stream().cuda_stream
    B, P, z = batch, 256, 20
    out = {}

    def rot(make, nbytes):
        sets = max(2, min(64, int(3 * 126e6 // nbytes) + 1))
        built = [make() for _ in range(sets)]
        for fn in built:
            fn()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for fn in built:
                fn()
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) / sets)
        return sorted(ts)[1]

    def kl_fwd():
        post, prior, eps = (torch.randn(B, P, 2 * z, device=device), torch.randn(B, P, 2 * z, device=device),
                            torch.randn(B, P, z, device=device))
        n1, n2 = torch.randn(B, P, z, device=device), torch.randn(B, P, z, device=device)
        xi, kl, grad = torch.empty(B, P, z, device=device), torch.empty(B, device=device), torch.empty(B, P, 6 * z, device=device)
        return lambda: _lib.call("davi_gauss_kl_fwd", post.data_ptr(), prior.data_ptr(), eps.data_ptr(), n1.data_ptr(),
                                 n2.data_ptr(), xi.data_ptr(), kl.data_ptr(), 0, 0, B, P, z, -5.0, 5.0, -1.0, 5.0,
                                 1e-3, 1e-3, grad.data_ptr(), stream)

    def kl_bwd():
        grad = torch.randn(B, P, 6 * z, device=device)
        dxi, dkl = torch.randn(B, P, z, device=device), torch.randn(B, device=device)
        dpost, dprior = torch.empty(B, P, 2 * z, device=device), torch.empty(B, P, 2 * z, device=device)
        return lambda: _lib.call("davi_gauss_kl_bwd_saved", grad.data_ptr(), dxi.data_ptr(), dkl.data_ptr(),
                                 dpost.data_ptr(), dprior.data_ptr(), B, P, z, stream)

    def dlm():
        params = torch.randn(B, 1024, 50, device=device)
        x = torch.rand(B, 1024, 3, device=device) * 2 - 1
        nll, g = torch.empty(B, device=device), torch.empty_like(params)
        return lambda: _lib.call("davi_dlm_nll_fwd", params.data_ptr(), x.data_ptr(), nll.data_ptr(), g.data_ptr(),
                                 B, 1024, -7.0, stream)

    for name, make, nbytes in (("gauss_kl_fwd+grad", kl_fwd, 4 * B * P * z * 14), ("gauss_kl_bwd_saved", kl_bwd, 4 * B * P * z * 11),
                               ("dlm_nll_fwd+grad", dlm, 4 * B * 1024 * (50 + 3 + 50))):
        t = rot(make, nbytes)
        out[name] = {"us": round(t * 1e3, 2), "MB": round(nbytes / 1e6, 2), "GBs": round(nbytes / t / 1e6)}
    return out


def conv_roofline(trace, device):
    """Tensor-pipe view of the model's own convolution launches (`trace`: one (kind, B, H, W, Cin, Cout, kh, kw) per
    davi_conv2d_fwd / davi_conv2d_wgrad call of a step, recorded by ops.conv_trace; the input gradient is a "fwd"
    call with the channel counts swapped).  Algorithmic flops 2 B H W Cin Cout kh kw per call over the CUDA-event time
    of the C-ABI calls, each distinct shape launched back to back on rotating operand sets whose footprint exceeds
    L2.  The kernels execute 3 tcgen05 TF32 MMAs per algorithmic product (3xTF32)."""
    import collections
    import torch
    from deep_attentive_vi_b200 import _lib
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    wsb = lib.davi_conv2d_workspace_bytes()
    ws = torch.zeros(wsb // 4, device=device)
    tot_f = tot_t = 0.0
    per_shape = []
    for (kind, B, H, W, Cin, Cout, kh, kw), count in sorted(collections.Counter(trace).items()):
        per_set = 4 * B * H * W * (Cin + Cout)
        sets = max(2, min(12, int(1.5 * 126e6 // per_set) + 1))
        x = [torch.randn(B, H, W, Cin, device=device) for _ in range(sets)]
        y = [torch.randn(B, H, W, Cout, device=device) for _ in range(sets)]
        w = torch.randn(Cout, kh, kw, Cin, device=device) * 0.05
        b = torch.randn(Cout, device=device)

        w_lo = torch.empty_like(w)
        _lib.call("davi_tf32_lo", w.data_ptr(), w_lo.data_ptr(), w.numel(), stream)

        def run(i):
            if kind == "fwd_pre":        # the trainer's convolutions: lo image of the weights kept per step, loaded by TMA
                _lib.call("davi_conv2d_fwd_pre", x[i].data_ptr(), w.data_ptr(), w_lo.data_ptr(), b.data_ptr(), y[i].data_ptr(),
                          B, H, W, Cin, Cout, kh, kw, Cin, Cout, 0, ws.data_ptr(), wsb, stream)
            elif kind == "fwd":
                _lib.call("davi_conv2d_fwd", x[i].data_ptr(), w.data_ptr(), b.data_ptr(), y[i].data_ptr(), B, H, W, Cin,
                          Cout, kh, kw, Cin, Cout, ws.data_ptr(), wsb, stream)
            else:
                _lib.call("davi_conv2d_wgrad", x[i].data_ptr(), y[i].data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, kh, kw,
                          Cin, Cout, ws.data_ptr(), wsb, stream)

        for i in range(min(sets, 3)):
            run(i)
        torch.cuda.synchronize()
        reps = max(sets, 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(i % sets)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        fl = 2.0 * B * H * W * Cin * Cout * kh * kw
        tot_f += fl * count
        tot_t += us * count
        per_shape.append({"k": kind, "B": B, "HW": [H, W], "Cin": Cin, "Cout": Cout, "ks": [kh, kw], "calls": count,
                          "us": round(us, 1), "TFLOPs": round(fl / us * 1e-6, 1)})
        del x, y
    return tot_f, tot_t / 1e3, per_shape


def nl_roofline(shapes, device):
    """Tensor-pipe view of the model's nonlocal attention calls (`shapes`: one (B, N, d, C) per call of a
    step, recorded by forward hooks).  Algorithmic flops -- forward 2 B N^2 (d + C), backward
    2 B N^2 (2 C + 3 d): the five GEMMs of the gradient plus one recomputation of S -- over the CUDA-event
    time of the C-ABI calls, each shape launched back to back on rotating operand sets with a footprint
    above 3x L2.  The kernels execute 3 tcgen05 TF32 MMAs per algorithmic product (3xTF32)."""
    import collections
    import torch
    from deep_attentive_vi_b200 import _lib
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    scale, gamma = torch.tensor(0.8, device=device), torch.tensor(0.6, device=device)
    tot_f = tot_t = 0.0
    per_shape = []
    for (B, N, d, C_alg), count in sorted(collections.Counter(shapes).items()):
        C = C_alg + (-C_alg) % 4          # the layer zero-pads odd channel counts (C = 138 in the pre-processor)
        rs = 2 * d + C
        per_set = 4 * B * N * (2 * rs + 4 * C)
        sets = max(2, min(12, int(3 * 126e6 // per_set) + 1))
        qkv = [torch.randn(B, N, rs, device=device) * 0.5 for _ in range(sets)]
        x = [torch.randn(B, N, C, device=device) for _ in range(sets)]
        go = [torch.randn(B, N, C, device=device) for _ in range(sets)]
        y = [torch.empty(B, N, C, device=device) for _ in range(sets)]
        o = [torch.empty(B, N, C, device=device) for _ in range(sets)]
        lse = [torch.empty(B, N, device=device) for _ in range(sets)]
        dqkv = [torch.empty(B, N, rs, device=device) for _ in range(2)]
        dgs = torch.empty(2, device=device)
        wsb = lib.davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
        ws = torch.empty(max(wsb // 4, 1), device=device)

        def fwd(i):
            q = qkv[i].data_ptr()
            _lib.call("davi_nonlocal_attn_fwd", q, q + 4 * d, q + 8 * d, x[i].data_ptr(), y[i].data_ptr(),
                      lse[i].data_ptr(), o[i].data_ptr(), B, N, d, C, rs, rs, rs, C, C, C, scale.data_ptr(),
                      gamma.data_ptr(), stream)

        def bwd(i):
            q, dq = qkv[i].data_ptr(), dqkv[i % 2].data_ptr()
            _lib.call("davi_nonlocal_attn_bwd", q, q + 4 * d, q + 8 * d, lse[i].data_ptr(), o[i].data_ptr(),
                      go[i].data_ptr(), dq, dq + 4 * d, dq + 8 * d, dgs.data_ptr(), B, N, d, C, rs, rs, rs, C, C,
                      scale.data_ptr(), gamma.data_ptr(), ws.data_ptr(), wsb, stream)

        flops = {"fwd": 2.0 * B * N * N * (d + C_alg), "bwd": 2.0 * B * N * N * (2 * C_alg + 3 * d)}
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            for i in range(sets):
                fn(i)                                   # forward first: the backward reads lse / o of its set
            ts = []
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(sets):
                    fn(i)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) / sets)
            t = sorted(ts)[1]
            tot_f += count * flops[name]
            tot_t += count * t
            per_shape.append({"k": name, "B": B, "N": N, "d": d, "C": C_alg, "calls": count, "us": round(t * 1e3, 1),
                              "TFLOPs": round(flops[name] / t / 1e9, 1)})
        del qkv, x, go, y, o, lse, dqkv
    return tot_f, tot_t, per_shape


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------

CONV_NOTE = {
    "3xtf32": "own tcgen05 implicit-GEMM kernels (csrc/conv_tc.cu: CTA pairs, stream-K, 3xTF32 error compensation built in "
              "shared memory / TMEM, fp32-grade) for stride-1 'same' convolutions; cuDNN TF32 with split operands over a "
              "tripled contraction axis for the few shapes outside their envelope (stride 2, 138-channel pre-processor)",
    "3xtf32-cudnn": "cuDNN on TF32 tensor cores with 3xTF32 error compensation (fp32-grade: operands split hi/lo by "
                    "davi_split_tf32, three products accumulated in fp32 over a tripled contraction axis)",
    "fp32": "cuDNN fp32 (CUDA cores)",
    "tf32": "cuDNN plain TF32 (opt-in; operands truncated to 10 mantissa bits, below the reference's fp32)",
}


def dist_setup():
    import torch
    import torch.distributed as dist
    from deep_attentive_vi_b200 import _lib, build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    _lib.load()
    torch.backends.cudnn.benchmark = True
    return world, rank, local, device


def dist_teardown(world, local):
    import torch
    import torch.distributed as dist
    if world > 1:
        # Tear-down: a captured CUDA graph holds NCCL kernels, and destroy_process_group() with such a
        # graph alive can block forever.  Synchronise, then leave without running the destructors.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_davi(args):
    import torch
    import torch.distributed as dist
    from deep_attentive_vi_b200 import _lib, ops
    from deep_attentive_vi_b200 import model as M
    from deep_attentive_vi_b200.util.trainer import Trainer

    world, rank, local, device = dist_setup()
    if args.conv_sm_reserve >= 0:
        _lib.set_option(_lib.OPT_CONV_SM_RESERVE, args.conv_sm_reserve)
    ops.conv_pre_lo = not args.no_pre_lo
    if args.plain_slices:
        ops.split_channels_fused = False
    wl = TRAIN_WORKLOADS[args.workload]
    flags = getattr(M, wl["flags"])
    B = args.batch or wl["batch"]
    g = torch.Generator().manual_seed(1000 + rank)
    nbuf = 4
    if wl["binary"]:      # Bernoulli(0.5) on 28x28, zero-padded to 32x32
        mk = lambda: torch.nn.functional.pad((torch.rand(B, 28, 28, 1, generator=g) < 0.5).float(), (0, 0, 2, 2, 2, 2))
    else:
        mk = lambda: torch.randint(0, 256, (B, *wl["shape"]), generator=g).float()
    host = [mk().pin_memory() for _ in range(nbuf)]
    dev_in = [h.to(device) for h in host]
    h2d_bytes = host[0].numel() * 4

    def sync_all():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def timed(loop_body, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            loop_body(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(convs, sampler=None, rooflines=False):
        """Whole train steps of the model with the host-side convolutions in precision `convs`
        (the hot-path kernels always compute in fp32): device-resident and end-to-end timings."""
        ops.set_conv_impl("cudnn" if convs == "3xtf32-cudnn" else "tcgen05")
        ops.set_conv_precision("3xtf32" if convs == "3xtf32-cudnn" else convs)
        torch.manual_seed(0)
        model = M.DeepVAE(**M.parse_flags(**flags)).to(device)
        trainer = Trainer(model, device=device, world_size=world, rank=rank, use_graph=not args.eager,
                          graph_warmup=min(3, max(1, args.warmup - 1)), train_size=50000, batch_size=B)
        last = {}

        def resident_step(s):
            trainer.train_step(dev_in[s % nbuf])

        def e2e_step(s):
            x = host[s % nbuf].to(device, non_blocking=True)          # H2D from pinned memory
            nll, kl = trainer.train_step(x)
            last["loss"] = float((nll.mean() + kl.mean()).item())      # D2H read of the step's loss

        from deep_attentive_vi_b200.layers.attention import NonlocalResNetBlock
        nl_shapes, hooks = [], []
        for mod in model.modules():
            if isinstance(mod, NonlocalResNetBlock):
                hooks.append(mod.register_forward_hook(
                    lambda m_, inp, out_: nl_shapes.append((inp[0].shape[0], inp[0].shape[1] * inp[0].shape[2],
                                                            m_._key_dim, inp[0].shape[3]))))
        for s in range(args.warmup):
            resident_step(s)
            for h in hooks:                                          # shapes of ONE step; no hooks under capture
                h.remove()
            hooks = []
        if sampler is not None:
            sampler.start()
        calls0 = _lib.launch_count()
        ms = timed(resident_step, args.steps)
        launches = _lib.launch_count() - calls0
        for s in range(min(args.warmup, 2)):
            e2e_step(s)
        ms_e2e = timed(e2e_step, args.steps)
        clocks = sampler.stop() if sampler is not None else None
        out = {"ms": ms, "ms_e2e": ms_e2e, "launches": launches, "loss": last.get("loss"), "clocks": clocks,
               "params": model.count_params(), "nl_shapes": nl_shapes}
        if rooflines and rank == 0:
            out["dw"] = dw_roofline_of_step(trainer, dev_in[0], device)
        del trainer, model
        torch.cuda.empty_cache()
        return out

    want_roofline = not args.no_roofline
    r = measure(args.convs, ClockSampler(local) if rank == 0 else None, rooflines=want_roofline)
    ms, ms_e2e, launches, clocks, nparams = r["ms"], r["ms_e2e"], r["launches"], r["clocks"], r["params"]
    last = {"loss": r["loss"]}
    alt = None
    if args.alt_convs and args.alt_convs != args.convs and world == 1:   # second trainer + graph: one GPU only
        alt = measure(args.alt_convs)

    value = B * world * args.steps / (ms / 1e3)
    e2e_val = B * world * args.steps / (ms_e2e / 1e3)
    cfg = train_config(wl, B, world)
    cfg.update({"params": nparams, "cuda_graph": not args.eager, "convs": CONV_NOTE[args.convs],
                "weights_lo_image": "once per step (davi_tf32_lo over all kernels, loaded by TMA)" if ops.conv_pre_lo else
                                    "derived on chip from every staged weight tile",
                "l2": "inputs (2 GB activations per step) exceed L2; the stand-alone roofline launches rotate over "
                      "operand sets with a total footprint > 3x L2 (every launch reads cold inputs)"})
    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "conv_dtype": args.convs, "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": last.get("loss")},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if alt is not None:
        # opt-in precision, measured in the same run for the record: NOT the headline
        line["alt_convs"] = {"convs": CONV_NOTE[args.alt_convs], "value": B * world * args.steps / (alt["ms"] / 1e3),
                             "unit": UNIT, "ms_per_step": alt["ms"] / args.steps,
                             "e2e_value": B * world * args.steps / (alt["ms_e2e"] / 1e3)}
    if rank == 0:
        if want_roofline:
            peak, which = measured_peak()
            dw = r["dw"]
            alg = dw["algorithmic_bytes"]
            ach_alone = alg / dw["stand_alone_ms"] / 1e6
            ach_step = alg / dw["in_step_ms"] / 1e6
            ach_kern = alg / dw["in_step_kernel_ms"] / 1e6 if dw["in_step_kernel_ms"] else None
            tr = ncu_traffic()
            line["roofline"] = {
                "bound": "hbm",
                "kernel": f"depth-wise attention: the {dw['launches']} launches one training step issues "
                          "(davi_dw_attn_slots_fwd / _bwd / davi_dw_attn_gather), recorded from the step and replayed",
                # headline fraction: the recorded launches replayed stand-alone on cold operands (rotating sets > 3x L2)
                "achieved": ach_alone, "peak": peak, "unit": "GB/s", "frac": ach_alone / peak,
                "frac_of_nominal_8TBs": ach_alone / 8000.0, "peak_source": which,
                # the same launches INSIDE an eager step, operands as the neighbouring kernels left them: kernel durations
                # from a CUPTI activity trace (`frac`), and CUDA-event pairs around each launch (`event_pairs`: an upper
                # bound on time, every pair also holds the launch latency)
                "in_step": {"achieved": ach_kern, "frac": ach_kern / peak if ach_kern else None,
                            "kernel_ms_per_step": dw["in_step_kernel_ms"], "kernels": dw["in_step_kernels"],
                            "timing": "kernel records of a CUPTI activity trace (torch.profiler) of one eager step",
                            "event_pairs": {"achieved": ach_step, "frac": ach_step / peak,
                                            "ms_per_step": dw["in_step_ms"],
                                            "ms_by_entry_point": dw["in_step_ms_by_entry_point"]}},
                "algorithmic_MB_per_step": alg / 1e6, "moved_MB_per_step": dw["moved_bytes"] / 1e6,
                "kernel_ms_per_step": dw["stand_alone_ms"],
                "traffic": tr["traffic_bytes_per_launch_sum"] if tr else None,
                "traffic_detail": tr,
            }
            if args.roofline_detail:
                with open(args.roofline_detail, "w") as f:
                    json.dump(dw["per_launch"], f)
            if args.dw_trace_out:
                with open(args.dw_trace_out, "w") as f:
                    json.dump(dw["trace"], f)
            tf, tt_nl, per_shape = nl_roofline(r["nl_shapes"], device)
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    mp = json.load(f)
                # kernels timed alone, back to back: the burst figure; the sustained one is printed beside it
                tf32_peak, tf32_sust = float(mp["bf16_tflops"]) / 2, float(mp["bf16_tflops_sustained"]) / 2
                tf32_src = ("half of the measured bf16 burst peak (MEASURED_PEAKS.json bf16_tflops; TF32 runs at half "
                            "the bf16 rate)")
            except Exception:
                tf32_peak, tf32_sust, tf32_src = 1125.0, 1125.0, "fallback: nominal dense TF32 (half of 2.25 PF bf16)"
            ach_tf = tf / tt_nl / 1e9
            line["roofline_nonlocal"] = {
                "bound": "tensor", "kernel": f"nonlocal attention fwd + bwd ({len(r['nl_shapes'])} calls of one step)",
                "achieved": ach_tf, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach_tf / tf32_peak,
                "executed_frac": 3 * ach_tf / tf32_peak, "frac_of_sustained_peak": ach_tf / tf32_sust,
                "peak_source": tf32_src,
                "note": "achieved = algorithmic flops / CUDA-event time; the kernels execute 3 TF32 MMAs per product "
                        "(3xTF32 error compensation, fp32-grade), so the tensor pipe runs at executed_frac",
                "kernel_ms_per_step": tt_nl, "per_shape": per_shape}
            line["roofline_elbo"] = {"bound": "hbm", "peak": peak, "unit": "GB/s", "kernels": elbo_roofline(B, device)}
            if dw.get("conv_trace"):
                cf, ct, cper = conv_roofline(dw["conv_trace"], device)
                ach_c = cf / ct / 1e9
                line["roofline_conv"] = {
                    "bound": "tensor", "kernel": f"convolutions on the own tcgen05 kernels: {len(dw['conv_trace'])} launches of one "
                                                 "step (forward, input gradient, weight gradient), conv_gemm_kernel<0/1>",
                    "achieved": ach_c, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach_c / tf32_peak,
                    "executed_frac": 3 * ach_c / tf32_peak, "executed_frac_of_nominal_1125": 3 * ach_c / 1125.0,
                    "peak_source": tf32_src, "kernel_ms_per_step": ct,
                    "ncu": "profiles/r02_ncu_conv_tc_summary.txt: tcgen05 TF32 op rate 57-59 % of the SM's 4096 ops/clk "
                           "(sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32 pct_of_peak_sustained_elapsed)",
                    "per_shape": sorted(cper, key=lambda e: -e["us"] * e["calls"])[:14]}
        if world == 1 and not args.no_cpu_baseline:
            # median of >= 5 steps after 2 warm-ups, all host threads, bounded to --ref-batch images
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cb = min(args.ref_batch, wl["batch"])
            ts = cpu_steps(wl["flags"], cb, 5, 2)
            med = sorted(ts)[len(ts) // 2]
            line["cpu_baseline"] = {"value": cb / med, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"median of 5 full train steps (after 2 warm-ups) of the same model at "
                                              f"batch {cb} (oracle/ port, torch CPU fp32, {cores} threads)"}
        print(json.dumps(line), flush=True)
    dist_teardown(world, local)
    return 0


# ---------------------------------------------------------------------------
# BASELINE configs[3]: depth-wise attention kernel sweep
# ---------------------------------------------------------------------------

def run_sweep4(args):
    """L = 4..32 slots x 16x16 / 32x32 maps x key_dim 8/20/32 x batch 40..320 through the stacked C-ABI entry
    points (davi_dw_attn_fwd / _bwd, the drop-in for att.py:135-164).  A "step" = one pass over every shape of
    the sweep, forward + backward; value = algorithmic GB/s over the pass.  Shapes above --sweep-max-gb per
    call are skipped (listed).  Inputs of consecutive launches differ and each call moves > L2."""
    import torch
    from deep_attentive_vi_b200 import _lib
    world, rank, local, device = dist_setup()
    stream = torch.cuda.current_stream().cuda_stream
    shapes, skipped = [], []
    for HW in (16, 32):
        for d, C in ((8, 72), (20, 276), (32, 288)):
            for n in (4, 8, 12, 16, 17, 24, 32):
                for B in (40, 80, 160, 320):
                    gb = 4 * B * HW * HW * (2 * n * (d + C) + 2 * d + C) / 1e9
                    (shapes if gb <= args.sweep_max_gb else skipped).append((B, n, HW, d, C))
    shapes = shapes[rank::world]                                     # independent kernels: shard the list
    biggest = max(4 * B * HW * HW * n * (d + C) for B, n, HW, d, C in shapes)
    pool = torch.randn(int(biggest * 2.2) // 4 + (1 << 22), device=device)     # operands carved out of one pool
    clocks = ClockSampler(local) if rank == 0 else None
    per, tot_b, tot_ms = [], 0.0, 0.0

    def one_pass(record):
        nonlocal tot_b, tot_ms
        for B, n, HW, d, C in shapes:
            P = HW * HW
            off = 0

            def take(k):
                nonlocal off
                t = pool[off:off + k]; off += (k + 63) // 64 * 64
                return t
            q, K, V, go = take(B * P * d), take(B * n * P * d), take(B * n * P * C), take(B * P * C)
            out, dq, dK, dV = take(B * P * C), take(B * P * d), take(B * n * P * d), take(B * n * P * C)
            fb = 4 * B * P * (n * (d + C) + d + C)
            bb = 4 * B * P * (2 * n * (d + C) + 2 * d + C)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            _lib.call("davi_dw_attn_fwd", q.data_ptr(), K.data_ptr(), V.data_ptr(), out.data_ptr(), 0, B, n, P, d, C,
                      d, n * P * d, P * d, d, n * P * C, P * C, C, C, 0, stream)
            e[1].record()
            _lib.call("davi_dw_attn_bwd", q.data_ptr(), K.data_ptr(), V.data_ptr(), go.data_ptr(), dq.data_ptr(),
                      dK.data_ptr(), dV.data_ptr(), 0, B, n, P, d, C, d, n * P * d, P * d, d, n * P * C, P * C, C, C,
                      0, stream)
            e[2].record()
            if record:
                e[2].synchronize()
                tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
                per.append({"B": B, "n": n, "HW": HW, "d": d, "C": C, "fwd_GBs": round(fb / tf / 1e6),
                            "bwd_GBs": round(bb / tb / 1e6)})
                tot_b += fb + bb; tot_ms += tf + tb

    for _ in range(max(1, args.warmup)):
        one_pass(False)
    torch.cuda.synchronize()
    if clocks:
        clocks.start()
    calls0 = _lib.launch_count()
    for _ in range(args.steps):
        one_pass(True)
    launches = _lib.launch_count() - calls0
    stats = torch.tensor([tot_b, tot_ms], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        gathered = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        tot_b = sum(float(t[0]) for t in gathered); tot_ms = max(float(t[1]) for t in gathered)
    if rank == 0:
        peak, which = measured_peak()
        ach = tot_b / tot_ms / 1e6
        best = {}
        for row in per[-len(shapes):]:
            best[(row["HW"], row["d"])] = max(best.get((row["HW"], row["d"]), 0), row["fwd_GBs"])
        line = {"metric": "dw_attn_sweep_GBs", "value": ach, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "depth-wise attention kernel sweep: n in {4,8,12,16,17,24,32} x {16x16,32x32} x "
                                       "(d,C) in {(8,72),(20,276),(32,288)} x B in {40,80,160,320}, fwd+bwd, stacked API",
                           "shapes": len(shapes) * world, "skipped_above_GB": [args.sweep_max_gb, len(skipped)],
                           "l2": "every call moves more than L2 and follows a different shape"},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "frac_of_nominal_8TBs": ach / 8000.0, "peak_source": which, "traffic": None},
                "e2e": {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "kernel sweep through the stacked C-ABI entry points: operands are device-resident by "
                                "definition (no host-side producer exists for them)"},
                "gpu_launches": launches, "clocks": clocks.stop() if clocks else None,
                "per_shape": per[-len(shapes):]}
        print(json.dumps(line), flush=True)
    dist_teardown(world, local)
    return 0


# ---------------------------------------------------------------------------
# BASELINE configs[4]: importance-sampled evaluation, sharded by sample
# ---------------------------------------------------------------------------

def run_is_eval(args):
    """K = --is-samples importance samples for a batch of 16 test images, rank r computing K / world of them in
    CUDA-graphed chunks of 25 (reference: util/eval.py:140-146 + model/vae.py:452-539); one all-gather of [B]
    partial logsumexps.  A "step" = one full K-sample estimate of the batch.  value = samples x images / s."""
    import torch
    import torch.distributed as dist
    from deep_attentive_vi_b200 import _lib, ops
    from deep_attentive_vi_b200 import model as M
    from deep_attentive_vi_b200.util.eval import GraphedChunks, importance_log_likelihood
    world, rank, local, device = dist_setup()
    ops.set_conv_impl("cudnn" if args.convs == "3xtf32-cudnn" else "tcgen05")
    ops.set_conv_precision("3xtf32" if args.convs == "3xtf32-cudnn" else args.convs)
    torch.manual_seed(0)                                      # same weights on every rank
    model = M.DeepVAE(**M.parse_flags(**M.CIFAR10_16L)).to(device).eval()
    B, K, chunk = 16, args.is_samples, 25
    host = torch.randint(0, 256, (B, 32, 32, 3), generator=torch.Generator().manual_seed(7)).float().pin_memory()
    x = host.to(device)
    chunks = GraphedChunks(model, x)
    for _ in range(max(1, min(args.warmup, 2))):
        importance_log_likelihood(model, x, 2 * chunk * world, chunk, world, rank, chunks=chunks)
    clocks = ClockSampler(local) if rank == 0 else None

    def timed(fn):
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    out = {}
    if clocks:
        clocks.start()
    calls0 = _lib.launch_count()
    ms = timed(lambda: out.__setitem__("lp", importance_log_likelihood(model, x, K, chunk, world, rank, chunks=chunks)))
    launches = _lib.launch_count() - calls0

    def e2e():
        xe = host.to(device, non_blocking=True)
        out["mean"] = float(importance_log_likelihood(model, xe, K, chunk, world, rank, chunks=chunks).mean().item())
    ms_e2e = timed(e2e)
    if rank == 0:
        line = {"metric": "cifar10_16layer_is_eval_sample_images_per_sec", "value": K * B * args.steps / (ms / 1e3),
                "unit": "sample-images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "conv_dtype": args.convs, "data": "synthetic",
                "config": {"workload": f"cifar10-16layer-attentive-vae-118.97M test NLL, {K} importance samples, batch 16, "
                                       "sharded by sample", "samples_per_rank": K // world, "chunk": chunk,
                           "cuda_graph": True, "convs": CONV_NOTE[args.convs]},
                "e2e": {"value": K * B * args.steps / (ms_e2e / 1e3), "unit": "sample-images/s",
                        "h2d_bytes_per_step": B * 32 * 32 * 3 * 4, "d2h_bytes_per_step": 4,
                        "mean_log_likelihood_nats": out.get("mean")},
                "gpu_launches": launches, "clocks": clocks.stop() if clocks else None}
        print(json.dumps(line), flush=True)
    dist_teardown(world, local)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="davi", choices=["davi", "reference"])
    ap.add_argument("--workload", default="cifar16", choices=["cifar16", "omniglot15", "sweep4", "is_eval"],
                    help="cifar16 = BASELINE's headline config (default); the others are BASELINE configs 0, 3, 4")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's, 40 / 16)")
    ap.add_argument("--ref-batch", type=int, default=4, help="images per step of the bounded CPU sample")
    ap.add_argument("--no-full-batch-step", action="store_true",
                    help="reference arm: skip the one extra CPU step at the workload's full batch size")
    ap.add_argument("--convs", default="3xtf32", choices=["3xtf32", "3xtf32-cudnn", "fp32", "tf32"],
                    help="precision of the host-side cuDNN convolutions (default: fp32-grade 3xTF32)")
    ap.add_argument("--alt-convs", default="", choices=["", "3xtf32", "3xtf32-cudnn", "fp32", "tf32"],
                    help="second precision measured for the record into 'alt_convs' ('' = skip)")
    ap.add_argument("--fp32-convs", action="store_true", help="same as --convs fp32")
    ap.add_argument("--eager", action="store_true", help="do not capture the step into a CUDA graph")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--plain-slices", action="store_true",
                    help="A/B: channel splits of the variational layers as plain slices (zero fill + copy + sum per piece on "
                         "the way back) instead of ops.split_channels (one concatenation)")
    ap.add_argument("--no-pre-lo", action="store_true",
                    help="A/B: the convolutions derive the lo image of the weights on chip instead of loading the per-step one")
    ap.add_argument("--conv-sm-reserve", type=int, default=-1,
                    help="SMs the convolution grids leave free for the NCCL kernels of the overlapped gradient exchange "
                         "(-1 = the trainer's default: 0 on one GPU)")
    ap.add_argument("--roofline-detail", default="", help="write the per-launch dw-attention timings here")
    ap.add_argument("--dw-trace-out", default="", help="write the recorded dw-attention launch structure here")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-max-gb", type=float, default=8.0)
    ap.add_argument("--is-samples", type=int, default=1000)
    args = ap.parse_args()
    if args.fp32_convs:
        args.convs = "fp32"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "sweep4":
        return run_sweep4(args)
    if args.workload == "is_eval":
        return run_is_eval(args)
    return run_davi(args)


if __name__ == "__main__":
    sys.exit(main())
