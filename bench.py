#!/usr/bin/env python
"""Benchmark of BASELINE.json's metric: CIFAR-10 16-layer attentive VAE train
images/sec (batch 40 per GPU, synthetic data, random-init weights).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, restated
                                                            # (oracle/, torch CPU, all host threads)

One JSON line on stdout (rank 0).  A "step" is one full optimisation step of the
model (forward, backward, gradient all-reduce, fused Adamax) on one batch of 40
synthetic images per GPU.  `value` is measured with the batches already resident
in HBM; `e2e` repeats the measurement through the public API with pinned-host
inputs copied H2D every step and the step's loss read back D2H.
`roofline` reports the depth-wise attention kernels (the memory-bound hot kernel
north_star names): algorithmic bytes / CUDA-event time of the model's 31 forward
+ 31 backward calls at this batch size, launched back to back on rotating operand
sets larger than L2.  `roofline_nonlocal` is the tensor-bound view of the 128
nonlocal attention calls.  Precision: the hot-path kernels compute in fp32 (3xTF32
on the tensor cores); the host-side cuDNN convolutions run in 3xTF32 by default
(`--convs`), and the opt-in plain-TF32 step is measured into `alt_convs` for the
record (one GPU only).
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

def oracle_setup(batch, seed=0):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags
    from oracle.model import OracleVAE
    kw = parse_flags(**CIFAR10_16L)
    torch.manual_seed(seed)
    m = DeepVAE(**kw)                       # CPU construction only: supplies shapes + random-init weights
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = args.ref_batch
    m, o, params, x = oracle_setup(batch)
    state = {"t": 0, "m": {}, "u": {}}
    for _ in range(max(1, min(args.warmup, 2))):
        oracle_step(m, o, params, x, state)
    t0 = time.time()
    for _ in range(args.steps):
        oracle_step(m, o, params, x, state)
    dt = (time.time() - t0) / args.steps
    val = batch / dt
    sample = f"{args.steps} full train steps (fwd+bwd+Adamax) of the same model at batch {batch} (fp32, torch CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "bounded CPU sample: batch %d per step" % batch},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# roofline of the depth-wise attention kernels (the model's 62 calls per step)
# ---------------------------------------------------------------------------

def dw_roofline(batch, device):
    """Achieved bandwidth of the model's 31 forward + 31 backward depth-wise attention launches.

    Timing: for every call shape the kernel is launched back to back on ROTATING operand sets whose
    total footprint exceeds 3x the 126 MB L2 (so every launch reads cold inputs: "inputs larger than
    L2"), all launches of a shape between ONE pair of CUDA events on the launching stream; the
    per-launch time is the mean.  (Timing single launches between their own event pair adds the
    ~2 us event granularity to kernels that run 13-45 us.)"""
    import torch
    from deep_attentive_vi_b200 import _lib
    L, d, HW, P = 16, 20, 16, 256
    stream = torch.cuda.current_stream().cuda_stream
    calls = [("dec", i, 276) for i in range(1, L)] + [("enc", L + 1 - i, 256) for i in range(L)]
    footprint = 3 * 126e6
    tot_b = tot_t = 0.0
    per_call = []
    for kind, n, C in calls:
        fb = 4 * batch * P * (n * (d + C) + d + C)
        bb = 4 * batch * P * (2 * n * (d + C) + 2 * d + C)
        sets = max(2, min(24, int(footprint // fb) + 1))
        mk = lambda *shape: [torch.randn(*shape, device=device) for _ in range(sets)]
        q, K, V, go = mk(batch, P, d), mk(batch, n, P, d), mk(batch, n, P, C), mk(batch, P, C)
        out = [torch.empty(batch, P, C, device=device) for _ in range(sets)]
        dq = [torch.empty_like(t) for t in q]
        dK = [torch.empty(batch, n, P, d, device=device) for _ in range(min(sets, 4))]
        dV = [torch.empty(batch, n, P, C, device=device) for _ in range(min(sets, 4))]

        def fwd(s):
            _lib.call("davi_dw_attn_fwd", q[s].data_ptr(), K[s].data_ptr(), V[s].data_ptr(), out[s].data_ptr(), 0,
                      batch, n, P, d, C, d, n * P * d, P * d, d, n * P * C, P * C, C, C, 0, stream)

        def bwd(s):
            g = s % len(dK)
            _lib.call("davi_dw_attn_bwd", q[s].data_ptr(), K[s].data_ptr(), V[s].data_ptr(), go[s].data_ptr(),
                      dq[s].data_ptr(), dK[g].data_ptr(), dV[g].data_ptr(), 0, batch, n, P, d, C,
                      d, n * P * d, P * d, d, n * P * C, P * C, C, C, 0, stream)

        for name, fn, nbytes in (("fwd", fwd, fb), ("bwd", bwd, bb)):
            for s in range(sets):                       # warm-up round (also touches every set once)
                fn(s)
            ts = []
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s in range(sets):
                    fn(s)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) / sets)
            t = sorted(ts)[1]
            tot_b += nbytes
            tot_t += t
            per_call.append({"k": name, "n": n, "C": C, "sets": sets, "us": round(t * 1e3, 1),
                             "GBs": round(nbytes / t / 1e6)})
        del q, K, V, go, out, dq, dK, dV
    return tot_b, tot_t, per_call


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
    "3xtf32": "cuDNN on TF32 tensor cores with 3xTF32 error compensation (fp32-grade: operands split hi/lo by "
              "davi_split_tf32, three products accumulated in fp32 over a tripled contraction axis)",
    "fp32": "cuDNN fp32 (CUDA cores)",
    "tf32": "cuDNN plain TF32 (opt-in; operands truncated to 10 mantissa bits, below the reference's fp32)",
}


def run_davi(args):
    import torch
    import torch.distributed as dist
    from deep_attentive_vi_b200 import _lib, build
    from deep_attentive_vi_b200.model import CIFAR10_16L, DeepVAE, parse_flags
    from deep_attentive_vi_b200.util.trainer import Trainer

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
    from deep_attentive_vi_b200 import ops
    torch.backends.cudnn.benchmark = True

    B = args.batch
    g = torch.Generator().manual_seed(1000 + rank)
    nbuf = 4
    host = [torch.randint(0, 256, (B, 32, 32, 3), generator=g).float().pin_memory() for _ in range(nbuf)]
    dev_in = [h.to(device) for h in host]

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

    def measure(convs, sampler=None):
        """Whole train steps of the model with the host-side convolutions in precision `convs`
        (the hot-path kernels always compute in fp32): device-resident and end-to-end timings."""
        ops.set_conv_precision(convs)
        torch.manual_seed(0)
        model = DeepVAE(**parse_flags(**CIFAR10_16L)).to(device)
        trainer = Trainer(model, device=device, world_size=world, rank=rank, use_graph=not args.eager,
                          graph_warmup=min(3, max(1, args.warmup - 1)))
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
        del trainer, model
        torch.cuda.empty_cache()
        return out

    r = measure(args.convs, ClockSampler(local) if rank == 0 else None)
    ms, ms_e2e, launches, clocks, nparams = r["ms"], r["ms_e2e"], r["launches"], r["clocks"], r["params"]
    last = {"loss": r["loss"]}
    alt = None
    if args.alt_convs and args.alt_convs != args.convs and world == 1:   # second trainer + graph: one GPU only
        alt = measure(args.alt_convs)

    value = B * world * args.steps / (ms / 1e3)
    e2e_val = B * world * args.steps / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world, "params": nparams,
                   "parallelism": f"dp{world}", "cuda_graph": not args.eager, "convs": CONV_NOTE[args.convs],
                   "l2": "inputs (2 GB activations per step) exceed L2; dw-attention roofline launches rotate over "
                         "operand sets with a total footprint > 3x L2 (every launch reads cold inputs)"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * 32 * 32 * 3 * 4,
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
        if not args.no_roofline:
            peak, which = measured_peak()
            tb, tt, per_call = dw_roofline(B, device)
            ach = tb / tt / 1e6
            line["roofline"] = {"bound": "hbm", "kernel": "dw_attn_fwd/bwd (62 calls of one step)",
                                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "frac_of_nominal_8TBs": ach / 8000.0, "peak_source": which,
                                # ncu --set full, largest decoder call (n=15, C=276, B=40), DRAM read+write bytes
                                # per launch next to the algorithmic bytes (profiles/r01_ncu_dw_attn_v1_summary.txt)
                                "traffic": 187.6e6 + 319.1e6,
                                "traffic_detail": {"launches": "dw_attn fwd + bwd, n=15", "fwd_dram_MB": 187.6,
                                                   "bwd_dram_MB": 319.1, "fwd_algorithmic_MB": 194.0,
                                                   "bwd_algorithmic_MB": 376.7,
                                                   "note": "bwd writes partly still in L2 at kernel end"},
                                "algorithmic_MB_per_step": tb / 1e6, "kernel_ms_per_step": tt}
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
                        "(3xTF32 error compensation, fp32-grade), so the tensor pipe runs at executed_frac; "
                        "ncu: profiles/r01_ncu_nonlocal_tc_v4_summary.txt",
                "kernel_ms_per_step": tt_nl, "per_shape": per_shape}
            if args.roofline_detail:
                with open(args.roofline_detail, "w") as f:
                    json.dump(per_call, f)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cb = args.ref_batch
            m, o, params, x = oracle_setup(cb)
            state = {"t": 0, "m": {}, "u": {}}
            oracle_step(m, o, params, x, state)
            t0 = time.time()
            n = 2
            for _ in range(n):
                oracle_step(m, o, params, x, state)
            dt = (time.time() - t0) / n
            line["cpu_baseline"] = {"value": cb / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} full train steps of the same model at batch {cb} "
                                              "(oracle/, torch CPU fp32, all host threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: a captured CUDA graph holds NCCL kernels, and destroy_process_group() with such a
        # graph alive can block forever.  Synchronise, then leave without running the destructors.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="davi", choices=["davi", "reference"])
    ap.add_argument("--batch", type=int, default=40, help="images per GPU (BASELINE config: 40)")
    ap.add_argument("--ref-batch", type=int, default=4, help="images per step of the bounded CPU sample")
    ap.add_argument("--convs", default="3xtf32", choices=["3xtf32", "fp32", "tf32"],
                    help="precision of the host-side cuDNN convolutions (default: fp32-grade 3xTF32)")
    ap.add_argument("--alt-convs", default="tf32", choices=["", "3xtf32", "fp32", "tf32"],
                    help="second precision measured for the record into 'alt_convs' ('' to skip)")
    ap.add_argument("--fp32-convs", action="store_true", help="same as --convs fp32")
    ap.add_argument("--eager", action="store_true", help="do not capture the step into a CUDA graph")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--roofline-detail", default="", help="write the per-call dw-attention timings here")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.fp32_convs:
        args.convs = "fp32"
    if args.impl == "reference":
        return run_reference(args)
    return run_davi(args)


if __name__ == "__main__":
    sys.exit(main())
