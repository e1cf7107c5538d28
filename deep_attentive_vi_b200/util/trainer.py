"""Training step -- replaces the reference's util/trainer.py (Keras compile/fit
under tf.distribute.MirroredStrategy, run.py:94-95, tr.py:149-187).

Semantics kept from the reference:
  * loss = mean_B(nll) + beta * mean_B(sum kl) + sum(regularisers), the means
    taken over the GLOBAL batch (tr.py:149-151,174); the third output (nelbo)
    has weight 0;
  * synchronous data parallelism, one replica per GPU, per-replica batch-norm
    statistics, ONE gradient sum-all-reduce per step (here: NCCL over NVLink);
  * Adamax(lr, beta_1=0.9, beta_2=0.999, epsilon=1e-3) (tr.py:226-231);
  * KL warm-up beta: linear 0.1 -> 1 over 16 epochs (tr.py:298-316).
B200-native plumbing: all parameters, gradients and Adamax moments live in four
flat fp32 arrays (476 MB each for the CIFAR-10 model), so the all-reduce is one
NCCL call on one buffer and the optimizer + l2-regulariser gradient is one fused
kernel pass per regulariser class (davi_adamax_step).
"""
import copy
import math

import torch
import torch.distributed as dist

from .. import _lib
from ..util.hparams import HParams


def get_default_train_hparams():
    """tr.py:207-233.  lr_decay_epochs / lr_steps_per_epoch carry the reference's defaults (1000 / 1563);
    like keras_train (tr.py:139-140) the Trainer overwrites them with epochs - lr_warmup_epochs and
    train_size / (num_gpus * batch_size) when it is told the training-set size."""
    return HParams(num_gpus=1, batch_size=256, epochs=400, learning_rate=1e-2,
                   learning_rate_schedule="cosine_warmup", lr_decay_epochs=1000, lr_steps_per_epoch=1563,
                   lr_warmup_epochs=5, lr_min_learning_rate=1e-4,
                   lr_first_decay_epochs=100, lr_t_mul=2.0, lr_m_mul=1.0, optimizer="adamax",
                   keras_lr_beta_1=0.9, keras_lr_beta_2=0.999, keras_lr_epsilon=1e-3,
                   kl_warmup_epochs=16, kl_warmup_smooth_start=0.1)


def regulariser_coefficients(model):
    """l2 strength of every trainable parameter, read off the model's OWN regulariser (the modules'
    reg_loss() methods: one source of truth, per-layer lambda_reg and disabled regularisers included).
    reg_loss is a quadratic form sum_i c_i p_i^2, so its gradient at p = 1 is 2 c_i.  Returns
    {parameter: c}; a parameter the regulariser does not touch gets 0."""
    params = [p for p in model.parameters() if p.requires_grad]
    saved = [p.detach().clone() for p in params]
    try:
        with torch.no_grad():
            for p in params:
                p.fill_(1.0)
        reg = model.reg_loss()
        grads = (torch.autograd.grad(reg, params, allow_unused=True) if torch.is_tensor(reg) and reg.requires_grad
                 else [None] * len(params))
    finally:
        with torch.no_grad():
            for p, v in zip(params, saved):
                p.copy_(v)
    out = {}
    for p, g in zip(params, grads):
        if g is None:
            out[p] = 0.0
            continue
        c = 0.5 * g.reshape(-1)
        lo, hi = float(c.min()), float(c.max())
        if hi - lo > 1e-6 * max(abs(hi), 1e-30):
            raise ValueError("regulariser strength varies inside one parameter tensor; the fused optimizer "
                             "applies one coefficient per tensor")
        out[p] = float(c[0])
    return out


def kl_beta(epoch, warmup_epochs=16, smooth_start=0.1):
    """KLRegularizerSchedule (tr.py:298-316)."""
    if warmup_epochs <= 0:
        return 1.0
    return min(1.0, smooth_start + (1.0 - smooth_start) * epoch / warmup_epochs)


def cosine_warmup_lr(step, steps_per_epoch, lr, min_lr, warmup_epochs, decay_epochs):
    """CosineWarmup.__call__ (util/learning_rate_schedule.py:229-268) at the callback's 1-based
    global_step (LearningRateScheduler increments it BEFORE the call, :60-68): linear per-step warm-up
    min_lr + step / warmup_steps * (lr - min_lr) while step <= warmup_steps, then a per-epoch cosine of
    local_epoch / decay_epochs; past decay_epochs the reference keeps its last value, which is min_lr."""
    warm = warmup_epochs * steps_per_epoch
    if step <= warm:
        return min_lr + float(step / warm) * (lr - min_lr)
    local_epoch = step // steps_per_epoch - warmup_epochs
    if local_epoch > decay_epochs:
        return min_lr
    alpha = min_lr / lr
    cosine = 0.5 * (1.0 + math.cos(math.pi * (local_epoch / decay_epochs)))
    return lr * ((1.0 - alpha) * cosine + alpha)


class GradBuckets:
    """Plan + bookkeeping of the bucketed gradient all-reduce (~25 MB buckets launched as the backward
    produces them).  Pure host logic, no device code.

    `sources`: {key: [(lo, hi), ...]} -- the ranges of the flat gradient array a gradient source writes once its
    tensor has arrived (a plain parameter: its own range; the effective kernel of a weight-normed convolution: the
    ranges of its v and log_g).  Buckets are contiguous slices of the flat array, cut at multiples of `bucket_floats`
    but never inside the hull of one source, so a source feeds exactly ONE bucket.  `arrived(key)` returns the
    bucket that has just become complete (or None); `remaining()` lists the incomplete ones in order."""

    def __init__(self, sources, total, bucket_floats):
        hulls = sorted((min(lo for lo, _ in r), max(hi for _, hi in r), k) for k, r in sources.items())
        merged = []                                   # overlapping hulls form one atom
        for lo, hi, k in hulls:
            if merged and lo < merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
                merged[-1][2].append(k)
            else:
                merged.append([lo, hi, [k]])
        self.bounds, self.members = [], []            # [(lo, hi)], [[key, ...]]
        start, keys = 0, []
        for lo, hi, ks in merged:
            if keys and hi - start > bucket_floats and lo > start:
                self.bounds.append((start, lo)); self.members.append(keys)
                start, keys = lo, []
            keys = keys + ks
        self.bounds.append((start, total)); self.members.append(keys)
        self.bucket_of = {k: b for b, ks in enumerate(self.members) for k in ks}
        self.reset()

    def reset(self):
        self.pending = [len(ks) for ks in self.members]
        self.seen = set()
        self.done = [False] * len(self.members)

    def arrived(self, key):
        if key in self.seen or key not in self.bucket_of:
            return None
        self.seen.add(key)
        b = self.bucket_of[key]
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.done[b]:
            self.done[b] = True
            return b
        return None

    def remaining(self):
        out = [b for b, d in enumerate(self.done) if not d]
        for b in out:
            self.done[b] = True
        return out


class FlatState:
    """Re-homes every trainable parameter into a flat fp32 array (grouped by regulariser strength)
    with flat gradient / Adamax-moment arrays of the same layout.  Gradients are NOT accumulated
    into views of the flat array (that costs one tiny add_ kernel per parameter and step, ~2 600
    launches): autograd deposits its own tensors in .grad and `gather_grads` copies them into the
    flat array with davi_multi_copy (one launch per 64 tensors)."""

    def __init__(self, model, device):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        coef = regulariser_coefficients(model)
        groups = {}
        for n, p in named:
            groups.setdefault(round(coef[p], 12), []).append((n, p))
        total = sum(p.numel() for _, p in named)
        pad = lambda k: (k + 63) // 64 * 64   # every tensor starts 256-byte aligned (cuDNN, float4 loads)
        size = sum(sum(pad(p.numel()) for _, p in g) for g in groups.values())
        self.P = torch.zeros(size, device=device)
        self.G = torch.zeros(size, device=device)
        self.M = torch.zeros(size, device=device)
        self.U = torch.zeros(size, device=device)
        self.segments = []   # (l2, offset, length)
        self.total = total
        self.slots = []      # (tensor that owns a .grad, flat gradient array, offset, numel, is_conv)
        off = 0
        for l2, plist in sorted(groups.items()):
            start = off
            for _, p in plist:
                n = p.numel()
                src = p.detach()
                if p.dim() == 4:   # conv kernels are kept channels_last (OHWI in memory)
                    o, i, h, w = p.shape
                    self.P[off:off + n].view(o, h, w, i).copy_(src.permute(0, 2, 3, 1))
                    p.data = self.P[off:off + n].view(o, h, w, i).permute(0, 3, 1, 2)
                else:
                    self.P[off:off + n].view(p.shape).copy_(src)
                    p.data = self.P[off:off + n].view(p.shape)
                p.grad = None
                self.slots.append((p, "G", off, n, p.dim() == 4))
                off += pad(n)
            self.segments.append((float(l2), start, off - start))
        assert off == size
        self._build_weight_norm(model, device, pad)

    def _build_weight_norm(self, model, device, pad):
        """Flat effective-weight array W (+ its gradient dW) and the device row tables of
        davi_weight_norm_fwd/bwd: every weight-normed conv reads its kernel from a leaf view of W."""
        from ..layers.weight_norm import WNConv2d
        mods = [m for m in model.modules() if isinstance(m, WNConv2d) and m.use_weight_norm]
        self.wn_rows = 0
        if not mods:
            return
        base = self.P.data_ptr()
        rv, rg, rw, rc = [], [], [], []
        woff = 0
        sizes = []
        self.wn_index = []      # per module: (W offset, W numel, v offset, log_g offset, rows, first row in the tables)
        for m in mods:
            o, i, h, w = m.v.shape
            cols = i * h * w
            v_off = (m.v.data_ptr() - base) // 4
            g_off = (m.log_g.data_ptr() - base) // 4
            self.wn_index.append((woff, o * cols, v_off, g_off, o, len(rv)))
            for r in range(o):
                rv.append(v_off + r * cols); rg.append(g_off + r); rw.append(woff + r * cols); rc.append(cols)
            sizes.append((m, woff, o, i, h, w))
            woff += pad(o * cols)
        # W | Wt: effective kernels and their transposed + flipped images (input gradients, same offsets) in ONE
        # allocation, so that ONE davi_tf32_lo launch per step writes the lo halves of both (Wlo | Wtlo, same offsets
        # again): the convolutions load those by TMA instead of deriving them from every weight tile they stage
        self.WW = torch.zeros(2 * woff, device=device)
        self.W, self.Wt = self.WW[:woff], self.WW[woff:]
        self.LL = torch.zeros(2 * woff, device=device)
        self.Wlo, self.Wtlo = self.LL[:woff], self.LL[woff:]
        self.dW = torch.zeros(woff, device=device)
        wt_entries = []                                  # (offset, Cout, taps, Cin) per kernel as the convolutions see it
        where = {id(m): (off, o, i, h, w) for m, off, o, i, h, w in sizes}
        # the q | k | v projections of a nonlocal block are ONE 1x1 convolution: when their three
        # effective kernels are adjacent in W (no alignment gap), the block reads them as one leaf
        # view [2d+C, C, 1, 1] instead of concatenating three tensors every step
        from ..layers.attention import NonlocalResNetBlock
        fused = set()
        self.qkv_blocks = []
        for blk in model.modules():
            if not isinstance(blk, NonlocalResNetBlock) or blk._pad:
                continue
            trio = [blk.query_conv, blk.key_conv, blk.value_conv]
            if not all(id(c) in where for c in trio):
                continue
            (oq, nq, C, _, _), (ok_, nk, _, _, _), (ov, nv, _, _, _) = (where[id(c)] for c in trio)
            if ok_ != oq + nq * C or ov != ok_ + nk * C:
                continue
            rows = nq + nk + nv
            wv = self.W[oq:oq + rows * C].view(rows, 1, 1, C).permute(0, 3, 1, 2)
            wv.requires_grad_(True)
            blk._qkv_w_override = wv
            blk._qkv_wt_override = self.Wt[oq:oq + rows * C].view(C, 1, 1, rows)
            blk._qkv_lo_override = (self.Wlo[oq:oq + rows * C].view(rows, 1, 1, C), self.Wtlo[oq:oq + rows * C].view(C, 1, 1, rows))
            wt_entries.append((oq, rows, 1, C))
            self.slots.append((wv, "dW", oq, rows * C, True))
            self.qkv_blocks.append(blk)
            fused.update(id(c) for c in trio)
        for m, off, o, i, h, w in sizes:
            if id(m) in fused:
                continue
            n = o * i * h * w
            wv = self.W[off:off + n].view(o, h, w, i).permute(0, 3, 1, 2)       # OHWI memory = channels_last
            wv.requires_grad_(True)
            m._w_override = wv
            m._wt_override = self.Wt[off:off + n].view(i, h, w, o)
            m._lo_override = (self.Wlo[off:off + n].view(o, h, w, i), self.Wtlo[off:off + n].view(i, h, w, o))
            wt_entries.append((off, o, h * w, i))
            self.slots.append((wv, "dW", off, n, True))
        t64 = lambda v: torch.tensor(v, dtype=torch.int64, device=device)
        self.wn_tables = (t64(rv), t64(rg), t64(rw), torch.tensor(rc, dtype=torch.int32, device=device))
        self.wn_rows = len(rv)
        self.wn_modules = [s[0] for s in sizes]
        # device tables of davi_conv2d_wt_batched: one launch per step refreshes every transposed kernel
        prefix, tiles = [], 0
        for _, o, taps, i in wt_entries:
            prefix.append(tiles)
            tiles += taps * (-(-o // 32)) * (-(-i // 32))
        t32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device)
        self.wt_tables = (t64([e[0] for e in wt_entries]), t32([e[1] for e in wt_entries]), t32([e[2] for e in wt_entries]),
                          t32([e[3] for e in wt_entries]), t64(prefix + [tiles]), len(wt_entries), tiles)

    def clear_grads(self):
        for t, _, _, _, _ in self.slots:
            t.grad = None

    def gather_grads(self, stream, slots=None):
        """Copy every deposited gradient (of `slots`, default all) into its place in G (parameters) or dW
        (effective conv kernels).  Conv gradients are stored like their weights: OHWI (channels_last) memory order."""
        for which, flat in (("G", self.G), ("dW", getattr(self, "dW", None))):
            ptrs, offs, sizes, keep = [], [], [], []
            for t, w, off, n, is_conv in (self.slots if slots is None else slots):
                g = t.grad
                if w != which or g is None:
                    continue
                if is_conv:
                    if not g.permute(0, 2, 3, 1).is_contiguous():
                        g = g.contiguous(memory_format=torch.channels_last)
                elif not g.is_contiguous():
                    g = g.contiguous()
                keep.append(g)
                ptrs.append(g.data_ptr()); offs.append(off); sizes.append(n)
            if not ptrs:
                continue
            order = sorted(range(len(ptrs)), key=lambda k: -sizes[k])       # similar sizes share a launch
            pt, pa = _lib.ptr_table([ptrs[k] for k in order])
            ot, oa = _lib.int_table([offs[k] for k in order])
            st_, sa = _lib.int_table([sizes[k] for k in order])
            _lib.call("davi_multi_copy", pt, ot, st_, len(order), flat.data_ptr(), stream)
            _lib.add_launches((len(order) + 63) // 64 - 1)

    def weight_norm_bwd(self, stream, row_lo=0, row_hi=None):
        """dW -> gradients of v and log_g for the table rows [row_lo, row_hi) (default: all)."""
        rv, rg, rw, rc = self.wn_tables
        row_hi = self.wn_rows if row_hi is None else row_hi
        if row_hi > row_lo:
            _lib.call("davi_weight_norm_bwd", self.P.data_ptr(), self.dW.data_ptr(), self.G.data_ptr(),
                      rv[row_lo:].data_ptr(), rg[row_lo:].data_ptr(), rw[row_lo:].data_ptr(), rc[row_lo:].data_ptr(),
                      row_hi - row_lo, stream)

    def plan_buckets(self, bucket_floats):
        """-> (GradBuckets, per bucket: slots to gather + weight-norm row range).  Sources are the tensors autograd
        deposits gradients in: plain parameters (but not the v / log_g of a convolution whose kernel is read from
        W: their gradient comes out of davi_weight_norm_bwd) and the leaf views of W."""
        wn_param_ranges = set()
        for woff, wn, v_off, g_off, o, row0 in getattr(self, "wn_index", []):
            wn_param_ranges.add(v_off); wn_param_ranges.add(g_off)
        sources, slot_of = {}, {}
        for slot in self.slots:
            t, which, off, n, _ = slot
            if which == "G":
                if off in wn_param_ranges:
                    continue
                sources[id(t)] = [(off, off + n)]
            else:
                ranges = []
                for woff, wn, v_off, g_off, o, row0 in self.wn_index:
                    if off <= woff < off + n:
                        ranges += [(v_off, v_off + wn), (g_off, g_off + o)]
                sources[id(t)] = ranges
            slot_of[id(t)] = slot
        plan = GradBuckets(sources, self.G.numel(), bucket_floats)
        per = []
        for (lo, hi), keys in zip(plan.bounds, plan.members):
            rows = [(row0, row0 + o) for woff, wn, v_off, g_off, o, row0 in getattr(self, "wn_index", [])
                    if lo <= v_off < hi]
            assert all(lo <= g_off < hi for woff, wn, v_off, g_off, o, row0 in getattr(self, "wn_index", [])
                       if lo <= v_off < hi), "a convolution's v and log_g must share a bucket"
            per.append({"slots": [slot_of[k] for k in keys], "rows": rows, "range": (lo, hi)})
        return plan, per

    def release(self):
        """Detach the model from the fused weight-norm path (it computes its kernels itself again)."""
        for m in getattr(self, "wn_modules", []):
            m._w_override = None
            m._wt_override = None
            m._lo_override = None
        for blk in getattr(self, "qkv_blocks", []):
            blk._qkv_w_override = None
            blk._qkv_wt_override = None
            blk._qkv_lo_override = None
        self.wn_rows = 0


class Trainer:
    """use_graph=True captures the whole optimisation step (forward, hand-written
    backward kernels, NCCL all-reduce, fused Adamax) into ONE CUDA graph after
    `graph_warmup` eager steps and replays it afterwards: the ~10^4 launches of a
    step then cost one host call.  Everything that changes between steps lives in
    device memory (input batch, KL weight beta, Adamax step size)."""

    def __init__(self, model, hparams=None, device="cuda", world_size=1, rank=0, train_size=None, batch_size=None,
                 use_graph=False, graph_warmup=3, use_schedule=True, bucket_mb=None):
        """train_size / batch_size (per GPU): when given, the schedule lengths are derived like keras_train does
        (tr.py:139-140): lr_steps_per_epoch = int(train_size / (num_gpus * batch_size)), lr_decay_epochs =
        epochs - lr_warmup_epochs; otherwise the hparams' own values are used."""
        self.model = model
        self.use_schedule = use_schedule
        self.hp = copy.deepcopy(hparams) if hparams is not None else get_default_train_hparams()
        self.device = torch.device(device)
        self.world, self.rank = world_size, rank
        self.state = FlatState(model, self.device)
        self.step_count = 0
        if train_size is not None:
            bs = batch_size if batch_size is not None else self.hp.batch_size
            self.hp["lr_steps_per_epoch"] = int(train_size / (world_size * bs))
            self.hp["lr_decay_epochs"] = self.hp.epochs - self.hp.lr_warmup_epochs
        self.steps_per_epoch = self.hp.lr_steps_per_epoch
        self.beta = 1.0
        self.lr = self.hp.learning_rate
        if use_schedule:      # epoch 0 of the reference run: lr = lr_min (warm-up), beta = 0.1 (KL warm-up)
            self.beta = kl_beta(0, self.hp.kl_warmup_epochs, self.hp.kl_warmup_smooth_start)
        for b in model.buffers():
            assert b.device.type == "cuda", "move the model to the GPU before building the Trainer"
        # device-resident step state (read by the kernels, so a captured graph sees fresh values)
        self.beta_dev = torch.full((), float(self.beta), device=self.device)
        self.hyper_dev = torch.tensor([0.0, 1.0], device=self.device)   # [lr / (1 - b1^t), grad_scale]
        self.lr_dev = torch.full((1,), float(self.lr), device=self.device)
        self.step_dev = torch.zeros(1, device=self.device)              # t, advanced on the device
        self._lr_pushed = float(self.lr)
        self.use_graph, self.graph_warmup = use_graph, graph_warmup
        self._graph = None
        self._static = None
        self.graph_launches = 0
        # capture needs every autograd node (AccumulateGrad included) to have been created on a
        # non-default stream: the eager warm-up steps of a graphed trainer run on this side stream
        self._side = torch.cuda.Stream(device=self.device) if use_graph else None
        # bucketed gradient all-reduce overlapped with the backward (world > 1; bucket_mb = 0 keeps the single
        # all-reduce after the backward, a value forces the bucketed path even on one rank -- tests)
        self.bucket_mb = (32 if world_size > 1 else 0) if bucket_mb is None else bucket_mb
        self._buckets = None
        if self.bucket_mb:
            self._setup_buckets()

    def _setup_buckets(self):
        st = self.state
        self._plan, self._bucket_work = st.plan_buckets(int(self.bucket_mb * 1e6 / 4))
        for b in self._bucket_work:                     # contiguous runs of weight-norm table rows
            runs = []
            for lo, hi in sorted(b["rows"]):
                if runs and runs[-1][1] == lo:
                    runs[-1][1] = hi
                else:
                    runs.append([lo, hi])
            b["rows"] = runs
        self._buckets = True
        self._bucketing = False
        self._works = []
        self.bucket_log = []                            # (bucket, sources still outstanding in OTHER buckets) per flush
        for t, _, _, _, _ in st.slots:
            if id(t) in self._plan.bucket_of:
                t.register_post_accumulate_grad_hook(self._grad_arrived)

    def _grad_arrived(self, t):
        if not self._bucketing:
            return
        b = self._plan.arrived(id(t))
        if b is not None:
            self._flush_bucket(b)

    def _flush_bucket(self, b):
        """Everything bucket b needs has arrived: gather its gradients into the flat array, finish the weight-norm
        chain rule for its convolutions and start its all-reduce while the backward goes on."""
        st, work = self.state, self._bucket_work[b]
        stream = torch.cuda.current_stream().cuda_stream
        st.gather_grads(stream, work["slots"])
        for lo, hi in work["rows"]:
            st.weight_norm_bwd(stream, lo, hi)
        self.bucket_log.append((b, sum(self._plan.pending)))
        if self.world > 1:
            lo, hi = work["range"]
            self._works.append(dist.all_reduce(st.G[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def set_epoch(self, epoch):
        self.beta = kl_beta(epoch, self.hp.kl_warmup_epochs, self.hp.kl_warmup_smooth_start)
        self.beta_dev.fill_(self.beta)

    def forward_backward(self, x, eps=None, noise=None):
        st = self.state
        st.G.zero_()                                                 # parameters that get no gradient this step
        if st.wn_rows:
            st.dW.zero_()                                            # ... and conv kernels likewise (no stale dW)
        st.clear_grads()
        stream = torch.cuda.current_stream().cuda_stream
        if st.wn_rows:                                               # all conv kernels: v, log_g -> W
            rv, rg, rw, rc = st.wn_tables
            _lib.call("davi_weight_norm_fwd", st.P.data_ptr(), st.W.data_ptr(), rv.data_ptr(), rg.data_ptr(),
                      rw.data_ptr(), rc.data_ptr(), st.wn_rows, stream)
            to, tc, tt, ti, tp, cnt, tiles = st.wt_tables                     # W -> Wt for every input gradient of the step
            _lib.call("davi_conv2d_wt_batched", st.W.data_ptr(), st.Wt.data_ptr(), to.data_ptr(), tc.data_ptr(),
                      tt.data_ptr(), ti.data_ptr(), tp.data_ptr(), cnt, tiles, stream)
            _lib.call("davi_tf32_lo", st.WW.data_ptr(), st.LL.data_ptr(), st.WW.numel(), stream)   # W | Wt -> Wlo | Wtlo
        nll, kl, nelbo = self.model(x, True, eps, noise)
        global_batch = x.shape[0] * self.world                       # tr.py:174
        loss = (nll.sum() + self.beta_dev * kl.sum()) * (1.0 / global_batch)
        if self._buckets:
            self._plan.reset()
            self._works, self.bucket_log = [], []
            self._bucketing = True
            try:
                loss.backward()                                      # buckets flush from the gradient hooks
            finally:
                self._bucketing = False
            for b in self._plan.remaining():                         # sources that received no gradient this step
                self._flush_bucket(b)
            return nll, kl
        loss.backward()
        st.gather_grads(stream)                                      # per-tensor gradients -> flat G / dW
        if st.wn_rows:                                               # dW -> grads of v and log_g
            st.weight_norm_bwd(stream)
        return nll, kl

    def all_reduce(self):
        """The step's only exchange: sum of all gradients (MirroredStrategy's implicit all-reduce).  Bucketed
        path: the per-bucket NCCL calls were started during the backward, here they are joined; otherwise one
        NCCL call on the flat gradient array."""
        if self._buckets:
            for w in self._works:
                w.wait()
            self._works = []
        elif self.world > 1:
            dist.all_reduce(self.state.G, op=dist.ReduceOp.SUM)

    def _advance_hyper(self):
        """Host side of the optimizer step: count it, evaluate the learning-rate schedule
        (per step, like the reference's Keras callback, util/learning_rate_schedule.py:55-79)
        and push a changed value (fill_ carries it in the launch arguments: no pinned-buffer race)."""
        hp = self.hp
        self.step_count += 1            # the reference's callback counts the step before asking the schedule
        if self.use_schedule and hp.learning_rate_schedule == "cosine_warmup":
            self.lr = cosine_warmup_lr(self.step_count, self.steps_per_epoch, hp.learning_rate,
                                       hp.lr_min_learning_rate, hp.lr_warmup_epochs, hp.lr_decay_epochs)
        if float(self.lr) != self._lr_pushed:
            self._lr_pushed = float(self.lr)
            self.lr_dev.fill_(self._lr_pushed)

    def apply(self):
        st, hp = self.state, self.hp
        # Keras Adamax bias correction lr / (1 - beta_1^t), t kept on the device
        self.step_dev.add_(1.0)
        corr = 1.0 - torch.pow(torch.full_like(self.step_dev, float(hp.keras_lr_beta_1)), self.step_dev)
        self.hyper_dev[0:1].copy_(self.lr_dev / corr)
        stream = torch.cuda.current_stream().cuda_stream
        for l2, off, n in st.segments:
            _lib.call("davi_adamax_step_dev", st.P[off:].data_ptr(), st.G[off:].data_ptr(),
                      st.M[off:].data_ptr(), st.U[off:].data_ptr(), n, self.hyper_dev.data_ptr(),
                      float(hp.keras_lr_beta_1), float(hp.keras_lr_beta_2), float(hp.keras_lr_epsilon),
                      l2, stream)

    def _step_body(self, x, eps=None, noise=None):
        nll, kl = self.forward_backward(x, eps, noise)
        self.all_reduce()
        self.apply()
        return nll, kl

    def _capture(self, x):
        self._static = {"x": torch.empty_like(x)}
        self._static["x"].copy_(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(g, stream=self._side):
            nll, kl = self._step_body(self._static["x"])
        self.graph_launches = _lib.launch_count() - before
        _lib.add_launches(-self.graph_launches)       # capture enqueued nothing; replays are counted below
        self._static["nll"], self._static["kl"] = nll, kl
        self._graph = g

    def train_step(self, x, eps=None, noise=None):
        """One optimisation step on this rank's shard x [B,H,W,C].  Returns the
        per-image nll and kl tensors of the shard (device, no host sync)."""
        self._advance_hyper()
        graph_ok = self.use_graph and eps is None and noise is None
        if not graph_ok:
            return self._step_body(x, eps, noise)
        if self.step_count <= self.graph_warmup:
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                out = self._step_body(x, eps, noise)
            cur.wait_stream(self._side)
            return out
        if self._graph is None:
            self._capture(x)
        self._static["x"].copy_(x, non_blocking=True)
        self._graph.replay()
        _lib.add_launches(self.graph_launches)
        # fresh tensors, like the eager path: the next replay overwrites the capture's static outputs
        return self._static["nll"].clone(), self._static["kl"].clone()
