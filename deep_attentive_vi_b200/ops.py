"""torch.autograd wrappers over the C ABI (include/davi.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape;
every forward and every gradient below is one or two launches of the hand-
written sm_100a kernels in ``csrc/`` through ``libdavi.so``.  There is no
fallback: CPU tensors, a missing library or a failing call raise.
"""
import os

import torch

from . import _lib

_lib_call = _lib.call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.DaviError(
                "davi ops need float32 CUDA tensors (no CPU fallback); got "
                f"{t.device}/{t.dtype}")


def _p(t):
    return 0 if t is None else t.data_ptr()


def _rows(t):
    """View `t` [..., w] as rows of w floats with ONE uniform row stride.
    Returns (tensor_kept_alive, row_stride).  Copies only if the view cannot be
    expressed that way or breaks the 16-byte rules of the ABI."""
    ok = t.stride(-1) == 1 and t.data_ptr() % 16 == 0
    if ok and t.dim() >= 2:
        rs = t.stride(-2)
        ok = rs % 4 == 0 and rs >= t.shape[-1]
        for k in range(t.dim() - 2, 0, -1):
            if t.shape[k - 1] != 1 and t.stride(k - 1) != t.stride(k) * t.shape[k]:
                ok = False
    if not ok:
        t = t.contiguous()
    return t, (t.stride(-2) if t.dim() >= 2 else t.shape[-1])


def _scalar_ptr(s, like):
    """Device pointer of a scalar parameter (None -> NULL = 1.0)."""
    if s is None:
        return None, 0
    if not torch.is_tensor(s):
        s = torch.tensor(float(s), device=like.device, dtype=torch.float32)
    s = s.reshape(()).contiguous()
    _require_cuda(s)
    return s, s.data_ptr()


# ---------------------------------------------------------------------------
# R1: depth-wise attention, stacked tensors (drop-in for DepthWiseAttention.apply,
# reference layers/attention.py:135-164)
# ---------------------------------------------------------------------------

class _DwAttnStacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, keys, values, scale):
        _require_cuda(query, keys, values, scale)
        B, n, H, W, d = keys.shape
        C = values.shape[-1]
        q = query.contiguous()
        k = keys.contiguous()
        v = values.contiguous()
        P = H * W
        out = torch.empty((B, H, W, C), device=q.device, dtype=torch.float32)
        s_t, s_ptr = _scalar_ptr(scale, q)
        _lib_call("davi_dw_attn_fwd", _p(q), _p(k), _p(v), _p(out), 0, B, n, P, d, C,
                  d, n * P * d, P * d, d, n * P * C, P * C, C, C, s_ptr, _stream())
        ctx.save_for_backward(q, k, v, s_t)
        ctx.dims = (B, n, P, d, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, s_t = ctx.saved_tensors
        B, n, P, d, C = ctx.dims
        dout = dout.contiguous()
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        dv = torch.empty_like(v)
        need_ds = s_t is not None and ctx.needs_input_grad[3]
        dscale = torch.empty((), device=q.device, dtype=torch.float32) if need_ds else None
        _lib_call("davi_dw_attn_bwd", _p(q), _p(k), _p(v), _p(dout), _p(dq), _p(dk), _p(dv),
                  _p(dscale), B, n, P, d, C, d, n * P * d, P * d, d, n * P * C, P * C, C, C,
                  _p(s_t), _stream())
        return dq, dk, dv, dscale


def dw_attn(query, keys, values, scale=None):
    """query [B,H,W,d], keys [B,n,H,W,d], values [B,n,H,W,C] -> [B,H,W,C]."""
    return _DwAttnStacked.apply(query, keys, values, scale)


def dw_attn_weights(query, keys, scale=None):
    """The softmax weights [B,H,W,n] (inspection / tests only, no gradient)."""
    _require_cuda(query, keys)
    B, n, H, W, d = keys.shape
    q, k = query.contiguous(), keys.contiguous()
    P = H * W
    dummy_v = torch.zeros((B, n, H, W, 4), device=q.device, dtype=torch.float32)
    out = torch.empty((B, H, W, 4), device=q.device, dtype=torch.float32)
    attn = torch.empty((B, H, W, n), device=q.device, dtype=torch.float32)
    s_t, s_ptr = _scalar_ptr(scale, q)
    _lib_call("davi_dw_attn_fwd", _p(q), _p(k), _p(dummy_v), _p(out), _p(attn), B, n, P, d, 4,
              d, n * P * d, P * d, d, n * P * 4, P * 4, 4, 4, s_ptr, _stream())
    return attn


# ---------------------------------------------------------------------------
# R1/R2/R3: depth-wise attention over a list of slots (no stack / split copies)
# ---------------------------------------------------------------------------

class SharedSlot:
    """A per-layer tensor [B,H,W,nv+d] (values || key along the channel axis) that several LATER
    layers attend to.  Its gradient is GATHERED lazily: the backward kernel of every
    consumer only saves its softmax weights a and scale*ds ([B,P,n], davi_dw_attn_slots_bwd attn_out /
    ds_out) and registers itself here; when the backward of the EARLIEST consumer in forward order runs
    (the last one to run; it takes the tensor as a real autograd input, "own"), ONE gather kernel
    (davi_dw_attn_gather) writes the gradient once as sum_i a^(i) dOut^(i) || sum_i scale ds^(i) q^(i).
    Neither the zero-filled slice gradients + add_n of TF autodiff nor a read-modify-write per consumer
    exist.  All later layers use the tensor detached ("acc")."""

    def __init__(self, tensor, nv):
        self.tensor = tensor if tensor.is_contiguous() else tensor.contiguous()
        self.nv = nv
        self.consumers = []     # (dout, q, q_sp, a, ds, slot index, n) of every backward that attended to it
        self.grad = None
        self.closed = False


def _slot_tables(spec, tensors, d, C):
    """Pointer / stride tables of the forward operands.  Returns (k_ptrs, v_ptrs, k_sps, v_sps, keep)."""
    kp, vp, ks, vs, keep = [], [], [], [], []
    it = iter(tensors)
    for item in spec:
        if item[0] == "kv":
            k, k_s = _rows(next(it))
            v, v_s = _rows(next(it))
            keep += [k, v]
            kp.append(k.data_ptr()); vp.append(v.data_ptr()); ks.append(k_s); vs.append(v_s)
        else:
            sh = item[1]
            if item[0] == "own":
                next(it)                                   # the autograd-visible alias of sh.tensor
            t = sh.tensor
            ct = t.shape[-1]
            assert ct == sh.nv + d and sh.nv == C, "shared slot must be values || key"
            keep.append(t)
            kp.append(t.data_ptr() + 4 * sh.nv); vp.append(t.data_ptr()); ks.append(ct); vs.append(ct)
    return kp, vp, ks, vs, keep


_GATHER_MAX_PROBLEMS, _GATHER_MAX_SLOTS = 32, 320      # kGatherMaxProblems / kGatherMaxSlots of dw_attn.cu

# Measurement hook (bench.py): when a list, every depth-wise attention launch appends the STRUCTURE of its
# call (shapes, per-slot strides, which gradients are written / skipped / gathered) so that exactly the
# launches of one training step can be replayed on synthetic operands (tools/dw_replay.py).
dw_trace = None


def _trace_slots(kp, vp, ks, vs):
    out = []
    for k, v, k_s, v_s in zip(kp, vp, ks, vs):
        koff = (k - v) // 4
        out.append({"k_sp": int(k_s), "v_sp": int(v_s), "koff": int(koff) if 0 <= koff < v_s else None})
    return out


def _gather_shared(owned, B, P, d, C):
    """davi_dw_attn_gather over the shared contexts whose last consumer just ran: one launch for as many
    contexts as fit its tables (the bottom-up tensors e[1..L] all finish in layer 0's backward)."""
    import ctypes
    batch, tot = [], 0
    batches = [batch]
    for sh in owned:
        m = len(sh.consumers)
        assert m <= _GATHER_MAX_SLOTS
        if len(batch) == _GATHER_MAX_PROBLEMS or tot + m > _GATHER_MAX_SLOTS:
            batch, tot = [], 0
            batches.append(batch)
        batch.append(sh); tot += m
    for batch in batches:
        ms, ov, ok, ovs, oks, vp, vs, kp, ks, wa, wd, ws = ([] for _ in range(12))
        for sh in batch:
            ct = sh.tensor.shape[-1]
            ms.append(len(sh.consumers))
            ov.append(sh.grad.data_ptr()); ok.append(sh.grad.data_ptr() + 4 * sh.nv); ovs.append(ct); oks.append(ct)
            for dout, q, q_sp, a, ds, j, n in sh.consumers:
                vp.append(dout.data_ptr()); vs.append(C); kp.append(q.data_ptr()); ks.append(q_sp)
                wa.append(a.data_ptr() + 4 * j); wd.append(ds.data_ptr() + 4 * j); ws.append(n)
        mt = (ctypes.c_int * len(ms))(*ms)
        tabs = [_lib.ptr_table(ov), _lib.ptr_table(ok), _lib.int_table(ovs), _lib.int_table(oks),
                _lib.ptr_table(vp), _lib.int_table(vs), _lib.ptr_table(kp), _lib.int_table(ks),
                _lib.ptr_table(wa), _lib.ptr_table(wd), _lib.int_table(ws)]
        _lib_call("davi_dw_attn_gather", len(ms), mt, *[t[0] for t in tabs], B, P, d, C, _stream())
        if dw_trace is not None:
            dw_trace.append({"op": "gather", "B": B, "P": P, "d": d, "C": C,
                             "problems": [{"ct": int(sh.tensor.shape[-1]), "nv": int(sh.nv),
                                           "consumers": [{"n": int(c[6]), "j": int(c[5]), "q_sp": int(c[2])}
                                                         for c in sh.consumers]} for sh in batch]})


class _DwAttnSlots(torch.autograd.Function):
    """inputs: query, scale, spec, then the autograd-visible tensors named by spec:
    ("kv",) -> key, value;  ("own", shared) -> shared.tensor;  ("acc", shared) -> nothing."""

    @staticmethod
    def forward(ctx, query, scale, spec, *tensors):
        _require_cuda(query, scale, *tensors)
        n = len(spec)
        B, H, W = query.shape[0], query.shape[1], query.shape[2]
        d = query.shape[-1]
        C = tensors[1].shape[-1] if spec[0][0] == "kv" else spec[0][1].nv
        P = H * W
        q, q_sp = _rows(query)
        kp, vp, ks, vs, keep = _slot_tables(spec, tensors, d, C)
        out = torch.empty((B, H, W, C), device=q.device, dtype=torch.float32)
        s_t, s_ptr = _scalar_ptr(scale, q)
        kt, ka = _lib.ptr_table(kp)
        vt, va = _lib.ptr_table(vp)
        kst, ksa = _lib.int_table(ks)
        vst, vsa = _lib.int_table(vs)
        _lib_call("davi_dw_attn_slots_fwd", _p(q), kt, vt, kst, vst, _p(out), 0, B, n, P, d, C,
                  q_sp, C, s_ptr, _stream())
        if dw_trace is not None:
            dw_trace.append({"op": "fwd", "B": B, "n": n, "P": P, "d": d, "C": C, "q_sp": int(q_sp),
                             "slots": _trace_slots(kp, vp, ks, vs)})
        ctx.save_for_backward(q, s_t, *keep)
        ctx.spec = spec
        ctx.dims = (B, n, P, d, C, q_sp, kp, vp, ks, vs)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, n, P, d, C, q_sp, kp, vp, ks, vs = ctx.dims
        spec = ctx.spec
        saved = ctx.saved_tensors
        q, s_t = saved[0], saved[1]
        dout = dout.contiguous()
        dev = q.device
        dq = torch.empty(q.shape, device=dev, dtype=torch.float32)   # dense rows
        dq_view, dq_sp = dq, d
        if q_sp != d:   # the query was a channel-slice view: give the kernel a dense copy of it
            q = q.contiguous(); q_sp = d
        need_ds = s_t is not None and ctx.needs_input_grad[1]
        dscale = torch.empty((), device=dev, dtype=torch.float32) if need_ds else None
        dkp, dvp, dks, dvs, grads, owned = [], [], [], [], [], []
        lead = dout.shape[:-1]
        a_out = ds_out = None
        if any(item[0] != "kv" for item in spec):
            a_out = torch.empty((B * P, n), device=dev, dtype=torch.float32)
            ds_out = torch.empty((B * P, n), device=dev, dtype=torch.float32)
        for j, item in enumerate(spec):
            if item[0] == "kv":
                dk = torch.empty((*lead, d), device=dev, dtype=torch.float32)
                dv = torch.empty((*lead, C), device=dev, dtype=torch.float32)
                grads += [dk, dv]
                dkp.append(dk.data_ptr()); dvp.append(dv.data_ptr()); dks.append(d); dvs.append(C)
                continue
            sh = item[1]
            assert not sh.closed, "shared context consumed after its owner's backward ran"
            ct = sh.tensor.shape[-1]
            if item[0] == "own":
                sh.grad = torch.empty_like(sh.tensor)                # every row is written once: no memset
                sh.closed = True
                grads.append(sh.grad)
                if not sh.consumers:                                 # sole consumer: this kernel writes it directly
                    dkp.append(sh.grad.data_ptr() + 4 * sh.nv); dvp.append(sh.grad.data_ptr())
                    dks.append(ct); dvs.append(ct)
                    continue
                owned.append(sh)
            sh.consumers.append((dout, q, q_sp, a_out, ds_out, j, n))
            dkp.append(0); dvp.append(0); dks.append(ct); dvs.append(ct)   # NULL: gathered later
        kt, ka = _lib.ptr_table(kp)
        vt, va = _lib.ptr_table(vp)
        kst, ksa = _lib.int_table(ks)
        vst, vsa = _lib.int_table(vs)
        dkt, dka = _lib.ptr_table(dkp)
        dvt, dva = _lib.ptr_table(dvp)
        dkst, dksa = _lib.int_table(dks)
        dvst, dvsa = _lib.int_table(dvs)
        _lib_call("davi_dw_attn_slots_bwd", _p(q), kt, vt, kst, vst, _p(dout), _p(dq), dkt, dvt, dkst, dvst,
                  _p(dscale), B, n, P, d, C, q_sp, C, _p(s_t), 0, _p(a_out), _p(ds_out), _stream())
        if dw_trace is not None:
            dw_trace.append({"op": "bwd", "B": B, "n": n, "P": P, "d": d, "C": C, "q_sp": int(q_sp),
                             "slots": _trace_slots(kp, vp, ks, vs), "save": a_out is not None,
                             "grads": [None if dv == 0 else {"k_sp": int(a), "v_sp": int(b),
                                                             "koff": int((dk - dv) // 4) if 0 <= (dk - dv) // 4 < b else None}
                                       for dk, dv, a, b in zip(dkp, dvp, dks, dvs)]})
        if owned:
            _gather_shared(owned, B, P, d, C)
            for sh in owned:
                sh.consumers = []                                    # release the consumers' tensors
        return (dq_view, dscale, None, *grads)


def dw_attn_ctx(query, slots, scale=None):
    """Depth-wise attention over a list of slots, each one of
        ("kv", key [B,H,W,d], value [B,H,W,C])   tensors of this layer (views allowed, no copies)
        ("own", SharedSlot) / ("acc", SharedSlot) a context shared with other layers (see SharedSlot)."""
    spec, tensors = [], []
    for item in slots:
        if item[0] == "kv":
            spec.append(("kv",)); tensors += [item[1], item[2]]
        elif item[0] == "own":
            spec.append(("own", item[1])); tensors.append(item[1].tensor)
        else:
            spec.append(("acc", item[1]))
    return _DwAttnSlots.apply(query, scale, tuple(spec), *tensors)


def dw_attn_slots(query, keys, values, scale=None):
    """query [B,H,W,d]; keys/values: lists of n tensors [B,H,W,d] / [B,H,W,C]
    (channel-slice views of wider tensors are passed without copying)."""
    assert len(keys) == len(values) and len(keys) >= 1
    return dw_attn_ctx(query, [("kv", k, v) for k, v in zip(keys, values)], scale)


# ---------------------------------------------------------------------------
# R2/R3 glue: LayerNorm (+ gelu residual)
# ---------------------------------------------------------------------------

class _LnGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, mode):
        _require_cuda(x, gamma, beta)
        C = x.shape[-1]
        xr, x_s = _rows(x)
        rows = xr.numel() // C
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        g, b = gamma.contiguous(), beta.contiguous()
        _lib_call("davi_ln_gelu_fwd", _p(xr), _p(g), _p(b), _p(y), rows, C, x_s, C, eps, mode,
                  _stream())
        ctx.save_for_backward(xr, g, b)
        ctx.cfg = (rows, C, x_s, eps, mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        xr, g, b = ctx.saved_tensors
        rows, C, x_s, eps, mode = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty(dy.shape, device=dy.device, dtype=torch.float32)
        dg = torch.empty_like(g)
        db = torch.empty_like(b)
        wsb = _lib.load().davi_ln_gelu_bwd_workspace_bytes(rows, C)
        ws = torch.empty(wsb // 4, device=dy.device, dtype=torch.float32)
        _lib_call("davi_ln_gelu_bwd", _p(xr), _p(g), _p(b), _p(dy), _p(dx), _p(dg), _p(db),
                  rows, C, x_s, C, C, eps, mode, _p(ws), wsb, _stream())
        return dx, dg, db, None, None


class _LnGeluRes(torch.autograd.Function):
    """y = base + res_gamma * (x + gelu(LN(x)))  (vl.py:408-412, 501-505) in one kernel each way."""

    @staticmethod
    def forward(ctx, x, gamma, beta, base, res_gamma, eps):
        _require_cuda(x, gamma, beta, base, res_gamma)
        C = x.shape[-1]
        xr, x_s = _rows(x)
        br, b_s = _rows(base)
        rows = xr.numel() // C
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        g, b = gamma.contiguous(), beta.contiguous()
        rg, rg_ptr = _scalar_ptr(res_gamma, x)
        _lib_call("davi_ln_gelu_res_fwd", _p(xr), _p(g), _p(b), _p(br), rg_ptr, _p(y), rows, C, x_s, b_s, C, eps,
                  _stream())
        ctx.save_for_backward(xr, g, b, rg)
        ctx.cfg = (rows, C, x_s, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        xr, g, b, rg = ctx.saved_tensors
        rows, C, x_s, eps = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty(dy.shape, device=dy.device, dtype=torch.float32)
        dg, db = torch.empty_like(g), torch.empty_like(b)
        drg = torch.empty((), device=dy.device, dtype=torch.float32)
        wsb = _lib.load().davi_ln_gelu_bwd_workspace_bytes(rows, C)
        ws = torch.empty(wsb // 4, device=dy.device, dtype=torch.float32)
        _lib_call("davi_ln_gelu_res_bwd", _p(xr), _p(g), _p(b), _p(rg), _p(dy), _p(dx), _p(dg), _p(db), _p(drg),
                  rows, C, x_s, C, C, eps, _p(ws), wsb, _stream())
        return dx, dg, db, dy, drg, None


def ln_gelu_res(x, gamma, beta, base, res_gamma, eps=1e-5):
    """base + res_gamma * (x + gelu(LN(x))): the post-attention glue of both attentions
    (reference vl.py:408-412 and 501-505) as ONE kernel forward and one backward."""
    return _LnGeluRes.apply(x, gamma, beta, base, res_gamma, eps)


def layer_norm(x, gamma, beta, eps=1e-5):
    """LayerNormalization over the last axis (reference vl.py:399, vae.py:347-348)."""
    return _LnGelu.apply(x, gamma, beta, eps, 0)


def ln_gelu_residual(x, gamma, beta, eps=1e-5):
    """x + gelu(LN(x)) (reference vl.py:394,409,502,503)."""
    return _LnGelu.apply(x, gamma, beta, eps, 1)


# ---------------------------------------------------------------------------
# R4: nonlocal attention core
# ---------------------------------------------------------------------------

class _NonlocalAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, x, scale, gamma):
        _require_cuda(q, k, v, x, scale, gamma)
        B = x.shape[0]
        C = x.shape[-1]
        d = q.shape[-1]
        N = x.numel() // (B * C)
        q, q_s = _rows(q); k, k_s = _rows(k); v, v_s = _rows(v); xr, x_s = _rows(x)
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        lse = torch.empty((B, N), device=x.device, dtype=torch.float32)
        # copy of O = softmax(.) v for the tensor-core gradient (E_i = <dy_i, O_i>, dgamma)
        need_o = any(ctx.needs_input_grad)
        o_save = torch.empty((B, N, C), device=x.device, dtype=torch.float32) if need_o else None
        s_t, s_ptr = _scalar_ptr(scale, x)
        g_t, g_ptr = _scalar_ptr(gamma, x)
        _lib_call("davi_nonlocal_attn_fwd", _p(q), _p(k), _p(v), _p(xr), _p(y), _p(lse), _p(o_save), B, N, d, C,
                  q_s, k_s, v_s, x_s, C, C, s_ptr, g_ptr, _stream())
        ctx.save_for_backward(q, k, v, lse, o_save, s_t, g_t)
        ctx.cfg = (B, N, d, C, q_s, k_s, v_s)
        return y

    @staticmethod
    def backward(ctx, dy):
        q, k, v, lse, o_save, s_t, g_t = ctx.saved_tensors
        B, N, d, C, q_s, k_s, v_s = ctx.cfg
        dy = dy.contiguous()
        lead = dy.shape[:-1]
        if q_s != d or k_s != d or v_s != C:   # dense gradients: use dense inputs
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
            q_s, k_s, v_s = d, d, C
        dq = torch.empty((*lead, d), device=dy.device, dtype=torch.float32)
        dk = torch.empty((*lead, d), device=dy.device, dtype=torch.float32)
        dv = torch.empty((*lead, C), device=dy.device, dtype=torch.float32)
        dgs = torch.empty(2, device=dy.device, dtype=torch.float32)
        wsb = _lib.load().davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
        ws = torch.empty(max(wsb // 4, 1), device=dy.device, dtype=torch.float32)
        _lib_call("davi_nonlocal_attn_bwd", _p(q), _p(k), _p(v), _p(lse), _p(o_save), _p(dy), _p(dq), _p(dk),
                  _p(dv), _p(dgs), B, N, d, C, q_s, k_s, v_s, C, C, _p(s_t), _p(g_t), _p(ws), wsb,
                  _stream())
        dscale = dgs[1] if (s_t is not None and ctx.needs_input_grad[4]) else None
        dgamma = dgs[0] if ctx.needs_input_grad[5] else None
        return dq, dk, dv, dy, dscale, dgamma


def nonlocal_attn(q, k, v, x, scale, gamma):
    """y = gamma * softmax(scale * q k^T) v + x over the H*W positions of each
    image (reference layers/attention.py:318-389).  q,k [B,H,W,d]; v,x [B,H,W,C]
    (channel-slice views allowed); scale (or None), gamma: 0-dim tensors."""
    C = x.shape[-1]
    if C % 4:
        # rows must be 16-byte multiples for the vector loads (the 138-channel
        # pre-processing blocks of the CIFAR-10 model): zero-pad the channel axis,
        # the pad columns of V contribute zeros and are sliced away again.
        pad = 4 - C % 4
        v = torch.nn.functional.pad(v, (0, pad))
        xp = torch.nn.functional.pad(x, (0, pad))
        return _NonlocalAttn.apply(q, k, v, xp, scale, gamma)[..., :C]
    return _NonlocalAttn.apply(q, k, v, x, scale, gamma)


class _NonlocalAttnPacked(torch.autograd.Function):
    """Same op on the PACKED projection output qkv [B,H,W,2d+C] (q | k | v along the channel axis):
    the kernels read the three column slices through strides and write dq | dk | dv straight into
    one packed gradient tensor, so neither forward slices nor backward zero-fill/add copies exist."""

    @staticmethod
    def forward(ctx, qkv, x, d, scale, gamma):
        _require_cuda(qkv, x, scale, gamma)
        B, C = x.shape[0], x.shape[-1]
        N = x.numel() // (B * C)
        rs = qkv.shape[-1]
        assert rs == 2 * d + C and d % 4 == 0 and C % 4 == 0
        qkv = qkv.contiguous()
        xr, x_s = _rows(x)
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        lse = torch.empty((B, N), device=x.device, dtype=torch.float32)
        o_save = torch.empty((B, N, C), device=x.device, dtype=torch.float32)
        s_t, s_ptr = _scalar_ptr(scale, x)
        g_t, g_ptr = _scalar_ptr(gamma, x)
        p = qkv.data_ptr()
        _lib_call("davi_nonlocal_attn_fwd", p, p + 4 * d, p + 8 * d, _p(xr), _p(y), _p(lse), _p(o_save), B, N, d, C,
                  rs, rs, rs, x_s, C, C, s_ptr, g_ptr, _stream())
        ctx.save_for_backward(qkv, lse, o_save, s_t, g_t)
        ctx.cfg = (B, N, d, C, rs)
        return y

    @staticmethod
    def backward(ctx, dy):
        qkv, lse, o_save, s_t, g_t = ctx.saved_tensors
        B, N, d, C, rs = ctx.cfg
        dy = dy.contiguous()
        dqkv = torch.empty_like(qkv)
        dgs = torch.empty(2, device=dy.device, dtype=torch.float32)
        wsb = _lib.load().davi_nonlocal_attn_bwd_workspace_bytes(B, N, d, C)
        ws = torch.empty(max(wsb // 4, 1), device=dy.device, dtype=torch.float32)
        p, g = qkv.data_ptr(), dqkv.data_ptr()
        _lib_call("davi_nonlocal_attn_bwd", p, p + 4 * d, p + 8 * d, _p(lse), _p(o_save), _p(dy), g, g + 4 * d,
                  g + 8 * d, _p(dgs), B, N, d, C, rs, rs, rs, C, C, _p(s_t), _p(g_t), _p(ws), wsb, _stream())
        dscale = dgs[1] if (s_t is not None and ctx.needs_input_grad[3]) else None
        dgamma = dgs[0] if ctx.needs_input_grad[4] else None
        return dqkv, dy, None, dscale, dgamma


def nonlocal_attn_packed(qkv, x, d, scale, gamma):
    """y = gamma * softmax(scale * q k^T) v + x with q | k | v packed along the last axis of
    qkv [B,H,W,2d+C] (the output of the fused 1x1 projection); x [B,H,W,C], C % 4 == 0."""
    return _NonlocalAttnPacked.apply(qkv, x, d, scale, gamma)


# ---------------------------------------------------------------------------
# R5/R6: bounded residual Gaussians, sample + KL (+ log q, log p)
# ---------------------------------------------------------------------------

class _GaussKl(torch.autograd.Function):
    """The forward launch also emits the gradient (davi_gauss_kl_fwd `grad`): the backward is one
    multiply-add pass over it (davi_gauss_kl_bwd_saved)."""

    @staticmethod
    def forward(ctx, post, prior, eps, noise_post, noise_prior, cfg):
        _require_cuda(post, prior, eps, noise_post, noise_prior)
        post_lo, post_hi, prior_lo, prior_hi, std_q, std_p, want_logp = cfg
        B = post.shape[0]
        z = eps.shape[-1]
        P = eps.numel() // (B * z)
        post = post.contiguous()
        prior = prior.contiguous() if prior is not None else None
        eps = eps.contiguous()
        noise_post = noise_post.contiguous() if noise_post is not None else None
        noise_prior = noise_prior.contiguous() if noise_prior is not None else None
        xi = torch.empty_like(eps)
        kl = torch.empty(B, device=post.device, dtype=torch.float32)
        logq = torch.empty_like(kl) if want_logp else None
        logp = torch.empty_like(kl) if want_logp else None
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        grad = torch.empty((B, P, 6 * z), device=post.device, dtype=torch.float32) if need else None
        _lib_call("davi_gauss_kl_fwd", _p(post), _p(prior), _p(eps), _p(noise_post), _p(noise_prior),
                  _p(xi), _p(kl), _p(logq), _p(logp), B, P, z, post_lo, post_hi, prior_lo, prior_hi,
                  std_q, std_p, _p(grad), _stream())
        ctx.save_for_backward(grad)
        ctx.cfg = (B, P, z, prior is not None, post.shape)
        if want_logp:
            ctx.mark_non_differentiable(logq, logp)
            return xi, kl, logq, logp
        return xi, kl

    @staticmethod
    def backward(ctx, dxi, dkl, *unused):
        (grad,) = ctx.saved_tensors
        B, P, z, has_prior, shape = ctx.cfg
        dxi = dxi.contiguous() if dxi is not None else None
        dkl = dkl.contiguous() if dkl is not None else None
        dpost = torch.empty(shape, device=grad.device, dtype=torch.float32)
        dprior = torch.empty_like(dpost) if has_prior else None
        _lib_call("davi_gauss_kl_bwd_saved", _p(grad), _p(dxi), _p(dkl), _p(dpost), _p(dprior), B, P, z, _stream())
        return dpost, dprior, None, None, None, None


def gauss_sample_kl(post, prior, eps, post_bounds, prior_bounds=(-5.0, 5.0), noise_post=None,
                    noise_prior=None, noise_std_post=0.0, noise_std_prior=0.0, compute_logp=False):
    """post/prior [B,H,W,2z] raw (mu || log-sigma); prior None for layer 0.
    Returns (xi [B,H,W,z], kl [B]) or (xi, kl, logq [B], logp [B])."""
    cfg = (float(post_bounds[0]), float(post_bounds[1]), float(prior_bounds[0]),
           float(prior_bounds[1]), float(noise_std_post), float(noise_std_prior), bool(compute_logp))
    return _GaussKl.apply(post, prior, eps, noise_post, noise_prior, cfg)


# ---------------------------------------------------------------------------
# R7/R8: data negative log-likelihoods with the gradient fused into the forward
# ---------------------------------------------------------------------------

class _DlmNll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, x, ls_low):
        _require_cuda(params, x)
        B = params.shape[0]
        P = params.numel() // (B * 50)
        params, x = params.contiguous(), x.contiguous()
        nll = torch.empty(B, device=params.device, dtype=torch.float32)
        need = ctx.needs_input_grad[0]
        g = torch.empty_like(params) if need else None
        _lib_call("davi_dlm_nll_fwd", _p(params), _p(x), _p(nll), _p(g), B, P, float(ls_low), _stream())
        ctx.save_for_backward(g)
        ctx.B = B
        return nll

    @staticmethod
    def backward(ctx, dnll):
        (g,) = ctx.saved_tensors
        out = torch.empty_like(g)
        dnll = dnll.contiguous()
        _lib_call("davi_scale_rows", _p(g), _p(dnll), _p(out), ctx.B, g.numel() // ctx.B, _stream())
        return out, None, None


def dlm_nll(params, x, log_scale_low_bound=-7.0):
    """params [B,H,W,50], x [B,H,W,3] in [-1,1] -> -log p(x) [B]
    (reference dlm.py:311-383, data_layer.py:179-181)."""
    return _DlmNll.apply(params, x, log_scale_low_bound)


class _BernoulliNll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, x, crop):
        _require_cuda(logits, x)
        B, H, W = logits.shape[0], logits.shape[1], logits.shape[2]
        logits, x = logits.contiguous(), x.contiguous()
        nll = torch.empty(B, device=logits.device, dtype=torch.float32)
        need = ctx.needs_input_grad[0]
        g = torch.empty_like(logits) if need else None
        _lib_call("davi_bernoulli_nll_fwd", _p(logits), _p(x), _p(nll), _p(g), B, H, W, int(crop),
                  _stream())
        ctx.save_for_backward(g)
        ctx.B = B
        return nll

    @staticmethod
    def backward(ctx, dnll):
        (g,) = ctx.saved_tensors
        out = torch.empty_like(g)
        dnll = dnll.contiguous()
        _lib_call("davi_scale_rows", _p(g), _p(dnll), _p(out), ctx.B, g.numel() // ctx.B, _stream())
        return out, None, None


def bernoulli_nll(logits, x, crop):
    """logits, x [B,H,W,1] -> -log p(x) [B] over the cropped window
    (reference bern.py:155-179)."""
    return _BernoulliNll.apply(logits, x, crop)


def is_logsumexp(log_w, num_samples):
    """log_w [S*B] sample-major -> logsumexp over samples [B] (reference vae.py:480-482)."""
    _require_cuda(log_w)
    log_w = log_w.contiguous()
    B = log_w.numel() // num_samples
    out = torch.empty(B, device=log_w.device, dtype=torch.float32)
    _lib_call("davi_is_logsumexp", _p(log_w), _p(out), num_samples, B, _stream())
    return out


# ---------------------------------------------------------------------------
# host-model glue (row f1): convolution with the bias gradient on a streaming column-sum kernel
# ---------------------------------------------------------------------------

def colsum_nhwc(g):
    """g: NCHW-shaped tensor in channels_last memory (or [rows, C]) -> per-channel sums [C]."""
    _require_cuda(g)
    if g.dim() == 4:
        g = g.permute(0, 2, 3, 1)
    if not g.is_contiguous():
        g = g.contiguous()
    C = g.shape[-1]
    out = torch.empty(C, device=g.device, dtype=torch.float32)
    if C % 4:
        return g.reshape(-1, C).sum(0)
    rows = g.numel() // C
    wsb = _lib.load().davi_colsum_workspace_bytes(rows, C)
    ws = torch.empty(wsb // 4, device=g.device, dtype=torch.float32)
    _lib_call("davi_colsum", _p(g), _p(out), rows, C, C, _p(ws), wsb, _stream())
    return out


class _ConvBias(torch.autograd.Function):
    """cuDNN convolution (through torch) whose bias gradient is davi_colsum instead of torch's
    generic reduction (~25 us -> ~4 us per layer at [40,16,16,316])."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding)
        return torch.nn.functional.conv2d(x, w, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.cfg
        s = (stride, stride) if isinstance(stride, int) else tuple(stride)
        p = (padding, padding) if isinstance(padding, int) else tuple(padding)
        gx, gw, _ = torch.ops.aten.convolution_backward(
            gy, x, w, [w.shape[0]], s, p, (1, 1), False, (0, 0), 1,
            (ctx.needs_input_grad[0], ctx.needs_input_grad[1], False))
        gb = colsum_nhwc(gy) if ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None


# Precision of the host-side cuDNN convolutions (row f1).  The reference's Conv2D layers are fp32:
#   "3xtf32" (default)  fp32-grade on the TF32 tensor cores: operands split by davi_split_tf32, the three
#                       compensated products accumulated by ONE cuDNN convolution over a tripled axis
#   "fp32"              cuDNN's CUDA-core fp32 kernels (the bit-level yardstick of the parity tests)
#   "tf32"              plain TF32 (13 operand mantissa bits dropped; opt-in, NOT reference precision)
CONV_PRECISIONS = ("3xtf32", "fp32", "tf32")
_conv_precision = "3xtf32"


def set_conv_precision(mode):
    global _conv_precision
    if mode not in CONV_PRECISIONS:
        raise ValueError(f"conv precision {mode!r}: expected one of {CONV_PRECISIONS}")
    _conv_precision = mode
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"


def get_conv_precision():
    return _conv_precision


def split_tf32(t, layout, pattern):
    """t: NCHW-shaped channels_last tensor (or any contiguous [..., C]) -> its three-part TF32 split.
    layout 0: channel axis tripled (same memory format); layout 1: batch axis tripled."""
    _require_cuda(t)
    if t.dim() == 4:
        nhwc = t.permute(0, 2, 3, 1)
        if not nhwc.is_contiguous():
            nhwc = nhwc.contiguous()
        B, H, W, C = nhwc.shape
        out = torch.empty((B, H, W, 3 * C) if layout == 0 else (3 * B, H, W, C), device=t.device, dtype=torch.float32)
        _lib_call("davi_split_tf32", _p(nhwc), _p(out), B * H * W, C, layout, pattern, _stream())
        return out.permute(0, 3, 1, 2)
    t = t.contiguous()
    C = t.shape[-1]
    rows = t.numel() // C
    out = torch.empty((rows, 3 * C) if layout == 0 else (3, rows, C), device=t.device, dtype=torch.float32)
    _lib_call("davi_split_tf32", _p(t), _p(out), rows, C, layout, pattern, _stream())
    return out


def split_tf32_dual(t, pattern_ch, pattern_pl):
    """Both layouts of split_tf32 from one read of t (NCHW-shaped, channels_last memory)."""
    _require_cuda(t)
    nhwc = t.permute(0, 2, 3, 1)
    if not nhwc.is_contiguous():
        nhwc = nhwc.contiguous()
    B, H, W, C = nhwc.shape
    ch = torch.empty((B, H, W, 3 * C), device=t.device, dtype=torch.float32)
    pl = torch.empty((3 * B, H, W, C), device=t.device, dtype=torch.float32)
    _lib_call("davi_split_tf32_dual", _p(nhwc), _p(ch), _p(pl), B * H * W, C, pattern_ch, pattern_pl, _stream())
    return ch.permute(0, 3, 1, 2), pl.permute(0, 3, 1, 2)


class _ConvBias3x(torch.autograd.Function):
    """fp32-grade convolution on the TF32 tensor cores.  With a ~= hi_a + lo_a (davi_split_tf32):
        y  = conv([x_hi | x_lo | x_hi], [w_hi | w_hi | w_lo])            contraction over 3 * C_in
        gx = conv_input([gy_hi | gy_lo | gy_hi], [w_hi ; w_hi ; w_lo])   contraction over 3 * C_out
        gw = conv_weight([x_hi ; x_lo ; x_hi], [gy_hi ; gy_hi ; gy_lo])  contraction over 3 * batch
    each ONE cuDNN TF32 call with fp32 accumulation (both parts are exact in TF32; lo carries the next
    11 bits).  One split kernel per tensor writes the layout the forward needs and the one the backward
    needs; the backward layouts are what is saved."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.cfg = (stride, padding)
        need_gx, need_gw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_gw:
            x3, x3p = split_tf32_dual(x, 0, 0)
        else:
            x3, x3p = split_tf32(x, 0, 0), None
        if need_gx:
            w3, w3p = split_tf32_dual(w, 1, 1)
        else:
            w3, w3p = split_tf32(w, 0, 1), None
        ctx.save_for_backward(x3p, w3p)
        ctx.shapes = (x.shape, w.shape)
        return torch.nn.functional.conv2d(x3, w3, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        x3p, w3p = ctx.saved_tensors
        (xs, ws), (stride, padding) = ctx.shapes, ctx.cfg
        s = (stride, stride) if isinstance(stride, int) else tuple(stride)
        p = (padding, padding) if isinstance(padding, int) else tuple(padding)
        need_gx, need_gw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = gw = None
        gb = colsum_nhwc(gy) if ctx.needs_input_grad[2] else None      # first: gy is still L2-resident
        if need_gx and need_gw:
            gy3, gy3p = split_tf32_dual(gy, 0, 1)
        else:
            gy3 = split_tf32(gy, 0, 0) if need_gx else None
            gy3p = split_tf32(gy, 1, 1) if need_gw else None
        # aten.convolution_backward takes shape and memory format of the result from its `input` / `weight`
        # argument: one plane of the saved split (same shape and channels_last layout as x / w) serves
        if need_gx:
            like_x = x3p[:xs[0]] if x3p is not None else torch.empty(xs, device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
            gx = torch.ops.aten.convolution_backward(gy3, like_x, w3p, [w3p.shape[0]], s, p, (1, 1), False,
                                                     (0, 0), 1, (True, False, False))[0]
        if need_gw:
            like_w = w3p[:ws[0]] if w3p is not None else torch.empty(ws, device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
            gw = torch.ops.aten.convolution_backward(gy3p, x3p, like_w, [ws[0]], s, p, (1, 1), False, (0, 0), 1,
                                                     (False, True, False))[1]
        return gx, gw, gb, None, None


# Implementation of the fp32-grade ("3xtf32") convolutions:
#   "tcgen05" (default)  the repo's own implicit-GEMM kernels (csrc/conv_tc.cu: davi_conv2d_fwd / _wgrad) for every
#                        shape inside their envelope (stride 1, 'same' padding, 16-byte-aligned channel counts);
#                        the few others (stride-2 cells, the 138-channel pre-processor) go through cuDNN below
#   "cudnn"              one cuDNN TF32 call per conv / dgrad / wgrad over a tripled contraction axis (_ConvBias3x)
CONV_IMPLS = ("tcgen05", "cudnn")
_conv_impl = "tcgen05"
conv_pre_lo = True     # use the caller's davi_tf32_lo images of the weights when it passes them (False: A/B measurements)
conv_trace = None      # measurement hook (bench.py): a list collects ("fwd" | "fwd_pre" | "wgrad", B, H, W, Cin, Cout, kh, kw) per launch
_conv_ws = {}          # device index -> zero-initialised workspace of the stream-K split (one per device: the
                       # model's convolutions are issued on one stream at a time)


def set_conv_impl(impl):
    global _conv_impl
    if impl not in CONV_IMPLS:
        raise ValueError(f"conv implementation {impl!r}: expected one of {CONV_IMPLS}")
    _conv_impl = impl


def get_conv_impl():
    return _conv_impl


def _conv_workspace(dev):
    ws = _conv_ws.get(dev.index)
    if ws is None:
        nbytes = _lib.load().davi_conv2d_workspace_bytes()
        ws = _conv_ws[dev.index] = torch.zeros(nbytes // 4, device=dev, dtype=torch.float32)
    return ws, ws.numel() * 4


def _nhwc(t):
    """NCHW-shaped tensor -> its [B,H,W,C] view, contiguous (a copy only if it was not channels_last)."""
    v = t.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _conv_tc_shape_ok(B, H, W, Cin, Cout, kh, kw, stride, padding):
    s = (stride, stride) if isinstance(stride, int) else tuple(stride)
    p = (padding, padding) if isinstance(padding, int) else tuple(padding)
    if s != (1, 1) or p != ((kh - 1) // 2, (kw - 1) // 2):
        return False
    return bool(_lib.load().davi_conv2d_supported(B, H, W, Cin, Cout, kh, kw))


def conv_tc_supported(x, w, stride, padding):
    """True when F.conv2d(x, w, stride, padding) (NCHW shapes) lies inside the envelope of the tcgen05 kernels."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    return _conv_tc_shape_ok(B, H, W, Cin, Cout, kh, kw, stride, padding)


def conv2d_tc_fwd(x, w, b, act=0, w_lo=None):
    """x [B,H,W,Cin] contiguous, w [Cout,kh,kw,Cin] contiguous (OHWI), b [Cout] or None -> y [B,H,W,Cout]
    = conv(swish(x) if act else x) + b.  w_lo: optional davi_tf32_lo image of w (same layout) that the caller keeps up
    to date; the kernel then loads it instead of deriving it from every staged weight tile (same bits out)."""
    B, H, W, Cin = x.shape
    Cout, kh, kw, _ = w.shape
    y = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.float32)
    ws, wsb = _conv_workspace(x.device)
    pre = w_lo is not None and conv_pre_lo
    if conv_trace is not None:
        conv_trace.append(("fwd_pre" if pre else "fwd", B, H, W, Cin, Cout, kh, kw))
    if pre:
        assert w_lo.shape == w.shape and w_lo.is_contiguous()
        _lib_call("davi_conv2d_fwd_pre", _p(x), _p(w), _p(w_lo), _p(b), _p(y), B, H, W, Cin, Cout, kh, kw, Cin, Cout, int(act),
                  _p(ws), wsb, _stream())
        return y
    _lib_call("davi_conv2d_fwd_ex", _p(x), _p(w), _p(b), _p(y), B, H, W, Cin, Cout, kh, kw, Cin, Cout, int(act), _p(ws), wsb,
              _stream())
    return y


def conv2d_tc_wgrad(x, gy, kh, kw, act=0):
    """x [B,H,W,Cin], gy [B,H,W,Cout] contiguous -> gw [Cout,kh,kw,Cin] = wgrad(swish(x) if act else x, gy)."""
    B, H, W, Cin = x.shape
    Cout = gy.shape[-1]
    gw = torch.empty((Cout, kh, kw, Cin), device=x.device, dtype=torch.float32)
    ws, wsb = _conv_workspace(x.device)
    if conv_trace is not None:
        conv_trace.append(("wgrad", B, H, W, Cin, Cout, kh, kw))
    _lib_call("davi_conv2d_wgrad_ex", _p(x), _p(gy), _p(gw), B, H, W, Cin, Cout, kh, kw, Cin, Cout, int(act), _p(ws), wsb,
              _stream())
    return gw


class _ConvTC(torch.autograd.Function):
    """The convolution of the ResNet cells on the repo's tcgen05 kernels (fp32-grade, 3xTF32 compensation on chip):
    y = conv(swish(x) if act else x, w) + b.  The activation is applied in the operand path of the forward (free: the
    lo warps have the slack) and recomputed there by the weight gradient (+5 us against the 9 us pass it replaces), so
    neither swish(x) nor a copy of it is ever stored; the input gradient = the forward kernel on the transposed +
    flipped weights (davi_conv2d_wt), times swish'(x) by ONE elementwise kernel (measured: the same product inside the
    convolution's epilogue costs +20 us per launch -- one CTA per SM has too few loads in flight); bias gradient =
    davi_colsum.  Saves x and w themselves -- no split copies."""

    @staticmethod
    def forward(ctx, x, w, b, act, wt, lo):
        xh, wh = _nhwc(x), _nhwc(w)
        ctx.save_for_backward(xh, wh)
        ctx.act = int(act)
        ctx.wt = wt         # [Cin,kh,kw,Cout]: the trainer's per-step davi_conv2d_wt_batched image of w, or None
        w_lo, ctx.wt_lo = lo if lo is not None else (None, None)      # davi_tf32_lo images of w and wt (trainer), or None
        return conv2d_tc_fwd(xh, wh, b, act=act, w_lo=w_lo).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xh, wh = ctx.saved_tensors
        act = ctx.act
        gyh = _nhwc(gy)
        Cout, kh, kw, Cin = wh.shape
        gx = gw = gb = None
        if ctx.needs_input_grad[2]:
            gb = colsum_nhwc(gyh.permute(0, 3, 1, 2))              # first: gy is still L2-resident (gyh: no second copy)
        if ctx.needs_input_grad[0]:
            wt, wt_lo = ctx.wt, ctx.wt_lo
            if wt is None:
                wt, wt_lo = torch.empty((Cin, kh, kw, Cout), device=wh.device, dtype=torch.float32), None
                _lib_call("davi_conv2d_wt", _p(wh), _p(wt), Cout, kh * kw, Cin, _stream())
            gx = conv2d_tc_fwd(gyh, wt, None, w_lo=wt_lo)
            if act:
                gx = torch.ops.aten.silu_backward(gx, xh)
            gx = gx.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = conv2d_tc_wgrad(xh, gyh, kh, kw, act=act).permute(0, 3, 1, 2)
        return gx, gw, gb, None, None, None


def _scale_add(y, add, alpha):
    """add + alpha * y (res.py:284-286) as ONE element-wise kernel (torch.add's alpha), not a product and a sum."""
    if add is None:
        return y if alpha == 1.0 else alpha * y
    if isinstance(alpha, (int, float)):
        return torch.add(add, y, alpha=alpha)
    return alpha * y + add


class _SplitChannels(torch.autograd.Function):
    """t[..., :s0], t[..., s0:s0+s1], ... as views; the gradient is ONE concatenation of the pieces' gradients.  (Plain
    slices cost a zero fill and a strided copy per piece plus the sums of the padded pieces on the way back: 8 launches
    for the three-way split of a decoder cell's output, vl.py:375-380.)"""

    @staticmethod
    def forward(ctx, t, *sizes):
        assert sum(sizes) == t.shape[-1]
        ctx.sizes = sizes
        outs, c = [], 0
        for n in sizes:
            outs.append(t[..., c:c + n])
            c += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        return (torch.cat(grads, dim=-1),) + (None,) * len(ctx.sizes)


split_channels_fused = os.environ.get("DAVI_PLAIN_SLICES", "") != "1"     # False: plain slices (A/B measurements)


def split_channels(t, *sizes):
    """Split the last axis of `t` into consecutive pieces of the given sizes (views, no copy)."""
    if split_channels_fused and t.requires_grad and torch.is_grad_enabled():
        return _SplitChannels.apply(t, *sizes)
    outs, c = [], 0
    for n in sizes:
        outs.append(t[..., c:c + n])
        c += n
    return tuple(outs)


def conv2d_bias(x, w, b, stride=1, padding=0, act=None, add=None, alpha=1.0, wt=None, lo=None):
    """add + alpha * F.conv2d(act(x), w, b) on NCHW views of channels_last tensors (act: None or "swish"; add: None or a
    tensor like the result), with the fast bias gradient, in the precision selected by set_conv_precision() (and, for
    "3xtf32", the implementation of set_conv_impl()): the ResNet node of res.py:613-633 and the block's
    shortcut + 0.1 * node of res.py:273-286.  wt: optional [Cin,kh,kw,Cout] davi_conv2d_wt image of w that the caller
    keeps up to date (the trainer: one batched launch per step); without it the input gradient computes its own.
    lo: optional (w_lo, wt_lo), the davi_tf32_lo images of w (OHWI memory) and of wt, likewise kept by the caller.
    The tcgen05 kernels take the activation inside (conv(swish(x)) is ONE kernel);
    residual and scale stay elementwise torch operations everywhere (fusing them into the convolution's epilogue was
    measured and dropped, see _ConvTC)."""
    _require_cuda(x, w, b)
    if act not in (None, "swish"):
        raise ValueError(f"conv2d_bias: activation {act!r} cannot be fused (None or 'swish')")
    if _conv_precision == "3xtf32" and _conv_impl == "tcgen05":
        if conv_tc_supported(x, w, stride, padding):
            y = _ConvTC.apply(x, w, b, 1 if act else 0, wt, lo if wt is not None else None)
            return _scale_add(y, add, alpha)
        # channel counts that are not 16-byte multiples (the 138-channel pre/post-processing cells, the 3-channel
        # stem, the 50-channel output head): zero-pad the channel axes, run the same kernels, slice the result
        # (autograd turns the pads into slices and the slice into a pad on the way back)
        pi, po = (-x.shape[1]) % 4, (-w.shape[0]) % 4
        if (pi or po) and _conv_tc_shape_ok(x.shape[0], x.shape[2], x.shape[3], x.shape[1] + pi, w.shape[0] + po,
                                            w.shape[2], w.shape[3], stride, padding):
            F = torch.nn.functional
            xp = F.pad(x.permute(0, 2, 3, 1), (0, pi)).permute(0, 3, 1, 2)
            wp = F.pad(w.permute(0, 2, 3, 1), (0, pi, 0, 0, 0, 0, 0, po)).permute(0, 3, 1, 2)
            bp = None if b is None else F.pad(b, (0, po))
            y = _ConvTC.apply(xp, wp, bp, 1 if act else 0, None, None)[:, :w.shape[0]]
            return _scale_add(y, add, alpha)
    if act:
        x = torch.nn.functional.silu(x)
    y = _conv2d_plain(x, w, b, stride, padding)
    return _scale_add(y, add, alpha)


def _conv2d_plain(x, w, b, stride, padding):
    if _conv_precision == "3xtf32":
        if not torch.backends.cudnn.allow_tf32:
            raise RuntimeError("conv precision '3xtf32' needs torch.backends.cudnn.allow_tf32 = True "
                               "(use ops.set_conv_precision)")
        return _ConvBias3x.apply(x, w, b, stride, padding)
    if b is None:
        return torch.nn.functional.conv2d(x, w, b, stride=stride, padding=padding)
    return _ConvBias.apply(x, w, b, stride, padding)
