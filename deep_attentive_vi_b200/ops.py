"""torch.autograd wrappers over the C ABI (include/davi.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape;
every forward and every gradient below is one or two launches of the hand-
written sm_100a kernels in ``csrc/`` through ``libdavi.so``.  There is no
fallback: CPU tensors, a missing library or a failing call raise.
"""
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

class _DwAttnSlots(torch.autograd.Function):
    """inputs: query, scale, n, then k_0..k_{n-1}, v_0..v_{n-1} (views allowed)."""

    @staticmethod
    def forward(ctx, query, scale, n, *kv):
        ks, vs = kv[:n], kv[n:]
        _require_cuda(query, scale, *ks, *vs)
        B, H, W, d = ks[0].shape
        C = vs[0].shape[-1]
        P = H * W
        q, q_sp = _rows(query)
        # every slot must share one (batch, pixel) stride pair
        ks = [_rows(t)[0] for t in ks]
        vs = [_rows(t)[0] for t in vs]
        k_sp, v_sp = ks[0].stride(-2), vs[0].stride(-2)
        if any(t.stride(-2) != k_sp for t in ks):
            ks = [t.contiguous() for t in ks]; k_sp = d
        if any(t.stride(-2) != v_sp for t in vs):
            vs = [t.contiguous() for t in vs]; v_sp = C
        out = torch.empty((B, H, W, C), device=q.device, dtype=torch.float32)
        s_t, s_ptr = _scalar_ptr(scale, q)
        kt, ka = _lib.ptr_table([t.data_ptr() for t in ks])
        vt, va = _lib.ptr_table([t.data_ptr() for t in vs])
        _lib_call("davi_dw_attn_slots_fwd", _p(q), kt, vt, _p(out), 0, B, n, P, d, C,
                  q_sp, P * k_sp, k_sp, P * v_sp, v_sp, C, s_ptr, _stream())
        ctx.save_for_backward(q, s_t, *ks, *vs)
        ctx.dims = (B, n, P, d, C, q_sp, k_sp, v_sp)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, n, P, d, C, q_sp, k_sp, v_sp = ctx.dims
        saved = ctx.saved_tensors
        q, s_t = saved[0], saved[1]
        ks, vs = saved[2:2 + n], saved[2 + n:]
        dout = dout.contiguous()
        dq = torch.empty(q.shape, device=q.device, dtype=torch.float32)
        dks = [torch.empty(t.shape, device=q.device, dtype=torch.float32) for t in ks]
        dvs = [torch.empty(t.shape, device=q.device, dtype=torch.float32) for t in vs]
        need_ds = s_t is not None and ctx.needs_input_grad[1]
        dscale = torch.empty((), device=q.device, dtype=torch.float32) if need_ds else None
        # gradients are dense, so they get their own strides: run the kernel on dense
        # copies of strided inputs only when needed (contiguous slots are the fast path)
        if q_sp != d or k_sp != d or v_sp != C:
            q = q.contiguous(); ks = [t.contiguous() for t in ks]; vs = [t.contiguous() for t in vs]
            q_sp, k_sp, v_sp = d, d, C
        kt, ka = _lib.ptr_table([t.data_ptr() for t in ks])
        vt, va = _lib.ptr_table([t.data_ptr() for t in vs])
        dkt, dka = _lib.ptr_table([t.data_ptr() for t in dks])
        dvt, dva = _lib.ptr_table([t.data_ptr() for t in dvs])
        _lib_call("davi_dw_attn_slots_bwd", _p(q), kt, vt, _p(dout), _p(dq), dkt, dvt, _p(dscale),
                  B, n, P, d, C, q_sp, P * k_sp, k_sp, P * v_sp, v_sp, C, _p(s_t), 0, _stream())
        return (dq, dscale, None, *dks, *dvs)


def dw_attn_slots(query, keys, values, scale=None):
    """query [B,H,W,d]; keys/values: lists of n tensors [B,H,W,d] / [B,H,W,C]
    (channel-slice views of wider tensors are passed without copying)."""
    n = len(keys)
    assert n == len(values) and n >= 1
    return _DwAttnSlots.apply(query, scale, n, *keys, *values)


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


# ---------------------------------------------------------------------------
# R5/R6: bounded residual Gaussians, sample + KL (+ log q, log p)
# ---------------------------------------------------------------------------

class _GaussKl(torch.autograd.Function):
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
        _lib_call("davi_gauss_kl_fwd", _p(post), _p(prior), _p(eps), _p(noise_post), _p(noise_prior),
                  _p(xi), _p(kl), _p(logq), _p(logp), B, P, z, post_lo, post_hi, prior_lo, prior_hi,
                  std_q, std_p, _stream())
        ctx.save_for_backward(post, prior, eps, noise_post, noise_prior)
        ctx.cfg = (B, P, z) + tuple(cfg[:6])
        if want_logp:
            ctx.mark_non_differentiable(logq, logp)
            return xi, kl, logq, logp
        return xi, kl

    @staticmethod
    def backward(ctx, dxi, dkl, *unused):
        post, prior, eps, noise_post, noise_prior = ctx.saved_tensors
        B, P, z, post_lo, post_hi, prior_lo, prior_hi, std_q, std_p = ctx.cfg
        dxi = dxi.contiguous() if dxi is not None else None
        dkl = dkl.contiguous() if dkl is not None else None
        dpost = torch.empty_like(post)
        dprior = torch.empty_like(prior) if prior is not None else None
        _lib_call("davi_gauss_kl_bwd", _p(post), _p(prior), _p(eps), _p(noise_post), _p(noise_prior),
                  _p(dxi), _p(dkl), _p(dpost), _p(dprior), B, P, z, post_lo, post_hi, prior_lo,
                  prior_hi, std_q, std_p, _stream())
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
