"""ctypes binding of the C ABI in ``include/davi.h`` (libdavi.so).

This is the only place the shared library is loaded.  There is no fallback of
any kind: if the library is missing or a call fails, a ``DaviError`` is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdavi.so")

_f = ctypes.c_void_p      # device pointer (float*)
_i = ctypes.c_int
_l = ctypes.c_int64
_r = ctypes.c_float
_s = ctypes.c_void_p      # cudaStream_t
_z = ctypes.c_size_t
_pp = ctypes.POINTER(ctypes.c_void_p)
_lp = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); mirrors include/davi.h declaration by declaration
SIGNATURES = {
    "davi_version": (_i, []),
    "davi_last_error": (_i, [ctypes.c_char_p, _z]),
    "davi_set_option": (_i, [_i, _i]),
    "davi_get_option": (_i, [_i]),
    "davi_dw_attn_fwd": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _l, _l, _l, _l, _l, _l, _l, _l, _f, _s]),
    "davi_dw_attn_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _l, _l, _l, _l, _l, _l, _l, _l, _f, _s]),
    "davi_dw_attn_slots_fwd": (_i, [_f, _pp, _pp, _lp, _lp, _f, _f, _i, _i, _i, _i, _i, _l, _l, _f, _s]),
    "davi_dw_attn_slots_bwd": (_i, [_f, _pp, _pp, _lp, _lp, _f, _f, _pp, _pp, _lp, _lp, _f, _i, _i, _i, _i, _i,
                                    _l, _l, _f, ctypes.c_uint32, _f, _f, _s]),
    "davi_dw_attn_gather": (_i, [_i, ctypes.POINTER(ctypes.c_int), _pp, _pp, _lp, _lp, _pp, _lp, _pp, _lp, _pp, _pp,
                                 _lp, _i, _i, _i, _i, _s]),
    "davi_ln_gelu_fwd": (_i, [_f, _f, _f, _f, _l, _i, _l, _l, _r, _i, _s]),
    "davi_ln_gelu_bwd_workspace_bytes": (_z, [_l, _i]),
    "davi_ln_gelu_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _l, _i, _l, _l, _l, _r, _i, _f, _z, _s]),
    "davi_ln_gelu_res_fwd": (_i, [_f, _f, _f, _f, _f, _f, _l, _i, _l, _l, _l, _r, _s]),
    "davi_ln_gelu_res_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _l, _i, _l, _l, _l, _r, _f, _z, _s]),
    "davi_nonlocal_attn_fwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _l, _l, _l, _l, _l, _l, _f, _f, _s]),
    "davi_nonlocal_attn_bwd_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "davi_nonlocal_attn_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _l, _l, _l, _l, _l, _f, _f, _f, _z, _s]),
    "davi_gauss_kl_fwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _r, _r, _r, _r, _r, _r, _f, _s]),
    "davi_gauss_kl_bwd_saved": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _s]),
    "davi_gauss_kl_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _r, _r, _r, _r, _r, _r, _s]),
    "davi_dlm_nll_fwd": (_i, [_f, _f, _f, _f, _i, _i, _r, _s]),
    "davi_bernoulli_nll_fwd": (_i, [_f, _f, _f, _f, _i, _i, _i, _i, _s]),
    "davi_scale_rows": (_i, [_f, _f, _f, _i, _l, _s]),
    "davi_is_logsumexp": (_i, [_f, _f, _i, _i, _s]),
    "davi_weight_norm_fwd": (_i, [_f, _f, _f, _f, _f, _f, _l, _s]),
    "davi_weight_norm_bwd": (_i, [_f, _f, _f, _f, _f, _f, _f, _l, _s]),
    "davi_colsum_workspace_bytes": (_z, [_l, _i]),
    "davi_colsum": (_i, [_f, _f, _l, _i, _l, _f, _z, _s]),
    "davi_multi_copy": (_i, [_pp, _lp, _lp, _i, _f, _s]),
    "davi_split_tf32": (_i, [_f, _f, _l, _i, _i, _i, _s]),
    "davi_split_tf32_dual": (_i, [_f, _f, _f, _l, _i, _i, _i, _s]),
    "davi_conv2d_workspace_bytes": (_z, []),
    "davi_fill_zero": (_i, [_f, _l, _s]),
    "davi_conv2d_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _lp]),
    "davi_conv2d_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "davi_conv2d_fwd": (_i, [_f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _f, _z, _s]),
    "davi_conv2d_fwd_ex": (_i, [_f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _i, _f, _z, _s]),
    "davi_tf32_lo": (_i, [_f, _f, _l, _s]),
    "davi_conv2d_fwd_pre": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _i, _f, _z, _s]),
    "davi_conv2d_wgrad_ex": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _i, _f, _z, _s]),
    "davi_conv2d_wgrad": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _l, _l, _f, _z, _s]),
    "davi_conv2d_wt": (_i, [_f, _f, _i, _i, _i, _s]),
    "davi_conv2d_wt_batched": (_i, [_f, _f, _f, _f, _f, _f, _f, _i, _l, _s]),
    "davi_adamax_step": (_i, [_f, _f, _f, _f, _l, _r, _r, _r, _r, _l, _r, _r, _s]),
    "davi_adamax_step_dev": (_i, [_f, _f, _f, _f, _l, _f, _r, _r, _r, _r, _s]),
}


# davi_set_option switches (include/davi.h)
OPT_DW_IMPL, OPT_NONLOCAL_IMPL, OPT_NL_DS, OPT_NL_BWD, OPT_CONV_DEBUG, OPT_CONV_SM_RESERVE = 0, 1, 2, 3, 4, 5


class DaviError(RuntimeError):
    pass


_lib = None


def load():
    """Load libdavi.so (building nothing: see build.py).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DaviError(
            f"{LIB_PATH} not found: build it with `python -m deep_attentive_vi_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    load().davi_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what):
    if rc != 0:
        raise DaviError(f"{what} failed (rc={rc}): {last_error()}")


# kernels launched by one successful call of each entry point (bench.py's
# `gpu_launches` claim is the sum over the timed region)
KERNELS_PER_CALL = {"davi_ln_gelu_bwd": 2, "davi_ln_gelu_res_bwd": 2, "davi_nonlocal_attn_bwd": 4, "davi_colsum": 2}
_launches = 0


def launch_count():
    return _launches


def add_launches(n):
    """Account for kernels launched without a Python-level call (CUDA-graph replays)."""
    global _launches
    _launches += n


# Measurement hook (bench.py): {"prefix": str, "events": []} brackets every call whose name starts with
# the prefix with a pair of CUDA events on the current stream (in-step kernel time of an eager step).
timed = None


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero code."""
    global _launches
    if timed is not None and name.startswith(timed["prefix"]):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(load(), name)(*args), name)
        e1.record()
        timed["events"].append((name, e0, e1))
    else:
        check(getattr(load(), name)(*args), name)
    _launches += KERNELS_PER_CALL.get(name, 1)


def set_option(option, value):
    check(load().davi_set_option(option, value), "davi_set_option")


def ptr_table(ptrs):
    """Host array of device pointers for the *_slots_* entry points."""
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    return ctypes.cast(arr, _pp), arr  # keep `arr` alive for the duration of the call


def int_table(vals):
    """Host array of int64 (per-slot strides)."""
    arr = (ctypes.c_int64 * len(vals))(*vals)
    return ctypes.cast(arr, _lp), arr
