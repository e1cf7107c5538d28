"""CPU-side checks of the drop-in boundary: libdavi.so loads, exports every
symbol include/davi.h declares, the ctypes table matches the header, and
argument validation fails loudly (no compute is launched without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "davi.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(davi_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ["davi_dw_attn_fwd", "davi_dw_attn_bwd", "davi_dw_attn_slots_fwd",
                 "davi_dw_attn_slots_bwd", "davi_nonlocal_attn_fwd", "davi_nonlocal_attn_bwd",
                 "davi_gauss_kl_fwd", "davi_gauss_kl_bwd", "davi_dlm_nll_fwd",
                 "davi_bernoulli_nll_fwd", "davi_ln_gelu_fwd", "davi_ln_gelu_bwd",
                 "davi_is_logsumexp", "davi_version", "davi_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol(libdavi):
    for name in _declared():
        assert hasattr(libdavi, name), f"{name} declared in davi.h but not exported"


def test_ctypes_table_matches_header(libdavi):
    from deep_attentive_vi_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    # argument counts: count commas of each prototype in the header
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        body = m.group(1).strip()
        n = 0 if body in ("", "void") else body.count(",") + 1
        assert n == len(args), f"{name}: header has {n} args, ctypes table {len(args)}"


def test_version_and_error_text(libdavi):
    assert libdavi.davi_version() == 200
    rc = libdavi.davi_dw_attn_fwd(None, None, None, None, None, 1, 1, 1, 4, 4,
                                  4, 4, 4, 4, 4, 4, 4, 4, None, None)
    assert rc == -1  # DAVI_EINVAL, nothing launched
    buf = ctypes.create_string_buffer(256)
    libdavi.davi_last_error(buf, 256)
    assert b"davi_dw_attn_fwd" in buf.value


def test_options_are_set_through_the_abi_not_the_environment(libdavi):
    """Implementation switches are explicit calls (atomics), never getenv."""
    from deep_attentive_vi_b200 import _lib
    assert libdavi.davi_get_option(_lib.OPT_DW_IMPL) == 0
    assert libdavi.davi_set_option(_lib.OPT_DW_IMPL, 1) == 0 and libdavi.davi_get_option(_lib.OPT_DW_IMPL) == 1
    assert libdavi.davi_set_option(_lib.OPT_DW_IMPL, 0) == 0
    assert libdavi.davi_set_option(99, 1) == -1
    csrc = os.path.join(ROOT, "deep_attentive_vi_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_validation_rejects_bad_shapes(libdavi):
    one = ctypes.c_void_p(16)  # aligned dummy, never dereferenced: validation comes first
    rc = libdavi.davi_nonlocal_attn_fwd(one, one, one, one, one, one, None, 1, 250, 32, 64,
                                        32, 32, 64, 64, 64, 64, None, one, None)
    assert rc == -1
    rc = libdavi.davi_dw_attn_fwd(one, one, one, one, None, 1, 33, 16, 20, 276,
                                  20, 1, 1, 20, 1, 1, 276, 276, None, None)
    assert rc == -1
    rc = libdavi.davi_ln_gelu_bwd(one, one, one, one, one, one, one, 8, 64, 64, 64, 64,
                                  1e-5, 1, None, 0, None)
    assert rc == -2  # DAVI_EWORKSPACE


def test_weight_image_entry_points_validate_before_touching_the_device(libdavi):
    """davi_tf32_lo / davi_conv2d_fwd_pre / davi_conv2d_wt_batched: null, aliased or misaligned arrays and bad
    activation codes are refused with DAVI_EINVAL (-1) and a message; an empty array is a no-op."""
    a, b, odd = ctypes.c_void_p(1024), ctypes.c_void_p(2048), ctypes.c_void_p(1028)
    assert libdavi.davi_tf32_lo(a, a, 16, None) == -1              # in place: the lo image is a separate array
    assert libdavi.davi_tf32_lo(a, None, 16, None) == -1
    assert libdavi.davi_tf32_lo(a, odd, 16, None) == -1            # 16-byte alignment
    assert libdavi.davi_tf32_lo(a, b, -1, None) == -1
    assert libdavi.davi_tf32_lo(a, b, 0, None) == 0                # nothing to do, nothing launched
    conv = lambda w_lo, act=0: libdavi.davi_conv2d_fwd_pre(a, b, w_lo, None, a, 1, 16, 16, 32, 32, 3, 3, 32, 32, act,
                                                           a, 1 << 20, None)
    assert conv(b) == -1                                           # w_lo must not alias w
    assert conv(None) == -1
    assert conv(ctypes.c_void_p(4096), act=2) == -1
    buf = ctypes.create_string_buffer(256)
    libdavi.davi_last_error(buf, 256)
    assert b"davi_conv2d_fwd_pre" in buf.value or b"act" in buf.value
    assert libdavi.davi_conv2d_wt_batched(a, a, a, a, a, a, a, 1, 1, None) == -1      # transposes go to a second array
    assert libdavi.davi_conv2d_wt_batched(a, b, a, a, a, a, a, 0, 0, None) == -1


def test_ops_refuse_cpu_tensors(libdavi):
    import pytest
    import torch
    from deep_attentive_vi_b200 import ops, _lib
    with pytest.raises(_lib.DaviError):
        ops.dw_attn(torch.zeros(1, 2, 2, 4), torch.zeros(1, 1, 2, 2, 4), torch.zeros(1, 1, 2, 2, 4))
    with pytest.raises(_lib.DaviError):
        ops.dlm_nll(torch.zeros(1, 2, 2, 50), torch.zeros(1, 2, 2, 3))


def test_host_convolution_refuses_cpu_tensors(libdavi):
    """ops.conv2d_bias is part of the GPU path like every other op: no silent CPU computation."""
    import pytest
    import torch
    from deep_attentive_vi_b200 import ops, _lib
    with pytest.raises(_lib.DaviError):
        ops.conv2d_bias(torch.zeros(1, 4, 4, 4), torch.zeros(4, 4, 1, 1), None)
