"""The TensorFlow side of the drop-in (tf_ops/), checked without TensorFlow:

* tf_ops/davi_tf_ops.cc is compiled with -fsyntax-only against the stand-in op-kernel headers of tf_ops/tf_stub/,
  so every C-ABI call site of the op kernels is type-checked against include/davi.h (and a deliberately broken
  copy must fail, which shows that the check has teeth);
* tf_ops/davi_tf.py's patched method bodies are EXECUTED inside the reference's own model code on the float64
  TensorFlow stand-in of tests/golden/tf_shim.py, with an op module that answers from the oracle: the patched
  model must reproduce the unpatched reference model (needs /root/reference, i.e. the build container).
CPU only."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_CC = os.path.join(ROOT, "tf_ops", "davi_tf_ops.cc")
SYNTAX = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-comment", "-Werror",
          "-I", os.path.join(ROOT, "tf_ops", "tf_stub"), "-I", os.path.join(ROOT, "include")]


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_op_kernels_type_check_against_the_c_abi(tmp_path):
    res = subprocess.run(SYNTAX + [SHIM_CC], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # every entry point the shim forwards to is declared in include/davi.h
    src = open(SHIM_CC).read()
    called = set(re.findall(r"\b(davi_[a-z0-9_]+)\(", src))
    declared = set(re.findall(r"\b(davi_[a-z0-9_]+)\(", open(os.path.join(ROOT, "include", "davi.h")).read()))
    assert called and called <= declared, called - declared
    assert {"davi_dw_attn_fwd", "davi_dw_attn_bwd", "davi_nonlocal_attn_fwd", "davi_nonlocal_attn_bwd",
            "davi_gauss_kl_fwd", "davi_gauss_kl_bwd_saved", "davi_dlm_nll_fwd", "davi_bernoulli_nll_fwd",
            "davi_ln_gelu_fwd", "davi_ln_gelu_bwd", "davi_ln_gelu_res_fwd", "davi_ln_gelu_res_bwd"} <= called
    # the check has teeth: dropping one argument of one call, or passing a const pointer as an output, fails
    broken = src.replace("P(xi), P(kl), P(lq), P(lp), B, Pn, z,", "P(xi), P(kl), P(lq), P(lp), B, Pn,", 1)
    assert broken != src
    bad = tmp_path / "broken_args.cc"
    bad.write_text(broken)
    assert subprocess.run(SYNTAX + [str(bad)], capture_output=True).returncode != 0
    broken = src.replace("davi_dlm_nll_fwd(P(p), P(x), P(nll), P(dp)", "davi_dlm_nll_fwd(P(p), P(x), P(nll), P(p)", 1)
    assert broken != src
    bad = tmp_path / "broken_const.cc"
    bad.write_text(broken)
    assert subprocess.run(SYNTAX + [str(bad)], capture_output=True).returncode != 0


PATCHED_MODEL = r'''
import os, sys, types
import numpy as np, torch
ROOT = sys.argv[1]
sys.path[:0] = [os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "tf_ops")]
import tf_shim
tf = tf_shim.install()
REF = "/root/reference/src"
sys.path.insert(0, REF)
_u = types.ModuleType("util"); _u.__path__ = [os.path.join(REF, "util")]; sys.modules["util"] = _u
import helpers, make_golden
from oracle import ops as R
import davi_tf

T = tf_shim.T
calls = {}
def count(name):
    calls[name] = calls.get(name, 0) + 1
def w(t):
    return t.as_subclass(T)

class OracleOps:
    """The op module tf.load_op_library would return, answering from the CPU oracle (float64)."""
    @staticmethod
    def davi_dw_attn(query, keys, values, scale):
        count("dw"); return w(R.depthwise_attention_apply(query, keys, values, scale))
    @staticmethod
    def davi_nonlocal_attn(q, k, v, x, scale, gamma):
        count("nl"); y = R.nonlocal_core(q, k, v, x, gamma, scale); return w(y), None, None
    @staticmethod
    def davi_gauss_kl(post, prior, eps, noise_post, noise_prior, post_lo, post_hi, prior_lo, prior_hi, has_prior):
        count("kl")
        assert noise_post.numel() == 0 and noise_prior.numel() == 0
        xi, kl, lq, lp = R.variational_sample_kl(post, eps, prior if has_prior else None, (post_lo, post_hi),
                                                 (prior_lo, prior_hi), compute_logp=True)
        return w(xi), w(kl), w(lq), w(lp), None
    @staticmethod
    def davi_dlm_nll(params, x, log_scale_low_bound):
        count("dlm"); return w(-R.dlm_log_prob(params, x, ls_low=log_scale_low_bound)), None
    @staticmethod
    def davi_bernoulli_nll(logits, x, crop):
        count("bern"); return w(-R.bernoulli_log_prob(logits, x, crop)), None

def run(flags, seed):
    torch.manual_seed(seed); tf_shim.REC.gen.manual_seed(seed)
    m = make_golden.build_reference_model(flags)
    x = helpers.synthetic_images({"image_info": make_golden.IMAGE_INFO[flags["dataset"]]}, 2, seed=seed).double()
    m(x, training=True)
    make_golden.randomize(m, seed + 1)
    out = []
    for training in (True, False):
        tf_shim.REC.gen.manual_seed(seed + 7); tf_shim.REC.training_noise = training
        out += [t.detach().clone() for t in m(x, training=training)]
    tf_shim.REC.gen.manual_seed(seed + 9); tf_shim.REC.training_noise = False
    out.append(m.infer_worker_body(x, 3).detach().clone())          # importance-sampled chunk: log q / log p path
    return out

for name, flags, seed in (("cifar", helpers.SMALL_CIFAR, 11), ("omniglot", helpers.SMALL_OMNIGLOT, 12)):
    run(flags, seed)              # the stand-in's first model build consumes global state: runs 2, 3, ... repeat exactly
    ref = run(flags, seed)
    assert all(torch.equal(a, b) for a, b in zip(ref, run(flags, seed)))
    davi_tf.install(tf=tf, mod=OracleOps)
    calls.clear()
    got = run(flags, seed)
    davi_tf.uninstall()
    assert calls.get("dw", 0) > 0 and calls.get("nl", 0) > 0 and calls.get("kl", 0) > 0, calls
    assert calls.get("dlm", 0) + calls.get("bern", 0) > 0, calls
    for a, b in zip(ref, got):
        err = (a - b).abs().max().item() / max(1.0, a.abs().max().item())
        assert err < 1e-9, (name, err)
    again = run(flags, seed)                                         # uninstall() restored the reference bodies
    assert all(torch.equal(a, b) for a, b in zip(ref, again))
    print(name, "ok", dict(calls))
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree only exists in the build container")
def test_patched_reference_model_reproduces_the_reference_on_the_tf_stand_in(tmp_path):
    script = tmp_path / "patched_model.py"
    script.write_text(PATCHED_MODEL)
    res = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "cifar ok" in res.stdout and "omniglot ok" in res.stdout
