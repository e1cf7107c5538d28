"""TensorFlow side of the drop-in: loads the custom-op library, registers the hand-written gradients and
swaps the bodies of the reference's hot-path methods (STATUS: TensorFlow is not installed in the build
container or on the GPU boxes, so the ops themselves cannot run here.  tests/test_tf_shim.py executes THIS
file's patched bodies on the float64 TensorFlow stand-in of tests/golden/tf_shim.py, with an op module that
answers from the oracle, against the reference's own unpatched bodies; the C++ op kernels are type-checked
against include/davi.h in the same test file; the arithmetic behind the C ABI is what every GPU test runs).

Usage inside the reference tree (src/ on sys.path, TF 2.x with a CUDA build):

    import davi_tf
    davi_tf.install()            # tf.load_op_library + gradient registration + the five patches below

`install()` replaces, in the reference's own classes, exactly the bodies the hot path covers:
  layers.attention.DepthWiseAttention.apply        -> DaviDwAttn        (att.py:135-164)
  layers.attention.NonlocalResNetBlock.call        -> DaviNonlocalAttn  (att.py:285-398; subsample /
                                                     use_layer_norm configurations keep the original body)
  GaussianDiag.sample / .log_prob + the KL site    -> DaviGaussKl       (gd.py:121-263, vl.py:511-542)
  DiscretizedLogisticMixture.log_prob              -> DaviDlmNll        (dlm.py:161-211,311-383)
  Bernoulli.log_prob                               -> DaviBernoulliNll  (bern.py:155-179)
hparams, constructor signatures and weights are untouched, so README command lines and checkpoints keep
working; `uninstall()` restores the original bodies.
"""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None          # the loaded op module (tf.load_op_library result, or a stand-in passed to install())
_tf = None
_saved = {}


def load(tf=None, mod=None):
    """Loads libdavi_tf.so (unless an op module is passed in) and registers the gradient functions."""
    global _mod, _tf
    if tf is None:
        import tensorflow as tf
    _tf = tf
    if mod is None:
        mod = tf.load_op_library(os.path.join(_HERE, "libdavi_tf.so"))
        _register_gradients()
    _mod = mod
    return mod


def _register_gradients():
    from tensorflow.python.framework import ops as tf_ops
    tf = _tf

    @tf_ops.RegisterGradient("DaviDwAttn")
    def _dw_attn_grad(op, dout):
        q, k, v, s = op.inputs
        return _mod.davi_dw_attn_grad(query=q, keys=k, values=v, scale=s, dout=dout)

    @tf_ops.RegisterGradient("DaviNonlocalAttn")
    def _nonlocal_grad(op, dy, _dlse, _do):
        q, k, v, x, scale, gamma = op.inputs
        dq, dk, dv, dgs = _mod.davi_nonlocal_attn_grad(q=q, k=k, v=v, lse=op.outputs[1], scale=scale,
                                                       gamma=gamma, dy=dy, o=op.outputs[2])
        return dq, dk, dv, dy, dgs[1], dgs[0]          # d x == dy (residual)

    @tf_ops.RegisterGradient("DaviGaussKl")
    def _gauss_kl_grad(op, dxi, dkl, _dlq, _dlp, _dgrad):
        # the forward launch already emitted the whole gradient (output 4): one multiply-add pass
        dpost, dprior = _mod.davi_gauss_kl_grad(grad=op.outputs[4], dxi=dxi, dkl=dkl,
                                                has_prior=op.get_attr("has_prior"))
        return dpost, dprior, None, None, None

    @tf_ops.RegisterGradient("DaviDlmNll")
    def _dlm_grad(op, dnll, _unused):
        g = op.outputs[1]                                # per-image d nll / d params from the forward launch
        return g * tf.reshape(dnll, [-1, 1, 1, 1]), None

    @tf_ops.RegisterGradient("DaviBernoulliNll")
    def _bern_grad(op, dnll, _unused):
        return op.outputs[1] * tf.reshape(dnll, [-1, 1, 1, 1]), None

    @tf_ops.RegisterGradient("DaviLnGelu")
    def _ln_gelu_grad(op, dy):
        x, gamma, beta = op.inputs
        return _mod.davi_ln_gelu_grad(x=x, gamma=gamma, beta=beta, dy=dy, mode=op.get_attr("mode"),
                                      epsilon=op.get_attr("epsilon"))

    @tf_ops.RegisterGradient("DaviConv2d")
    def _conv2d_grad(op, dy):
        x, w, _b = op.inputs
        dx = _mod.davi_conv2d_input_grad(dy=dy, kernel_ohwi=w)
        dw = _mod.davi_conv2d_filter_grad(x=x, dy=dy, kh=int(w.shape[1]), kw=int(w.shape[2]))
        return dx, dw, tf.reduce_sum(dy, axis=[0, 1, 2])

    @tf_ops.RegisterGradient("DaviConv2dPre")
    def _conv2d_pre_grad(op, dy):
        # kernel_lo is a function of the kernel (DaviTf32Lo) that carries no gradient of its own: the filter gradient is
        # the one of the plain convolution.  (swish_input: the caller multiplies dx by swish'(x), like ops._ConvTC.)
        x, w, _w_lo, _b = op.inputs
        dx = _mod.davi_conv2d_input_grad(dy=dy, kernel_ohwi=w)
        if op.get_attr("swish_input"):
            sig = tf.sigmoid(x)
            dx = dx * (sig * (1.0 + x * (1.0 - sig)))
            x = x * sig
        dw = _mod.davi_conv2d_filter_grad(x=x, dy=dy, kh=int(w.shape[1]), kw=int(w.shape[2]))
        return dx, dw, None, tf.reduce_sum(dy, axis=[0, 1, 2])

    @tf_ops.RegisterGradient("DaviLnGeluRes")
    def _ln_gelu_res_grad(op, dy):
        x, gamma, beta, base, res_gamma = op.inputs
        dx, dgamma, dbeta, dres = _mod.davi_ln_gelu_res_grad(x=x, gamma=gamma, beta=beta, res_gamma=res_gamma,
                                                             dy=dy, epsilon=op.get_attr("epsilon"))
        return dx, dgamma, dbeta, dy, dres               # d base == dy


# ---------------------------------------------------------------------------------------------------------
# patched bodies (each keeps the signature and the side effects of the reference method it replaces)
# ---------------------------------------------------------------------------------------------------------
def _alignment_scale(af):
    """SoftmaxAlignment creates `scale` lazily in build() (att.py:84-96)."""
    tf = _tf
    if not getattr(af, "_built", True):
        af.build()
    scale = getattr(af, "scale", None)
    return tf.constant(1.0, tf.float32) if scale is None else scale


def depthwise_attention_apply(self, query, keys, values):
    """Body for layers.attention.DepthWiseAttention.apply (att.py:135-164): query [B,W,H,d],
    keys [B,n,W,H,d], values [B,n,W,H,C] -> [B,W,H,C]."""
    return _mod.davi_dw_attn(query=query, keys=keys, values=values, scale=_alignment_scale(self._alignment_function))


def nonlocal_block_call(self, inputs):
    """Body for layers.attention.NonlocalResNetBlock.call (att.py:285-398): the three 1x1 projections stay the
    layer's own Keras convolutions (NHWC, no Permute / reshape copies), everything between them and the
    gamma-residual is one fused op.  Configurations the kernel does not cover keep the original body."""
    if self._subsample or self._nonlocop_hparams.use_layer_norm:
        return _saved["NonlocalResNetBlock.call"](self, inputs)
    q = self.query_conv(inputs)
    k = self.key_conv(inputs)
    v = self.value_conv(inputs)
    y, _, _ = _mod.davi_nonlocal_attn(q=q, k=k, v=v, x=inputs, scale=_alignment_scale(self._alignment_function),
                                      gamma=self.gamma)
    return y


class _FusedKl(object):
    """What GaussianDiag._p[-1] becomes after a fused sample: carries the op's per-image results to the two
    places variational_layer.py reads them (tfp kl_divergence, vl.py:540; GaussianDiag.log_prob, vl.py:517)."""

    def __init__(self, dist, kl, logq, logp):
        self._dist, self.davi_kl, self.davi_logq, self.davi_logp = dist, kl, logq, logp

    def __getattr__(self, name):                 # everything else (sample, log_prob, mean ...) is the TFP object's
        return getattr(self._dist, name)


def gaussian_sample(self, condition, training=True, num_samples=1, initial_params=None):
    """Body for GaussianDiag.sample (gd.py:224-240).  The residual posterior (initial_params = raw prior
    parameters, vl.py:511) with one sample per image is the fused path: bounds, residual, reparameterised
    sample, KL and both log-probabilities come from ONE launch; the distribution objects the reference
    builds in call() are kept (they cost nothing unless used) so that every other consumer still works."""
    tf = _tf
    self.call(condition=condition, training=training, initial_params=initial_params)
    if num_samples != 1:
        return self._p[-1].sample(sample_shape=tf.cast([num_samples], tf.int32))
    hp = self._hparams
    has_prior = initial_params is not None
    # raw posterior increments of this call: (delta mu, delta log sigma) incl. the train-time noise (gd.py:158-169)
    post = tf.concat([self._delta_mu[-1], self._delta_sigma_diag[-1]], axis=-1)
    prior = initial_params if has_prior else tf.zeros_like(post)
    eps = tf.random.normal(tf.shape(self._delta_mu[-1]), dtype=post.dtype)
    # the prior's own bounds are its hparams'; the reference bounds the re-used raw prior parameters of the
    # POSTERIOR object with the posterior's bounds, but the KL is taken against prior_network._p[0], which was
    # bounded with the prior's (vl.py:533-540) -- the op takes both pairs
    # (gaussian_params below tags the prior's raw parameters with the prior's bounds)
    plo, phi = getattr(initial_params, "_davi_bounds", (hp.log_scale_low_bound, hp.log_scale_upper_bound))
    none = tf.zeros([0], post.dtype)             # the GaussianNoise of call() is already inside the raw parameters
    xi, kl, logq, logp, _ = _mod.davi_gauss_kl(post=post, prior=prior, eps=eps, noise_post=none, noise_prior=none,
                                               post_lo=hp.log_scale_low_bound, post_hi=hp.log_scale_upper_bound,
                                               prior_lo=plo, prior_hi=phi, has_prior=has_prior)
    self._p[-1] = _FusedKl(self._p[-1], kl, logq, logp)
    return tf.expand_dims(xi, 0)                 # [num_samples = 1, B, H, W, z] like MVNDiag.sample([1])


def gaussian_params(self, network=-1):
    """Body for GaussianDiag.params (gd.py:266-269): the same tensor, tagged with the bounds of the distribution
    that owns it -- the posterior's fused launch needs the PRIOR's bounds for the prior side of the KL."""
    out = _saved["GaussianDiag.params"](self, network)
    try:
        out._davi_bounds = (self._hparams.log_scale_low_bound, self._hparams.log_scale_upper_bound)
    except AttributeError:                       # a tensor type without instance attributes: same-bounds default
        pass
    return out


def gaussian_log_prob(self, x, condition, training=True, initial_params=None, compute_dist=True):
    """Body for GaussianDiag.log_prob (gd.py:242-263): log q of the sample the fused launch just drew."""
    p = self._p[-1] if not compute_dist and getattr(self, "_p", None) else None
    if isinstance(p, _FusedKl):
        return p.davi_logq
    return _saved["GaussianDiag.log_prob"](self, x, condition, training=training, initial_params=initial_params,
                                           compute_dist=compute_dist)


def kl_divergence(q, p, *args, **kwargs):
    """Stands in for tfp.distributions.kl_divergence at vl.py:540: [B] from the fused launch, shaped [B,1,...] so
    that the reference's reduce_sum over the map axes (vl.py:542) returns it unchanged."""
    tf = _tf
    if isinstance(q, _FusedKl):
        rank = len(q._dist.batch_shape)          # [B, H, W]
        return tf.reshape(q.davi_kl, [-1] + [1] * (rank - 1))
    return _saved["kl_divergence"](q, p, *args, **kwargs)


def dlm_log_prob(self, x, condition, training=True, crop_size=0):
    """Body for DiscretizedLogisticMixture.log_prob (dlm.py:311-383): the parameter network is the layer's own,
    the reshape scramble of create_params (dlm.py:161-211) is resolved inside the kernel."""
    tf = _tf
    self.call(inputs=condition, training=training)
    shape = [int(s) for s in self._shape]
    if self._hparams.num_logistic_mix != 5 or len(shape) != 3 or shape[-1] != 3:
        return _saved["DiscretizedLogisticMixture.log_prob"](self, x, condition, training=training, crop_size=crop_size)
    x = tf.reshape(x, [-1] + shape)
    nll, _ = _mod.davi_dlm_nll(params=self._params, x=x, log_scale_low_bound=self._hparams.log_scale_low_bound)
    return -nll


def bernoulli_log_prob(self, x, condition, training=True, crop_size=0):
    """Body for Bernoulli.log_prob (bern.py:155-179)."""
    tf = _tf
    x = tf.reshape(x, [-1] + [int(s) for s in self._shape])
    self.call(inputs=condition, training=training)
    nll, _ = _mod.davi_bernoulli_nll(logits=self._params, x=x, crop=int(crop_size))
    return -nll


def install(tf=None, mod=None):
    """Monkey-patch the reference classes (run once, before building DeepVAE).  `mod` (tests): an object with the
    op functions, used in place of tf.load_op_library."""
    load(tf, mod)
    from layers import attention as att
    from model import variational_layer as vl
    from stat_tools import bernoulli as bern
    from stat_tools import discretized_logistic_mixture as dlm
    from stat_tools import gaussian_diag as gd

    def swap(owner, name, fn, key):
        if key not in _saved:
            _saved[key] = getattr(owner, name)
        setattr(owner, name, fn)

    swap(att.DepthWiseAttention, "apply", depthwise_attention_apply, "DepthWiseAttention.apply")
    swap(att.NonlocalResNetBlock, "call", nonlocal_block_call, "NonlocalResNetBlock.call")
    swap(gd.GaussianDiag, "sample", gaussian_sample, "GaussianDiag.sample")
    swap(gd.GaussianDiag, "log_prob", gaussian_log_prob, "GaussianDiag.log_prob")
    swap(gd.GaussianDiag, "params", gaussian_params, "GaussianDiag.params")
    swap(vl.tfp.distributions, "kl_divergence", kl_divergence, "kl_divergence")
    swap(dlm.DiscretizedLogisticMixture, "log_prob", dlm_log_prob, "DiscretizedLogisticMixture.log_prob")
    swap(bern.Bernoulli, "log_prob", bernoulli_log_prob, "Bernoulli.log_prob")
    _saved["_owners"] = {"DepthWiseAttention.apply": (att.DepthWiseAttention, "apply"),
                         "NonlocalResNetBlock.call": (att.NonlocalResNetBlock, "call"),
                         "GaussianDiag.sample": (gd.GaussianDiag, "sample"),
                         "GaussianDiag.log_prob": (gd.GaussianDiag, "log_prob"),
                         "GaussianDiag.params": (gd.GaussianDiag, "params"),
                         "kl_divergence": (vl.tfp.distributions, "kl_divergence"),
                         "DiscretizedLogisticMixture.log_prob": (dlm.DiscretizedLogisticMixture, "log_prob"),
                         "Bernoulli.log_prob": (bern.Bernoulli, "log_prob")}


def uninstall():
    for key, (owner, name) in _saved.pop("_owners", {}).items():
        setattr(owner, name, _saved.pop(key))
