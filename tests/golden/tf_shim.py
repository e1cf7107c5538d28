"""Minimal eager TensorFlow / Keras / TFP / tfa stand-in, backed by torch CPU float64.

Purpose: the reference (ifiaposto/Deep_Attentive_VI, /root/reference) is pure Python on
TensorFlow 2.3, which is not installable here.  `install()` registers just enough of the
`tensorflow`, `tensorflow_probability` and `tensorflow_addons` module trees in sys.modules
for the reference's OWN source files (src/layers, src/stat_tools, src/model) to import and
run unmodified.  The golden vectors under tests/golden/ are produced by executing that
code through this shim (make_golden.py), so they pin the reference's Python logic --
reshapes, channel maps, split/stack order, layer-norm placement, residual wiring --
while the library arithmetic underneath (conv, batch-norm, softmax, MVN KL ...) is this
file's restatement of the documented TF 2.3 / TFP 0.11 / tfa 0.13 semantics in float64.

Test infrastructure only: nothing in deep_attentive_vi_b200/ imports this file.
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


# ---------------------------------------------------------------------------
# tensors and shapes
# ---------------------------------------------------------------------------
class TensorShape(tuple):
    def __new__(cls, dims=()):
        if isinstance(dims, TensorShape):
            return dims
        if dims is None:
            dims = ()
        return super().__new__(cls, tuple(None if d is None else int(d) for d in dims))

    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return len(self)

    ndims = rank

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TensorShape(r) if isinstance(i, slice) else r


class T(torch.Tensor):
    """torch tensor whose .shape speaks TensorFlow (as_list / rank)."""

    @property
    def shape(self):
        return TensorShape(torch.Tensor.size(self))

    def get_shape(self):
        return self.shape

    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))

    def assign(self, v):
        with torch.no_grad():
            self.copy_(_t(v))
        return self


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        r = x
    elif isinstance(x, (bool, np.bool_)):
        r = torch.tensor(bool(x))
    elif isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        r = torch.stack([_t(v) for v in x])
    else:
        r = torch.as_tensor(np.asarray(x))
    if r.is_floating_point() and r.dtype != DT:
        r = r.to(DT)
    if dtype is not None and dtype in (torch.int32, torch.int64, torch.bool):
        r = r.to(dtype)
    return r.as_subclass(T)


def _ints(v):
    if isinstance(v, torch.Tensor):
        return [int(i) for i in v.reshape(-1).tolist()]
    out = []
    for i in v:
        out.append(int(i) if not isinstance(i, (list, tuple, torch.Tensor)) else _ints(i))
    return out


def _flat_ints(v):
    """shape arguments arrive as nested mixes of ints, lists, TensorShapes and int tensors"""
    out = []
    if isinstance(v, torch.Tensor):
        return [int(i) for i in v.reshape(-1).tolist()]
    if isinstance(v, (int, np.integer)):
        return [int(v)]
    for i in v:
        out.extend(_flat_ints(i))
    return out


def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, torch.Tensor):
        axis = axis.tolist()
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


# ---------------------------------------------------------------------------
# random draws are recorded so that the product under test can be fed the same ones
# ---------------------------------------------------------------------------
class Recorder:
    def __init__(self):
        self.gen = torch.Generator().manual_seed(1234)
        self.normal_draws = []        # (tag, tensor) in call order
        self.training_noise = True    # GaussianNoise active (Keras learning phase)

    def randn(self, shape, tag):
        z = torch.randn(*shape, generator=self.gen, dtype=DT)
        self.normal_draws.append((tag, z))
        return z.as_subclass(T)


REC = Recorder()


# ---------------------------------------------------------------------------
# tf.* functions
# ---------------------------------------------------------------------------
class _NameScope:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def convert_to_tensor(x, dtype=None, name=None, **kw):
    return _t(x, dtype if isinstance(dtype, torch.dtype) else None)


def reshape(x, shape, name=None):
    return _t(x).reshape(_flat_ints(shape)).as_subclass(T)


def shape(x, **kw):
    return list(_t(x).size())


def reduce_sum(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return (x.sum() if axis is None else x.sum(dim=_axis(axis), keepdim=keepdims)).as_subclass(T)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return (x.mean() if axis is None else x.mean(dim=_axis(axis), keepdim=keepdims)).as_subclass(T)


def reduce_max(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return (x.max() if axis is None else x.amax(dim=_axis(axis), keepdim=keepdims)).as_subclass(T)


def reduce_logsumexp(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    if axis is None:
        return torch.logsumexp(x.reshape(-1), 0).as_subclass(T)
    return torch.logsumexp(x, dim=_axis(axis), keepdim=keepdims).as_subclass(T)


def concat(values, axis, name=None):
    if isinstance(values, torch.Tensor):          # array_ops.concat: a single tensor is wrapped in a list -> identity
        return _t(values)
    return torch.cat([_t(v) for v in values], dim=int(axis)).as_subclass(T)


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=int(axis)).as_subclass(T)


def unstack(value, num=None, axis=0, name=None):
    return [v.as_subclass(T) for v in torch.unbind(_t(value), dim=int(axis))]


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    value = _t(value)
    if isinstance(num_or_size_splits, (int, np.integer)):
        return [v.as_subclass(T) for v in torch.chunk(value, int(num_or_size_splits), dim=int(axis))]
    return [v.as_subclass(T) for v in torch.split(value, _flat_ints(num_or_size_splits), dim=int(axis))]


def tile(x, multiples, name=None):
    return _t(x).repeat(*_flat_ints(multiples)).as_subclass(T)


def transpose(x, perm=None, **kw):
    x = _t(x)
    return (x.permute(*_flat_ints(perm)) if perm is not None else x.t()).as_subclass(T)


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(int(axis)).as_subclass(T)


def squeeze(x, axis=None, name=None):
    x = _t(x)
    return (x.squeeze() if axis is None else x.squeeze(int(axis))).as_subclass(T)


def where(cond, x=None, y=None, name=None):
    return torch.where(_t(cond), _t(x), _t(y)).as_subclass(T)


def cast(x, dtype, name=None):
    if isinstance(x, (list, tuple)) and dtype in (torch.int32, torch.int64):
        return _flat_ints(x)
    x = _t(x)
    if isinstance(dtype, str):
        dtype = DT
    return (x.to(DT) if dtype.is_floating_point else x.to(dtype)).as_subclass(T)


def zeros(shape, dtype=None, name=None):
    return torch.zeros(_flat_ints(shape), dtype=DT).as_subclass(T)


def zeros_like(x, **kw):
    return torch.zeros_like(_t(x)).as_subclass(T)


def constant(v, dtype=None, **kw):
    return _t(v)


def maximum(a, b, name=None):
    return torch.maximum(_t(a), _t(b)).as_subclass(T)


def minimum(a, b, name=None):
    return torch.minimum(_t(a), _t(b)).as_subclass(T)


def _un(fn):
    return lambda x, name=None: fn(_t(x)).as_subclass(T)


def _bin(fn):
    return lambda a, b, name=None: fn(_t(a), _t(b)).as_subclass(T)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None, **kw):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b).as_subclass(T)


def matvec(a, b, name=None, **kw):
    return torch.matmul(_t(a), _t(b).unsqueeze(-1)).squeeze(-1).as_subclass(T)


def l2_normalize(x, axis=None, epsilon=1e-12, name=None):
    x = _t(x)
    ss = x.square().sum(dim=_axis(axis), keepdim=True)
    return (x * torch.rsqrt(torch.clamp(ss, min=epsilon))).as_subclass(T)      # tf.nn.l2_normalize


def moments(x, axes, keepdims=False, **kw):
    x = _t(x)
    m = x.mean(dim=_axis(axes), keepdim=keepdims)
    v = x.var(dim=_axis(axes), unbiased=False, keepdim=keepdims)
    return m.as_subclass(T), v.as_subclass(T)


def norm(x, axis=None, **kw):
    return torch.linalg.vector_norm(_t(x), dim=_axis(axis)).as_subclass(T)


def one_hot(idx, depth, axis=-1, dtype=None, **kw):
    return F.one_hot(_t(idx).long(), int(depth)).to(DT).as_subclass(T)


def argmax(x, axis=None, **kw):
    return torch.argmax(_t(x), dim=_axis(axis)).as_subclass(T)


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=None, seed=None, name=None):
    u = torch.rand(_flat_ints(shape), generator=REC.gen, dtype=DT)
    return (minval + (maxval - minval) * u).as_subclass(T)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    """tf.random.normal, drawn from the recorder's stream like every other normal draw (used by the
    TF-side drop-in tf_ops/davi_tf.py, not by the reference)."""
    return (mean + stddev * REC.randn(tuple(_flat_ints(shape)), "random_normal")).as_subclass(T)


def add_n(xs, name=None):
    r = _t(xs[0])
    for v in xs[1:]:
        r = r + _t(v)
    return r.as_subclass(T)


def cond(pred, true_fn, false_fn, name=None):
    return true_fn() if bool(pred) else false_fn()


def swish(x):
    x = _t(x)
    return (x * torch.sigmoid(x)).as_subclass(T)


def elu(x, alpha=1.0):
    return F.elu(_t(x), alpha).as_subclass(T)


def gelu(x, approximate=True):
    """tfa.activations.gelu (0.13): approximate=True is the default -> tanh form."""
    x = _t(x)
    if approximate:
        return (0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))).as_subclass(T)
    return (0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))).as_subclass(T)


def _same_pad(n, k, s):
    """TF 'SAME': out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total // 2."""
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def conv2d_nhwc(x, kernel, strides, padding):
    """x [B,H,W,Cin], kernel [kh,kw,Cin,Cout] (TF layout)."""
    x = _t(x).permute(0, 3, 1, 2)
    w = _t(kernel).permute(3, 2, 0, 1)
    kh, kw = w.shape[2], w.shape[3]
    sh, sw = (strides, strides) if isinstance(strides, int) else strides
    if padding.lower() == "same":
        (pt, pb), (pl, pr) = _same_pad(x.shape[2], kh, sh), _same_pad(x.shape[3], kw, sw)
        x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x.as_subclass(torch.Tensor), w.as_subclass(torch.Tensor), None, stride=(sh, sw))
    return y.permute(0, 2, 3, 1).as_subclass(T)


def resize_bilinear(x, size, align_corners=False, **kw):
    y = F.interpolate(_t(x).permute(0, 3, 1, 2).as_subclass(torch.Tensor), size=tuple(_flat_ints(size)),
                      mode="bilinear", align_corners=bool(align_corners))
    return y.permute(0, 2, 3, 1).as_subclass(T)


def resize_nearest(x, size, align_corners=False, **kw):
    assert not align_corners
    # src = floor(dst * in / out): torch 'nearest'
    y = F.interpolate(_t(x).permute(0, 3, 1, 2).as_subclass(torch.Tensor), size=tuple(_flat_ints(size)),
                      mode="nearest")
    return y.permute(0, 2, 3, 1).as_subclass(T)


# ---------------------------------------------------------------------------
# initializers / regularizers
# ---------------------------------------------------------------------------
class VarianceScaling:
    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        self.scale, self.mode, self.distribution, self.seed = scale, mode, distribution, seed

    def __call__(self, shape, dtype=None, **kw):
        shape = _flat_ints(shape)
        if len(shape) < 1:
            fan_in = fan_out = 1
        elif len(shape) == 1:
            fan_in = fan_out = shape[0]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:                                       # Keras _compute_fans for conv kernels
            rf = int(np.prod(shape[:-2]))
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
        s = self.scale / max(1.0, n)
        assert self.distribution == "uniform"
        lim = math.sqrt(3.0 * s)
        u = torch.rand(shape, generator=REC.gen, dtype=DT)
        return ((2.0 * u - 1.0) * lim).as_subclass(T)


class _Ones:
    def __call__(self, shape, dtype=None, **kw):
        return torch.ones(_flat_ints(shape), dtype=DT).as_subclass(T)


class _Zeros:
    def __call__(self, shape, dtype=None, **kw):
        return torch.zeros(_flat_ints(shape), dtype=DT).as_subclass(T)


class _Constant:
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=None, **kw):
        return (torch.zeros(_flat_ints(shape), dtype=DT) + _t(self.value)).as_subclass(T)


class _L2:
    def __init__(self, l=0.01, l2=None):
        self.l2 = l if l2 is None else l2

    def __call__(self, w):
        return self.l2 * _t(w).square().sum()


class _L1:
    def __init__(self, l=0.01, l1=None):
        self.l1 = l if l1 is None else l1

    def __call__(self, w):
        return self.l1 * _t(w).abs().sum()


def _get_regularizer(r):
    if r is None or callable(r):
        return r
    if r == "l2":
        return _L2(0.01)          # Keras default strength for the STRING 'l2'
    if r == "l1":
        return _L1(0.01)
    raise ValueError(r)


def _get_initializer(i):
    if i is None:
        return None
    if callable(i):
        return i
    return {"zeros": _Zeros(), "ones": _Ones()}[i]


# ---------------------------------------------------------------------------
# Keras layers
# ---------------------------------------------------------------------------
def walk_weights(obj, prefix="", seen=None, include_modules=True):
    """All weights reachable from `obj` through attributes / lists / dicts, as
    (attribute_path, name, tensor, regularizer, trainable); every layer is visited once.
    include_modules=False mimics Keras, which does not track variables held by a bare tf.Module."""
    seen = set() if seen is None else seen
    out = []
    if id(obj) in seen:
        return out
    seen.add(id(obj))
    if isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            out += walk_weights(v, f"{prefix}.{i}" if prefix else str(i), seen, include_modules)
        return out
    if isinstance(obj, dict):
        for k, v in obj.items():
            out += walk_weights(v, f"{prefix}.{k}" if prefix else str(k), seen, include_modules)
        return out
    if not isinstance(obj, (Module, Distribution)):
        return out
    if isinstance(obj, Module) and not isinstance(obj, Layer) and not include_modules:
        return out
    for name, v, reg, tr in getattr(obj, "_weights", []):
        out.append((prefix, name, v, reg, tr))
    for k, v in list(vars(obj).items()):
        if k in ("_weights",):
            continue
        if isinstance(v, (Module, Distribution, list, tuple, dict)):
            out += walk_weights(v, f"{prefix}.{k}" if prefix else k, seen, include_modules)
    return out


class Module:
    def __init__(self, name=None):
        self._name = name

    @property
    def name(self):
        return getattr(self, "_name", None) or type(self).__name__

    @property
    def trainable_weights(self):
        return [w for _, _, w, _, tr in walk_weights(self, include_modules=False) if tr]


class Layer(Module):
    def __init__(self, trainable=True, name=None, dtype=None, dynamic=False, **kwargs):
        self._name = name
        self._dtype = dtype
        self.built = False
        self.trainable = trainable
        self._weights = []                      # (name, tensor, regularizer, trainable)
        self.input_spec = None

    @property
    def dtype(self):
        return torch.float32 if getattr(self, "_dtype", None) is None else self._dtype

    def _ensure(self):
        if not hasattr(self, "_weights"):
            Layer.__init__(self)

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True,
                   **kwargs):
        self._ensure()
        if dtype is torch.bool:
            v = torch.zeros(_flat_ints(shape if shape is not None else ()), dtype=torch.bool).as_subclass(T)
        else:
            init = _get_initializer(initializer) or _Zeros()
            v = init(shape if shape is not None else ())
        self._weights.append((name, v, _get_regularizer(regularizer), trainable))
        return v

    def build(self, input_shape=None):
        self.built = True

    def _track_trackable(self, *a, **k):
        pass

    def __call__(self, inputs, *args, **kwargs):
        self._ensure()
        if not self.built:
            if isinstance(inputs, (list, tuple)) and len(inputs) and isinstance(inputs[0], torch.Tensor):
                shp = [TensorShape(i.size()) for i in inputs]
            elif isinstance(inputs, torch.Tensor):
                shp = TensorShape(inputs.size())
            else:
                shp = None
            self.build(shp)
            self.built = True
        if "training" in kwargs:
            # Keras base_layer.__call__: a `training` keyword is dropped when call() does not take one
            # (the reference relies on it: vae.py:468 passes training=False to ConcatConv2D.call(inputs))
            import inspect
            ps = inspect.signature(self.call).parameters
            if "training" not in ps and not any(p.kind == p.VAR_KEYWORD for p in ps.values()):
                kwargs.pop("training")
        return self.call(inputs, *args, **kwargs)

    def call(self, inputs, *a, **k):
        return inputs

    def compute_output_shape(self, s):
        return s


class Model(Layer):
    def __init__(self, name=None, dtype=None, **kw):
        Layer.__init__(self, name=name if isinstance(name, str) else None, dtype=dtype)

    def build(self, input_shape=None):
        self.built = True

    def count_params(self):
        return 0


class Wrapper(Layer):
    def __init__(self, layer, **kwargs):
        Layer.__init__(self, **kwargs)
        self.layer = layer

    def build(self, input_shape=None):
        self.built = True


class Lambda(Layer):
    def __init__(self, function, **kw):
        Layer.__init__(self, **kw)
        self.function = function

    def call(self, inputs, *a, **k):
        return self.function(inputs)


class Sequential(Layer):
    def __init__(self, layers=None, **kw):
        Layer.__init__(self, **kw)
        self.layers = list(layers or [])
        self.built = True

    def call(self, x, training=None, **k):
        import inspect
        for l in self.layers:
            if "training" in inspect.signature(l.call).parameters:
                x = l(x, training=training)
            else:
                x = l(x)
        return x


class Permute(Layer):
    def __init__(self, dims, **kw):
        Layer.__init__(self, **kw)
        self.dims = tuple(dims)

    def call(self, x):
        return _t(x).permute(0, *self.dims).as_subclass(T)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 **kw):
        Layer.__init__(self, **{k: v for k, v in kw.items() if k in ("name", "dtype")})
        self.filters = int(filters)
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.strides = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.padding, self.activation, self.use_bias = padding, activation, use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.kernel = self.bias = None
        self.last_output = None

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (*self.kernel_size, cin, self.filters),
                                      initializer=self.kernel_initializer, regularizer=self.kernel_regularizer)
        if self.use_bias:
            self.bias = self.add_weight("bias", (self.filters,), initializer=self.bias_initializer,
                                        regularizer=self.bias_regularizer)
        self.built = True

    def call(self, x):
        y = conv2d_nhwc(x, self.kernel, self.strides, self.padding)
        if self.bias is not None:
            y = y + _t(self.bias)
        if self.activation is not None:
            y = self.activation(y)
        self.last_output = y
        return y

    def compute_output_shape(self, s):
        return TensorShape(list(s[:-1]) + [self.filters])


class BatchNormalization(Layer):
    """tf.keras.layers.BatchNormalization, last axis; training=True uses the batch's biased variance."""

    def __init__(self, momentum=0.99, epsilon=1e-3, gamma_regularizer=None, beta_regularizer=None, **kw):
        Layer.__init__(self)
        self.momentum, self.epsilon = momentum, epsilon
        self.gr, self.br = gamma_regularizer, beta_regularizer

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (c,), initializer=_Ones(), regularizer=self.gr)
        self.beta = self.add_weight("beta", (c,), initializer=_Zeros(), regularizer=self.br)
        self.moving_mean = self.add_weight("moving_mean", (c,), initializer=_Zeros(), trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (c,), initializer=_Ones(), trainable=False)
        self.built = True

    def call(self, x, training=None):
        x = _t(x)
        if training:
            dims = tuple(range(x.dim() - 1))
            mean, var = x.mean(dim=dims), x.var(dim=dims, unbiased=False)
        else:
            mean, var = self.moving_mean, self.moving_variance
        return ((x - mean) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta).as_subclass(T)


class LayerNormalization(Layer):
    """tf.keras.layers.LayerNormalization(axis=-1): moments with biased variance, rsqrt(var + eps)."""

    def __init__(self, axis=-1, epsilon=1e-3, gamma_regularizer=None, beta_regularizer=None, **kw):
        Layer.__init__(self)
        self.epsilon, self.gr, self.br = epsilon, gamma_regularizer, beta_regularizer

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (c,), initializer=_Ones(), regularizer=self.gr)
        self.beta = self.add_weight("beta", (c,), initializer=_Zeros(), regularizer=self.br)
        self.built = True

    def call(self, x):
        x = _t(x)
        mean = x.mean(dim=-1, keepdim=True)
        var = x.var(dim=-1, unbiased=False, keepdim=True)
        return ((x - mean) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta).as_subclass(T)


class GaussianNoise(Layer):
    def __init__(self, stddev, **kw):
        Layer.__init__(self)
        self.stddev = stddev

    def call(self, x, training=None):
        x = _t(x)
        if not REC.training_noise:
            return x
        return (x + self.stddev * REC.randn(tuple(x.size()), "gaussian_noise")).as_subclass(T)


class MaxPool2D(Layer):
    def __init__(self, pool_size=(2, 2), **kw):
        Layer.__init__(self)
        self.pool = pool_size

    def call(self, x):
        y = F.max_pool2d(_t(x).permute(0, 3, 1, 2).as_subclass(torch.Tensor), self.pool)
        return y.permute(0, 2, 3, 1).as_subclass(T)


class GlobalAveragePooling2D(Layer):
    def call(self, x):
        return _t(x).mean(dim=(1, 2)).as_subclass(T)


class Add(Layer):
    def call(self, xs):
        return add_n(xs)


class Softmax(Layer):
    def __init__(self, axis=-1, **kw):
        Layer.__init__(self)
        self.axis = axis

    def call(self, x):
        return torch.softmax(_t(x), dim=self.axis).as_subclass(T)


class InputSpec:
    def __init__(self, *a, **k):
        pass


class VariableLayer(Layer):
    """tfp.layers.VariableLayer: returns its variable whatever the input."""

    def __init__(self, shape, dtype=None, initializer="zeros", regularizer=None, **kw):
        Layer.__init__(self)
        self._var = self.add_weight("constant", _flat_ints(shape), initializer=initializer, regularizer=regularizer)
        self.built = True

    def call(self, _):
        return self._var


# ---------------------------------------------------------------------------
# tensorflow_probability
# ---------------------------------------------------------------------------
FULLY_REPARAMETERIZED, NOT_REPARAMETERIZED = "FULLY_REPARAMETERIZED", "NOT_REPARAMETERIZED"


class Distribution:
    def __init__(self, dtype=None, reparameterization_type=None, validate_args=False, allow_nan_stats=True,
                 parameters=None, name=None, **kw):
        self._dist_name = name
        self._dtype = dtype
        self.reparameterization_type = reparameterization_type

    @property
    def name(self):
        return getattr(self, "_dist_name", None) or type(self).__name__

    @property
    def trainable_weights(self):
        return [w for _, _, w, _, tr in walk_weights(self, include_modules=False) if tr]


class MultivariateNormalDiag(Distribution):
    def __init__(self, loc=None, scale_diag=None, **kw):
        Distribution.__init__(self, reparameterization_type=FULLY_REPARAMETERIZED, name="MVNDiag")
        self.loc, self.scale = _t(loc), _t(scale_diag)

    @property
    def batch_shape(self):                             # TFP: every axis but the event (last) one
        return TensorShape(torch.broadcast_shapes(self.loc.shape, self.scale.shape)[:-1])

    def sample(self, sample_shape=(), seed=None):
        ss = _flat_ints(sample_shape) if not isinstance(sample_shape, int) else [sample_shape]
        eps = REC.randn(tuple(ss) + tuple(self.loc.size()), "mvn_sample")
        return (self.loc + self.scale * eps).as_subclass(T)

    def log_prob(self, x):
        z = (_t(x) - self.loc) / self.scale
        lp = -0.5 * z.square() - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)
        return lp.sum(dim=-1).as_subclass(T)           # event = last axis


def kl_divergence(q, p, **kw):
    """KL(q || p) of two MultivariateNormalDiag: closed form, summed over the event (last) axis."""
    t = torch.log(p.scale) - torch.log(q.scale) + (q.scale.square() + (q.loc - p.loc).square()) / (
        2.0 * p.scale.square()) - 0.5
    return t.sum(dim=-1).as_subclass(T)


class BernoulliDist(Distribution):
    def __init__(self, logits=None, probs=None, dtype=None, **kw):
        Distribution.__init__(self, name="Bernoulli")
        self.logits = _t(logits)

    def log_prob(self, x):
        x = _t(x)
        return (x * self.logits - F.softplus(self.logits)).as_subclass(T)   # = -sigmoid_cross_entropy_with_logits

    def sample(self, sample_shape=(), seed=None):
        u = torch.rand(tuple(self.logits.size()), generator=REC.gen, dtype=DT)
        return (u < torch.sigmoid(self.logits)).to(DT).as_subclass(T)


# ---------------------------------------------------------------------------
# module tree
# ---------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "__davi_shim__", False):
        return sys.modules["tensorflow"]

    activations = types.SimpleNamespace(
        tanh=_un(torch.tanh), relu=lambda x: F.relu(_t(x)).as_subclass(T), softmax=lambda x: torch.softmax(_t(x), -1),
        sigmoid=_un(torch.sigmoid), softplus=lambda x: F.softplus(_t(x)).as_subclass(T), exponential=_un(torch.exp),
        elu=elu, swish=swish)
    layers = _mod("tensorflow.keras.layers", Layer=Layer, Wrapper=Wrapper, Lambda=Lambda, Permute=Permute,
                  Conv2D=Conv2D, Convolution2D=Conv2D, Conv2DTranspose=Conv2D, Dense=Conv2D,
                  BatchNormalization=BatchNormalization, LayerNormalization=LayerNormalization,
                  GaussianNoise=GaussianNoise, MaxPool2D=MaxPool2D, GlobalAveragePooling2D=GlobalAveragePooling2D,
                  Add=Add, InputSpec=InputSpec, LeakyReLU=lambda *a, **k: None, Softmax=Softmax)
    initializers = _mod("tensorflow.keras.initializers", VarianceScaling=VarianceScaling, Constant=_Constant)
    regularizers = _mod("tensorflow.keras.regularizers", l1=_L1, l2=_L2)
    backend = types.SimpleNamespace(
        count_params=lambda p: int(p.numel()), cast=cast, permute_dimensions=lambda x, pattern: transpose(x, pattern), print_tensor=lambda x, *a, **k: x,
        cast_to_floatx=lambda x: _t(x))
    keras = _mod("tensorflow.keras", layers=layers, activations=activations, initializers=initializers,
                 regularizers=regularizers, backend=backend, Model=Model, Sequential=Sequential)

    math_ns = types.SimpleNamespace(
        log=_un(torch.log), exp=_un(torch.exp), tanh=_un(torch.tanh), multiply=_bin(torch.mul), add=_bin(torch.add),
        subtract=_bin(torch.sub), pow=lambda x, y, name=None: torch.pow(_t(x), y).as_subclass(T),
        reduce_logsumexp=reduce_logsumexp, scalar_mul=lambda s, x, name=None: (s * _t(x)).as_subclass(T),
        not_equal=_bin(torch.ne), equal=_bin(torch.eq), negative=_un(torch.neg), argmax=argmax, add_n=add_n,
        sqrt=_un(torch.sqrt), square=_un(torch.square))
    nn_ns = types.SimpleNamespace(
        softplus=lambda x, name=None: F.softplus(_t(x)).as_subclass(T), sigmoid=_un(torch.sigmoid),
        softmax=lambda x, axis=-1, name=None: torch.softmax(_t(x), axis).as_subclass(T), moments=moments,
        l2_normalize=l2_normalize, swish=swish, elu=elu)
    linalg = types.SimpleNamespace(matmul=matmul, matvec=matvec)
    image_v1 = types.SimpleNamespace(resize_bilinear=resize_bilinear, resize_nearest_neighbor=resize_nearest)
    random_ns = types.SimpleNamespace(uniform=random_uniform, normal=random_normal)
    dtypes_ns = types.SimpleNamespace(cast=cast)

    common = dict(
        float32=torch.float32, float64=torch.float64, int32=torch.int32, int64=torch.int64, bool=torch.bool,
        name_scope=_NameScope, convert_to_tensor=convert_to_tensor, reshape=reshape, shape=shape,
        reduce_sum=reduce_sum, reduce_mean=reduce_mean, reduce_max=reduce_max, concat=concat, stack=stack,
        unstack=unstack, split=split, tile=tile, transpose=transpose, expand_dims=expand_dims, squeeze=squeeze,
        where=where, cast=cast, zeros=zeros, zeros_like=zeros_like, constant=constant, maximum=maximum,
        minimum=minimum, exp=_un(torch.exp), log=_un(torch.log), sqrt=_un(torch.sqrt), square=_un(torch.square),
        norm=norm, one_hot=one_hot, cond=cond, TensorShape=TensorShape, Module=Module,
        zeros_initializer=_Zeros, ones_initializer=_Ones, math=math_ns, nn=nn_ns, linalg=linalg, random=random_ns,
        dtypes=dtypes_ns, keras=keras, Tensor=T,
        dot=lambda a, b: torch.matmul(_t(a), _t(b)).as_subclass(T),
        while_loop=None)

    tf = _mod("tensorflow", __davi_shim__=True, **common)
    v1 = _mod("tensorflow.compat.v1", image=image_v1, log=_un(torch.log), **{k: v for k, v in common.items()
                                                                          if k not in ("log",)})
    tf.compat = _mod("tensorflow.compat", v1=v1, v2=tf)
    sys.modules["tensorflow.compat.v2"] = tf
    tf.compat.v2 = tf
    tf.compat.v1 = v1

    # tensorflow.python.* bits the reference imports directly
    adv = _mod("tensorflow.python.keras.layers.advanced_activations", Softmax=Softmax)
    _mod("tensorflow.python.keras.layers", advanced_activations=adv)
    _mod("tensorflow.python.keras")
    init_ops = _mod("tensorflow.python.ops.init_ops", ones_initializer=_Ones, zeros_initializer=_Zeros)
    init_ops_v2 = _mod("tensorflow.python.ops.init_ops_v2", Constant=_Constant)
    math_ops = _mod("tensorflow.python.ops.math_ops", matmul=matmul)
    nn_mod = _mod("tensorflow.python.ops.nn", swish=swish)
    _mod("tensorflow.python.ops", init_ops=init_ops, init_ops_v2=init_ops_v2, math_ops=math_ops, nn=nn_mod)
    _mod("tensorflow.python.util.tf_export", keras_export=lambda *a, **k: (lambda cls: cls))
    _mod("tensorflow.python.util")
    _mod("tensorflow.python")

    # tensorflow_probability
    tshape_util = _mod("tensorflow_probability.python.internal.tensorshape_util",
                       rank=lambda s: len(s) if not isinstance(s, int) else 0)
    reparam = _mod("tensorflow_probability.python.internal.reparameterization",
                   FULLY_REPARAMETERIZED=FULLY_REPARAMETERIZED, NOT_REPARAMETERIZED=NOT_REPARAMETERIZED)
    prefer_static = _mod("tensorflow_probability.python.internal.prefer_static",
                         concat=lambda vals, axis=0: _flat_ints(vals))
    _mod("tensorflow_probability.python.internal", tensorshape_util=tshape_util, reparameterization=reparam,
         prefer_static=prefer_static)
    dist_mod = _mod("tensorflow_probability.python.distributions.distribution", Distribution=Distribution)
    _mod("tensorflow_probability.python.distributions", distribution=dist_mod)
    _mod("tensorflow_probability.python")
    tfd = types.SimpleNamespace(MultivariateNormalDiag=MultivariateNormalDiag, kl_divergence=kl_divergence,
                                Bernoulli=BernoulliDist, FULLY_REPARAMETERIZED=FULLY_REPARAMETERIZED,
                                Distribution=Distribution)
    _mod("tensorflow_probability", distributions=tfd, bijectors=types.SimpleNamespace(),
         layers=types.SimpleNamespace(VariableLayer=VariableLayer))

    # tensorflow_addons
    _mod("tensorflow_addons", activations=types.SimpleNamespace(gelu=gelu))
    return tf
