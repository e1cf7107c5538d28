"""Host-side restatement (PyTorch, NHWC) of the reference's conv cells
(layers/resnet.py).  Dense convolutions / batch-norm run through cuDNN via
torch: they are outside the hot-path scope but are needed to
drive the attention + ELBO kernels end to end with the reference's shapes.
Keras builds lazily from the input shape; here every constructor takes
`in_channels` (and `in_size` where a resize needs it) explicitly.
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..util.hparams import HParams
from .attention import NonlocalResNetBlock
from .util import ACTIVATION, swish
from .weight_norm import WNConv2d, to_nchw, to_nhwc

BN_EPS = 1e-5


def _wn(hp, cin, cout, k, strides=1, use_bias=None, padding="same"):
    return WNConv2d(cin, cout, kernel_size=k, strides=strides,
                    use_bias=hp.use_bias if use_bias is None else use_bias, padding=padding,
                    use_weight_norm=hp.use_weight_norm, lambda_reg=hp.lambda_reg)


def _reg(mods):
    r = 0.0
    for m in mods:
        if m is not None and hasattr(m, "reg_loss"):
            r = r + m.reg_loss()
    return r


class BatchNorm(nn.Module):
    """tf.keras.layers.BatchNormalization(momentum, epsilon=1e-5) (res.py:577-582);
    plain per-replica statistics (no sync-BN, matching MirroredStrategy)."""

    def __init__(self, C, momentum=0.9, lambda_reg=1.5e-4):
        super().__init__()
        self.bn = nn.BatchNorm2d(C, eps=BN_EPS, momentum=1.0 - momentum)
        self.lambda_reg = lambda_reg

    def forward(self, x, training):
        y = F.batch_norm(to_nchw(x), self.bn.running_mean, self.bn.running_var, self.bn.weight,
                         self.bn.bias, training, self.bn.momentum, self.bn.eps)
        return to_nhwc(y)

    def reg_loss(self):
        return self.lambda_reg * (self.bn.weight.square().sum() + self.bn.bias.square().sum())


class FactorizedReduce(nn.Module):
    """res.py:62-135: swish, then four stride-2 1x1 convs on shifted views, concatenated."""

    def __init__(self, num_filters, hparams, in_channels):
        super().__init__()
        q = num_filters // 4
        outs = [q, q, q, num_filters - 3 * q]
        self.convs = nn.ModuleList([_wn(hparams, in_channels, o, 1, strides=2, padding="valid") for o in outs])

    def forward(self, x):
        out = swish(x)
        views = [out, out[:, 1:, 1:, :], out[:, :, 1:, :], out[:, 1:, :, :]]       # res.py:128-131
        return torch.cat([c(v.contiguous()) for c, v in zip(self.convs, views)], dim=-1)

    def reg_loss(self):
        return _reg(self.convs)


class SqueezeAndExciteBlock(nn.Module):
    """res.py:288-359 (use_se is False in every published config)."""

    def __init__(self, hparams, in_channels):
        super().__init__()
        ratio = hparams.se_ratio if hparams.exists("se_ratio") else 16
        mid = max(in_channels // ratio, 4)
        self.fc1 = _wn(hparams, in_channels, mid, 1, use_bias=False)
        self.fc2 = _wn(hparams, mid, in_channels, 1, use_bias=False)

    def forward(self, x):
        s = x.mean(dim=(1, 2), keepdim=True)
        return x * torch.sigmoid(self.fc2(F.relu(self.fc1(s))))

    def reg_loss(self):
        return _reg([self.fc1, self.fc2])


class Conv2D(nn.Module):
    """res.py:538-666: BN -> [nonlocal] -> (activation -> wn-conv) x num_times -> [SE]."""

    def __init__(self, hparams, in_channels):
        super().__init__()
        hp = self._hparams = hparams
        self._num_times = hp.num_times if hp.exists("num_times") else 1            # res.py:547
        self._activation = ACTIVATION[hp.activation] if hp.activation else (lambda t: t)
        self._batch_norm_layer = BatchNorm(in_channels, hp.bn_momentum, hp.lambda_reg) if hp.use_batch_norm else None
        filters = hp.filters if hp.filters >= 0 else in_channels                    # res.py:586-587
        self.out_channels = filters
        convs = []
        for i in range(self._num_times):                                            # res.py:589-601
            last = i == self._num_times - 1
            convs.append(_wn(hp, in_channels, filters if last else in_channels,
                             hp.kernel_size if last else 1, strides=hp.strides))
        self._conv2d_layers = nn.ModuleList(convs)
        self._se_layer = SqueezeAndExciteBlock(hp, filters) if hp.use_se else None
        self._nonlocal_layer = (NonlocalResNetBlock(hp, hp.nonlocop_hparams, in_channels)
                                if hp.use_nonlocal else None)

    def forward(self, inputs, training=True, residual=None):
        """residual = (shortcut, scale): return shortcut + scale * node(inputs) (res.py:284-286); without a squeeze-
        and-excite layer behind it the LAST convolution takes the sum into its epilogue."""
        t = inputs
        if self._batch_norm_layer is not None:
            t = self._batch_norm_layer(t, training)                                 # res.py:619
        if self._nonlocal_layer is not None:
            t = self._nonlocal_layer(t)                                             # res.py:621-622
        fuse_act = self._hparams.activation == "swish"                              # conv(swish(t)) as ONE kernel
        last = len(self._conv2d_layers) - 1
        for i, conv in enumerate(self._conv2d_layers):                              # res.py:624-627
            add, alpha = (residual if (residual is not None and i == last and self._se_layer is None) else (None, 1.0))
            if fuse_act:
                t = conv(t, act="swish", add=add, alpha=alpha)
            else:
                t = conv(self._activation(t), add=add, alpha=alpha)
        if self._se_layer is not None:
            t = self._se_layer(t)
            if residual is not None:
                t = residual[0] + residual[1] * t
        return t

    def reg_loss(self):
        return _reg([self._batch_norm_layer, self._nonlocal_layer, self._se_layer, *self._conv2d_layers])

    @staticmethod
    def get_default_hparams():
        return HParams(layer_type="conv2d", kernel_size=3, num_times=2, strides=1, use_bias=True,
                       kernel_regularizer="l2", bias_regularizer="l2", lambda_reg=1.5e-4, use_se=False,
                       use_weight_norm=True, use_data_init=False, activation="", activation_alpha=0.0,
                       activation_max_value=1e3, activation_threshold=0, activation_kwargs="",
                       activation_result="", use_batch_norm=False, bn_momentum=0.9,
                       batch_norm_regularizer="l2", use_nonlocal=False,
                       nonlocop_hparams=NonlocalResNetBlock.get_default_hparams(), filters=-1)


class ConcatConv2D(nn.Module):
    """res.py:361-439: concat -> wn-conv -> [nonlocal]."""

    def __init__(self, hparams, in_channels, num_filters=None):
        super().__init__()
        ca, cb = in_channels
        self._hparams = hparams
        self._num_filters = ca if num_filters is None else num_filters              # res.py:385
        self._conv2d_layer = _wn(hparams, ca + cb, self._num_filters, hparams.kernel_size, use_bias=True)
        self._nonlocal_layer = (NonlocalResNetBlock(hparams, hparams.nonlocop_hparams, self._num_filters)
                                if hparams.use_nonlocal else None)
        self.out_channels = self._num_filters

    def forward(self, inputs, training=True):
        res = self._conv2d_layer(torch.cat([inputs[0], inputs[1]], dim=-1))        # res.py:411-413
        if self._nonlocal_layer is not None:
            res = self._nonlocal_layer(res)
        return res

    def reg_loss(self):
        return _reg([self._conv2d_layer, self._nonlocal_layer])

    @staticmethod
    def get_default_hparams():
        return HParams(layer_type="concat_conv2d", kernel_size=1, use_bias=True, kernel_regularizer="l2",
                       bias_regularizer="l2", lambda_reg=1.5e-4, use_weight_norm=True, use_data_init=False,
                       use_nonlocal=False, nonlocop_hparams=NonlocalResNetBlock.get_default_hparams())


class Conv2DSum(nn.Module):
    """res.py:442-534: conv_a(a) + conv(b) -> [nonlocal]."""

    def __init__(self, hparams, in_channels, num_filters=None):
        super().__init__()
        ca, cb = in_channels
        self._hparams = hparams
        self._num_filters = ca if num_filters is None else num_filters              # res.py:464
        self._conv2d_layer = _wn(hparams, cb, self._num_filters, hparams.kernel_size, use_bias=True)
        self._conv2d_layer_a = (_wn(hparams, ca, self._num_filters, hparams.kernel_size, use_bias=True)
                                if self._num_filters != ca else None)               # res.py:476-487
        self._nonlocal_layer = (NonlocalResNetBlock(hparams, hparams.nonlocop_hparams, self._num_filters)
                                if hparams.use_nonlocal else None)
        self.out_channels = self._num_filters

    def forward(self, inputs, training=True):
        a, b = inputs
        b = self._conv2d_layer(b)                                                   # res.py:505
        if self._conv2d_layer_a is not None:
            a = self._conv2d_layer_a(a)                                             # res.py:507
        res = a + b
        if self._nonlocal_layer is not None:
            res = self._nonlocal_layer(res)
        return res

    def reg_loss(self):
        return _reg([self._conv2d_layer, self._conv2d_layer_a, self._nonlocal_layer])

    @staticmethod
    def get_default_hparams():
        return HParams(layer_type="conv2d_sum", kernel_size=1, use_bias=True, kernel_regularizer="l2",
                       bias_regularizer="l2", lambda_reg=1.5e-4, use_weight_norm=True, use_data_init=False,
                       use_nonlocal=False, nonlocop_hparams=NonlocalResNetBlock.get_default_hparams())


def _resize_bilinear_align_corners(x, size):
    """tf.compat.v1.image.resize_bilinear(align_corners=True) (res.py:232)."""
    return to_nhwc(F.interpolate(to_nchw(x), size=size, mode="bilinear", align_corners=True))


def _resize_nearest(x, size):
    """tf.compat.v1.image.resize_nearest_neighbor(align_corners=False) (res.py:252):
    src = floor(dst * in/out), which is torch's 'nearest'."""
    return to_nhwc(F.interpolate(to_nchw(x), size=size, mode="nearest"))


class ResNetBlock(nn.Module):
    """res.py:139-286: shortcut(x) + 0.1 * node_n(...node_1(x))."""

    def __init__(self, scale, hparams, num_nodes, node_type, in_channels, in_size):
        super().__init__()
        hp = copy.deepcopy(hparams)
        self._scale, self._num_nodes = scale, num_nodes
        if scale < 0:                                                                # res.py:197-206
            hp["filters"], hp["strides"] = abs(scale) * in_channels, abs(scale)
        elif scale > 0:
            hp["filters"], hp["strides"] = in_channels // 2, 1
        else:
            hp["filters"], hp["strides"] = in_channels, 1
        self.out_channels = hp["filters"]
        self._up_size = None
        if scale == 0:
            self._shortcut = None
        elif scale < 0:
            self._shortcut = FactorizedReduce(hp["filters"], hp, in_channels)       # res.py:215
        else:
            dim = int(in_size * abs(scale))
            self._up_size = (dim, dim)
            shp = copy.deepcopy(hp)                                                  # res.py:223-229
            shp["use_bias"], shp["activation"], shp["kernel_size"] = False, "", 1
            shp["use_batch_norm"], shp["use_se"] = False, False
            self._shortcut = Conv2D(shp, in_channels)
        nodes = []
        nhp = copy.deepcopy(hp)
        nhp["use_se"] = hp.use_se if num_nodes == 1 else False
        nodes.append(LAYER[node_type](nhp, in_channels))                            # res.py:246-254
        for i in range(num_nodes - 1):
            nhp = copy.deepcopy(hp)
            nhp["strides"] = 1
            nhp["use_se"] = hp.use_se if i == num_nodes - 2 else False
            nodes.append(LAYER[node_type](nhp, hp["filters"]))
        self._nodes = nn.ModuleList(nodes)

    def forward(self, inputs, training=True):
        if self._scale == 0:
            shortcut = inputs
        elif self._scale < 0:
            shortcut = self._shortcut(inputs)
        else:
            shortcut = self._shortcut(_resize_bilinear_align_corners(inputs, self._up_size), training)
        t = inputs
        last = len(self._nodes) - 1
        for i, node in enumerate(self._nodes):
            if i == 0 and self._scale > 0:
                t = _resize_nearest(t, self._up_size)
            if i == last and isinstance(node, Conv2D):
                return node(t, training, residual=(shortcut, 0.1))                   # res.py:284-286, in the conv's epilogue
            t = node(t, training)
        return 0.1 * t + shortcut                                                    # res.py:284-286

    def reg_loss(self):
        return _reg([self._shortcut, *self._nodes])


class ResNet(nn.Module):
    """res.py:668-807: optional stem conv, then num_blocks ResNetBlocks."""

    def __init__(self, hparams, in_channels, in_size):
        super().__init__()
        hp = self._hparams = copy.deepcopy(hparams)
        scales = hp.scale if isinstance(hp.scale, list) else [hp.scale] * hp.num_blocks   # res.py:694-695
        c, size = in_channels, in_size
        self._stem_layer = None
        if hp.use_stem:
            self._stem_layer = _wn(hp, c, hp.num_filters, hp.stem_kernel_size, use_bias=True)  # res.py:711-717
            c = hp.num_filters
        blocks = []
        for k in range(hp.num_blocks):
            blk = ResNetBlock(scales[k], hp, hp.num_nodes, hp.node_type, c, size)
            c = blk.out_channels
            if scales[k] < 0:
                size = -(-size // abs(scales[k]))
            elif scales[k] > 0:
                size = int(size * scales[k])
            blocks.append(blk)
        self._residual_blocks = nn.ModuleList(blocks)
        self.out_channels, self.out_size = c, size

    def forward(self, inputs, training=True):
        t = inputs
        if self._stem_layer is not None:
            t = self._stem_layer(t)
        for blk in self._residual_blocks:
            t = blk(t, training)
        return t

    def reg_loss(self):
        return _reg([self._stem_layer, *self._residual_blocks])

    @staticmethod
    def get_default_hparams():
        return HParams(layer_type="resnet", node_type="conv2d", num_nodes=2, use_stem=True, kernel_size=3,
                       stem_kernel_size=3, num_blocks=2, scale=-2, activation="relu", activation_alpha=0.0,
                       activation_max_value=1e3, activation_threshold=0, activation_kwargs="",
                       output_activation_kwargs="", activation_result="", num_filters=16, use_bias=True,
                       kernel_regularizer="l2", bias_regularizer="l2", batch_norm_regularizer="l2",
                       lambda_reg=1.5e-4, use_se=False, use_weight_norm=True, use_data_init=False,
                       use_batch_norm=True, bn_momentum=0.9, use_nonlocal=False,
                       nonlocop_hparams=NonlocalResNetBlock.get_default_hparams())


LAYER = {"conv2d": Conv2D, "conv2d_sum": Conv2DSum, "concat_conv2d": ConcatConv2D, "resnet": ResNet}   # res.py:810
