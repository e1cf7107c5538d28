"""Weight-normalised convolution: host-side restatement of the reference's
WeightNorm(tf.keras.layers.Conv2D) (layers/weight_norm.py:126-143,225-239).

kernel = l2_normalize(v, over all axes but the output filter) * exp(log_g);
log_g is initialised to log||v|| + 1e-2 on the first call (_init_norm).
Tensors are NHWC; the convolution itself is cuDNN through torch on a
channels_last view (dense convs are outside the hot-path scope).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .util import deep_vae_init_


def to_nchw(x):  # [B,H,W,C] contiguous -> NCHW view in channels_last memory format (no copy)
    return x.permute(0, 3, 1, 2)


def to_nhwc(y):  # inverse; copies only if cuDNN returned a non-channels_last tensor
    return y.permute(0, 2, 3, 1).contiguous()


class WNConv2d(nn.Module):
    def __init__(self, cin, cout, kernel_size=1, strides=1, use_bias=True, padding="same",
                 use_weight_norm=True, lambda_reg=1.5e-4):
        super().__init__()
        k = kernel_size
        self.k, self.strides, self.padding, self.use_weight_norm = k, strides, padding, use_weight_norm
        v = torch.empty(cout, cin, k, k)
        deep_vae_init_(v, fan_in=k * k * cin)                                   # res.py:596
        self.v = nn.Parameter(v.contiguous(memory_format=torch.channels_last))
        if use_weight_norm:
            norm = self.v.detach().flatten(1).norm(dim=1)
            self.log_g = nn.Parameter(torch.log(norm) + 1e-2)                    # weight_norm.py:139-143
        if use_bias:
            self.bias = nn.Parameter(deep_vae_init_(torch.empty(cout), fan_in=cout))  # res.py:597
        else:
            self.bias = None
        self.lambda_reg = lambda_reg
        # set by util/trainer.py: a leaf view into the flat effective-weight array that ONE fused
        # kernel (davi_weight_norm_fwd) fills for all convolutions of the model before the forward
        self._w_override = None
        self._wt_override = None    # likewise: [Cin,kh,kw,Cout] transposed + flipped image of the kernel (input gradient)
        self._lo_override = None    # likewise: (davi_tf32_lo image of the kernel [Cout,kh,kw,Cin], of its transposed image)

    def kernel(self):
        if not self.use_weight_norm:
            return self.v
        if self._w_override is not None:
            return self._w_override
        # tf.nn.l2_normalize: v * rsqrt(max(sum v^2, 1e-12))                     weight_norm.py:133-134
        ss = self.v.square().sum(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12)
        return self.v * (torch.exp(self.log_g).view(-1, 1, 1, 1) * torch.rsqrt(ss))

    def forward(self, x, act=None, add=None, alpha=1.0):
        """x [B,H,W,Cin] -> add + alpha * conv(act(x)) [B,H',W',Cout].  act (None | "swish"), add (a tensor like the
        result) and alpha are the neighbours of the convolution inside a ResNet node (res.py:613-633, 273-286); the
        tcgen05 kernels take them inside (ops.conv2d_bias)."""
        xc = to_nchw(x)
        k, s = self.k, self.strides
        addc = None if add is None else to_nchw(add)
        wt = self._wt_override if (self.use_weight_norm and self._w_override is not None) else None
        lo = self._lo_override if wt is not None else None
        if self.padding == "same":
            H, W = x.shape[1], x.shape[2]
            # TF 'same': total = max((ceil(n/s)-1)*s + k - n, 0); before = total//2, rest after
            def pads(n):
                out = -(-n // s)
                tot = max((out - 1) * s + k - n, 0)
                return tot // 2, tot - tot // 2
            (pt, pb), (pl, pr) = pads(H), pads(W)
            if pt == pb and pl == pr:
                y = ops.conv2d_bias(xc, self.kernel(), self.bias, stride=s, padding=(pt, pl), act=act, add=addc, alpha=alpha,
                                    wt=wt, lo=lo)
            else:
                if act == "swish":
                    xc = F.silu(xc)
                elif act is not None:
                    raise ValueError(act)
                y = ops.conv2d_bias(F.pad(xc, (pl, pr, pt, pb)), self.kernel(), self.bias, stride=s, add=addc, alpha=alpha)
        else:
            y = ops.conv2d_bias(xc, self.kernel(), self.bias, stride=s, act=act, add=addc, alpha=alpha)
        return to_nhwc(y)

    def reg_loss(self):
        """l2(lambda_reg) on v, g and bias (res.py:564-566, weight_norm.py:213)."""
        r = self.v.square().sum()
        if self.use_weight_norm:
            r = r + self.log_g.square().sum()
        if self.bias is not None:
            r = r + self.bias.square().sum()
        return self.lambda_reg * r
