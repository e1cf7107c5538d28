"""Initialiser / activation tables of the reference (layers/util.py:53-84)."""
import math

import torch
import torch.nn.functional as F


def deep_vae_init_(t, fan_in):
    """DeepVaeInit = VarianceScaling(scale=0.33, mode='fan_in', distribution='uniform')
    (layers/util.py:53-58): U(-l, l) with l = sqrt(3 * 0.33 / fan_in)."""
    lim = math.sqrt(3.0 * 0.33 / max(1.0, float(fan_in)))
    with torch.no_grad():
        return t.uniform_(-lim, lim)


def swish(x):
    """layers/util.py:62-69."""
    return F.silu(x)


ACTIVATION = {            # layers/util.py:72-81
    "tanh": torch.tanh,
    "relu": F.relu,
    "sigmoid": torch.sigmoid,
    "softplus": F.softplus,
    "exp": torch.exp,
    "elu": F.elu,
    "swish": swish,
}

# Regularisation strengths (SURVEY.md appendix B "regularisation quirk"): objects
# built through REGULARIZER['l2'](l=lambda_reg) use lambda_reg; places that pass
# the bare string 'l2' get Keras' default l2(0.01).
KERAS_DEFAULT_L2 = 0.01
