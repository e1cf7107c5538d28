"""Bernoulli data layer -- host mirror of the reference's stat_tools/bernoulli.py."""
import copy

import torch.nn as nn

from .. import ops
from ..layers import LAYER
from ..util.hparams import HParams


class Bernoulli(nn.Module):
    def __init__(self, shape, hparams=None, network_hparams=None, in_channels=None):
        super().__init__()
        self._shape = list(shape)
        self._hparams = copy.deepcopy(hparams) if hparams is not None else self.get_default_hparams()
        nh = copy.deepcopy(network_hparams)
        nh["filters"] = self._shape[-1]
        self._params_network = LAYER[nh.layer_type](nh, in_channels)

    def call(self, inputs, training):
        return self._params_network(inputs, training)              # bern.py:128

    def nll(self, x, condition, training=True, crop_size=0):
        """-log_prob with the border crop (bern.py:155-179) -> [B]."""
        return ops.bernoulli_nll(self.call(condition, training), x, crop_size)

    def log_prob(self, x, condition, training=True, crop_size=0):
        return -self.nll(x, condition, training, crop_size)

    def reg_loss(self):
        return self._params_network.reg_loss()

    @staticmethod
    def get_default_hparams():
        return HParams()                                           # bern.py:183-184
