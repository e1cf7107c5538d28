"""Discretized logistic mixture data layer -- host mirror of the reference's
stat_tools/discretized_logistic_mixture.py; log_prob is one fused kernel that
also emits the gradient w.r.t. the 50-channel parameter tensor."""
import copy

import torch.nn as nn

from .. import ops
from ..layers import LAYER
from ..util.hparams import HParams


class DiscretizedLogisticMixture(nn.Module):
    def __init__(self, shape, hparams=None, network_hparams=None, in_channels=None):
        super().__init__()
        if hparams is None:
            hparams = self.get_default_hparams()
        self._hparams = copy.deepcopy(hparams)
        self._shape = list(shape)
        if self._hparams.num_logistic_mix != 5 or self._shape[-1] != 3:
            raise NotImplementedError("the fused kernel covers the published 5-mixture RGB configuration")
        nh = copy.deepcopy(network_hparams)
        C, M = self._shape[-1], self._hparams.num_logistic_mix
        nh["filters"] = M * C * 2 + M * 3 + M                      # dlm.py:129-133
        self._params_network = LAYER[nh.layer_type](nh, in_channels)

    def call(self, inputs, training):
        return self._params_network(inputs, training)              # dlm.py:153

    def nll(self, x, condition, training=True, crop_size=0):
        """-log_prob (dlm.py:311-383 negated as in data_layer.py:181) -> [B]."""
        return ops.dlm_nll(self.call(condition, training), x, self._hparams.log_scale_low_bound)

    def log_prob(self, x, condition, training=True, crop_size=0):
        return -self.nll(x, condition, training, crop_size)

    def reg_loss(self):
        return self._params_network.reg_loss()

    @staticmethod
    def get_default_hparams():
        return HParams(num_logistic_mix=5, high=255, low=0, log_scale_low_bound=-7.0,
                       log_scale_upper_bound=7.0)                  # dlm.py:419-424
