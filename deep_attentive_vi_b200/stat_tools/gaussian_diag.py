"""Bounded (residual) diagonal Gaussian layer -- host mirror of the reference's
stat_tools/gaussian_diag.py.  The parameter network is a host-side conv cell;
bounding, the residual posterior, sampling, the KL and the importance-sampling
log-probs all run in ONE fused kernel (ops.gauss_sample_kl)."""
import copy

import torch
import torch.nn as nn

from .. import ops
from ..layers import LAYER
from ..util.hparams import HParams


class GaussianDiag(nn.Module):
    def __init__(self, shape, hparams=None, network_hparams=None, in_channels=None):
        super().__init__()
        self._shape = list(shape)
        self._hparams = copy.deepcopy(hparams)
        nh = copy.deepcopy(network_hparams)
        nh["filters"] = 2 * self._shape[-1]                      # gd.py:112-117: mean and log-scale
        self._network_hparams = nh
        self._params_network = LAYER[nh.layer_type](nh, in_channels)   # gd.py:119
        self._raw = None

    @property
    def bounds(self):
        return (self._hparams.log_scale_low_bound, self._hparams.log_scale_upper_bound)

    @property
    def noise_std(self):
        return self._hparams.noise_stddev if self._hparams.noise_stddev > 0 else 0.0   # gd.py:165

    def call(self, condition, training):
        """Runs the parameter network (gd.py:158) and keeps the RAW (mu || log-sigma)
        tensor; bounds are applied inside the fused kernel."""
        self._raw = self._params_network(condition, training)
        return self._raw

    def params(self):
        """gd.py:267-269 (raw parameters; the train-time noise on the log-scale is
        injected by the kernel from the noise tensor handed to sample_kl)."""
        return self._raw

    def sample_kl(self, condition, training, eps, prior=None, prior_noise=None, noise=None,
                  compute_logp=False):
        """Posterior side of one variational layer: gd.py:224-240 (sample),
        vl.py:533-542 (KL vs `prior`, or N(0,I) when prior is None), vl.py:513-527.

        `prior` is the prior GaussianDiag of the same layer, already `call`-ed."""
        post_raw = self.call(condition, training)
        std_q = self.noise_std if (training and noise is not None) else 0.0
        if prior is None:
            return ops.gauss_sample_kl(post_raw, None, eps, self.bounds, (-5.0, 5.0),
                                       noise if std_q > 0 else None, None, std_q, 0.0, compute_logp)
        std_p = prior.noise_std if (training and prior_noise is not None) else 0.0
        return ops.gauss_sample_kl(post_raw, prior.params(), eps, self.bounds, prior.bounds,
                                   noise if std_q > 0 else None, prior_noise if std_p > 0 else None,
                                   std_q, std_p, compute_logp)

    def reg_loss(self):
        return self._params_network.reg_loss()

    @staticmethod
    def get_default_hparams():
        return HParams(log_scale_low_bound=-5.0, log_scale_upper_bound=5.0, noise_stddev=-1.0)   # gd.py:276-283
