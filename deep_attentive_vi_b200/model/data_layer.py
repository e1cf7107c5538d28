"""Data distribution head -- host mirror of the reference's model/data_layer.py."""
import copy

import torch.nn as nn

from ..layers import LAYER
from ..stat_tools import DISTRIBUTION


class DataLayer(nn.Module):
    def __init__(self, preproc_hparams=None, data_dist_hparams=None, network_hparams=None,
                 class_dist=None, image_info=None, num_proc_blocks=1, in_channels=None, in_size=None):
        super().__init__()
        self._image_info = image_info
        self._num_proc_blocks = num_proc_blocks
        hp = copy.deepcopy(preproc_hparams)
        hp["num_filters"] = preproc_hparams["num_filters"] * num_proc_blocks            # data_layer.py:91
        layers, c, size = [], in_channels, in_size
        for s in range(num_proc_blocks):                                                # data_layer.py:92-101
            net = LAYER[preproc_hparams.layer_type](hp, c, size)
            c, size = net.out_channels, net.out_size
            layers.append(net)
            if s + 1 < num_proc_blocks:
                hp = copy.deepcopy(hp)
                hp["num_filters"] = hp["num_filters"] / 2
        self._preproc_layer = nn.ModuleList(layers)
        w = image_info["shape"][0] + image_info["pad_size"] * 2                         # data_layer.py:105-107
        h = image_info["shape"][1] + image_info["pad_size"] * 2
        self._data_dist = DISTRIBUTION[class_dist]([w, h, image_info["shape"][-1]], data_dist_hparams,
                                                   network_hparams, c)

    def decode(self, latent_codes, x=None, training=True):
        """Negative conditional log-likelihood [B] (data_layer.py:157-181)."""
        for net in self._preproc_layer:
            latent_codes = net(latent_codes, training)
        return self._data_dist.nll(x, latent_codes, training, crop_size=self._image_info["pad_size"])

    def reg_loss(self):
        r = self._data_dist.reg_loss()
        for net in self._preproc_layer:
            r = r + net.reg_loss()
        return r
