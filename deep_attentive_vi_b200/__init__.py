"""deep_attentive_vi_b200 -- B200 (sm_100a) implementation of the attention and
ELBO-reduction hot path of the Deep Attentive VAE (ifiaposto/Deep_Attentive_VI).

Layout
------
csrc/            hand-written CUDA kernels + the C ABI (include/davi.h) -> libdavi.so
_lib.py          ctypes binding of that ABI (the drop-in boundary)
ops.py           torch.autograd wrappers (hand-written gradients)
layers/ model/ stat_tools/ util/
                 host-side mirror of the reference's src/ tree: same class names,
                 hparams and registries, bodies routed through the kernels
"""
__version__ = "0.1.0"
