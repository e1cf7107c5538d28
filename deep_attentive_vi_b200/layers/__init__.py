from .attention import ALIGNMENT_FUNCTION, DepthWiseAttention, NonlocalResNetBlock, SoftmaxAlignment  # noqa: F401
from .resnet import (LAYER, BatchNorm, ConcatConv2D, Conv2D, Conv2DSum, FactorizedReduce, ResNet,  # noqa: F401
                     ResNetBlock, SqueezeAndExciteBlock)
from .util import ACTIVATION, deep_vae_init_, swish  # noqa: F401
from .weight_norm import WNConv2d  # noqa: F401
