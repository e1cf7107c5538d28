from .data_layer import DataLayer  # noqa: F401
from .util import CIFAR10_16L, OMNIGLOT_15L, image_info, parse_flags  # noqa: F401
from .vae import DeepVAE  # noqa: F401
from .variational_layer import VariationalLayer  # noqa: F401
