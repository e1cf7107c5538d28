from .bernoulli import Bernoulli
from .discretized_logistic_mixture import DiscretizedLogisticMixture
from .gaussian_diag import GaussianDiag

DISTRIBUTION = {            # reference stat_tools/__init__.py:51
    "gaussian_diag": GaussianDiag,
    "discretized_logistic_mixture": DiscretizedLogisticMixture,
    "bernoulli": Bernoulli,
}
