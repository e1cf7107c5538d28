import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def libdavi():
    """Build (if stale) and load the C-ABI library."""
    from deep_attentive_vi_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deep_attentive_vi_b200 import build
    build.build()
    return torch.device("cuda:0")


@pytest.fixture(params=["3xtf32", "fp32"])
def convs(request):
    """Precision of the host-side cuDNN convolutions: the product default (fp32-grade 3xTF32 on the
    tensor cores) and cuDNN's CUDA-core fp32.  Parity tests run in both, to the same tolerance."""
    from deep_attentive_vi_b200 import ops
    ops.set_conv_precision(request.param)
    yield request.param
    ops.set_conv_precision("3xtf32")


@pytest.fixture
def fp32_convs():
    from deep_attentive_vi_b200 import ops
    ops.set_conv_precision("fp32")
    yield
    ops.set_conv_precision("3xtf32")
