import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import relearn_b200 as R

    c = R.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["tcgen05", "ffma"])
def pass_kernel(request):
    """Run a GPU test once per full-batch pass kernel: tensor cores (the default) and the FP32-pipe twin."""
    from relearn_b200 import _lib as L

    k = L.RL_PASS_KERNEL_TCGEN05 if request.param == "tcgen05" else L.RL_PASS_KERNEL_FFMA
    L.check(L.lib().rl_pass_kernel_select(k))
    yield request.param
    L.check(L.lib().rl_pass_kernel_select(L.RL_PASS_KERNEL_TCGEN05))
