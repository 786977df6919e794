"""Shared pytest configuration.

``-m "not gpu"`` tests run anywhere (oracles, host-side runtime, C-ABI symbol
check, world_size-2 gloo sharding); ``-m gpu`` tests are the parity tests
proper and need a B200.  Nothing here reads /root/reference at run time.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# fixtures, --devices and the markers of the reference's plugin (device / force_autotune / ...)
pytest_plugins = ["katsdpsigproc_b200.pytest_plugin"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(autouse=True)
def _stub_autotune(patch_autotune):
    """As in the reference's suite: templates take their ``test=`` tuning instead of consulting
    the sqlite cache, unless the test is marked ``force_autotune``."""


@pytest.fixture(scope="session")
def golden():
    """Reference-generated fixtures (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "reference_host.npz")
    with np.load(path) as data:
        return {k: data[k] for k in data.files}


@pytest.fixture(scope="session")
def abs_mode():
    """Amplitude rule of this host's numpy (R1); -1 if neither known rule matches."""
    from oracle import contract

    return contract.detect_abs_mode()


@pytest.fixture(scope="session")
def context():
    """A device context; only requested by ``gpu`` tests, so failures are loud."""
    from katsdpsigproc_b200 import accel

    return accel.create_some_context(interactive=False)


@pytest.fixture
def command_queue(context):
    return context.create_command_queue()
