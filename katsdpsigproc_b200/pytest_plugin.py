"""pytest plugin for projects that test operations built on this runtime.

The fixtures, command-line option and markers of the reference's plugin
(``pytest_plugin.py:30-131``), so that a test suite written against it runs against this package
by loading ``katsdpsigproc_b200.pytest_plugin`` instead (``pytest_plugins = [...]`` in a
``conftest.py`` or ``-p`` on the command line):

``device``
    parametrised over the CUDA devices found (``--devices=first-per-api`` - the default -
    takes the first, ``all`` every one, ``none`` skips); a test that needs a device and finds
    none is reported as ``xfail`` without being run.  ``@pytest.mark.cuda_only(
    min_compute_capability=(major, minor))`` and ``@pytest.mark.device_filter(fn)`` narrow the
    list; ``@pytest.mark.opencl_only`` leaves it empty (there is no OpenCL backend here).
``context``, ``command_queue``
    a context on that device, made current for the duration of the test, and a queue on it.
``patch_autotune``
    replaces :func:`katsdpsigproc_b200.tune.autotuner_impl` by ``stub_autotuner`` (templates take
    the ``test=`` value of their ``@autotuner``) or, for tests marked
    ``@pytest.mark.force_autotune``, by ``force_autotuner``; ``context`` depends on it.
"""

from __future__ import annotations

from typing import Any, Generator, List

import pytest

from . import accel, tune


@pytest.fixture
def patch_autotune(request: Any, monkeypatch: Any) -> None:
    impl = tune.stub_autotuner
    if request.node.get_closest_marker("force_autotune"):
        impl = tune.force_autotuner
    monkeypatch.setattr(tune, "autotuner_impl", impl)


@pytest.fixture
def context(device: Any, patch_autotune: None) -> Generator[Any, None, None]:
    with device.make_context() as ctx:
        yield ctx


@pytest.fixture
def command_queue(context: Any) -> Any:
    return context.create_command_queue()


def pytest_addoption(parser: Any) -> None:
    group = parser.getgroup("katsdpsigproc")
    group.addoption("--devices", choices=["first-per-api", "all", "none"], default="first-per-api",
                    help="Select which devices to use for testing")


def pytest_configure(config: Any) -> None:
    config.addinivalue_line("markers", "force_autotune: unconditionally run autotuning")
    config.addinivalue_line("markers", "cuda_only: run test only on CUDA devices")
    config.addinivalue_line("markers", "opencl_only: run test only on OpenCL devices")
    config.addinivalue_line("markers", "device_filter(filter): run test only on devices matching 'filter'")


def _matching_devices(definition: Any) -> List[Any]:
    try:
        devices = list(accel.candidate_devices())
    except Exception:              # no driver / no library on this machine
        devices = []
    for marker in definition.iter_markers("cuda_only"):
        min_cc = marker.kwargs.get("min_compute_capability", (0, 0))
        devices = [d for d in devices if d.is_cuda and d.compute_capability >= min_cc]
    if definition.get_closest_marker("opencl_only") is not None:
        devices = [d for d in devices if not d.is_cuda]
    for marker in definition.iter_markers("device_filter"):
        devices = [d for d in devices if marker.args[0](d)]
    return devices


def pytest_generate_tests(metafunc: Any) -> None:
    if "device" not in metafunc.fixturenames:
        return
    option = metafunc.config.getoption("devices")
    if option == "none":
        skip = pytest.mark.skip(reason="--devices=none passed on command line")
        metafunc.parametrize("device", [pytest.param(None, marks=skip)])
        return
    devices = _matching_devices(metafunc.definition)
    if option == "first-per-api":
        devices = devices[:1]      # one API (CUDA) here
    if not devices:
        missing = pytest.mark.xfail(reason="No matching device found", run=False)
        metafunc.parametrize("device", [pytest.param(None, marks=missing)])
    else:
        metafunc.parametrize("device", devices, ids=[f"{d.name} ({d.platform_name})" for d in devices])
