"""CPU oracles for the RFI flagging hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``katsdpsigproc_b200`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker or as the reported CPU
baseline -- never as the thing shipped.

Two tiers ("Oracle"):

``oracle.host_numpy``
    numpy/pandas restatement of the reference's ``katsdpsigproc.rfi.host``
    classes (float64, same library calls, same operation order).  Pinned
    against the reference's own known-answer tests and against outputs of the
    real reference generated in the build container
    (``tests/golden/make_golden.py``).

``oracle.contract``
    plain-C restatement (``contract.c``, built with gcc) of the *float32
    device contract*: the exact arithmetic the CUDA kernels promise (rules
    R1-R10 of SURVEY.md section 8(c')).  The CUDA path must match it bit for
    bit; it must match ``host_numpy`` as the rules say (selections exact,
    noise within 1 ulp, flags exact outside a counted 1e-6 band).
"""


import importlib
import os
import sys
from typing import Any, Optional


def reference_host() -> Optional[Any]:
    """The reference's own ``katsdpsigproc.rfi.host`` module from ``oracle/_ref`` (placed there
    by ``oracle/make_ref.py`` in the build container), or ``None`` if it is not there."""
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if not os.path.isfile(os.path.join(root, "katsdpsigproc", "rfi", "host.py")):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        return importlib.import_module("katsdpsigproc.rfi.host")
    except Exception:
        return None


def reference_twodflag() -> Optional[Any]:
    """The reference's own ``katsdpsigproc.rfi.twodflag`` module (numba) from ``oracle/_ref``, or
    ``None`` if it is not there or numba is missing."""
    if reference_host() is None:
        return None
    try:
        return importlib.import_module("katsdpsigproc.rfi.twodflag")
    except Exception:
        return None
