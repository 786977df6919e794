"""Call signatures of the host-style flagger stages.

The reference's ``rfi/host.py`` holds both these interfaces (:41-125) and numpy
implementations of them (:128-273).  Only the interfaces live here: they are the
base classes of the ``*HostFromDevice`` wrappers in :mod:`.device`, which run the
CUDA kernels behind a numpy-in / numpy-out call.  There is deliberately no CPU
implementation in this package.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import numpy as np


class AbstractBackgroundHost(ABC):
    @abstractmethod
    def __call__(self, vis: np.ndarray, flags: Optional[np.ndarray] = None) -> np.ndarray:
        """Deviations (channels x baselines, float32) of ``vis`` from its background."""


class AbstractNoiseEstHost(ABC):
    @abstractmethod
    def __call__(self, deviations: np.ndarray) -> np.ndarray:
        """Per-baseline noise estimate from channels x baselines deviations."""


class AbstractThresholdHost(ABC):
    @abstractmethod
    def __call__(self, deviations: np.ndarray, noise: np.ndarray) -> np.ndarray:
        """uint8 flags (channels x baselines)."""


class AbstractFlaggerHost(ABC):
    @abstractmethod
    def __call__(self, vis: np.ndarray, input_flags: Optional[np.ndarray] = None) -> np.ndarray:
        """uint8 flags (channels x baselines) for complex visibilities."""


class ReferenceHostClass:
    """Descriptor behind ``Template.host_class`` (reference ``rfi/device.py:174,380,504,679,834``).

    In the reference every device template names the numpy class that computes the same thing on
    the CPU, and the reference's own device tests build their expected values with
    ``template.host_class(...)`` (``test/rfi/test_background.py:89``, ``test_noise_est.py:57``,
    ``test_threshold.py:65``).  This package ships no CPU implementation, so the attribute
    resolves, on access, to the class of that name in the REFERENCE package
    (``katsdpsigproc.rfi.host``) when it can be imported - the situation of somebody re-pointing
    the reference's tests at this package, or of this repository's tests, which put the
    reference's own module on the path - and raises ``ImportError`` otherwise.  Nothing on the
    device path ever touches it.
    """

    def __init__(self, name: str) -> None:
        self.name = name

    def __get__(self, obj: object, owner: type) -> type:
        try:
            from katsdpsigproc.rfi import host as reference_host
        except ImportError as exc:
            raise ImportError(
                f"host_class is the reference's katsdpsigproc.rfi.host.{self.name}: install the "
                "reference package (or put it on sys.path) to use it; katsdpsigproc_b200 has no "
                "CPU implementation") from exc
        return getattr(reference_host, self.name)
