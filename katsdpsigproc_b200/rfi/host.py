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
