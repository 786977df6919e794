"""MaskedSum operation (mirror of reference ``src/katsdpsigproc/maskedsum.py:33-162``)."""

from __future__ import annotations

from typing import Any, Mapping, Optional, Tuple

import numpy as np

from . import _capi, accel
from . import tune
from ._launch import FixedTuning, launch, ptr


class MaskedSumTemplate(FixedTuning):
    """Per-column sum over rows of ``mask[row] * src[row, col]`` (or of the amplitudes).

    Accumulation is float64 with one final rounding; the reference's float32 fma
    chain is order-dependent and only pinned to 1e-6 (``test/test_maskedsum.py:67``).
    """

    _TUNING = {"size": 16}

    @classmethod
    @tune.autotuner(test={"size": 16})
    def autotune(cls, context: Any, use_amplitudes: bool) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, use_amplitudes: bool = False,
                 tuning: Optional[Mapping[str, Any]] = None,
                 abs_mode: Optional[int] = None) -> None:
        self.context = context
        self.use_amplitudes = use_amplitudes
        self.abs_mode = _capi.default_abs_mode() if abs_mode is None else abs_mode
        self._init_tuning(context, tuning, use_amplitudes)

    def instantiate(self, command_queue: Any, shape: Tuple[int, int],
                    allocator: Optional[accel.AbstractAllocator] = None) -> "MaskedSum":
        return MaskedSum(self, command_queue, shape, allocator)


class MaskedSum(accel.Operation):
    """Concrete MaskedSum.  Slots: **src** (rows x cols complex64), **mask** (rows, float32),
    **dest** (cols; complex64, or float32 with ``use_amplitudes``)."""

    def __init__(self, template: MaskedSumTemplate, command_queue: Any, shape: Tuple[int, int],
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.shape = tuple(shape)
        self.slots["src"] = accel.IOSlot((shape[0], accel.Dimension(shape[1])), np.complex64)
        self.slots["mask"] = accel.IOSlot((shape[0],), np.float32)
        self.slots["dest"] = accel.IOSlot(
            (accel.Dimension(shape[1]),), np.float32 if template.use_amplitudes else np.complex64)

    def _run(self) -> None:
        src = self.buffer("src")
        launch(self.command_queue, "ksp_maskedsum", ptr(src), ptr(self.buffer("mask")),
               ptr(self.buffer("dest")), src.shape[0], src.shape[1], src.padded_shape[1],
               int(self.template.use_amplitudes), self.template.abs_mode)

    def parameters(self) -> Mapping[str, Any]:
        return {"shape": self.shape, "use_amplitudes": self.template.use_amplitudes}
