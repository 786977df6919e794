"""Percentile5 operation (mirror of reference ``src/katsdpsigproc/percentile.py:34-217``)."""

from __future__ import annotations

from typing import Any, Mapping, Optional, Tuple

import numpy as np

from . import _capi, accel
from . import tune
from ._launch import FixedTuning, launch, ptr


class Percentile5Template(FixedTuning):
    """Per-row minimum, maximum and 25th/75th/50th "lower" percentiles of the amplitudes.

    Parameters
    ----------
    context
        Context the operation will run in
    max_columns
        Upper bound on the number of columns processed per row.  The reference
        keeps a row in registers and so limits this to a few thousand
        (``percentile.py:75,104``); here it is only validated.
    is_amplitude
        ``True``: float32 input (assumed non-negative).  ``False``: complex64
        input; amplitudes follow ``abs_mode`` and the results are pure
        selections of those amplitudes (so they equal ``np.percentile(np.abs(x),
        ..., method="lower")`` bit for bit, which the reference's squared-
        amplitude ranking only matches to 1e-6).
    tuning
        Accepted and ignored
    abs_mode
        ``_capi.ABS_NUMPY`` / ``_capi.ABS_HYPOT`` (default from the environment)
    """

    _TUNING = {"size": 1024, "wgsy": 1}

    @classmethod
    @tune.autotuner(test={"size": 1024, "wgsy": 1})
    def autotune(cls, context: Any, max_columns: int, is_amplitude: bool) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, max_columns: int, is_amplitude: bool = True,
                 tuning: Optional[Mapping[str, Any]] = None,
                 abs_mode: Optional[int] = None) -> None:
        self.context = context
        self.max_columns = max_columns
        self.is_amplitude = is_amplitude
        self.abs_mode = _capi.default_abs_mode() if abs_mode is None else abs_mode
        self._init_tuning(context, tuning, max_columns, is_amplitude)

    def instantiate(self, command_queue: Any, shape: Tuple[int, int],
                    column_range: Optional[Tuple[int, int]] = None,
                    allocator: Optional[accel.AbstractAllocator] = None) -> "Percentile5":
        return Percentile5(self, command_queue, shape, column_range, allocator)


class Percentile5(accel.Operation):
    """Concrete Percentile5.  Slots: **src** (rows x cols, float32 or complex64),
    **dest** (5 x rows, float32: min, max, 25 %, 75 %, 50 %)."""

    def __init__(self, template: Percentile5Template, command_queue: Any, shape: Tuple[int, int],
                 column_range: Optional[Tuple[int, int]],
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        if column_range is None:
            column_range = (0, shape[1])
        if column_range[1] <= column_range[0]:
            raise ValueError("column range is empty")
        if column_range[0] < 0 or column_range[1] > shape[1]:
            raise IndexError("column range is out of range")
        if column_range[1] - column_range[0] > template.max_columns:
            raise ValueError("columns exceeds max_columns")
        self.template = template
        self.shape = tuple(shape)
        self.column_range = tuple(column_range)
        src_type = np.float32 if template.is_amplitude else np.complex64
        row_dim = accel.Dimension(shape[0])
        self.slots["src"] = accel.IOSlot((row_dim, accel.Dimension(shape[1])), src_type)
        self.slots["dest"] = accel.IOSlot((5, row_dim), np.float32)

    def _run(self) -> None:
        src = self.buffer("src")
        dest = self.buffer("dest")
        launch(self.command_queue, "ksp_percentile5", ptr(src), ptr(dest), src.shape[0],
               src.padded_shape[1], dest.padded_shape[1], self.column_range[0],
               self.column_range[1] - self.column_range[0], int(self.template.is_amplitude),
               self.template.abs_mode)

    def parameters(self) -> Mapping[str, Any]:
        return {
            "max_columns": self.template.max_columns,
            "is_amplitude": self.template.is_amplitude,
            "shape": self.shape,
            "column_range": self.column_range,
        }
