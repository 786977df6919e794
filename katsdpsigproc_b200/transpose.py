"""2-D transpose operation (mirror of reference ``src/katsdpsigproc/transpose.py:39-174``)."""

from __future__ import annotations

from typing import Any, Mapping, Optional, Tuple

import numpy as np

from . import accel
from . import tune
from ._launch import FixedTuning, launch, ptr


class TransposeTemplate(FixedTuning):
    """Transpose of a 2-D array of any 1/2/4/8/16-byte element type.

    Parameters
    ----------
    context
        Context the operation will run in
    dtype
        Element type
    ctype
        C name of the type; kept for compatibility with the reference signature
        (``transpose.py:60-66``), not used: the kernel only needs the element size
    tuning
        Accepted and ignored (see :class:`._launch.FixedTuning`)
    """

    _TUNING = {"block": 32, "vtx": 1, "vty": 4}

    @classmethod
    @tune.autotuner(test={"block": 32, "vtx": 1, "vty": 4})
    def autotune(cls, context: Any, dtype: Any, ctype: str) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, dtype: Any, ctype: str = "",
                 tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self.dtype = np.dtype(dtype)
        self.ctype = ctype
        if self.dtype.itemsize not in (1, 2, 4, 8, 16):
            raise ValueError("element size must be 1, 2, 4, 8 or 16 bytes")
        self._init_tuning(context, tuning, self.dtype, ctype)

    def instantiate(self, command_queue: Any, shape: Tuple[int, int],
                    allocator: Optional[accel.AbstractAllocator] = None) -> "Transpose":
        return Transpose(self, command_queue, shape, allocator)


class Transpose(accel.Operation):
    """Concrete transpose.  Slots: **src** (rows x cols), **dest** (cols x rows)."""

    def __init__(self, template: TransposeTemplate, command_queue: Any, shape: Tuple[int, int],
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.shape = tuple(shape)
        self.slots["src"] = accel.IOSlot(self.shape, template.dtype)
        self.slots["dest"] = accel.IOSlot((self.shape[1], self.shape[0]), template.dtype)

    def _run(self) -> None:
        src = self.buffer("src")
        dest = self.buffer("dest")
        launch(self.command_queue, "ksp_transpose", ptr(dest), ptr(src), src.shape[0],
               src.shape[1], dest.padded_shape[1], src.padded_shape[1],
               self.template.dtype.itemsize)

    def parameters(self) -> Mapping[str, Any]:
        return {"dtype": self.template.dtype, "ctype": self.template.ctype, "shape": self.shape}
