"""Fill operation (mirror of reference ``src/katsdpsigproc/fill.py:31-148``)."""

from __future__ import annotations

import ctypes
from typing import Any, Mapping, Optional, Tuple

import numpy as np

from . import accel
from . import tune
from ._launch import FixedTuning, launch, ptr


class FillTemplate(FixedTuning):
    """Set every element of a device array - padding included - to one value.

    ``ctype`` is kept for signature compatibility (the reference pastes it into the kernel
    source); the pre-built kernel treats an element as ``dtype.itemsize`` opaque bytes.
    To fill with zeros, :meth:`~katsdpsigproc_b200.accel.DeviceArray.zero` is cheaper.
    """

    _TUNING = {"wgs": 256}

    @classmethod
    @tune.autotuner(test={"wgs": 256})
    def autotune(cls, context: Any, dtype: Any, ctype: str) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, dtype: Any, ctype: str,
                 tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self.dtype = np.dtype(dtype)
        self.ctype = ctype
        if self.dtype.itemsize not in (1, 2, 4, 8, 16):
            raise ValueError("Fill supports element sizes of 1, 2, 4, 8 and 16 bytes")
        self._init_tuning(context, tuning, self.dtype, ctype)

    def instantiate(self, command_queue: Any, shape: Tuple[int, ...],
                    allocator: Optional[accel.AbstractAllocator] = None) -> "Fill":
        return Fill(self, command_queue, shape, allocator)


class Fill(accel.Operation):
    """Concrete Fill.  Slot **data**: the array to fill.  :meth:`set_value` chooses the value
    (default: the dtype's zero)."""

    def __init__(self, template: FillTemplate, command_queue: Any, shape: Tuple[int, ...],
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.shape = tuple(shape)
        self.slots["data"] = accel.IOSlot(shape, template.dtype)
        self.value = template.dtype.type()

    def set_value(self, value: Any) -> None:
        self.value = self.template.dtype.type(value)

    def _run(self) -> None:
        data = self.buffer("data")
        elements = int(np.prod(data.padded_shape))
        raw = np.asarray(self.value, dtype=self.template.dtype).tobytes()
        launch(self.command_queue, "ksp_fill", ptr(data), elements,
               ctypes.c_char_p(raw), self.template.dtype.itemsize)

    def parameters(self) -> Mapping[str, Any]:
        return {"dtype": self.template.dtype, "ctype": self.template.ctype,
                "shape": self.slots["data"].shape, "value": self.value}
