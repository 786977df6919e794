"""HReduce operation (mirror of reference ``src/katsdpsigproc/reduce.py:38-214``)."""

from __future__ import annotations

import ctypes
import weakref
from typing import Any, Mapping, Optional, Tuple

import numpy as np

from . import _capi, accel
from . import tune
from ._launch import FixedTuning, launch, ptr


class HReduceTemplate(FixedTuning):
    """Reduce every row of a 2-D array over a column range with a user-supplied operator.

    As in the reference the operator is C source text: ``op`` is an expression combining
    ``a`` and ``b`` (``"a + b"``, ``"max(a, b)"``), ``identity`` an expression for its
    identity, ``extra_code`` anything they need, ``ctype`` the element type.  The kernel is
    compiled when the template is constructed (NVRTC, sm_100a); compile errors raise
    :exc:`RuntimeError` with the compiler log.  Only commutative, associative operators.
    """

    _TUNING = {"wgsx": 32, "wgsy": 8}

    @classmethod
    @tune.autotuner(test={"wgsx": 32, "wgsy": 8})
    def autotune(cls, context: Any, dtype: Any, ctype: str, op: str, identity: str, extra_code: str) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, dtype: Any, ctype: str, op: str, identity: str,
                 extra_code: str = "", tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self.dtype = np.dtype(dtype)
        self.ctype = ctype
        self.op = op
        self.identity = identity
        self.extra_code = extra_code
        self._init_tuning(context, tuning, self.dtype, ctype, op, identity, extra_code)
        self.wgsx = self.tuning["wgsx"]
        self.wgsy = self.tuning["wgsy"]
        context._make_current()
        handle = ctypes.c_void_p()
        code = _capi.load().ksp_hreduce_create(
            ctype.encode(), op.encode(), identity.encode(), extra_code.encode(),
            self.dtype.itemsize, ctypes.byref(handle))
        if code != 0:
            log = (_capi.load().ksp_jit_log() or b"").decode(errors="replace")
            raise RuntimeError(f"HReduce kernel could not be built: "
                               f"{_capi.error_string(code)}\n{log}")
        self._handle = handle
        self._finalizer = weakref.finalize(self, _capi.load().ksp_hreduce_destroy, handle)

    def instantiate(self, command_queue: Any, shape: Tuple[int, int],
                    column_range: Optional[Tuple[int, int]] = None,
                    allocator: Optional[accel.AbstractAllocator] = None) -> "HReduce":
        return HReduce(self, command_queue, shape, column_range, allocator)


class HReduce(accel.Operation):
    """Concrete HReduce.  Slots: **src** (rows x columns), **dest** (rows)."""

    def __init__(self, template: HReduceTemplate, command_queue: Any, shape: Tuple[int, int],
                 column_range: Optional[Tuple[int, int]] = None,
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        if len(shape) != 2:
            raise ValueError("shape must be 2-dimensional")
        if column_range is None:
            column_range = (0, shape[1])
        if column_range[0] < 0 or column_range[1] > shape[1]:
            raise ValueError("column range overflows the array")
        if column_range[0] >= column_range[1]:
            raise ValueError("column range is empty")
        super().__init__(command_queue, allocator)
        self.template = template
        self.column_range = (int(column_range[0]), int(column_range[1]))
        rows = accel.Dimension(shape[0], template.wgsy)
        self.slots["src"] = accel.IOSlot((rows, shape[1]), template.dtype)
        self.slots["dest"] = accel.IOSlot((rows,), template.dtype)

    def _run(self) -> None:
        src = self.buffer("src")
        launch(self.command_queue, "ksp_hreduce", self.template._handle, ptr(src),
               ptr(self.buffer("dest")), src.shape[0], src.padded_shape[1],
               self.column_range[0], self.column_range[1] - self.column_range[0])

    def parameters(self) -> Mapping[str, Any]:
        return {"dtype": self.template.dtype, "ctype": self.template.ctype,
                "shape": self.slots["src"].shape, "column_range": self.column_range,
                "op": self.template.op, "identity": self.template.identity,
                "extra_code": self.template.extra_code}
