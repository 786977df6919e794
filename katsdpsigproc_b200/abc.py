"""Abstract interfaces of the device backend.

The reference defines its backend contract in ``src/katsdpsigproc/abc.py`` (:43-466) so that
operations work with either PyCUDA or PyOpenCL.  There is a single backend here
(:mod:`.cuda`, over the C-ABI runtime shims), but the interfaces are kept so that code which
type-checks against them (``isinstance(ctx, AbstractContext)``) or supplies its own queue /
event doubles in tests keeps working.  Kernel compilation and launching by name
(``compile`` / ``get_kernel`` / ``enqueue_kernel``) are not part of the contract: kernels are
ahead-of-time sm_100a code behind ``include/ksp_b200.h``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Sequence, Tuple

import numpy as np


class AbstractEvent(ABC):
    """A point in a command queue that can be waited for and timed."""

    @abstractmethod
    def wait(self) -> None:
        """Block until everything enqueued before the event has completed."""

    @abstractmethod
    def time_since(self, prior_event: "AbstractEvent") -> float:
        """Seconds from ``prior_event`` to this event (waits for both)."""

    @abstractmethod
    def time_till(self, next_event: "AbstractEvent") -> float:
        """Seconds from this event to ``next_event`` (waits for both)."""


class AbstractDevice(ABC):
    """One compute device."""

    @abstractmethod
    def make_context(self) -> "AbstractContext":
        ...

    @property
    @abstractmethod
    def name(self) -> str:
        ...

    @property
    @abstractmethod
    def platform_name(self) -> str:
        ...

    @property
    @abstractmethod
    def driver_version(self) -> str:
        ...

    @property
    @abstractmethod
    def simd_group_size(self) -> int:
        ...

    is_cuda: bool
    is_gpu: bool
    is_accelerator: bool
    is_cpu: bool

    @classmethod
    @abstractmethod
    def get_devices(cls) -> Sequence["AbstractDevice"]:
        ...

    @classmethod
    @abstractmethod
    def get_devices_by_platform(cls) -> Sequence[Sequence["AbstractDevice"]]:
        ...


class AbstractContext(ABC):
    """Allocation and queue factory for one device; ``with context:`` makes it current."""

    @property
    @abstractmethod
    def device(self) -> AbstractDevice:
        ...

    @abstractmethod
    def allocate_raw(self, n_bytes: int) -> Any:
        """Untyped device memory."""

    @abstractmethod
    def allocate(self, shape: Tuple[int, ...], dtype: Any, raw: Any = None) -> Any:
        """A typed device buffer, optionally over ``raw`` memory."""

    @abstractmethod
    def allocate_pinned(self, shape: Tuple[int, ...], dtype: Any) -> np.ndarray:
        """Page-locked host memory as a numpy array."""

    @abstractmethod
    def create_command_queue(self, profile: bool = False) -> "AbstractCommandQueue":
        ...

    @abstractmethod
    def __enter__(self) -> "AbstractContext":
        ...

    @abstractmethod
    def __exit__(self, *exc: Any) -> None:
        ...


class AbstractCommandQueue(ABC):
    """An in-order queue of asynchronous device work."""

    context: AbstractContext

    @abstractmethod
    def enqueue_read_buffer(self, buffer: Any, data: Any, blocking: bool = True) -> None:
        ...

    @abstractmethod
    def enqueue_write_buffer(self, buffer: Any, data: Any, blocking: bool = True) -> None:
        ...

    @abstractmethod
    def enqueue_copy_buffer_rect(self, src_buffer: Any, dest_buffer: Any, src_origin: int,
                                 dest_origin: int, shape: Sequence[int],
                                 src_strides: Sequence[int], dest_strides: Sequence[int]) -> None:
        ...

    @abstractmethod
    def enqueue_read_buffer_rect(self, buffer: Any, data: Any, buffer_origin: int,
                                 data_origin: int, shape: Sequence[int],
                                 buffer_strides: Sequence[int], data_strides: Sequence[int],
                                 blocking: bool = True) -> None:
        ...

    @abstractmethod
    def enqueue_write_buffer_rect(self, buffer: Any, data: Any, buffer_origin: int,
                                  data_origin: int, shape: Sequence[int],
                                  buffer_strides: Sequence[int], data_strides: Sequence[int],
                                  blocking: bool = True) -> None:
        ...

    @abstractmethod
    def enqueue_zero_buffer(self, buffer: Any) -> None:
        ...

    @abstractmethod
    def enqueue_marker(self) -> AbstractEvent:
        ...

    @abstractmethod
    def enqueue_wait_for_events(self, events: Sequence[AbstractEvent]) -> None:
        ...

    @abstractmethod
    def flush(self) -> None:
        ...

    @abstractmethod
    def finish(self) -> None:
        ...
