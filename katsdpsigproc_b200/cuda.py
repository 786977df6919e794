"""CUDA backend over the C-ABI runtime shims (no PyCUDA, no runtime compilation).

Mirrors the behaviour of the reference's PyCUDA backend (reference
``src/katsdpsigproc/cuda.py``) for the parts the flagging path uses:
``Device`` (:86-158), ``Context`` (:163-255), ``CommandQueue`` (:257-479) and
``Event`` (:71-84).  Everything is stream-ordered and asynchronous unless a
call is documented as blocking, exactly as there.  What is deliberately absent:
``Context.compile`` / ``enqueue_kernel`` (kernels are ahead-of-time compiled
sm_100a code reached through ``_capi``), SVM/managed memory and the tuning
queue.
"""

from __future__ import annotations

import ctypes
import threading
from ctypes import byref, c_float, c_int, c_size_t, c_void_p
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .abc import AbstractCommandQueue, AbstractContext, AbstractDevice, AbstractEvent


class RawAllocation:
    """An untyped device allocation (what ``Context.allocate_raw`` returns)."""

    def __init__(self, device_index: int, n_bytes: int) -> None:
        self.n_bytes = int(n_bytes)
        self.device_index = device_index
        ptr = c_void_p()
        _capi.call("ksp_device_set", device_index)
        _capi.call("ksp_malloc", byref(ptr), c_size_t(max(self.n_bytes, 1)))
        self.ptr = ptr.value or 0

    def __int__(self) -> int:
        return self.ptr

    def __del__(self) -> None:
        ptr, self.ptr = getattr(self, "ptr", 0), 0
        if ptr:
            try:
                _capi.load().ksp_device_set(self.device_index)
                _capi.load().ksp_free(c_void_p(ptr))
            except Exception:  # interpreter shutdown
                pass


class ExternalAllocation:
    """Device memory owned by someone else (a torch / cupy tensor, another library).

    Has the interface of :class:`RawAllocation`, so it can back a ``DeviceArray`` through the
    ``raw=`` argument without a copy; ``owner`` is only kept alive, never freed here.
    """

    def __init__(self, ptr: int, n_bytes: int, device_index: int = 0, owner: Any = None) -> None:
        self.ptr = int(ptr)
        self.n_bytes = int(n_bytes)
        self.device_index = device_index
        self.owner = owner

    def __int__(self) -> int:
        return self.ptr


class DeviceBuffer:
    """A typed view of device memory: the backend buffer held by ``DeviceArray``.

    Plays the role of ``pycuda.gpuarray.GPUArray`` in the reference
    (``cuda.py:193-200``): it only knows its pointer, shape and dtype.
    """

    def __init__(self, raw: RawAllocation, shape: Tuple[int, ...], dtype: np.dtype) -> None:
        self.raw = raw
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    @property
    def ptr(self) -> int:
        return self.raw.ptr


class _PinnedBlock:
    """Page-locked host memory; numpy arrays made from it keep it alive (``base``)."""

    def __init__(self, device_index: int, n_bytes: int) -> None:
        ptr = c_void_p()
        _capi.call("ksp_device_set", device_index)
        _capi.call("ksp_host_alloc", byref(ptr), c_size_t(max(n_bytes, 1)))
        self.ptr = ptr.value
        self.n_bytes = max(n_bytes, 1)
        self.__array_interface__ = {
            "shape": (self.n_bytes,),
            "typestr": "|u1",
            "data": (self.ptr, False),
            "version": 3,
        }

    def __del__(self) -> None:
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _capi.load().ksp_host_free(c_void_p(ptr))
            except Exception:
                pass


class Event(AbstractEvent):
    """A marker in a command queue (reference ``cuda.py:71-84``)."""

    def __init__(self, stream: int, device_index: int) -> None:
        handle = c_void_p()
        _capi.call("ksp_device_set", device_index)
        _capi.call("ksp_event_create", byref(handle), 1)
        self._handle = handle.value
        _capi.call("ksp_event_record", c_void_p(self._handle), c_void_p(stream))

    def wait(self) -> None:
        _capi.call("ksp_event_synchronize", c_void_p(self._handle))

    def time_since(self, prior_event: "Event") -> float:
        """Seconds between ``prior_event`` and this event (both are waited for)."""
        prior_event.wait()
        self.wait()
        ms = c_float()
        _capi.call("ksp_event_elapsed_ms", c_void_p(prior_event._handle), c_void_p(self._handle),
                   byref(ms))
        return 1e-3 * ms.value

    def time_till(self, next_event: "Event") -> float:
        return next_event.time_since(self)

    def __del__(self) -> None:
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            try:
                _capi.load().ksp_event_destroy(c_void_p(handle))
            except Exception:
                pass


class Device(AbstractDevice):
    """One CUDA device (reference ``cuda.py:86-158``)."""

    def __init__(self, index: int) -> None:
        self.index = index
        cc_major, cc_minor, sms, warp, l2 = c_int(), c_int(), c_int(), c_int(), c_int()
        total = c_size_t()
        _capi.call("ksp_device_attributes", index, byref(cc_major), byref(cc_minor), byref(sms),
                   byref(warp), byref(total), byref(l2))
        self._cc = (cc_major.value, cc_minor.value)
        self.sm_count = sms.value
        self._warp = warp.value
        self.total_memory = total.value
        self.l2_bytes = l2.value

    def make_context(self) -> "Context":
        return Context(self)

    @property
    def name(self) -> str:
        buf = ctypes.create_string_buffer(256)
        _capi.call("ksp_device_name", self.index, buf, 256)
        return buf.value.decode()

    @property
    def platform_name(self) -> str:
        return "CUDA"

    @property
    def driver_version(self) -> str:
        rt, drv = c_int(), c_int()
        _capi.call("ksp_versions", byref(rt), byref(drv))
        return f"CUDA:{rt.value} Driver:{drv.value}"

    is_cuda = True
    is_gpu = True
    is_accelerator = False
    is_cpu = False

    @property
    def simd_group_size(self) -> int:
        return self._warp

    @property
    def compute_capability(self) -> Tuple[int, int]:
        return self._cc

    @classmethod
    def get_devices(cls) -> List["Device"]:
        count = c_int()
        _capi.call("ksp_device_count", byref(count))
        return [cls(i) for i in range(count.value)]

    @classmethod
    def get_devices_by_platform(cls) -> List[List["Device"]]:
        return [cls.get_devices()]


class Context(AbstractContext):
    """Allocation and queue factory for one device (reference ``cuda.py:163-255``).

    The CUDA runtime's primary context of the device is used, so a ``Context``
    is only a device index; ``with context:`` makes that device current for the
    calling thread as ``push``/``pop`` do in the reference.
    """

    def __init__(self, device: Device) -> None:
        self._device = device
        self._local = threading.local()

    @property
    def device(self) -> Device:
        return self._device

    def _make_current(self) -> None:
        _capi.call("ksp_device_set", self._device.index)

    def allocate_raw(self, n_bytes: int) -> RawAllocation:
        return RawAllocation(self._device.index, n_bytes)

    def allocate(self, shape: Tuple[int, ...], dtype: Any, raw: Optional[RawAllocation] = None
                 ) -> DeviceBuffer:
        dtype = np.dtype(dtype)
        n_bytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if raw is None:
            raw = self.allocate_raw(n_bytes)
        return DeviceBuffer(raw, shape, dtype)

    def allocate_pinned(self, shape: Tuple[int, ...], dtype: Any) -> np.ndarray:
        dtype = np.dtype(dtype)
        n_bytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        block = _PinnedBlock(self._device.index, n_bytes)
        flat = np.asarray(block)[:n_bytes]  # flat.base is the block: it lives as long as any view
        return flat.view(dtype).reshape(shape)

    def create_command_queue(self, profile: bool = False) -> "CommandQueue":
        return CommandQueue(self)

    def create_tuning_command_queue(self) -> "CommandQueue":
        return CommandQueue(self)

    def __enter__(self) -> "Context":
        prev = c_int()
        _capi.call("ksp_device_get", byref(prev))
        stack = getattr(self._local, "stack", None)
        if stack is None:
            stack = self._local.stack = []
        stack.append(prev.value)
        self._make_current()
        return self

    def __exit__(self, *exc) -> None:
        _capi.call("ksp_device_set", self._local.stack.pop())



class CommandQueue(AbstractCommandQueue):
    """An in-order stream (reference ``cuda.py:257-479``)."""

    def __init__(self, context: Context) -> None:
        self.context = context
        handle = c_void_p()
        context._make_current()
        _capi.call("ksp_stream_create", byref(handle))
        self._stream = handle.value

    @property
    def stream(self) -> int:
        """The ``cudaStream_t`` as an integer (pass it as ``void *`` to the C ABI)."""
        return self._stream

    def _current(self) -> None:
        self.context._make_current()

    @staticmethod
    def _host_ptr(data: np.ndarray) -> int:
        if not data.flags.c_contiguous:
            raise ValueError("host array must be C-contiguous")
        return data.ctypes.data

    def enqueue_read_buffer(self, buffer: DeviceBuffer, data: np.ndarray, blocking: bool = True
                            ) -> None:
        """Device -> host copy of the whole buffer."""
        if data.nbytes != buffer.nbytes:
            raise ValueError(f"host array has {data.nbytes} bytes, the buffer {buffer.nbytes}")
        self._current()
        _capi.call("ksp_memcpy_async", c_void_p(self._host_ptr(data)), c_void_p(buffer.ptr),
                   c_size_t(buffer.nbytes), _capi.D2H, c_void_p(self._stream))
        if blocking:
            self.finish()

    def enqueue_write_buffer(self, buffer: DeviceBuffer, data: np.ndarray, blocking: bool = True
                             ) -> None:
        """Host -> device copy of the whole buffer."""
        if data.nbytes != buffer.nbytes:
            raise ValueError(f"host array has {data.nbytes} bytes, the buffer {buffer.nbytes}")
        self._current()
        _capi.call("ksp_memcpy_async", c_void_p(buffer.ptr), c_void_p(self._host_ptr(data)),
                   c_size_t(buffer.nbytes), _capi.H2D, c_void_p(self._stream))
        if blocking:
            self.finish()

    def _rect(self, dst: int, src: int, dst_origin: int, src_origin: int, shape: Sequence[int],
              dst_strides: Sequence[int], src_strides: Sequence[int], kind: int) -> None:
        """Copy a <= 3-D byte rectangle; shape/strides are fastest-axis FIRST, in bytes."""
        self._current()
        shape = list(shape) + [1] * (3 - len(shape))
        dst_strides = list(dst_strides) + [0] * (3 - len(dst_strides))
        src_strides = list(src_strides) + [0] * (3 - len(src_strides))
        if len(shape) > 3:
            raise ValueError("at most 3 dimensions")
        if dst_strides[0] != 1 or src_strides[0] != 1:
            raise ValueError("the fastest axis must be contiguous")
        for z in range(shape[2]):
            d = dst + dst_origin + z * dst_strides[2]
            s = src + src_origin + z * src_strides[2]
            if shape[1] == 1:
                _capi.call("ksp_memcpy_async", c_void_p(d), c_void_p(s), c_size_t(shape[0]), kind,
                           c_void_p(self._stream))
            else:
                _capi.call("ksp_memcpy_2d_async", c_void_p(d), c_size_t(dst_strides[1]),
                           c_void_p(s), c_size_t(src_strides[1]), c_size_t(shape[0]),
                           c_size_t(shape[1]), kind, c_void_p(self._stream))

    def enqueue_copy_buffer_rect(self, src_buffer: DeviceBuffer, dest_buffer: DeviceBuffer,
                                 src_origin: int, dest_origin: int, shape: Sequence[int],
                                 src_strides: Sequence[int], dest_strides: Sequence[int]) -> None:
        self._rect(dest_buffer.ptr, src_buffer.ptr, dest_origin, src_origin, shape, dest_strides,
                   src_strides, _capi.D2D)

    def enqueue_read_buffer_rect(self, buffer: DeviceBuffer, data: np.ndarray, buffer_origin: int,
                                 data_origin: int, shape: Sequence[int],
                                 buffer_strides: Sequence[int], data_strides: Sequence[int],
                                 blocking: bool = True) -> None:
        self._rect(self._host_ptr(data), buffer.ptr, data_origin, buffer_origin, shape,
                   data_strides, buffer_strides, _capi.D2H)
        if blocking:
            self.finish()

    def enqueue_write_buffer_rect(self, buffer: DeviceBuffer, data: np.ndarray, buffer_origin: int,
                                  data_origin: int, shape: Sequence[int],
                                  buffer_strides: Sequence[int], data_strides: Sequence[int],
                                  blocking: bool = True) -> None:
        self._rect(buffer.ptr, self._host_ptr(data), buffer_origin, data_origin, shape,
                   buffer_strides, data_strides, _capi.H2D)
        if blocking:
            self.finish()

    def enqueue_zero_buffer(self, buffer: DeviceBuffer) -> None:
        self._current()
        _capi.call("ksp_memset_async", c_void_p(buffer.ptr), 0, c_size_t(buffer.nbytes),
                   c_void_p(self._stream))

    def enqueue_marker(self) -> Event:
        self._current()
        return Event(self._stream, self.context.device.index)

    def enqueue_wait_for_events(self, events: Sequence[Event]) -> None:
        self._current()
        for event in events:
            _capi.call("ksp_stream_wait_event", c_void_p(self._stream), c_void_p(event._handle))

    def flush(self) -> None:
        pass

    def finish(self) -> None:
        self._current()
        _capi.call("ksp_stream_synchronize", c_void_p(self._stream))

    def __del__(self) -> None:
        stream, self._stream = getattr(self, "_stream", None), None
        if stream:
            try:
                _capi.load().ksp_stream_destroy(c_void_p(stream))
            except Exception:
                pass
