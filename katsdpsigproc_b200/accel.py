"""Operation runtime: device/host arrays, padded dimensions, slots and operations.

Host-side mirror of the parts of the reference's ``katsdpsigproc.accel`` that the
RFI-flagging path uses (reference ``src/katsdpsigproc/accel.py``):

===========================  =====================================
this module                  reference ``accel.py``
===========================  =====================================
``divup`` / ``roundup``      :113-120
``create_some_context`` ...  :211-365
``HostArray``                :368-456
``DeviceArray``              :459-925
``DeviceAllocator``          :1057-1093
``Dimension``                :1115-1294
``IOSlotBase`` / ``IOSlot``  :1297-1502
``CompoundIOSlot``           :1505-1554
``AliasIOSlot``              :1557-1608
``Operation``                :1611-1756
``OperationSequence``        :1759-1835
===========================  =====================================

Names, argument meaning and error behaviour follow the reference so that its
tests (``test/test_accel.py``) read the same against this module.  What is
absent on purpose: ``build()`` and the Mako lexer (kernels are ahead-of-time
sm_100a code behind the C ABI), SVM arrays, the OpenCL backend and the graphviz
visualiser.
"""

from __future__ import annotations

import itertools
import os
import sys
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Any, Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np

from . import cuda

have_cuda = True
have_opencl = False

_Index = Any


def divup(x: int, y: int) -> int:
    """Quotient of ``x / y`` rounded towards +infinity."""
    return -((-x) // y)


def roundup(x: int, y: int) -> int:
    """Smallest multiple of ``y`` that is not less than ``x``."""
    return divup(x, y) * y


# ----------------------------------------------------------------------------- devices
def all_devices() -> List[cuda.Device]:
    return list(cuda.Device.get_devices())


def _env_index(name: str) -> Optional[int]:
    try:
        value = int(os.environ[name])
    except (KeyError, ValueError):
        return None
    return value if value >= 0 else None


def candidate_devices(device_filter: Optional[Callable[[cuda.Device], bool]] = None
                      ) -> Sequence[cuda.Device]:
    """Devices considered by :func:`create_some_context`.

    ``KATSDPSIGPROC_DEVICE`` (index among all devices) or, failing that,
    ``CUDA_DEVICE`` narrows the choice to one device; an out-of-range index
    raises :exc:`RuntimeError`.  ``PYOPENCL_CTX`` is not consulted: there is no
    OpenCL backend.
    """
    devices = all_devices()
    if device_filter is not None:
        devices = [d for d in devices if device_filter(d)]
    chosen = _env_index("KATSDPSIGPROC_DEVICE")
    if chosen is None:
        chosen = _env_index("CUDA_DEVICE")
    if chosen is not None:
        if chosen >= len(devices):
            raise RuntimeError("Out-of-range device selected")
        devices = [devices[chosen]]
    return devices


def create_some_context(interactive: bool = True,
                        device_filter: Optional[Callable[[cuda.Device], bool]] = None
                        ) -> cuda.Context:
    """Create a single-device context, choosing the device automatically (or by prompt)."""
    devices = candidate_devices(device_filter)
    if not devices:
        raise RuntimeError("No compute devices found")
    device = devices[0]
    if interactive and len(devices) > 1 and sys.stdin.isatty():
        print("Select device:")
        for i, dev in enumerate(devices):
            print(f"    [{i}]: {dev.name} ({dev.platform_name})")
        print()
        try:
            choice = int(input("Enter selection: "))
            if choice < 0:
                raise IndexError
            device = devices[choice]
        except (ValueError, IndexError):
            raise RuntimeError("Invalid device number") from None
    return device.make_context()


# ----------------------------------------------------------------------------- host arrays
class HostArray(np.ndarray):
    """C-ordered, optionally padded, optionally pinned host array.

    Only the object returned by the constructor is a *safe* transfer target
    (:meth:`safe`): it is the origin-anchored slice of a contiguous allocation
    of ``padded_shape``.  Views of it lose that guarantee.
    """

    _owner: Optional[np.ndarray]
    padded_shape: Optional[Tuple[int, ...]]

    def __new__(cls, shape: Tuple[int, ...], dtype: Any,
                padded_shape: Optional[Tuple[int, ...]] = None,
                context: Optional[cuda.Context] = None) -> "HostArray":
        shape = tuple(shape)
        padded_shape = shape if padded_shape is None else tuple(padded_shape)
        assert len(padded_shape) == len(shape)
        assert all(p >= s for p, s in zip(padded_shape, shape))
        if context is None:
            storage = np.empty(padded_shape, dtype)
        else:
            storage = context.allocate_pinned(padded_shape, dtype)
        window = storage[tuple(slice(0, n) for n in shape)] if shape else storage
        self = window.view(cls)
        self._owner = storage
        self.padded_shape = padded_shape
        return self

    def __array_finalize__(self, obj: Any) -> None:
        if obj is None:
            return
        # a view or a template copy: nobody vouches for its anchoring
        self._owner = None
        self.padded_shape = obj.padded_shape if isinstance(obj, HostArray) else None

    @classmethod
    def safe(cls, obj: np.ndarray) -> bool:
        """Can ``obj`` be used directly in a transfer?"""
        return getattr(obj, "_owner", None) is not None

    @classmethod
    def padded_view(cls, obj: np.ndarray) -> Optional[np.ndarray]:
        """The whole padded allocation behind ``obj`` (``None`` if not :meth:`safe`)."""
        return getattr(obj, "_owner", None)


# ----------------------------------------------------------------------------- device arrays
def _index_to_rect(index: _Index, shape: Sequence[int], strides: Sequence[int]
                   ) -> Tuple[int, Tuple[int, ...], Tuple[int, ...]]:
    """Resolve a (restricted) numpy index expression to (byte origin, shape, byte strides)."""
    if not isinstance(index, tuple):
        index = (index,)
    origin = 0
    out_shape: List[int] = []
    out_strides: List[int] = []
    axis = 0
    for item in index:
        if item is np.newaxis:
            out_shape.append(1)
            out_strides.append(0)
            continue
        if not isinstance(item, (slice, int, np.integer)):
            raise TypeError(f"Invalid type in slice: {type(item)}")
        if axis >= len(shape):
            raise IndexError("Too many axes in index expression")
        if isinstance(item, slice):
            start, stop, step = item.indices(shape[axis])
            if step <= 0:
                raise IndexError("Only positive strides are supported")
            count = (stop - start) // step
            if count <= 0:
                raise IndexError("Empty slice selection")
            origin += start * strides[axis]
            out_shape.append(count)
            out_strides.append(step * strides[axis])
        else:
            pos = int(item)
            if pos < 0:
                pos += shape[axis]
            if not 0 <= pos < shape[axis]:
                raise IndexError("Index out of range")
            origin += pos * strides[axis]
        axis += 1
    out_shape.extend(shape[axis:])
    out_strides.extend(strides[axis:])
    return origin, tuple(out_shape), tuple(out_strides)


class DeviceArray:
    """C-order device array whose rows may be padded (reference ``accel.py:459-925``)."""

    def __init__(self, context: cuda.Context, shape: Tuple[int, ...], dtype: Any,
                 padded_shape: Optional[Tuple[int, ...]] = None, raw: Any = None) -> None:
        shape = tuple(shape)
        padded_shape = shape if padded_shape is None else tuple(padded_shape)
        assert len(shape) == len(padded_shape)
        assert all(p >= s for p, s in zip(padded_shape, shape))
        self._shape = shape
        self._dtype = np.dtype(dtype)
        self.padded_shape = padded_shape
        self.context = context
        self.buffer = context.allocate(padded_shape, dtype, raw)

    @classmethod
    def wrap(cls, context: cuda.Context, ptr: int, shape: Tuple[int, ...], dtype: Any,
             padded_shape: Optional[Tuple[int, ...]] = None, owner: Any = None) -> "DeviceArray":
        """A ``DeviceArray`` over device memory that already exists (extension: zero-copy hand-over
        from a GPU producer, e.g. ``DeviceArray.wrap(ctx, t.data_ptr(), t.shape, np.complex64,
        owner=t)`` for a torch tensor).  ``owner`` is kept alive as long as the array."""
        padded = tuple(shape) if padded_shape is None else tuple(padded_shape)
        n_bytes = int(np.prod(padded, dtype=np.int64)) * np.dtype(dtype).itemsize
        raw = cuda.ExternalAllocation(ptr, n_bytes, context.device.index, owner)
        return cls(context, shape, dtype, padded, raw=raw)

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    @property
    def dtype(self) -> np.dtype:
        return self._dtype

    @property
    def ndim(self) -> int:
        return len(self._shape)

    @property
    def strides(self) -> Tuple[int, ...]:
        """Byte strides, as numpy would report them for the padded allocation."""
        strides = []
        step = self._dtype.itemsize
        for extent in reversed(self.padded_shape):
            strides.append(step)
            step *= extent
        return tuple(reversed(strides))

    @property
    def ptr(self) -> int:
        """Device address (what the C ABI takes)."""
        return self.buffer.ptr

    # -- whole-array transfers
    def _copyable(self, ary: np.ndarray) -> bool:
        return (HostArray.safe(ary) and ary.dtype == self.dtype and ary.shape == self.shape
                and ary.padded_shape == self.padded_shape)  # type: ignore[attr-defined]

    def empty_like(self) -> HostArray:
        """A pinned host array with this array's shape, dtype and padding."""
        return HostArray(self.shape, self.dtype, self.padded_shape, context=self.context)

    def asarray_like(self, ary: np.ndarray) -> HostArray:
        """``ary`` itself if it can be transferred directly, else a staged copy of it."""
        assert ary.shape == self.shape
        if self._copyable(ary):
            return ary  # type: ignore[return-value]
        staged = self.empty_like()
        np.copyto(staged, ary, casting="no")
        return staged

    def set(self, command_queue: cuda.CommandQueue, ary: np.ndarray) -> None:
        """Synchronous host -> device copy."""
        staged = self.asarray_like(ary)
        command_queue.enqueue_write_buffer(self.buffer, HostArray.padded_view(staged))

    def set_async(self, command_queue: cuda.CommandQueue, ary: np.ndarray) -> None:
        staged = self.asarray_like(ary)
        command_queue.enqueue_write_buffer(self.buffer, HostArray.padded_view(staged),
                                           blocking=False)

    def _target(self, ary: Optional[np.ndarray]) -> HostArray:
        if ary is None or not self._copyable(ary):
            return self.empty_like()
        return ary  # type: ignore[return-value]

    def get(self, command_queue: cuda.CommandQueue, ary: Optional[np.ndarray] = None
            ) -> np.ndarray:
        """Synchronous device -> host copy; returns the array actually written."""
        target = self._target(ary)
        command_queue.enqueue_read_buffer(self.buffer, HostArray.padded_view(target))
        return target

    def get_async(self, command_queue: cuda.CommandQueue, ary: Optional[np.ndarray] = None
                  ) -> np.ndarray:
        target = self._target(ary)
        command_queue.enqueue_read_buffer(self.buffer, HostArray.padded_view(target),
                                          blocking=False)
        return target

    # -- region transfers
    @classmethod
    def _canonical_slice(cls, region: _Index, shape: Sequence[int], strides: Sequence[int]):
        return _index_to_rect(region, shape, strides)

    @classmethod
    def _region_transfer_params(cls, src: Any, dest: Any, src_region: _Index, dest_region: _Index):
        """(src origin, dest origin, shape, src strides, dest strides), all in bytes with the
        fastest axis FIRST and contiguous runs merged."""
        if src.dtype != dest.dtype:
            raise TypeError(f"dtypes do not match ({src.dtype} and {dest.dtype})")
        s_origin, s_shape, s_strides = _index_to_rect(src_region, src.shape, src.strides)
        d_origin, d_shape, d_strides = _index_to_rect(dest_region, dest.shape, dest.strides)
        if s_shape != d_shape:
            raise ValueError("Source and destination shapes for the copy do not match")
        shape = [src.dtype.itemsize]
        s_out, d_out = [1], [1]
        for extent, ss, ds in zip(reversed(s_shape), reversed(s_strides), reversed(d_strides)):
            if extent == 1:
                continue
            if ss == shape[-1] * s_out[-1] and ds == shape[-1] * d_out[-1]:
                shape[-1] *= extent          # this axis continues the run below it
            else:
                shape.append(extent)
                s_out.append(ss)
                d_out.append(ds)
        return s_origin, d_origin, tuple(shape), tuple(s_out), tuple(d_out)

    @classmethod
    def _transfer_region(cls, func: Callable[..., None], buffer1: Any, buffer2: Any, origin1: int,
                         origin2: int, shape: Tuple[int, ...], strides1: Tuple[int, ...],
                         strides2: Tuple[int, ...], **kwargs: Any) -> None:
        if len(shape) <= 3:
            func(buffer1, buffer2, origin1, origin2, shape, strides1, strides2, **kwargs)
            return
        for i in range(shape[-1]):       # peel the slowest axis
            cls._transfer_region(func, buffer1, buffer2, origin1 + i * strides1[-1],
                                 origin2 + i * strides2[-1], shape[:-1], strides1[:-1],
                                 strides2[:-1], **kwargs)

    def copy_region(self, command_queue: cuda.CommandQueue, dest: "DeviceArray",
                    src_region: _Index, dest_region: _Index) -> None:
        """Device -> device copy of a sub-region (ints, positive-step slices, ``np.newaxis``)."""
        so, do, shape, ss, ds = self._region_transfer_params(self, dest, src_region, dest_region)
        self._transfer_region(command_queue.enqueue_copy_buffer_rect, self.buffer, dest.buffer,
                              so, do, shape, ss, ds)

    def get_region(self, command_queue: cuda.CommandQueue, ary: np.ndarray, device_region: _Index,
                   ary_region: _Index, blocking: bool = True) -> None:
        if not HostArray.safe(ary):
            raise ValueError("Target region is not suitable for device-to-host copy")
        do, ao, shape, ds, as_ = self._region_transfer_params(self, ary, device_region, ary_region)
        self._transfer_region(command_queue.enqueue_read_buffer_rect, self.buffer,
                              HostArray.padded_view(ary), do, ao, shape, ds, as_,
                              blocking=blocking)

    def set_region(self, command_queue: cuda.CommandQueue, ary: np.ndarray, device_region: _Index,
                   ary_region: _Index, blocking: bool = True) -> None:
        if not HostArray.safe(ary):
            piece = ary[ary_region]
            staged = HostArray(piece.shape, ary.dtype, context=self.context)
            np.copyto(staged, piece, casting="no")
            ary, ary_region = staged, np.s_[()]
        ao, do, shape, as_, ds = self._region_transfer_params(ary, self, ary_region, device_region)
        self._transfer_region(command_queue.enqueue_write_buffer_rect, self.buffer,
                              HostArray.padded_view(ary), do, ao, shape, ds, as_,
                              blocking=blocking)

    def zero(self, command_queue: cuda.CommandQueue) -> None:
        """Asynchronous memset of the whole (padded) allocation."""
        command_queue.enqueue_zero_buffer(self.buffer)


class AbstractAllocator(ABC):
    context: Any

    @abstractmethod
    def allocate(self, shape: Tuple[int, ...], dtype: Any,
                 padded_shape: Optional[Tuple[int, ...]] = None, raw: Any = None) -> DeviceArray:
        ...

    @abstractmethod
    def allocate_raw(self, n_bytes: int) -> Any:
        ...


class DeviceAllocator(AbstractAllocator):
    """Allocates :class:`DeviceArray` objects from a context."""

    def __init__(self, context: cuda.Context) -> None:
        self.context = context

    def allocate(self, shape, dtype, padded_shape=None, raw=None) -> DeviceArray:
        return DeviceArray(self.context, shape, dtype, padded_shape, raw)

    def allocate_raw(self, n_bytes: int) -> Any:
        return self.context.allocate_raw(n_bytes)


# ----------------------------------------------------------------------------- dimensions
def _is_power2(value: int) -> bool:
    return value > 0 and value & (value - 1) == 0


class _Requirement:
    """The shared state of a set of linked dimensions."""

    __slots__ = ("size", "min_padded_size", "alignment", "alignment_hint", "exact", "frozen")

    def __init__(self, size: int, min_padded_size: int, alignment: int, exact: bool) -> None:
        self.size = size
        self.min_padded_size = min_padded_size
        self.alignment = alignment
        self.alignment_hint = alignment
        self.exact = exact
        self.frozen = False

    def required(self) -> int:
        padded = roundup(self.min_padded_size, self.alignment)
        # a dimension smaller than the hint is not worth blowing up to it
        if not self.exact and padded >= self.alignment_hint:
            padded = roundup(padded, self.alignment_hint)
        return padded

    def valid(self, padded: int) -> bool:
        if self.exact:
            return padded == self.required()
        return padded >= self.min_padded_size and padded % self.alignment == 0


class Dimension:
    """Padding/alignment requirements of one axis; linkable (union-find) and freezable.

    ``min_padded_round`` only sets a minimum (``roundup(size, round)``); ``alignment``
    (a power of two) constrains every acceptable padded size; ``align_dtype`` is a
    hint that rows of that dtype should start on ``ALIGN_BYTES`` boundaries;
    ``exact`` forbids padding.
    """

    ALIGN_BYTES = 128
    _is_power2 = staticmethod(_is_power2)

    def __init__(self, size: int, min_padded_round: Optional[int] = None,
                 min_padded_size: Optional[int] = None, alignment: int = 1,
                 align_dtype: Any = None, exact: bool = False) -> None:
        if min_padded_size is None:
            min_padded_size = size if min_padded_round is None else roundup(size, min_padded_round)
        if not _is_power2(alignment):
            raise ValueError("alignment is not a power of 2")
        if min_padded_size < size:
            raise ValueError("padded size is less than size")
        self._up: Optional[Dimension] = None
        self._req: Optional[_Requirement] = _Requirement(size, min_padded_size, alignment, exact)
        if align_dtype is not None:
            self.add_align_dtype(align_dtype)

    def _root(self) -> "Dimension":
        node = self
        while node._up is not None:
            node = node._up
        # path compression
        walk = self
        while walk._up is not None:
            walk._up, walk = node, walk._up
        return node

    def _state(self) -> _Requirement:
        req = self._root()._req
        assert req is not None
        return req

    @property
    def size(self) -> int:
        return self._state().size

    @property
    def min_padded_size(self) -> int:
        return self._state().min_padded_size

    @property
    def alignment(self) -> int:
        return self._state().alignment

    @property
    def alignment_hint(self) -> int:
        return self._state().alignment_hint

    @property
    def exact(self) -> bool:
        return self._state().exact

    @property
    def frozen(self) -> bool:
        return self._state().frozen

    def required_padded_size(self) -> int:
        return self._state().required()

    def valid(self, padded_size: int) -> bool:
        return self._state().valid(padded_size)

    def add_align_dtype(self, dtype: Any) -> None:
        req = self._state()
        if req.frozen:
            raise ValueError("cannot modify a frozen requirement")
        itemsize = np.dtype(dtype).itemsize
        if _is_power2(itemsize):
            req.alignment_hint = max(req.alignment_hint, self.ALIGN_BYTES // itemsize)

    def link(self, other: "Dimension") -> None:
        """Merge the requirements of ``self`` and ``other`` (and everything linked to them).

        Raises :exc:`ValueError` -- leaving both untouched -- if either is frozen, the
        sizes differ, or an exact requirement cannot be met.
        """
        mine, theirs = self._root(), other._root()
        a, b = mine._req, theirs._req
        assert a is not None and b is not None
        if a.frozen or b.frozen:
            raise ValueError("cannot link frozen requirements")
        if mine is theirs:
            return
        if a.size != b.size:
            raise ValueError("sizes are incompatible")
        if a.exact and not b.valid(a.required()):
            raise ValueError("linked requirement is unsatisfiable")
        if b.exact and not b.valid(b.required()):
            raise ValueError("linked requirement is unsatisfiable")
        a.min_padded_size = max(a.min_padded_size, b.min_padded_size)
        a.alignment = max(a.alignment, b.alignment)
        a.alignment_hint = max(a.alignment_hint, b.alignment_hint)
        a.exact = a.exact or b.exact
        theirs._up = mine
        theirs._req = None

    def freeze(self) -> None:
        self._state().frozen = True


# ----------------------------------------------------------------------------- slots
class IOSlotBase(ABC):
    """Input/output slot of an operation; slots form trees that share storage and only a
    root may be bound or allocated."""

    def __init__(self) -> None:
        self.is_root = True

    def check_root(self) -> None:
        if not self.is_root:
            raise ValueError("not a root slot")

    @abstractmethod
    def required_bytes(self) -> int:
        ...

    @abstractmethod
    def is_bound(self) -> bool:
        ...

    def attachable(self) -> bool:
        return self.is_root and not self.is_bound()

    @abstractmethod
    def _allocate(self, allocator: AbstractAllocator, raw: Any = None, *, bind: bool) -> Any:
        ...

    def allocate(self, allocator: AbstractAllocator, raw: Any = None, *, bind: bool = True) -> Any:
        """Allocate (and by default bind) storage that meets the requirements."""
        self.check_root()
        return self._allocate(allocator, raw, bind=bind)

    @abstractmethod
    def _allocate_host(self, context: Any) -> HostArray:
        ...

    def allocate_host(self, context: Any) -> HostArray:
        self.check_root()
        return self._allocate_host(context)


class IOSlot(IOSlotBase):
    """Typed, shaped slot; each axis is a :class:`Dimension` (ints are wrapped)."""

    def __init__(self, dimensions: Tuple[Union[Dimension, int], ...], dtype: Any) -> None:
        super().__init__()
        self.dimensions = tuple(d if isinstance(d, Dimension) else Dimension(d)
                                for d in dimensions)
        self.shape = tuple(d.size for d in self.dimensions)
        self.dtype = np.dtype(dtype)
        if len(self.dimensions) > 1:
            self.dimensions[-1].add_align_dtype(self.dtype)
        self.buffer: Optional[DeviceArray] = None

    def is_bound(self) -> bool:
        return self.buffer is not None

    def validate(self, buffer: DeviceArray) -> None:
        if buffer.dtype != self.dtype:
            raise TypeError("dtype does not match")
        if len(buffer.shape) != len(self.shape):
            raise ValueError("number of dimensions does not match")
        for size, padded, dim in zip(buffer.shape, buffer.padded_shape, self.dimensions):
            if size != dim.size:
                raise ValueError("size does not match")
            if padded != dim.required_padded_size():
                raise ValueError("padded size does not match")

    def _bind(self, buffer: Optional[DeviceArray]) -> None:
        if buffer is not None:
            self.validate(buffer)
        self.buffer = buffer
        for dim in self.dimensions:
            dim.freeze()

    def bind(self, buffer: Optional[DeviceArray]) -> None:
        """Attach ``buffer`` (validated unless ``None``); freezes the dimensions."""
        self.check_root()
        self._bind(buffer)

    def required_padded_shape(self) -> Tuple[int, ...]:
        return tuple(d.required_padded_size() for d in self.dimensions)

    def required_bytes(self) -> int:
        return int(np.prod(self.required_padded_shape(), dtype=np.int64)) * self.dtype.itemsize

    def _allocate(self, allocator: AbstractAllocator, raw: Any = None, *, bind: bool = True
                  ) -> DeviceArray:
        buffer = allocator.allocate(self.shape, self.dtype, self.required_padded_shape(), raw=raw)
        if bind:
            self._bind(buffer)
        return buffer

    def _allocate_host(self, context: Any) -> HostArray:
        return HostArray(self.shape, self.dtype, self.required_padded_shape(), context=context)


class CompoundIOSlot(IOSlot):
    """One buffer feeding several child slots of equal shape and dtype; its requirements are
    the union of theirs."""

    def __init__(self, children: Iterable[IOSlot]) -> None:
        self.children = list(children)
        if not self.children:
            raise ValueError("empty child list")
        first = self.children[0]
        for child in self.children:
            if not child.attachable():
                raise ValueError("child is not attachable")
            if child.shape != first.shape:
                raise ValueError("inconsistent shapes")
            if child.dtype != first.dtype:
                raise TypeError("inconsistent dtypes")
            if any(dim.frozen for dim in child.dimensions):
                raise ValueError("child has frozen dimensions")
        for child in self.children:
            for mine, theirs in zip(first.dimensions, child.dimensions):
                mine.link(theirs)
        super().__init__(first.dimensions, first.dtype)
        for child in self.children:
            child.is_root = False

    def _bind(self, buffer: Optional[DeviceArray]) -> None:
        super()._bind(buffer)
        for child in self.children:
            child._bind(buffer)


class AliasIOSlot(IOSlotBase):
    """One raw allocation backing several (differently typed) children that are never live at
    the same time."""

    def __init__(self, children: Iterable[IOSlotBase]) -> None:
        super().__init__()
        self.children = list(children)
        self.raw: Any = None
        if not self.children:
            raise ValueError("empty child list")
        if not all(child.attachable() for child in self.children):
            raise ValueError("child is not attachable")
        for child in self.children:
            child.is_root = False

    def is_bound(self) -> bool:
        return self.raw is not None

    def required_bytes(self) -> int:
        return max(child.required_bytes() for child in self.children)

    def _allocate_host(self, context: Any) -> HostArray:
        return HostArray((self.required_bytes(),), np.uint8, context=context)

    def _allocate(self, allocator: AbstractAllocator, raw: Any = None, *, bind: bool = True) -> Any:
        if raw is None:
            raw = allocator.allocate_raw(self.required_bytes())
        if bind:
            for child in self.children:
                child._allocate(allocator, raw, bind=True)
            self.raw = raw
        return raw


# ----------------------------------------------------------------------------- operations
class Operation(ABC):
    """A device operation with named slots; subclasses fill ``slots`` and implement ``_run``."""

    def __init__(self, command_queue: cuda.CommandQueue,
                 allocator: Optional[AbstractAllocator] = None) -> None:
        if allocator is None:
            allocator = DeviceAllocator(command_queue.context)
        elif allocator.context is not None and allocator.context is not command_queue.context:
            raise ValueError("command_queue and allocator have different contexts")
        self.slots: Dict[str, IOSlotBase] = {}
        self.hidden_slots: Dict[str, IOSlotBase] = {}
        self.command_queue = command_queue
        self.allocator = allocator
        self.is_root = True

    def bind(self, **kwargs: Optional[DeviceArray]) -> None:
        for name, buffer in kwargs.items():
            slot = self.slots[name]
            if not isinstance(slot, IOSlot):
                raise TypeError(f"Slot {slot} is not an IOSlot")
            slot.bind(buffer)

    def ensure_bound(self, name: str) -> None:
        slot = self.slots[name]
        if not slot.is_bound():
            slot.allocate(self.allocator)

    def ensure_all_bound(self) -> None:
        for slot in self.slots.values():
            if not slot.is_bound():
                slot.allocate(self.allocator)

    def buffer(self, name: str) -> DeviceArray:
        slot = self.slots.get(name)
        if slot is None:
            slot = self.hidden_slots.get(name)
        if slot is None:
            raise KeyError("no slot named " + name)
        if not isinstance(slot, IOSlot):
            raise TypeError("slot " + name + " is an alias slot")
        if slot.buffer is None:
            raise ValueError("slot " + name + " has no buffer bound")
        return slot.buffer

    def required_bytes(self) -> int:
        return sum(slot.required_bytes() for slot in self.slots.values())

    def parameters(self) -> Mapping[str, Any]:
        return {}

    @abstractmethod
    def _run(self) -> Any:
        ...

    def __call__(self, **kwargs: Optional[DeviceArray]) -> Any:
        self.bind(**kwargs)
        self.ensure_all_bound()
        return self._run()


class OperationSequence(Operation):
    """Named child operations run in order, with their slots re-exported as ``child:slot`` and
    optionally merged (``compounds``) or overlaid in memory (``aliases``)."""

    def __init__(self, command_queue: cuda.CommandQueue,
                 operations: Iterable[Tuple[str, Operation]],
                 compounds: Optional[Mapping[str, Iterable[str]]] = None,
                 aliases: Optional[Mapping[str, Iterable[str]]] = None,
                 allocator: Optional[AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.operations: "OrderedDict[str, Operation]" = OrderedDict(operations)
        for op_name, op in self.operations.items():
            if op.command_queue is not command_queue:
                raise ValueError("child has a different command queue to the parent")
            if not op.is_root:
                raise ValueError("child already has another parent")
            for slot_name, slot in op.slots.items():
                self.slots[f"{op_name}:{slot_name}"] = slot
        for name, members in (compounds or {}).items():
            found = self._extract_slots(members, False)
            if found:
                if not all(isinstance(slot, IOSlot) for slot in found):
                    raise TypeError(f"Children of {name} must all be IOSlots")
                self.slots[name] = CompoundIOSlot(found)  # type: ignore[arg-type]
        for name, members in (aliases or {}).items():
            found = self._extract_slots(members, True)
            if found:
                self.slots[name] = AliasIOSlot(found)
        for op in self.operations.values():
            op.is_root = False

    def _extract_slots(self, names: Iterable[str], add_to_hidden: bool) -> List[IOSlotBase]:
        taken: List[IOSlotBase] = []
        for name in names:
            slot = self.slots.pop(name, None)
            if slot is None:
                continue          # names that do not exist are skipped
            taken.append(slot)
            if add_to_hidden:
                assert name not in self.hidden_slots
                self.hidden_slots[name] = slot
        return taken

    def _run(self) -> None:
        for op in self.operations.values():
            op()
