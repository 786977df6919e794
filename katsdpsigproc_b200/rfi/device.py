"""RFI flagging on the device: background, noise estimate, thresholds and the flagger.

Mirror of the reference's ``src/katsdpsigproc/rfi/device.py`` -- same template /
``instantiate`` / slot names -- over the sm_100a kernels of ``libksp_b200.so``:

==================================  ==========================  =======================
class                               reference ``rfi/device.py``  C-ABI entry point
==================================  ==========================  =======================
BackgroundMedianFilterDevice        :141-333                    ksp_background_median_filter
NoiseEstMADDevice                   :363-472                    ksp_madnz
NoiseEstMADTDevice                  :475-614                    ksp_madnz_t
ThresholdSimpleDevice               :654-809                    ksp_threshold_simple
ThresholdSumDevice                  :812-995                    ksp_threshold_sum
FlaggerDevice                       :998-1166                   the five above, or ksp_flagger
*HostFromDevice                     :113-138,336-360,617-651,   --
                                    1169-1222
==================================  ==========================  =======================

Arithmetic follows the reference's HOST classes (``rfi/host.py``), which is what the
north star asks to match, not the float32 shortcuts of the reference's device
kernels; DESIGN.md lists the differences (even-count medians, threshold formula,
band edges).

Data layout: all arrays are C-order with padded rows.  "Transposed" operations
(``transposed = True``) take baseline-major arrays (baselines x channels).
"""

from __future__ import annotations

import enum
from abc import ABC, abstractmethod
import ctypes
from ctypes import byref, c_double, c_size_t
from typing import Any, List, Mapping, Optional, Tuple, Union

import numpy as np

from .. import _capi, accel, transpose, tune
from .._launch import FixedTuning, launch, ptr
from . import host

DEFAULT_THRESHOLD_FALLOFF = 1.2


class BackgroundFlags(enum.Enum):
    NONE = 0
    CHANNEL = 1
    FULL = 2

    def __bool__(self) -> bool:
        return self is not BackgroundFlags.NONE


# ----------------------------------------------------------------------------- interfaces
class AbstractBackgroundDevice(accel.Operation):
    pass


class AbstractBackgroundDeviceTemplate(ABC):
    context: Any
    use_flags: BackgroundFlags

    @abstractmethod
    def instantiate(self, command_queue: Any, channels: int, baselines: int, *,
                    allocator: Optional[accel.AbstractAllocator] = None
                    ) -> AbstractBackgroundDevice:
        ...


class AbstractNoiseEstDevice(accel.Operation):
    transposed: bool


class AbstractNoiseEstDeviceTemplate(ABC):
    context: Any
    transposed: bool

    @abstractmethod
    def instantiate(self, command_queue: Any, channels: int, baselines: int, *,
                    allocator: Optional[accel.AbstractAllocator] = None) -> AbstractNoiseEstDevice:
        ...


class AbstractThresholdDevice(accel.Operation):
    transposed: bool


class AbstractThresholdDeviceTemplate(ABC):
    context: Any
    transposed: bool

    @abstractmethod
    def instantiate(self, command_queue: Any, channels: int, baselines: int, *args: Any,
                    allocator: Optional[accel.AbstractAllocator] = None, **kwargs: Any
                    ) -> AbstractThresholdDevice:
        ...


# ----------------------------------------------------------------------------- background
class BackgroundHostFromDevice(host.AbstractBackgroundHost):
    """numpy in, numpy out around a background template (reference :113-138)."""

    def __init__(self, template: AbstractBackgroundDeviceTemplate, command_queue: Any) -> None:
        self.template = template
        self.command_queue = command_queue

    def __call__(self, vis: np.ndarray, flags: Optional[np.ndarray] = None) -> np.ndarray:
        if flags is not None and not self.template.use_flags:
            raise TypeError("flags were provided but not included in the template")
        if flags is None and self.template.use_flags:
            raise TypeError("flags were expected but not provided")
        channels, baselines = vis.shape
        fn = self.template.instantiate(self.command_queue, channels, baselines)
        fn.ensure_all_bound()
        fn.buffer("vis").set(self.command_queue, vis)
        if flags is not None:
            fn.buffer("flags").set(self.command_queue, flags)
        fn()
        return fn.buffer("deviations").get(self.command_queue)


class BackgroundMedianFilterDeviceTemplate(FixedTuning, AbstractBackgroundDeviceTemplate):
    """Sliding median along channels; deviations = amplitude - median.

    Parameters
    ----------
    context
        Context the operation will run in
    width
        Window width in channels (odd, at most 63; 13 has a dedicated kernel)
    is_amplitude
        Input is float32 amplitudes rather than complex64 visibilities
    use_flags
        :class:`BackgroundFlags` (or bool: ``True`` means ``CHANNEL``): flagged
        samples take no part in any median and produce a zero deviation
    tuning
        Accepted and ignored
    abs_mode
        How complex amplitudes are rounded: ``_capi.ABS_NUMPY`` reproduces
        ``np.abs`` of numpy on AVX-512F hosts, ``_capi.ABS_HYPOT`` is the correctly
        rounded ``hypot`` of other hosts (extension; default from
        ``KATSDPSIGPROC_B200_ABS_MODE``)
    """

    host_class = host.ReferenceHostClass("BackgroundMedianFilterHost")   # reference rfi/device.py: same attribute

    _TUNING = {"wgs": 128, "csplit": 4}

    @classmethod
    @tune.autotuner(test={"wgs": 128, "csplit": 4})
    def autotune(cls, context: Any, width: int, is_amplitude: bool, use_flags: "BackgroundFlags") -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, width: int, is_amplitude: bool = False,
                 use_flags: Union[BackgroundFlags, bool] = BackgroundFlags.NONE,
                 tuning: Optional[Mapping[str, Any]] = None,
                 abs_mode: Optional[int] = None) -> None:
        if use_flags is True:
            use_flags = BackgroundFlags.CHANNEL
        elif use_flags is False:
            use_flags = BackgroundFlags.NONE
        if not isinstance(use_flags, BackgroundFlags):
            raise TypeError("use_flags must be an instance of BackgroundFlags or bool")
        if width < 1 or width % 2 == 0:
            raise ValueError("width must be odd and positive")
        if width > _capi.MAX_WIDTH:
            raise ValueError(f"width must be at most {_capi.MAX_WIDTH}")
        self.context = context
        self.width = width
        self.is_amplitude = is_amplitude
        self.use_flags = use_flags
        self.abs_mode = _capi.default_abs_mode() if abs_mode is None else abs_mode
        self._init_tuning(context, tuning, width, is_amplitude, use_flags)

    def instantiate(self, command_queue: Any, channels: int, baselines: int,
                    allocator: Optional[accel.AbstractAllocator] = None
                    ) -> "BackgroundMedianFilterDevice":
        return BackgroundMedianFilterDevice(self, command_queue, channels, baselines, allocator)


class BackgroundMedianFilterDevice(AbstractBackgroundDevice):
    """Slots: **vis** (channels x baselines, complex64 or float32), **flags** (channels x
    baselines or channels, uint8; only with ``use_flags``), **deviations** (float32)."""

    def __init__(self, template: BackgroundMedianFilterDeviceTemplate, command_queue: Any,
                 channels: int, baselines: int,
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.channels = channels
        self.baselines = baselines
        vis_type = np.float32 if template.is_amplitude else np.complex64
        # one shared Dimension: vis, deviations and full flags get the same row stride
        dims = (channels, accel.Dimension(baselines))
        self.slots["vis"] = accel.IOSlot(dims, vis_type)
        self.slots["deviations"] = accel.IOSlot(dims, np.float32)
        if template.use_flags is BackgroundFlags.FULL:
            self.slots["flags"] = accel.IOSlot(dims, np.uint8)
        elif template.use_flags is BackgroundFlags.CHANNEL:
            self.slots["flags"] = accel.IOSlot((channels,), np.uint8)

    def _run(self) -> None:
        vis = self.buffer("vis")
        deviations = self.buffer("deviations")
        flags = self.buffer("flags") if self.template.use_flags else None
        flags_stride = flags.padded_shape[1] if flags is not None and flags.ndim == 2 else 0
        launch(self.command_queue, "ksp_background_median_filter", ptr(vis), ptr(deviations),
               ptr(flags) if flags is not None else None, self.channels, self.baselines,
               vis.padded_shape[1], deviations.padded_shape[1], flags_stride, self.template.width,
               int(self.template.is_amplitude), self.template.use_flags.value,
               self.template.abs_mode)

    def parameters(self) -> Mapping[str, Any]:
        return {
            "width": self.template.width,
            "use_flags": self.template.use_flags.name,
            "channels": self.channels,
            "baselines": self.baselines,
        }


# ----------------------------------------------------------------------------- noise
class NoiseEstHostFromDevice(host.AbstractNoiseEstHost):
    """numpy in, numpy out around a noise-estimate template (reference :336-360)."""

    def __init__(self, template: AbstractNoiseEstDeviceTemplate, command_queue: Any) -> None:
        self.template = template
        self.command_queue = command_queue

    def __call__(self, deviations: np.ndarray) -> np.ndarray:
        channels, baselines = deviations.shape
        if self.template.transposed:
            deviations = deviations.T
        fn = self.template.instantiate(self.command_queue, channels, baselines)
        fn.ensure_all_bound()
        fn.buffer("deviations").set(self.command_queue, deviations)
        fn()
        return fn.buffer("noise").get(self.command_queue)


class NoiseEstMADDeviceTemplate(FixedTuning, AbstractNoiseEstDeviceTemplate):
    """``noise = 1.4826 * median(|deviations| != 0)`` per baseline, channel-major input."""

    host_class = host.ReferenceHostClass("NoiseEstMADHost")   # reference rfi/device.py: same attribute

    transposed = False
    _TUNING = {"wgsx": 32, "wgsy": 32}

    @classmethod
    @tune.autotuner(test={"wgsx": 32, "wgsy": 32})
    def autotune(cls, context: Any) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self._init_tuning(context, tuning)

    def instantiate(self, command_queue: Any, channels: int, baselines: int,
                    allocator: Optional[accel.AbstractAllocator] = None) -> "NoiseEstMADDevice":
        return NoiseEstMADDevice(self, command_queue, channels, baselines, allocator)


class NoiseEstMADDevice(AbstractNoiseEstDevice):
    """Slots: **deviations** (channels x baselines, float32), **noise** (baselines, float32) and,
    for ``channels >= TRANSPOSE_FROM``, **scratch_t** (baselines x channels, float32).

    Long rows go through the baseline-major kernel: one tiled transposition into ``scratch_t``
    (0.35 ms for 32768 x 8320) and the one-pass streaming select (0.26 ms), against 1.21 ms for
    the channel-major kernel, which needs three passes over device memory (one radix digit of 32
    baselines' keys per pass is what fits the shared memory).  Short rows use the channel-major
    kernel directly.  The result is the same selection either way.
    """

    transposed = False
    TRANSPOSE_FROM = 2048

    def __init__(self, template: NoiseEstMADDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.channels = channels
        self.baselines = baselines
        self.slots["noise"] = accel.IOSlot((baselines,), np.float32)
        self.slots["deviations"] = accel.IOSlot((channels, accel.Dimension(baselines)), np.float32)
        self.via_transpose = channels >= self.TRANSPOSE_FROM
        if self.via_transpose:
            self.slots["scratch_t"] = accel.IOSlot((baselines, accel.Dimension(channels)), np.float32)

    def _run(self) -> None:
        deviations = self.buffer("deviations")
        if self.via_transpose:
            scratch_t = self.buffer("scratch_t")
            launch(self.command_queue, "ksp_transpose", ptr(scratch_t), ptr(deviations), self.channels,
                   self.baselines, scratch_t.padded_shape[1], deviations.padded_shape[1], 4)
            launch(self.command_queue, "ksp_madnz_t", ptr(scratch_t), ptr(self.buffer("noise")),
                   self.channels, self.baselines, scratch_t.padded_shape[1])
            return
        launch(self.command_queue, "ksp_madnz", ptr(deviations), ptr(self.buffer("noise")),
               self.channels, self.baselines, deviations.padded_shape[1])

    def parameters(self) -> Mapping[str, Any]:
        return {"channels": self.channels, "baselines": self.baselines}


class NoiseEstMADTDeviceTemplate(FixedTuning, AbstractNoiseEstDeviceTemplate):
    """Same statistic on baseline-major input (the efficient layout).

    ``max_channels`` is kept from the reference signature (``rfi/device.py:507-512``),
    where it sizes a register array; here any channel count up to it works, and rows
    longer than 49 152 channels fall back from shared memory to L2.
    """

    host_class = host.ReferenceHostClass("NoiseEstMADHost")   # reference rfi/device.py: same attribute

    transposed = True
    _TUNING = {"wgsx": 1024}

    @classmethod
    @tune.autotuner(test={"wgsx": 1024})
    def autotune(cls, context: Any, max_channels: int) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, max_channels: int,
                 tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self.max_channels = max_channels
        self._init_tuning(context, tuning, max_channels)

    def instantiate(self, command_queue: Any, channels: int, baselines: int,
                    allocator: Optional[accel.AbstractAllocator] = None) -> "NoiseEstMADTDevice":
        return NoiseEstMADTDevice(self, command_queue, channels, baselines, allocator)


class NoiseEstMADTDevice(AbstractNoiseEstDevice):
    """Slots: **deviations** (baselines x channels, float32), **noise** (baselines, float32)."""

    transposed = True

    def __init__(self, template: NoiseEstMADTDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        if channels > template.max_channels:
            raise ValueError("channels exceeds max_channels")
        self.template = template
        self.channels = channels
        self.baselines = baselines
        self.slots["noise"] = accel.IOSlot((baselines,), np.float32)
        self.slots["deviations"] = accel.IOSlot((baselines, channels), np.float32)

    def _run(self) -> None:
        deviations = self.buffer("deviations")
        launch(self.command_queue, "ksp_madnz_t", ptr(deviations), ptr(self.buffer("noise")),
               self.channels, self.baselines, deviations.padded_shape[1])

    def parameters(self) -> Mapping[str, Any]:
        return {
            "max_channels": self.template.max_channels,
            "channels": self.channels,
            "baselines": self.baselines,
        }


# ----------------------------------------------------------------------------- thresholds
class ThresholdHostFromDevice(host.AbstractThresholdHost):
    """numpy in, numpy out around a threshold template (reference :617-651)."""

    def __init__(self, template: AbstractThresholdDeviceTemplate, command_queue: Any,
                 *args: Any, **kwargs: Any) -> None:
        self.template = template
        self.command_queue = command_queue
        self.args = args
        self.kwargs = kwargs

    def __call__(self, deviations: np.ndarray, noise: np.ndarray) -> np.ndarray:
        channels, baselines = deviations.shape
        transposed = self.template.transposed
        if transposed:
            deviations = deviations.T
        fn = self.template.instantiate(self.command_queue, channels, baselines, *self.args,
                                       **self.kwargs)
        fn.ensure_all_bound()
        fn.buffer("deviations").set(self.command_queue, deviations)
        fn.buffer("noise").set(self.command_queue, noise)
        fn()
        flags = fn.buffer("flags").get(self.command_queue)
        return flags.T if transposed else flags


class ThresholdSimpleDeviceTemplate(FixedTuning, AbstractThresholdDeviceTemplate):
    """``flag = deviation > n_sigma * noise[baseline]``, either memory order."""

    host_class = host.ReferenceHostClass("ThresholdSimpleHost")   # reference rfi/device.py: same attribute

    _TUNING = {"wgsx": 256, "wgsy": 1}

    @classmethod
    @tune.autotuner(test={"wgsx": 256, "wgsy": 1})
    def autotune(cls, context: Any) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, transposed: bool, flag_value: int = 1,
                 tuning: Optional[Mapping[str, Any]] = None) -> None:
        self.context = context
        self.transposed = transposed
        self.flag_value = flag_value
        self._init_tuning(context, tuning)

    def instantiate(self, command_queue: Any, channels: int, baselines: int, n_sigma: float,
                    allocator: Optional[accel.AbstractAllocator] = None
                    ) -> "ThresholdSimpleDevice":
        return ThresholdSimpleDevice(self, command_queue, channels, baselines, n_sigma, allocator)


class ThresholdSimpleDevice(AbstractThresholdDevice):
    """Slots: **deviations** (float32), **noise** (baselines, float32), **flags** (uint8);
    2-D slots are channels x baselines, or baselines x channels when transposed."""

    def __init__(self, template: ThresholdSimpleDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, n_sigma: float,
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.channels = channels
        self.baselines = baselines
        self.n_sigma = n_sigma
        self.transposed = template.transposed
        shape = (baselines, channels) if self.transposed else (channels, baselines)
        dims = (accel.Dimension(shape[0]), accel.Dimension(shape[1]))
        self.slots["deviations"] = accel.IOSlot(dims, np.float32)
        self.slots["noise"] = accel.IOSlot((dims[0] if self.transposed else dims[1],), np.float32)
        self.slots["flags"] = accel.IOSlot(dims, np.uint8)

    def _run(self) -> None:
        deviations = self.buffer("deviations")
        flags = self.buffer("flags")
        launch(self.command_queue, "ksp_threshold_simple", ptr(deviations),
               ptr(self.buffer("noise")), ptr(flags), deviations.shape[0], deviations.shape[1],
               deviations.padded_shape[1], flags.padded_shape[1], c_double(self.n_sigma),
               self.template.flag_value, int(self.transposed))

    def parameters(self) -> Mapping[str, Any]:
        return {
            "n_sigma": self.n_sigma,
            "flag_value": self.template.flag_value,
            "transposed": self.transposed,
            "channels": self.channels,
            "baselines": self.baselines,
        }


class ThresholdSumDeviceTemplate(FixedTuning, AbstractThresholdDeviceTemplate):
    """Offringa SumThreshold along channels with windows 1, 2, ..., 2^(n_windows-1).

    Takes transposed (baseline-major) data.  ``n_windows`` is at most 11 (windows up to 1024);
    up to 7 (windows up to 64) run on the fast kernel.
    """

    host_class = host.ReferenceHostClass("ThresholdSumHost")   # reference rfi/device.py: same attribute

    transposed = True
    _TUNING = {"wgs": 1024, "vt": 32}

    @classmethod
    @tune.autotuner(test={"wgs": 1024, "vt": 32})
    def autotune(cls, context: Any, n_windows: int) -> Mapping[str, Any]:
        """Nothing to search (the library fixes the launch geometry for sm_100a); the answer
        is cached under the reference's key layout all the same (see :mod:`katsdpsigproc_b200.tune`)."""
        return dict(cls._TUNING)

    def __init__(self, context: Any, n_windows: int = 4, flag_value: int = 1,
                 tuning: Optional[Mapping[str, Any]] = None) -> None:
        if n_windows < 1:
            raise ValueError("n_windows must be at least 1")
        if n_windows > _capi.MAX_WINDOWS:
            raise ValueError(f"n_windows must be at most {_capi.MAX_WINDOWS}")
        self.context = context
        self.n_windows = n_windows
        self.flag_value = flag_value
        self._init_tuning(context, tuning, n_windows)

    def instantiate(self, command_queue: Any, channels: int, baselines: int, n_sigma: float,
                    threshold_falloff: float = DEFAULT_THRESHOLD_FALLOFF,
                    allocator: Optional[accel.AbstractAllocator] = None) -> "ThresholdSumDevice":
        return ThresholdSumDevice(self, command_queue, channels, baselines, n_sigma,
                                  threshold_falloff, allocator)


def _window_scales(n_windows: int, threshold_falloff: float):
    """falloff^-i, computed in Python floats exactly as the reference host does
    (``rfi/host.py:215``); the kernel forms float32((n_sigma * noise) * scale)."""
    values = [pow(threshold_falloff, -i) for i in range(n_windows)]
    return values, (c_double * len(values))(*values)


class ThresholdSumDevice(AbstractThresholdDevice):
    """Slots: **deviations** (baselines x channels, float32), **noise** (baselines, float32),
    **flags** (baselines x channels, uint8, same row stride as deviations in elements)."""

    transposed = True

    def __init__(self, template: ThresholdSumDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, n_sigma: float,
                 threshold_falloff: float = DEFAULT_THRESHOLD_FALLOFF,
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.template = template
        self.channels = channels
        self.baselines = baselines
        self.n_sigma_base = float(n_sigma)
        self.threshold_falloff = float(threshold_falloff)
        scales, self._scales = _window_scales(template.n_windows, threshold_falloff)
        self.n_sigma = [np.float32(n_sigma * s) for s in scales]   # as the reference reports it
        dims = (baselines, accel.Dimension(channels))
        self.slots["deviations"] = accel.IOSlot(dims, np.float32)
        self.slots["noise"] = accel.IOSlot((baselines,), np.float32)
        self.slots["flags"] = accel.IOSlot(dims, np.uint8)

    def _run(self) -> None:
        deviations = self.buffer("deviations")
        flags = self.buffer("flags")
        launch(self.command_queue, "ksp_threshold_sum", ptr(deviations), ptr(self.buffer("noise")),
               ptr(flags), self.channels, self.baselines, deviations.padded_shape[1],
               flags.padded_shape[1], self.template.n_windows, c_double(self.n_sigma_base),
               self._scales, self.template.flag_value)

    def parameters(self) -> Mapping[str, Any]:
        return {
            "n_sigma": self.n_sigma,
            "flag_value": self.template.flag_value,
            "channels": self.channels,
            "baselines": self.baselines,
        }


# ----------------------------------------------------------------------------- flagger
class FusedFlaggerDevice(accel.Operation):
    """Median background + MAD noise + SumThreshold as one operation (``ksp_flagger``).

    Default ("chunked" form): four launches per chunk of baselines (background written
    baseline-major, noise, thresholds with bit-packed flags, expansion to channel-major bytes),
    up to four chunks in flight; a chunk's deviations live in ``scratch`` and DO go through
    device memory (written once, read twice: about 20 bytes of traffic per visibility).
    ``chunk_baselines < 0`` ("dataflow" form; width 13, up to 7 window sizes, channels a multiple
    of 32): ONE persistent kernel per dump whose work items (background tiles, noise rows,
    threshold rows, flag expansion tiles) hand the deviations on through a ring of a few strips of
    32 baselines that stays in the L2 cache, so device memory sees the visibilities once and the
    flags once - less traffic, but measured slower on B200 than the chunked form.

    Slots: **vis**, **flags** (input flags; only with ``use_flags``), **noise**,
    **out_flags** (channels x baselines uint8) and **scratch** (uint8 bytes).
    """

    def __init__(self, background: BackgroundMedianFilterDeviceTemplate,
                 threshold: ThresholdSumDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, n_sigma: float,
                 threshold_falloff: float = DEFAULT_THRESHOLD_FALLOFF, chunk_baselines: int = 0,
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        super().__init__(command_queue, allocator)
        self.background_template = background
        self.threshold_template = threshold
        self.channels = channels
        self.baselines = baselines
        self.n_sigma = float(n_sigma)
        self.threshold_falloff = float(threshold_falloff)
        bl_dim = accel.Dimension(baselines)
        dims = (channels, bl_dim)
        self.slots["vis"] = accel.IOSlot(
            dims, np.float32 if background.is_amplitude else np.complex64)
        if background.use_flags is BackgroundFlags.FULL:
            self.slots["flags"] = accel.IOSlot(dims, np.uint8)
        elif background.use_flags is BackgroundFlags.CHANNEL:
            self.slots["flags"] = accel.IOSlot((channels,), np.uint8)
        self.slots["noise"] = accel.IOSlot((baselines,), np.float32)
        self.slots["out_flags"] = accel.IOSlot(dims, np.uint8)

        params = _capi.FlaggerParams()
        params.channels, params.baselines = channels, baselines
        params.width = background.width
        params.is_amplitude = int(background.is_amplitude)
        params.flag_mode = background.use_flags.value
        params.abs_mode = background.abs_mode
        params.n_windows = threshold.n_windows
        params.flag_value = threshold.flag_value
        params.n_sigma = self.n_sigma
        scales, _ = _window_scales(threshold.n_windows, threshold_falloff)
        for i in range(_capi.MAX_WINDOWS):
            params.scales[i] = scales[i] if i < len(scales) else 0.0
        params.chunk_baselines = chunk_baselines
        # the default chunk depends on the L2 size of the queue's device
        getattr(command_queue.context, "_make_current", lambda: None)()
        lib = _capi.load()
        self.scratch_bytes = int(lib.ksp_flagger_scratch_bytes(byref(params)))
        self.chunk_baselines = int(lib.ksp_flagger_chunk_baselines(byref(params)))
        self.dataflow = bool(lib.ksp_flagger_is_dataflow(byref(params)))
        self._params = params
        self.slots["scratch"] = accel.IOSlot((max(self.scratch_bytes, 16),), np.uint8)

    def _run(self) -> None:
        vis = self.buffer("vis")
        out = self.buffer("out_flags")
        in_flags = self.buffer("flags") if "flags" in self.slots else None
        p = self._params
        p.vis_stride = vis.padded_shape[1]
        p.flags_stride = out.padded_shape[1]
        p.input_flags_stride = (in_flags.padded_shape[1]
                                if in_flags is not None and in_flags.ndim == 2 else 0)
        scratch = self.buffer("scratch")
        launch(self.command_queue, "ksp_flagger", byref(p), ptr(vis),
               ptr(in_flags) if in_flags is not None else None, ptr(self.buffer("noise")),
               ptr(out), ptr(scratch), c_size_t(self.scratch_bytes))

    def stats(self) -> Mapping[str, int]:
        """Diagnostics of the last run (waits for the queue): per-kind SM cycles, wait cycles,
        items, noise fallbacks of the dataflow kernel; raises if it abandoned the launch."""
        out = (ctypes.c_ulonglong * len(_capi.DF_STAT_NAMES))()
        launch(self.command_queue, "ksp_flagger_stats", byref(self._params),
               ptr(self.buffer("scratch")), out, len(out))
        return dict(zip(_capi.DF_STAT_NAMES, (int(v) for v in out)))

    def parameters(self) -> Mapping[str, Any]:
        return {
            "width": self.background_template.width,
            "use_flags": self.background_template.use_flags.name,
            "n_windows": self.threshold_template.n_windows,
            "n_sigma": self.n_sigma,
            "threshold_falloff": self.threshold_falloff,
            "flag_value": self.threshold_template.flag_value,
            "chunk_baselines": self.chunk_baselines,
            "dataflow": self.dataflow,
            "channels": self.channels,
            "baselines": self.baselines,
        }


class FlaggerDeviceTemplate:
    """Background + noise estimate + threshold (reference :998-1059).

    Parameters
    ----------
    background, noise_est, threshold
        Templates of the three stages; they must share a context
    fused
        Extension.  ``None`` (default): use the fused single-operation path
        whenever the combination is the standard one (median background, MAD
        noise in either layout, SumThreshold); ``False``: always run the
        reference's sequence of stage operations with ``deviations`` /
        ``deviations_t`` / ``flags_t`` as real slots; ``True``: require fusion.
    """

    def __init__(self, background: AbstractBackgroundDeviceTemplate,
                 noise_est: AbstractNoiseEstDeviceTemplate,
                 threshold: AbstractThresholdDeviceTemplate,
                 fused: Optional[bool] = None) -> None:
        self.background = background
        self.noise_est = noise_est
        self.threshold = threshold
        context = background.context
        assert noise_est.context is context
        assert threshold.context is context
        self.context = context
        fusable = (isinstance(background, BackgroundMedianFilterDeviceTemplate)
                   and isinstance(noise_est, (NoiseEstMADTDeviceTemplate, NoiseEstMADDeviceTemplate))
                   and isinstance(threshold, ThresholdSumDeviceTemplate))
        if fused and not fusable:
            raise ValueError("only median background + MAD noise + SumThreshold can be fused")
        self.fused = fusable if fused is None else bool(fused)
        self.transpose_deviations: Optional[transpose.TransposeTemplate] = None
        self.transpose_flags: Optional[transpose.TransposeTemplate] = None
        if noise_est.transposed or threshold.transposed:
            self.transpose_deviations = transpose.TransposeTemplate(context, np.float32, "float")
        if threshold.transposed:
            self.transpose_flags = transpose.TransposeTemplate(context, np.uint8, "unsigned char")

    def instantiate(self, command_queue: Any, channels: int, baselines: int,
                    background_args: Mapping[str, Any] = {},
                    noise_est_args: Mapping[str, Any] = {},
                    threshold_args: Mapping[str, Any] = {},
                    allocator: Optional[accel.AbstractAllocator] = None) -> "FlaggerDevice":
        return FlaggerDevice(self, command_queue, channels, baselines, background_args,
                             noise_est_args, threshold_args, allocator)


class FlaggerDevice(accel.OperationSequence):
    """The whole flagger.

    Slots: **vis** (channels x baselines, complex64 or float32), **input_flags**
    (only if the background uses flags), **noise** (baselines, float32), **flags**
    (channels x baselines, uint8).  Unfused instances also expose the temporaries
    **deviations**, **deviations_t**, **flags_t** exactly as the reference does
    (``rfi/device.py:1081-1091``); fused instances have **scratch** instead.
    Input flags are neither copied to the output nor ever set in it.
    """

    def __init__(self, template: FlaggerDeviceTemplate, command_queue: Any, channels: int,
                 baselines: int, background_args: Mapping[str, Any] = {},
                 noise_est_args: Mapping[str, Any] = {}, threshold_args: Mapping[str, Any] = {},
                 allocator: Optional[accel.AbstractAllocator] = None) -> None:
        self.template = template
        self.channels = channels
        self.baselines = baselines
        self.background = template.background.instantiate(
            command_queue, channels, baselines, allocator=allocator, **background_args)
        self.noise_est = template.noise_est.instantiate(
            command_queue, channels, baselines, allocator=allocator, **noise_est_args)
        self.threshold = template.threshold.instantiate(
            command_queue, channels, baselines, allocator=allocator, **threshold_args)
        self.transpose_deviations: Optional[transpose.Transpose] = None
        self.transpose_flags: Optional[transpose.Transpose] = None
        self.fused_op: Optional[FusedFlaggerDevice] = None

        operations: List[Tuple[str, accel.Operation]] = []
        if template.fused:
            # the stage operations above stay reachable for introspection (parameters())
            # but are not part of the sequence and own no memory
            self.fused_op = FusedFlaggerDevice(
                template.background, template.threshold, command_queue, channels, baselines,
                self.threshold.n_sigma_base, self.threshold.threshold_falloff,
                allocator=allocator)
            operations.append(("fused", self.fused_op))
            compounds = {
                "vis": ["fused:vis"],
                "input_flags": ["fused:flags"],
                "noise": ["fused:noise"],
                "flags": ["fused:out_flags"],
                "scratch": ["fused:scratch"],
            }
        else:
            noise_est_suffix = "_t" if self.noise_est.transposed else ""
            threshold_suffix = "_t" if self.threshold.transposed else ""
            compounds = {
                "vis": ["background:vis"],
                "input_flags": ["background:flags"],
                "deviations": ["background:deviations", "transpose_deviations:src"],
                "deviations_t": ["transpose_deviations:dest"],
                "noise": ["noise_est:noise", "threshold:noise"],
                "flags_t": ["transpose_flags:src"],
                "flags": ["transpose_flags:dest"],
            }
            compounds["deviations" + noise_est_suffix].append("noise_est:deviations")
            compounds["deviations" + threshold_suffix].append("threshold:deviations")
            compounds["flags" + threshold_suffix].append("threshold:flags")
            operations.append(("background", self.background))
            if template.transpose_deviations:
                self.transpose_deviations = template.transpose_deviations.instantiate(
                    command_queue, (channels, baselines))
                operations.append(("transpose_deviations", self.transpose_deviations))
            operations.append(("noise_est", self.noise_est))
            operations.append(("threshold", self.threshold))
            if template.transpose_flags:
                self.transpose_flags = template.transpose_flags.instantiate(
                    command_queue, (baselines, channels))
                operations.append(("transpose_flags", self.transpose_flags))
        super().__init__(command_queue, operations, compounds, allocator=allocator)

    def stats(self) -> Mapping[str, int]:
        """Diagnostics of the fused operation's last run (see FusedFlaggerDevice.stats)."""
        return self.fused_op.stats() if self.fused_op is not None else {}

    def parameters(self) -> Mapping[str, Any]:
        return {
            "fused": self.fused_op is not None,
            "dataflow": self.fused_op is not None and self.fused_op.dataflow,
            "channels": self.channels,
            "baselines": self.baselines,
        }


class FlaggerHostFromDevice(host.AbstractFlaggerHost):
    """numpy in, numpy out around a flagger template (reference :1169-1222); allocates on
    every call, so it is a convenience for tests, not a fast path."""

    def __init__(self, template: FlaggerDeviceTemplate, command_queue: Any,
                 background_args: Mapping[str, Any] = {}, noise_est_args: Mapping[str, Any] = {},
                 threshold_args: Mapping[str, Any] = {}) -> None:
        self.template = template
        self.command_queue = command_queue
        self.background_args = dict(background_args)
        self.noise_est_args = dict(noise_est_args)
        self.threshold_args = dict(threshold_args)

    def __call__(self, vis: np.ndarray, input_flags: Optional[np.ndarray] = None) -> np.ndarray:
        if input_flags is not None and not self.template.background.use_flags:
            raise TypeError("channel flags were provided but not included in the template")
        if input_flags is None and self.template.background.use_flags:
            raise TypeError("channel flags were expected but not provided")
        channels, baselines = vis.shape
        fn = self.template.instantiate(self.command_queue, channels, baselines,
                                       self.background_args, self.noise_est_args,
                                       self.threshold_args)
        fn.ensure_all_bound()
        fn.buffer("vis").set(self.command_queue, vis)
        if input_flags is not None:
            fn.buffer("input_flags").set(self.command_queue, input_flags)
        fn()
        return fn.buffer("flags").get(self.command_queue)
