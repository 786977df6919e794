"""numpy-in / numpy-out helpers that drive the C ABI directly (``include/ksp_b200.h``).

Used by the ``-m gpu`` parity tests and by ``tools/``; every call goes
host -> device -> kernel -> host through ``libksp_b200.so``.  Buffers may be
given non-trivial row padding (``pad``) so that strides are exercised.
"""

from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_size_t, c_void_p
from typing import Optional, Tuple

import numpy as np

from katsdpsigproc_b200 import _capi


class Dev:
    """A device allocation holding a (padded) C-order 2-D or 1-D array."""

    def __init__(self, shape, dtype, pad: int = 0, fill: Optional[int] = None) -> None:
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.stride = (self.shape[-1] + pad) if len(self.shape) == 2 else 1
        self.padded = (self.shape[0], self.stride) if len(self.shape) == 2 else self.shape
        self.nbytes = int(np.prod(self.padded, dtype=np.int64)) * self.dtype.itemsize
        ptr = c_void_p()
        _capi.call("ksp_malloc", byref(ptr), c_size_t(max(self.nbytes, 16)))
        self.ptr = ptr.value
        if fill is not None:
            _capi.call("ksp_memset_async", c_void_p(self.ptr), fill, c_size_t(self.nbytes), None)

    @classmethod
    def from_host(cls, a: np.ndarray, pad: int = 0) -> "Dev":
        a = np.asarray(a)
        d = cls(a.shape, a.dtype, pad)
        full = np.zeros(d.padded, a.dtype)
        if a.ndim == 2:
            full[:, : a.shape[1]] = a
        else:
            full[...] = a
        full = np.ascontiguousarray(full)
        if d.nbytes:
            _capi.call("ksp_memcpy_async", c_void_p(d.ptr), c_void_p(full.ctypes.data),
                       c_size_t(d.nbytes), _capi.H2D, None)
            _capi.call("ksp_stream_synchronize", None)
        return d

    def get(self) -> np.ndarray:
        full = np.empty(self.padded, self.dtype)
        if self.nbytes:
            _capi.call("ksp_memcpy_async", c_void_p(full.ctypes.data), c_void_p(self.ptr),
                       c_size_t(self.nbytes), _capi.D2H, None)
        _capi.call("ksp_stream_synchronize", None)
        if len(self.shape) == 2:
            return np.ascontiguousarray(full[:, : self.shape[1]])
        return full

    @property
    def p(self) -> c_void_p:
        return c_void_p(self.ptr)

    def __del__(self) -> None:
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _capi.load().ksp_free(c_void_p(ptr))
            except Exception:
                pass


def scales(n_windows: int, falloff: float):
    vals = [pow(falloff, -i) for i in range(max(n_windows, 1))]
    return (c_double * len(vals))(*vals)


def sync() -> None:
    _capi.call("ksp_stream_synchronize", None)


def transpose(a: np.ndarray, pad_in: int = 0, pad_out: int = 0) -> np.ndarray:
    src = Dev.from_host(a, pad_in)
    dst = Dev((a.shape[1], a.shape[0]), a.dtype, pad_out, fill=0)
    _capi.call("ksp_transpose", None, dst.p, src.p, a.shape[0], a.shape[1], dst.stride, src.stride,
               a.dtype.itemsize)
    return dst.get()


def _flag_args(flags: Optional[np.ndarray], pad: int):
    if flags is None:
        return None, _capi.FLAGS_NONE, 0, None
    flags = np.ascontiguousarray(flags, np.uint8)
    if flags.ndim == 1:
        d = Dev.from_host(flags)
        return d.p, _capi.FLAGS_CHANNEL, 0, d
    d = Dev.from_host(flags, pad)
    return d.p, _capi.FLAGS_FULL, d.stride, d


def background(vis: np.ndarray, width: int, flags: Optional[np.ndarray] = None,
               amplitudes: bool = False, abs_mode: int = 0, transposed: bool = False,
               pad: int = 0) -> np.ndarray:
    """Deviations (channels, baselines) -- or (baselines, channels) when ``transposed``."""
    vis = np.ascontiguousarray(vis, np.float32 if amplitudes else np.complex64)
    channels, baselines = vis.shape
    dvis = Dev.from_host(vis, pad)
    fp, mode, fstride, keep = _flag_args(flags, pad)
    if transposed:
        out = Dev((baselines, channels), np.float32, pad, fill=0xFF)
        name = "ksp_background_median_filter_t"
    else:
        out = Dev((channels, baselines), np.float32, pad, fill=0xFF)
        name = "ksp_background_median_filter"
    _capi.call(name, None, dvis.p, out.p, fp, channels, baselines, dvis.stride, out.stride, fstride,
               int(width), int(amplitudes), mode, int(abs_mode))
    return out.get()


def madnz(dev: np.ndarray, transposed: bool, pad: int = 0) -> np.ndarray:
    dev = np.ascontiguousarray(dev, np.float32)
    d = Dev.from_host(dev, pad)
    if transposed:
        baselines, channels = dev.shape
        name = "ksp_madnz_t"
    else:
        channels, baselines = dev.shape
        name = "ksp_madnz"
    noise = Dev((baselines,), np.float32, fill=0)
    _capi.call(name, None, d.p, noise.p, channels, baselines, d.stride)
    return noise.get()


def threshold_sum(dev_t: np.ndarray, noise: np.ndarray, n_sigma: float, n_windows: int = 4,
                  falloff: float = 1.2, flag_value: int = 1, pad: int = 0) -> np.ndarray:
    """Baseline-major in, baseline-major out."""
    dev_t = np.ascontiguousarray(dev_t, np.float32)
    baselines, channels = dev_t.shape
    d = Dev.from_host(dev_t, pad)
    n = Dev.from_host(np.ascontiguousarray(noise, np.float32))
    out = Dev((baselines, channels), np.uint8, pad, fill=0x55)
    _capi.call("ksp_threshold_sum", None, d.p, n.p, out.p, channels, baselines, d.stride,
               out.stride, int(n_windows), c_double(n_sigma), scales(n_windows, falloff),
               int(flag_value))
    return out.get()


def threshold_simple(dev: np.ndarray, noise: np.ndarray, n_sigma: float, flag_value: int = 1,
                     transposed: bool = False, pad: int = 0) -> np.ndarray:
    dev = np.ascontiguousarray(dev, np.float32)
    rows, cols = dev.shape
    d = Dev.from_host(dev, pad)
    n = Dev.from_host(np.ascontiguousarray(noise, np.float32))
    out = Dev((rows, cols), np.uint8, pad, fill=0x55)
    _capi.call("ksp_threshold_simple", None, d.p, n.p, out.p, rows, cols, d.stride, out.stride,
               c_double(n_sigma), int(flag_value), int(transposed))
    return out.get()


def percentile5(src: np.ndarray, column_range: Optional[Tuple[int, int]] = None, abs_mode: int = 0,
                pad: int = 0) -> np.ndarray:
    is_amp = not np.iscomplexobj(src)
    src = np.ascontiguousarray(src, np.float32 if is_amp else np.complex64)
    rows, cols = src.shape
    if column_range is None:
        column_range = (0, cols)
    d = Dev.from_host(src, pad)
    out = Dev((5, rows), np.float32, pad, fill=0)
    _capi.call("ksp_percentile5", None, d.p, out.p, rows, d.stride, out.stride, column_range[0],
               column_range[1] - column_range[0], int(is_amp), int(abs_mode))
    return out.get()


def masked_sum(src: np.ndarray, mask: np.ndarray, use_amplitudes: bool = False, abs_mode: int = 0,
               pad: int = 0) -> np.ndarray:
    src = np.ascontiguousarray(src, np.complex64)
    rows, cols = src.shape
    d = Dev.from_host(src, pad)
    m = Dev.from_host(np.ascontiguousarray(mask, np.float32))
    out = Dev((cols,), np.float32 if use_amplitudes else np.complex64, fill=0)
    _capi.call("ksp_maskedsum", None, d.p, m.p, out.p, rows, cols, d.stride, int(use_amplitudes),
               int(abs_mode))
    return out.get()


def flagger_params(channels: int, baselines: int, vis_stride: int, flags_stride: int,
                   input_flags_stride: int = 0, width: int = 13, amplitudes: bool = False,
                   flag_mode: int = 0, abs_mode: int = 0, n_windows: int = 4, flag_value: int = 1,
                   n_sigma: float = 11.0, falloff: float = 1.2, chunk_baselines: int = 0
                   ) -> _capi.FlaggerParams:
    p = _capi.FlaggerParams()
    p.channels, p.baselines = channels, baselines
    p.vis_stride, p.flags_stride, p.input_flags_stride = vis_stride, flags_stride, input_flags_stride
    p.width, p.is_amplitude, p.flag_mode, p.abs_mode = width, int(amplitudes), flag_mode, abs_mode
    p.n_windows, p.flag_value, p.n_sigma = n_windows, flag_value, n_sigma
    for i in range(_capi.MAX_WINDOWS):
        p.scales[i] = pow(falloff, -i) if i < n_windows else 0.0
    p.chunk_baselines = chunk_baselines
    return p


def flagger(vis: np.ndarray, input_flags: Optional[np.ndarray] = None, *, width: int = 13,
            n_sigma: float = 11.0, n_windows: int = 4, falloff: float = 1.2, flag_value: int = 1,
            amplitudes: bool = False, abs_mode: int = 0, chunk_baselines: int = 0, pad: int = 0
            ) -> Tuple[np.ndarray, np.ndarray]:
    """Fused flagger: (flags u8 (channels, baselines), noise f32)."""
    vis = np.ascontiguousarray(vis, np.float32 if amplitudes else np.complex64)
    channels, baselines = vis.shape
    dvis = Dev.from_host(vis, pad)
    fp, mode, fstride, keep = _flag_args(input_flags, pad)
    flags = Dev((channels, baselines), np.uint8, pad, fill=0x55)
    noise = Dev((baselines,), np.float32, fill=0)
    p = flagger_params(channels, baselines, dvis.stride, flags.stride, fstride, width, amplitudes,
                       mode, abs_mode, n_windows, flag_value, n_sigma, falloff, chunk_baselines)
    n_scratch = _capi.load().ksp_flagger_scratch_bytes(byref(p))
    scratch = Dev((max(n_scratch, 16),), np.uint8)
    _capi.call("ksp_flagger", None, byref(p), dvis.p, fp, noise.p, flags.p, scratch.p,
               c_size_t(n_scratch))
    stats = (ctypes.c_ulonglong * len(_capi.DF_STAT_NAMES))()
    _capi.call("ksp_flagger_stats", None, byref(p), scratch.p, stats, len(stats))  # raises if abandoned
    LAST_FLAGGER_STATS.clear()
    LAST_FLAGGER_STATS.update(zip(_capi.DF_STAT_NAMES, (int(v) for v in stats)))
    LAST_FLAGGER_STATS["dataflow"] = int(_capi.load().ksp_flagger_is_dataflow(byref(p)))
    return flags.get(), noise.get()


# diagnostics of the most recent flagger() call (ksp_flagger_stats + whether the dataflow form ran)
LAST_FLAGGER_STATS: dict = {}


def selection_fallbacks(reset: bool = False) -> int:
    """Rows (x ranks) that ksp_madnz_t / ksp_percentile5 redid with the radix select so far."""
    count = ctypes.c_ulonglong(0)
    _capi.call("ksp_selection_fallback_count", None, ctypes.byref(count), int(reset))
    return int(count.value)
