"""ctypes binding of ``oracle/contract.c`` (the float32 device contract).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The shared object is
built on first use with gcc (``oracle/Makefile``) into ``oracle/_build/``.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

ABS_NUMPY = 0
ABS_HYPOT = 1
FLAGS_NONE, FLAGS_CHANNEL, FLAGS_FULL = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcontract.so")
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    """Compile contract.c if needed; return the path of the shared object."""
    src = os.path.join(_HERE, "contract.c")
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ko_abs.restype = ctypes.c_float
        _lib.ko_abs.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _scales(n_windows: int, falloff: float) -> np.ndarray:
    # computed in Python exactly as the reference does (rfi/host.py:215)
    return np.array([pow(falloff, -i) for i in range(max(n_windows, 1))], np.float64)


def detect_abs_mode() -> int:
    """Which amplitude rule does this host's ``np.abs(complex64)`` follow? (R1)"""
    rs = np.random.RandomState(12345)
    v = (rs.standard_normal(4096) + 1j * rs.standard_normal(4096)).astype(np.complex64)
    ref = np.abs(v)
    for mode in (ABS_NUMPY, ABS_HYPOT):
        if np.array_equal(amplitude(v, mode), ref):
            return mode
    return -1


def amplitude(vis: np.ndarray, abs_mode: int = ABS_NUMPY) -> np.ndarray:
    vis = np.ascontiguousarray(vis, np.complex64)
    out = np.empty(vis.shape, np.float32)
    lib().ko_amplitude(_p(vis), _p(out), ctypes.c_long(vis.size), ctypes.c_int(abs_mode))
    return out


def background(
    vis: np.ndarray,
    width: int,
    flags: Optional[np.ndarray] = None,
    amplitudes: bool = False,
    abs_mode: int = ABS_NUMPY,
) -> np.ndarray:
    """float32 deviations, (channels, baselines)."""
    vis = np.ascontiguousarray(vis, np.float32 if amplitudes else np.complex64)
    channels, baselines = vis.shape
    mode, fstride = FLAGS_NONE, 0
    if flags is not None:
        flags = np.ascontiguousarray(flags, np.uint8)
        if flags.ndim == 1:
            mode = FLAGS_CHANNEL
        else:
            mode, fstride = FLAGS_FULL, baselines
    dev = np.empty((channels, baselines), np.float32)
    rc = lib().ko_background(
        _p(vis), int(amplitudes), _p(flags), mode, ctypes.c_long(fstride), _p(dev),
        ctypes.c_long(channels), ctypes.c_long(baselines), ctypes.c_long(baselines),
        ctypes.c_long(baselines), int(width), int(abs_mode),
    )
    if rc:
        raise ValueError("width must be odd and in [1, 255]")
    return dev


def noise_mad(dev: np.ndarray, transposed: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """(noise float32, median float32) per baseline.

    ``dev`` is (channels, baselines), or (baselines, channels) if ``transposed``.
    """
    dev = np.ascontiguousarray(dev, np.float32)
    if transposed:
        baselines, channels = dev.shape
        sc, sb = 1, channels
    else:
        channels, baselines = dev.shape
        sc, sb = baselines, 1
    noise = np.empty(baselines, np.float32)
    med = np.empty(baselines, np.float32)
    lib().ko_noise_mad(
        _p(dev), ctypes.c_long(channels), ctypes.c_long(baselines), ctypes.c_long(sc),
        ctypes.c_long(sb), _p(noise), _p(med),
    )
    return noise, med


def threshold_sum(
    dev: np.ndarray,
    noise: np.ndarray,
    n_sigma: float,
    n_windows: int = 4,
    threshold_falloff: float = 1.2,
    flag_value: int = 1,
    transposed: bool = False,
) -> np.ndarray:
    dev = np.ascontiguousarray(dev, np.float32)
    noise = np.ascontiguousarray(noise, np.float32)
    if transposed:
        baselines, channels = dev.shape
        sc, sb = 1, channels
    else:
        channels, baselines = dev.shape
        sc, sb = baselines, 1
    flags = np.empty(dev.shape, np.uint8)
    scales = _scales(n_windows, threshold_falloff)
    lib().ko_threshold_sum(
        _p(dev), _p(noise), _p(flags), ctypes.c_long(channels), ctypes.c_long(baselines),
        ctypes.c_long(sc), ctypes.c_long(sb), ctypes.c_long(sc), ctypes.c_long(sb),
        int(n_windows), ctypes.c_double(n_sigma), _p(scales), int(flag_value),
    )
    return flags


def threshold_simple(
    dev: np.ndarray, noise: np.ndarray, n_sigma: float, flag_value: int = 1,
    transposed: bool = False,
) -> np.ndarray:
    dev = np.ascontiguousarray(dev, np.float32)
    noise = np.ascontiguousarray(noise, np.float32)
    if transposed:
        baselines, channels = dev.shape
        sc, sb = 1, channels
    else:
        channels, baselines = dev.shape
        sc, sb = baselines, 1
    flags = np.empty(dev.shape, np.uint8)
    lib().ko_threshold_simple(
        _p(dev), _p(noise), _p(flags), ctypes.c_long(channels), ctypes.c_long(baselines),
        ctypes.c_long(sc), ctypes.c_long(sb), ctypes.c_long(sc), ctypes.c_long(sb),
        ctypes.c_double(n_sigma), int(flag_value),
    )
    return flags


def flagger(
    vis: np.ndarray,
    input_flags: Optional[np.ndarray] = None,
    *,
    width: int = 13,
    n_sigma: float = 11.0,
    n_windows: int = 4,
    threshold_falloff: float = 1.2,
    flag_value: int = 1,
    amplitudes: bool = False,
    abs_mode: int = ABS_NUMPY,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Composed contract flagger: returns (flags u8, deviations f32, noise f32).

    ``n_windows=0`` selects the simple (per-sample) threshold.
    """
    vis = np.ascontiguousarray(vis, np.float32 if amplitudes else np.complex64)
    channels, baselines = vis.shape
    mode = FLAGS_NONE
    if input_flags is not None:
        input_flags = np.ascontiguousarray(input_flags, np.uint8)
        mode = FLAGS_CHANNEL if input_flags.ndim == 1 else FLAGS_FULL
    dev = np.empty((channels, baselines), np.float32)
    noise = np.empty(baselines, np.float32)
    flags = np.empty((channels, baselines), np.uint8)
    scales = _scales(n_windows, threshold_falloff)
    rc = lib().ko_flagger(
        _p(vis), int(amplitudes), _p(input_flags), mode, _p(dev), _p(noise), _p(flags),
        ctypes.c_long(channels), ctypes.c_long(baselines), int(width), int(n_windows),
        ctypes.c_double(n_sigma), _p(scales), int(flag_value), int(abs_mode),
    )
    if rc:
        raise ValueError("bad flagger arguments")
    return flags, dev, noise


def percentile5(
    src: np.ndarray, column_range: Optional[Tuple[int, int]] = None, abs_mode: int = ABS_NUMPY
) -> np.ndarray:
    is_amp = not np.iscomplexobj(src)
    src = np.ascontiguousarray(src, np.float32 if is_amp else np.complex64)
    rows, cols = src.shape
    if column_range is None:
        column_range = (0, cols)
    dest = np.empty((5, rows), np.float32)
    lib().ko_percentile5(
        _p(src), int(is_amp), ctypes.c_long(rows), ctypes.c_long(cols),
        ctypes.c_long(column_range[0]), ctypes.c_long(column_range[1] - column_range[0]),
        _p(dest), ctypes.c_long(rows), int(abs_mode),
    )
    return dest


def masked_sum(
    src: np.ndarray, mask: np.ndarray, use_amplitudes: bool = False, abs_mode: int = ABS_NUMPY
) -> np.ndarray:
    src = np.ascontiguousarray(src, np.complex64)
    mask = np.ascontiguousarray(mask, np.float32)
    rows, cols = src.shape
    dest = np.empty(cols, np.float32 if use_amplitudes else np.complex64)
    lib().ko_maskedsum(
        _p(src), _p(mask), ctypes.c_long(rows), ctypes.c_long(cols), ctypes.c_long(cols),
        int(use_amplitudes), _p(dest), int(abs_mode),
    )
    return dest
