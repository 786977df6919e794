"""ctypes binding of ``libksp_b200.so`` (the C ABI declared in ``include/ksp_b200.h``).

The library is built in-tree by ``katsdpsigproc_b200/csrc/Makefile`` (see
``__graft_entry__.build``).  There is no fallback of any kind: if the shared
object is missing or a call fails, an exception is raised.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p
from typing import Optional

ABS_NUMPY = 0
ABS_HYPOT = 1
FLAGS_NONE, FLAGS_CHANNEL, FLAGS_FULL = 0, 1, 2
MAX_WINDOWS = 11
MAX_WIDTH = 63
H2D, D2H, D2D = 1, 2, 3

# KSP_B200_LIB points at an alternative build of the library (kernel tuning experiments)
LIB_PATH = os.environ.get("KSP_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "_lib", "libksp_b200.so")


class KspError(RuntimeError):
    """A C-ABI call returned a non-zero code."""

    def __init__(self, code: int, what: str, message: str) -> None:
        super().__init__(f"{what} failed with code {code}: {message}")
        self.code = code


class FlaggerParams(ctypes.Structure):
    """Mirror of ``ksp_flagger_params``."""

    _fields_ = [
        ("channels", c_int64),
        ("baselines", c_int64),
        ("vis_stride", c_int64),
        ("flags_stride", c_int64),
        ("input_flags_stride", c_int64),
        ("width", c_int),
        ("is_amplitude", c_int),
        ("flag_mode", c_int),
        ("abs_mode", c_int),
        ("n_windows", c_int),
        ("flag_value", c_int),
        ("n_sigma", c_double),
        ("scales", c_double * MAX_WINDOWS),
        ("chunk_baselines", c_int64),
    ]


TWOD_MAX_WINDOWS = 16
TWOD_MAX_WINDOW = 64
TWOD_MAX_CHUNKS = 64
TWOD_MAX_ITERATIONS = 63


class TwodflagParams(ctypes.Structure):
    """Mirror of ``ksp_twodflag_params``."""

    _fields_ = [
        ("n_time", c_int64), ("n_freq", c_int64), ("n_bl", c_int64),
        ("is_complex", c_int), ("average_freq", c_int),
        ("n_windows_time", c_int), ("n_windows_freq", c_int),
        ("windows_time", c_int * TWOD_MAX_WINDOWS), ("windows_freq", c_int * TWOD_MAX_WINDOWS),
        ("tf_time", c_double * TWOD_MAX_WINDOWS), ("tf_freq", c_double * TWOD_MAX_WINDOWS),
        ("outlier_nsigma", c_double), ("background_reject", c_double),
        ("background_iterations", c_int),
        ("r_time", c_int * (TWOD_MAX_ITERATIONS + 1)), ("r_freq", c_int * (TWOD_MAX_ITERATIONS + 1)),
        ("time_extend", c_int), ("freq_extend", c_int),
        ("n_chunks", c_int),
        ("chunk_ends", c_int64 * (TWOD_MAX_CHUNKS + 1)),
        ("flag_all_time_frac", c_double), ("flag_all_freq_frac", c_double),
    ]


# name -> (restype, argtypes); every symbol declared in include/ksp_b200.h
PROTOTYPES = {
    "ksp_abi_version": (c_int, []),
    "ksp_error_string": (c_char_p, [c_int]),
    "ksp_device_count": (c_int, [POINTER(c_int)]),
    "ksp_device_pci_bus_id": (c_int, [c_int, ctypes.c_char_p, c_int]),
    "ksp_device_name": (c_int, [c_int, c_char_p, c_int]),
    "ksp_device_attributes": (
        c_int,
        [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t),
         POINTER(c_int)],
    ),
    "ksp_device_set": (c_int, [c_int]),
    "ksp_device_get": (c_int, [POINTER(c_int)]),
    "ksp_versions": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "ksp_malloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "ksp_free": (c_int, [c_void_p]),
    "ksp_host_alloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "ksp_host_free": (c_int, [c_void_p]),
    "ksp_stream_create": (c_int, [POINTER(c_void_p)]),
    "ksp_stream_destroy": (c_int, [c_void_p]),
    "ksp_stream_synchronize": (c_int, [c_void_p]),
    "ksp_stream_query": (c_int, [c_void_p]),
    "ksp_stream_wait_event": (c_int, [c_void_p, c_void_p]),
    "ksp_event_create": (c_int, [POINTER(c_void_p), c_int]),
    "ksp_event_destroy": (c_int, [c_void_p]),
    "ksp_event_record": (c_int, [c_void_p, c_void_p]),
    "ksp_event_synchronize": (c_int, [c_void_p]),
    "ksp_event_elapsed_ms": (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    "ksp_memcpy_async": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "ksp_memcpy_2d_async": (
        c_int,
        [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_void_p],
    ),
    "ksp_memset_async": (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    "ksp_transpose": (
        c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int]),
    "ksp_background_median_filter": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
         c_int, c_int, c_int, c_int],
    ),
    "ksp_background_median_filter_t": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
         c_int, c_int, c_int, c_int],
    ),
    "ksp_madnz_t": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]),
    "ksp_madnz": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]),
    "ksp_threshold_simple": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_double,
         c_int, c_int],
    ),
    "ksp_threshold_sum": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int,
         c_double, POINTER(c_double), c_int],
    ),
    "ksp_percentile5": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int],
    ),
    "ksp_maskedsum": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int]),
    "ksp_fill": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]),
    "ksp_hreduce_create": (
        c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, c_size_t,
                POINTER(c_void_p)]),
    "ksp_hreduce": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64]),
    "ksp_hreduce_destroy": (c_int, [c_void_p]),
    "ksp_jit_log": (ctypes.c_char_p, []),
    "ksp_kernel_launch_count": (c_int, [POINTER(ctypes.c_ulonglong)]),
    "ksp_profile_enable": (c_int, [c_int]),
    "ksp_profile_read": (c_int, [POINTER(c_double), POINTER(c_int), c_int]),
    "ksp_selection_fallback_count": (c_int, [c_void_p, POINTER(ctypes.c_ulonglong), c_int]),
    "ksp_flagger_scratch_bytes": (c_size_t, [POINTER(FlaggerParams)]),
    "ksp_flagger_chunk_baselines": (c_int64, [POINTER(FlaggerParams)]),
    "ksp_flagger": (
        c_int,
        [c_void_p, POINTER(FlaggerParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_size_t],
    ),
    "ksp_flagger_stats": (
        c_int, [c_void_p, POINTER(FlaggerParams), c_void_p, POINTER(ctypes.c_ulonglong), c_int]),
    "ksp_flagger_is_dataflow": (c_int, [POINTER(FlaggerParams)]),
    "ksp_twodflag_scratch_bytes": (c_size_t, [POINTER(TwodflagParams), c_int64]),
    "ksp_twodflag_resident_baselines": (c_int, []),
    "ksp_twodflag_phases": (c_int, [POINTER(ctypes.c_ulonglong), c_int, c_int]),
    "ksp_twodflag": (
        c_int,
        [c_void_p, POINTER(TwodflagParams), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int64],
    ),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach prototypes.  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C katsdpsigproc_b200/csrc`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def error_string(code: int) -> str:
    return load().ksp_error_string(code).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise KspError(code, what, error_string(code))


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise on failure."""
    check(getattr(load(), name)(*args), name)


def kernel_launch_count() -> int:
    """Kernels launched by the library in this process so far."""
    count = ctypes.c_ulonglong()
    call("ksp_kernel_launch_count", byref(count))
    return count.value


STAGE_NAMES = ("background", "noise", "threshold", "expand_flags")


# indices of ksp_flagger_stats (KSP_DF_STAT_*)
DF_STAT_NAMES = ("cycles_background", "cycles_noise", "cycles_threshold", "cycles_expand",
                 "cycles_wait", "items", "fallbacks", "error")


def profile_enable(on: bool) -> None:
    call("ksp_profile_enable", int(on))


def profile_read() -> dict:
    """{stage: (milliseconds, launches)} accumulated by ``ksp_flagger`` since the last read."""
    ms = (c_double * len(STAGE_NAMES))()
    n = (c_int * len(STAGE_NAMES))()
    call("ksp_profile_read", ms, n, len(STAGE_NAMES))
    return {name: (ms[i], n[i]) for i, name in enumerate(STAGE_NAMES)}


def default_abs_mode() -> int:
    """Amplitude rule used when a template does not say otherwise.

    ``KATSDPSIGPROC_B200_ABS_MODE`` = ``numpy`` (default: equals ``np.abs`` of numpy on
    AVX-512F hosts) or ``hypot`` (correctly rounded, equals numpy elsewhere).
    """
    value = os.environ.get("KATSDPSIGPROC_B200_ABS_MODE", "numpy").lower()
    if value in ("numpy", "0"):
        return ABS_NUMPY
    if value in ("hypot", "1"):
        return ABS_HYPOT
    raise ValueError("KATSDPSIGPROC_B200_ABS_MODE must be 'numpy' or 'hypot'")
