"""2-D RFI flagger: ``SumThresholdFlagger`` of the reference (``rfi/twodflag.py:894-1118``) on the GPU.

Same constructor, same ``get_flags(data, flags)`` with ``(time, frequency, baseline)`` arrays, same
flags - the reference's is a numba CPU implementation whose float64 running sums fix the result
to the last bit; ``csrc/twodflag.cu`` keeps those recurrences in the reference's order and takes
its parallelism from the baselines (and, inside a baseline, from the rows / columns that the
reference loops over independently).  Parameter conditioning (window clipping, frequency-chunk
boundaries, box radii of the Gaussian approximations, ``rho ** log2(window)``) is done here in
Python with the same expressions as the reference's ``__init__`` / ``_get_flags``
(``twodflag.py:951-1026,341,527``) and handed to the C ABI as plain numbers.

There is no CPU implementation in this module: without the CUDA library it fails on import of
the library, not silently.
"""

from __future__ import annotations

import concurrent.futures
import ctypes
from ctypes import byref, c_size_t
from typing import Any, Optional, Sequence

import numpy as np

from .. import _capi, accel


_STAGING_THREADS = 8
_pool: Optional[concurrent.futures.ThreadPoolExecutor] = None


def _staging_pool() -> concurrent.futures.ThreadPoolExecutor:
    """Threads that copy large inputs into the pinned staging arrays (created on first use)."""
    global _pool
    if _pool is None:
        _pool = concurrent.futures.ThreadPoolExecutor(_STAGING_THREADS, thread_name_prefix="twodflag-stage")
    return _pool


def _as_min_dtype(value: int) -> np.generic:
    """The smallest unsigned integer type that holds ``value`` (reference ``twodflag.py:28-42``;
    kept because ``time_extend`` / ``freq_extend`` / ``average_freq`` are public attributes)."""
    if value >= 0:
        for dtype in (np.uint8, np.uint16, np.uint32, np.uint64):
            if value <= np.iinfo(dtype).max:
                return dtype(value)
    raise ValueError("Value must be a non-negative integer that fits in 64 bits")


class SumThresholdFlagger:
    """Flagger that detects spikes along both the frequency and the time axis (SumThreshold,
    Offringa 2010).  Parameters as the reference's class of the same name
    (``twodflag.py:917-949``); ``context`` (extension, keyword only) is the device context to run
    on - by default one is created on first use with :func:`accel.create_some_context`.
    """

    def __init__(self, outlier_nsigma: float = 4.5, windows_time: Sequence[int] = (1, 2, 4, 8),
                 windows_freq: Sequence[int] = (1, 2, 4, 8), background_reject: float = 2.0,
                 background_iterations: int = 1, spike_width_time: float = 12.5,
                 spike_width_freq: float = 10.0, time_extend: int = 3, freq_extend: int = 3,
                 freq_chunks: int = 10, average_freq: int = 1, flag_all_time_frac: float = 0.6,
                 flag_all_freq_frac: float = 0.8, rho: float = 1.3, *, context: Any = None) -> None:
        self.outlier_nsigma = outlier_nsigma
        self.windows_time = windows_time
        # scale the frequency windows by the averaging, drop duplicates (twodflag.py:969-971)
        scaled = np.ceil(np.array(windows_freq, dtype=np.float32) / average_freq)
        self.windows_freq = np.unique(scaled.astype(np.int_))
        self.background_reject = background_reject
        self.background_iterations = background_iterations
        self.spike_width_time = spike_width_time
        self.spike_width_freq = spike_width_freq / average_freq
        self.time_extend = _as_min_dtype(time_extend)
        self.freq_extend = _as_min_dtype(freq_extend)
        self.freq_chunks = freq_chunks
        self.average_freq = _as_min_dtype(average_freq)
        self.flag_all_time_frac = flag_all_time_frac
        self.flag_all_freq_frac = flag_all_freq_frac
        self.rho = rho
        self._context = context
        self._queue: Any = None
        self._buffers: dict = {}

    # ------------------------------------------------------------------ parameters for the C ABI
    def _params(self, shape: Sequence[int], is_complex: bool) -> "_capi.TwodflagParams":
        n_time, n_freq, n_bl = (int(x) for x in shape)
        average_freq = int(self.average_freq)
        averaged = (n_freq + average_freq - 1) // average_freq
        chunk_ends = np.linspace(0, averaged, self.freq_chunks + 1).astype(np.int_)
        # clipped to the data (the time windows against shape[1], as the reference does)
        windows_time = [int(w) for w in self.windows_time if w <= n_freq]
        windows_freq = [int(w) for w in self.windows_freq if w <= averaged]
        p = _capi.TwodflagParams()
        p.n_time, p.n_freq, p.n_bl = n_time, n_freq, n_bl
        p.is_complex = int(is_complex)
        p.average_freq = average_freq
        for name, windows in (("time", windows_time), ("freq", windows_freq)):
            if len(windows) > _capi.TWOD_MAX_WINDOWS:
                raise ValueError(f"at most {_capi.TWOD_MAX_WINDOWS} window sizes per axis")
            if windows and max(windows) > _capi.TWOD_MAX_WINDOW:
                raise ValueError(f"window sizes up to {_capi.TWOD_MAX_WINDOW} are supported")
            setattr(p, f"n_windows_{name}", len(windows))
            for i, w in enumerate(windows):
                getattr(p, f"windows_{name}")[i] = w
                getattr(p, f"tf_{name}")[i] = pow(self.rho, np.log2(w))       # twodflag.py:527
        p.outlier_nsigma = self.outlier_nsigma
        p.background_reject = self.background_reject
        if not 1 <= self.background_iterations <= _capi.TWOD_MAX_ITERATIONS:
            raise ValueError(f"background_iterations must be 1 .. {_capi.TWOD_MAX_ITERATIONS}")
        p.background_iterations = self.background_iterations
        passes = 4
        for extend_factor in range(1, self.background_iterations + 1):
            sigma = extend_factor * np.array((self.spike_width_time, self.spike_width_freq))
            r = (0.5 * np.sqrt(12.0 * sigma**2 / passes + 1)).astype(np.int_)   # twodflag.py:341
            p.r_time[extend_factor] = int(r[0])
            p.r_freq[extend_factor] = int(r[1])
        p.time_extend = int(self.time_extend)
        p.freq_extend = int(self.freq_extend)
        if not 1 <= self.freq_chunks <= _capi.TWOD_MAX_CHUNKS:
            raise ValueError(f"freq_chunks must be 1 .. {_capi.TWOD_MAX_CHUNKS}")
        p.n_chunks = self.freq_chunks
        for i, end in enumerate(chunk_ends):
            p.chunk_ends[i] = int(end)
        p.flag_all_time_frac = self.flag_all_time_frac
        p.flag_all_freq_frac = self.flag_all_freq_frac
        return p

    def _ensure_queue(self) -> Any:
        if self._queue is None:
            if self._context is None:
                self._context = accel.create_some_context(interactive=False)
            self._queue = self._context.create_command_queue()
        return self._queue

    # ------------------------------------------------------------------ the reference's entry point
    def get_flags(self, data: np.ndarray, flags: np.ndarray, pool: Any = None,
                  chunk_size: Optional[int] = None, is_multiprocess: Optional[bool] = None
                  ) -> np.ndarray:
        """Flags for ``data`` (``(time, frequency, baseline)``, complex64 visibilities or real
        magnitudes); ``flags`` of the same shape marks samples to ignore.  Returns a bool array of
        that shape.  ``pool`` and ``is_multiprocess`` (the reference's CPU executors) are accepted
        and ignored; ``chunk_size`` is the number of baselines in flight on the device at a time
        (default: enough to fill it, within ~8 GB of scratch).
        """
        if data.shape != flags.shape:
            raise ValueError("Shape mismatch")
        if data.ndim != 3:
            raise ValueError("data has wrong number of dimensions")
        is_complex = np.iscomplexobj(data)
        if data.size == 0:
            return np.empty(data.shape, np.bool_)
        p = self._params(data.shape, is_complex)
        queue = self._ensure_queue()
        context = queue.context
        lib = _capi.load()
        context._make_current()
        per_baseline = int(lib.ksp_twodflag_scratch_bytes(byref(p), 1))
        if per_baseline == 0:
            raise ValueError("parameters outside the supported range")
        n_bl = data.shape[2]
        if not chunk_size:
            chunk_size = max(1, min(n_bl, int(lib.ksp_twodflag_resident_baselines()),
                                    (8 << 30) // per_baseline))
        chunk_size = int(min(chunk_size, n_bl))
        buf = self._buffers_for(context, data.shape, np.complex64 if is_complex else np.float32,
                                per_baseline * chunk_size)
        # into the pinned staging arrays (the conversions of dtype / layout happen in these copies)
        h_flags = buf["h_flags"].view(np.bool_)

        def stage(rows: slice) -> None:
            np.copyto(buf["h_data"][rows], data[rows], casting="unsafe")
            if flags.dtype == np.bool_:
                np.copyto(h_flags[rows], flags[rows])
            else:
                np.not_equal(flags[rows], 0, out=h_flags[rows])

        # Time slices: staged by the pool's threads (numpy copies release the GIL) and uploaded one
        # by one as they become ready, so the upload of the first slices runs under the staging of
        # the later ones; small inputs take the plain path.
        n_time = data.shape[0]
        stream = ctypes.c_void_p(queue.stream)
        row_items = int(np.prod(data.shape[1:]))

        def upload(rows: slice) -> None:
            for dev, host in ((buf["d_data"], buf["h_data"]), (buf["d_flags"], buf["h_flags"])):
                step = row_items * host.dtype.itemsize
                _capi.call("ksp_memcpy_async", ctypes.c_void_p(dev.buffer.ptr + rows.start * step),
                           ctypes.c_void_p(host.ctypes.data + rows.start * step),
                           c_size_t((rows.stop - rows.start) * step), _capi.H2D, stream)

        large = data.nbytes >= (64 << 20) and n_time > 1
        slices = [slice(None)]
        if large:
            bounds = np.linspace(0, n_time, min(n_time, 2 * _STAGING_THREADS) + 1).astype(int)
            slices = [slice(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:])]
            for rows, done in zip(slices, [_staging_pool().submit(stage, rows) for rows in slices]):
                done.result()
                upload(rows)
        else:
            stage(slice(None))
            buf["d_data"].set_async(queue, buf["h_data"])
            buf["d_flags"].set_async(queue, buf["h_flags"])
        _capi.call("ksp_twodflag", stream, byref(p),
                   ctypes.c_void_p(buf["d_data"].buffer.ptr), ctypes.c_void_p(buf["d_flags"].buffer.ptr),
                   ctypes.c_void_p(buf["d_out"].buffer.ptr), ctypes.c_void_p(buf["d_scratch"].buffer.ptr),
                   c_size_t(per_baseline * chunk_size), ctypes.c_int64(chunk_size))
        buf["d_out"].get_async(queue, buf["h_out"])
        queue.finish()
        # the kernel writes 0 / 1; the caller gets a copy it owns
        h_out = buf["h_out"].view(np.bool_)
        if not large:
            return np.array(h_out)
        out = np.empty(data.shape, np.bool_)
        list(_staging_pool().map(lambda rows: np.copyto(out[rows], h_out[rows]), slices))
        return out

    def _buffers_for(self, context: Any, shape: Sequence[int], dtype: Any, scratch_bytes: int) -> dict:
        """Device arrays and pinned staging arrays for one input shape, kept from call to call
        (``get_flags`` is called again and again with blocks of the same shape; allocating
        pinned and device memory costs more than the flagging)."""
        key = (tuple(shape), np.dtype(dtype).str)
        buf = self._buffers.get("arrays")
        if buf is None or buf["key"] != key:
            self._buffers.clear()
            buf = {"key": key}
            buf["h_data"] = accel.HostArray(tuple(shape), dtype, context=context)
            buf["h_flags"] = accel.HostArray(tuple(shape), np.uint8, context=context)
            buf["h_out"] = accel.HostArray(tuple(shape), np.uint8, context=context)
            buf["d_data"] = accel.DeviceArray(context, tuple(shape), dtype)
            buf["d_flags"] = accel.DeviceArray(context, tuple(shape), np.uint8)
            buf["d_out"] = accel.DeviceArray(context, tuple(shape), np.uint8)
            buf["d_scratch"] = None
            self._buffers["arrays"] = buf
        if buf["d_scratch"] is None or buf["d_scratch"].shape[0] < scratch_bytes:
            buf["d_scratch"] = None                                # release before the larger allocation
            buf["d_scratch"] = accel.DeviceArray(context, (scratch_bytes,), np.uint8)
        return buf

        p = self._params(data.shape, is_complex)
        queue = self._ensure_queue()
        context = queue.context
        lib = _capi.load()
        context._make_current()
        per_baseline = int(lib.ksp_twodflag_scratch_bytes(byref(p), 1))
        if per_baseline == 0:
            raise ValueError("parameters outside the supported range")
        n_bl = data.shape[2]
        if not chunk_size:
            chunk_size = max(1, min(n_bl, int(lib.ksp_twodflag_resident_baselines()), (8 << 30) // per_baseline))
        chunk_size = int(min(chunk_size, n_bl))
        d_data = accel.DeviceArray(context, host_data.shape, host_data.dtype)
        d_flags = accel.DeviceArray(context, host_flags.shape, np.uint8)
        d_out = accel.DeviceArray(context, host_flags.shape, np.uint8)
        d_scratch = accel.DeviceArray(context, (per_baseline * chunk_size,), np.uint8)
        d_data.set(queue, host_data)
        d_flags.set(queue, host_flags)
        _capi.call("ksp_twodflag", ctypes.c_void_p(queue.stream), byref(p),
                   ctypes.c_void_p(d_data.buffer.ptr), ctypes.c_void_p(d_flags.buffer.ptr),
                   ctypes.c_void_p(d_out.buffer.ptr), ctypes.c_void_p(d_scratch.buffer.ptr),
                   c_size_t(per_baseline * chunk_size), ctypes.c_int64(chunk_size))
        result = d_out.get(queue)
        np.not_equal(result, 0, out=out)
        return out
