"""numpy restatement of the reference's 2-D flagger ``rfi/twodflag.py`` (SumThresholdFlagger).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference implements this flagger in numba (``twodflag.py:67-890``); the arithmetic that
decides the results - which values are float32, which accumulators float64, in which ORDER the
running sums are formed - is restated here with explicit dtypes, one function per reference
function, vectorised across the axis the reference loops over independently and serial along the
axis whose order matters.  Every function is pinned bit for bit against the reference itself
(``tests/test_oracle_twodflag.py``: imported from ``oracle/_ref`` where numba is available, and
against ``tests/golden/reference_twodflag.npz`` everywhere).  The CUDA implementation
(``csrc/twodflag.cu``) follows THIS file operation for operation.

Shapes: a baseline is a ``(time, frequency)`` float32 array with a bool flag array.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

MAD_NORMAL = 1.4826        # reference rfi/__init__.py:31


def abs_hypot(x: np.ndarray) -> np.ndarray:
    """Magnitudes as numba's ``np.abs`` gives them (twodflag.py:103): for complex64 the correctly
    rounded hypot, for real input the plain absolute value."""
    if np.iscomplexobj(x):
        return np.hypot(x.real.astype(np.float64), x.imag.astype(np.float64)).astype(np.float32)
    return np.abs(x).astype(np.float32)


def average_freq(in_data: np.ndarray, in_flags: np.ndarray, factor: int
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """twodflag.py:68-116.  ``(time, freq, baseline)`` -> ``(baseline, time, ceil(freq / factor))``
    float32 means of the unflagged, non-NaN magnitudes of each group of ``factor`` channels
    (summed in channel order in float32, divided by the count), 0 and flagged if there are none."""
    n_time, n_freq, n_bl = in_data.shape
    a_freq = (n_freq + factor - 1) // factor
    mag = abs_hypot(in_data)
    use = (~in_flags.astype(np.bool_)) & ~np.isnan(mag)
    total = np.zeros((n_time, a_freq, n_bl), np.float32)
    count = np.zeros((n_time, a_freq, n_bl), np.int64)
    for j in range(n_freq):                                   # float32 sum in channel order
        jo = j // factor
        total[:, jo] = np.where(use[:, j], total[:, jo] + mag[:, j], total[:, jo])
        count[:, jo] += use[:, j]
    flags = count == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        # the weight has the dtype of `factor` (uint8 .. ): float32 / small int is a float32 division
        avg = np.where(flags, np.float32(0), total / count.astype(np.float32)).astype(np.float32)
    return np.ascontiguousarray(avg.transpose(2, 0, 1)), np.ascontiguousarray(flags.transpose(2, 0, 1))


def _median_f32(values: np.ndarray) -> np.float32:
    """np.median of a non-empty float32 vector as numba computes it: the middle element, or the
    float32 mean of the two middle ones."""
    s = np.sort(values)
    n = s.size
    if n & 1:
        return s[n // 2]
    return np.float32((s[n // 2 - 1] + s[n // 2]) * np.float32(0.5))


def time_median(data: np.ndarray, flags: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """twodflag.py:120-158: per channel the median over time of the unflagged values; channels
    without any become 0 and flagged.  Returns ``(1, freq)`` arrays."""
    n_time, n_freq = data.shape
    out = np.zeros((1, n_freq), np.float32)
    out_flags = np.zeros((1, n_freq), np.bool_)
    for f in range(n_freq):
        vals = data[~flags[:, f], f]
        if vals.size == 0:
            out_flags[0, f] = True
        else:
            out[0, f] = _median_f32(vals)
    return out, out_flags


def median_abs(data: np.ndarray, flags: np.ndarray) -> float:
    """twodflag.py:162-173: median of |data| over the unflagged elements, as float64 (NaN if none)."""
    vals = np.abs(data[~flags])
    if vals.size == 0:
        return float("nan")
    return float(_median_f32(vals.astype(np.float32)))


def median_abs_axis0(data: np.ndarray, flags: np.ndarray) -> np.ndarray:
    """twodflag.py:177-196: the same along axis 0, float32, first axis kept with length 1."""
    rest = data.shape[1:]
    out = np.empty((1,) + rest, np.float32)
    d2 = data.reshape(data.shape[0], -1)
    f2 = flags.reshape(flags.shape[0], -1)
    flat = out.reshape(-1)
    for j in range(d2.shape[1]):
        vals = np.abs(d2[~f2[:, j], j])
        flat[j] = np.nan if vals.size == 0 else _median_f32(vals)
    return out


def linearly_interpolate_nans(data: np.ndarray) -> None:
    """twodflag.py:200-251, in place, row by row: NaNs replaced by linear interpolation between
    their valid neighbours (``start + k * grad`` with ``grad`` in float64), the ends by the nearest
    valid value, an all-NaN row by zeros."""
    for row in data:
        n = row.size
        good = np.flatnonzero(~np.isnan(row))
        if good.size == 0:
            row[:] = 0
            continue
        row[:good[0]] = row[good[0]]
        row[good[-1] + 1:] = row[good[-1]]
        for a, b in zip(good[:-1], good[1:]):
            if b > a + 1:
                start = row[a]
                # (row[b] - start) is a float32 subtraction in the reference, the quotient float64
                grad = np.float64(np.float32(row[b] - start)) / np.float64(b - a)
                k = np.arange(1, b - a, dtype=np.float64)
                row[a + 1:b] = (np.float64(start) + k * grad).astype(np.float32)


def f32_pow_int(base: int, exponent: int) -> np.float32:
    """``np.float32(base) ** exponent`` as numba evaluates it: binary exponentiation with every
    product rounded to float32 (for exponent 4: the square of the square)."""
    result, b, k = np.float32(1), np.float32(base), int(exponent)
    while k:
        if k & 1:
            result = np.float32(result * b)
        k >>= 1
        if k:
            b = np.float32(b * b)
    return result


def box_gaussian_filter1d(data: np.ndarray, r: int, out: np.ndarray, passes: int) -> None:
    """twodflag.py:255-309 along axis 0: ``passes`` running box sums of width ``2 r + 1`` (zeros
    outside the array), the accumulator in float64, every pass stored as float32, the result
    divided in float32 by ``float32(2 r + 1) ** passes`` (:func:`f32_pow_int`).  May run in place."""
    K = passes
    if data.shape[0] == 0 or K == 0:
        out[...] = data
        return
    n = data.shape[0]
    d = 2 * r + 1
    padding = r * K
    padded = np.empty((n + padding,) + data.shape[1:], np.float32)
    padded[:padding] = 0
    padded[padding:] = data
    prev_start = padding
    for p in range(1, K + 1):
        s = np.zeros(data.shape[1:], np.float64)
        start = padding - 2 * r * p
        stop = start + n + 2 * padding
        start = max(start, 0)
        stop = min(stop, padded.shape[0])
        tail = min(stop, padded.shape[0] - 2 * r)
        for i in range(prev_start, min(start + 2 * r, padded.shape[0])):
            s = s + padded[i]
        for i in range(start, tail):
            s = s + padded[i + 2 * r]
            prev = padded[i].copy()
            padded[i] = s.astype(np.float32)
            s = s - prev
        for i in range(tail, stop):
            prev = padded[i].copy()
            padded[i] = s.astype(np.float32)
            s = s - prev
        prev_start = start
    out[...] = padded[:n] / f32_pow_int(d, K)


def box_radius(sigma: np.ndarray, passes: int) -> np.ndarray:
    """twodflag.py:341."""
    return (0.5 * np.sqrt(12.0 * np.asarray(sigma, np.float64) ** 2 / passes + 1)).astype(np.int_)


def box_gaussian_filter(data: np.ndarray, sigma: Sequence[float], out: np.ndarray, passes: int = 4) -> None:
    """twodflag.py:313-356: time axis first (if its radius is > 0), then frequency."""
    r = box_radius(np.asarray(sigma), passes)
    src = data
    copied = False
    if r[0] > 0:
        box_gaussian_filter1d(src, int(r[0]), out, passes)
        src = out
        copied = True
    if r[1] > 0:
        tmp = np.empty_like(out.T)
        box_gaussian_filter1d(np.ascontiguousarray(src.T), int(r[1]), tmp, passes)
        out[...] = tmp.T
        copied = True
    if not copied:
        out[...] = data


def masked_gaussian_filter(data: np.ndarray, flags: np.ndarray, sigma: Sequence[float],
                           out: np.ndarray, passes: int = 4) -> None:
    """twodflag.py:360-400: filtered (data with flagged samples zeroed) / filtered (weights),
    NaN where the filtered weight is exactly 0."""
    weight = (~flags.astype(np.bool_)).astype(np.float32)
    out[...] = np.where(flags, np.float32(0), data)
    box_gaussian_filter(weight, sigma, weight, passes)
    box_gaussian_filter(out, sigma, out, passes)
    with np.errstate(invalid="ignore", divide="ignore"):
        out[...] = np.where(weight == 0, np.float32(np.nan), out / weight)


def get_background2d(data: np.ndarray, flags: np.ndarray, iterations: int, spike_width: np.ndarray,
                     reject_threshold: float, freq_chunk_ends: np.ndarray) -> np.ndarray:
    """twodflag.py:404-463."""
    flags = flags.astype(np.bool_).copy()
    background = np.empty_like(data)
    for extend_factor in range(iterations, 0, -1):
        masked_gaussian_filter(data, flags, extend_factor * np.asarray(spike_width, np.float64), background)
        for c in range(len(freq_chunk_ends) - 1):
            sub = (slice(None), slice(int(freq_chunk_ends[c]), int(freq_chunk_ends[c + 1])))
            residual = np.abs(data[sub] - background[sub])         # float32
            background[sub] = residual
            threshold = median_abs(residual, flags[sub])            # float64
            threshold *= MAD_NORMAL * reject_threshold
            with np.errstate(invalid="ignore"):
                flags[sub] |= residual.astype(np.float64) > threshold
    masked_gaussian_filter(data, flags, np.asarray(spike_width, np.float64), background)
    linearly_interpolate_nans(background)
    return background


def _convolve_flags(in_values: np.ndarray, scale: np.float32, threshold: np.ndarray,
                    out_flags: np.ndarray, window: int) -> None:
    """twodflag.py:467-489 along axis 0: flag ``v * scale > threshold`` and smear every hit over
    the ``window`` samples of its window."""
    with np.errstate(invalid="ignore"):
        hit = in_values * np.float64(scale) > threshold.astype(np.float64)
    n = out_flags.shape[0]
    cum = np.zeros((n + 1,) + out_flags.shape[1:], np.int64)
    # hit[i] covers samples i .. i + window - 1
    smeared = np.zeros(out_flags.shape, np.bool_)
    for i in range(hit.shape[0]):
        smeared[i:i + window] |= hit[i]
    out_flags |= smeared


def sum_threshold1d(input_data: np.ndarray, input_flags: np.ndarray, output_flags: np.ndarray,
                    windows: Sequence[int], outlier_nsigma: float, rho: float, chunks: np.ndarray) -> None:
    """twodflag.py:493-560 along axis 0 (the other axes are independent)."""
    max_window = int(np.max(windows)) if len(windows) else 1
    n = input_data.shape[0]
    for ci in range(len(chunks) - 1):
        c0, c1 = int(chunks[ci]), int(chunks[ci + 1])
        threshold = median_abs_axis0(input_data[c0:c1], input_flags[c0:c1])   # float32
        scale = outlier_nsigma * MAD_NORMAL                                    # float64
        threshold = np.where(np.isnan(threshold), np.float32(np.inf),
                             (threshold.astype(np.float64) * scale).astype(np.float32))
        p0 = max(c0 - max_window + 1, 0)
        p1 = min(c1 + max_window - 1, n)
        padded = input_data[p0:p1]
        pos = np.zeros(padded.shape, np.bool_)
        neg = np.zeros(padded.shape, np.bool_)
        for window in windows:
            window = int(window)
            tf = pow(rho, np.log2(window))
            this = (threshold.astype(np.float64) / tf).astype(np.float32)      # (1, ...) float32
            limit = this[0]
            clamped = padded.astype(np.float32).copy()
            clamped = np.where(pos & (clamped > limit), limit, clamped)
            clamped = np.where(~(pos & (padded > limit)) & neg & (padded < -limit), -limit, clamped)
            cum = np.zeros((padded.shape[0] + 1,) + padded.shape[1:], np.float64)
            for i in range(padded.shape[0]):                                   # float64, in order
                cum[i + 1] = cum[i] + clamped[i]
            avg = cum[window:] - cum[:-window]
            rolling = np.float32(1.0 / window)
            _convolve_flags(avg, rolling, this, pos, window)
            _convolve_flags(avg, -rolling, this, neg, window)
        output_flags[c0:c1] = (pos | neg)[c0 - p0:c1 - p0]


def sum_threshold(input_data: np.ndarray, input_flags: np.ndarray, axis: int, windows: Sequence[int],
                  outlier_nsigma: float, rho: float, chunks: Optional[np.ndarray] = None) -> np.ndarray:
    """twodflag.py:564-631."""
    if chunks is None:
        chunks = np.array([0, input_data.shape[axis]])
    out = np.empty(input_data.shape, np.bool_)
    if axis == 1:
        tmp = np.empty(input_data.T.shape, np.bool_)
        sum_threshold1d(np.ascontiguousarray(input_data.T), np.ascontiguousarray(input_flags.T), tmp,
                        windows, outlier_nsigma, rho, chunks)
        out[...] = tmp.T
    elif axis == 0:
        sum_threshold1d(input_data, input_flags, out, windows, outlier_nsigma, rho, chunks)
    else:
        raise ValueError("axis must be 0 or 1")
    return out


def combine_flags(spec_flags: np.ndarray, time_flags: np.ndarray, freq_flags: np.ndarray,
                  time_extend: int) -> np.ndarray:
    """twodflag.py:691-722: OR the three sources and smear over ``time_extend`` dumps."""
    n_time, n_freq = time_flags.shape
    flag = spec_flags[0][None, :] | time_flags | freq_flags
    cum = np.zeros((n_time + 1, n_freq), np.int64)
    cum[1:] = np.cumsum(flag, axis=0)
    lo = -(int(time_extend) // 2)
    hi = lo + int(time_extend)
    t = np.arange(n_time)
    t0 = np.maximum(t + lo, 0)
    t1 = np.minimum(t + hi, n_time)
    return cum[t0] != cum[t1]


def unaverage_freq(flags: np.ndarray, freq_extend: int, average_freq_: int, flag_all_time_frac: float,
                   flag_all_freq_frac: float, orig_freq: int) -> np.ndarray:
    """twodflag.py:726-764."""
    n_time = flags.shape[0]
    rep = flags[:, np.arange(orig_freq) // int(average_freq_)]
    cum = np.zeros((n_time, orig_freq + 1), np.int64)
    cum[:, 1:] = np.cumsum(rep, axis=1)
    lo = -(int(freq_extend) // 2)
    hi = lo + int(freq_extend)
    f = np.arange(orig_freq)
    f0 = np.maximum(f + lo, 0)
    f1 = np.minimum(f + hi, orig_freq)
    out = cum[:, f1] != cum[:, f0]
    per_time = out.sum(axis=1)
    per_freq = out.sum(axis=0)                     # counted BEFORE whole rows are filled in
    out[per_time > flag_all_freq_frac * orig_freq, :] = True
    out[:, per_freq > n_time * flag_all_time_frac] = True
    return out


def get_baseline_flags(data: np.ndarray, flags: np.ndarray, orig_freq: int, outlier_nsigma: float,
                       windows_time: Sequence[int], windows_freq: Sequence[int], background_reject: float,
                       background_iterations: int, spike_width_time: float, spike_width_freq: float,
                       time_extend: int, freq_extend: int, freq_chunk_ends: np.ndarray, average_freq_: int,
                       flag_all_time_frac: float, flag_all_freq_frac: float, rho: float) -> np.ndarray:
    """twodflag.py:768-881 for one baseline (``data``, ``flags`` are modified)."""
    spec_data, spec_flags = time_median(data, flags)
    spec_background = get_background2d(spec_data, spec_flags, background_iterations,
                                       np.array((0.0, spike_width_freq)), background_reject, freq_chunk_ends)
    spec_data -= spec_background
    spec_flags = sum_threshold(spec_data, spec_flags, 1, windows_freq, outlier_nsigma, rho, freq_chunk_ends)
    flags |= spec_flags
    background = get_background2d(data, flags, background_iterations,
                                  np.array((spike_width_time, spike_width_freq)), background_reject,
                                  freq_chunk_ends)
    data -= background
    time_flags = sum_threshold(data, flags, 0, windows_time, outlier_nsigma, rho)
    flags |= time_flags
    freq_flags = sum_threshold(data, flags, 1, windows_freq, outlier_nsigma, rho, freq_chunk_ends)
    combined = combine_flags(spec_flags, time_flags, freq_flags, time_extend)
    return unaverage_freq(combined, freq_extend, average_freq_, flag_all_time_frac, flag_all_freq_frac,
                          orig_freq)


class SumThresholdFlagger:
    """twodflag.py:894-1026 (parameter conditioning) + :635-688 (the per-batch driver)."""

    def __init__(self, outlier_nsigma=4.5, windows_time=(1, 2, 4, 8), windows_freq=(1, 2, 4, 8),
                 background_reject=2.0, background_iterations=1, spike_width_time=12.5,
                 spike_width_freq=10.0, time_extend=3, freq_extend=3, freq_chunks=10, average_freq=1,
                 flag_all_time_frac=0.6, flag_all_freq_frac=0.8, rho=1.3):
        self.outlier_nsigma = outlier_nsigma
        self.windows_time = list(windows_time)
        wf = np.ceil(np.array(windows_freq, dtype=np.float32) / average_freq)
        self.windows_freq = np.unique(wf.astype(np.int_))
        self.background_reject = background_reject
        self.background_iterations = background_iterations
        self.spike_width_time = spike_width_time
        self.spike_width_freq = spike_width_freq / average_freq
        self.time_extend = int(time_extend)
        self.freq_extend = int(freq_extend)
        self.freq_chunks = freq_chunks
        self.average_freq = int(average_freq)
        self.flag_all_time_frac = flag_all_time_frac
        self.flag_all_freq_frac = flag_all_freq_frac
        self.rho = rho

    def get_flags(self, data: np.ndarray, flags: np.ndarray) -> np.ndarray:
        n_time, n_freq, n_bl = data.shape
        a_freq = (n_freq + self.average_freq - 1) // self.average_freq
        chunk_ends = np.linspace(0, a_freq, self.freq_chunks + 1).astype(np.int_)
        windows_time = np.array([w for w in self.windows_time if w <= n_freq], np.int_)   # (sic: twodflag.py:1003)
        windows_freq = np.array([w for w in self.windows_freq if w <= a_freq], np.int_)
        avg, avg_flags = average_freq(data, flags, self.average_freq)
        out = np.empty(data.shape, np.bool_)
        nan_in = np.isnan(data)
        for bl in range(n_bl):
            bl_flags = get_baseline_flags(
                avg[bl], avg_flags[bl], n_freq, self.outlier_nsigma, windows_time, windows_freq,
                self.background_reject, self.background_iterations, self.spike_width_time,
                self.spike_width_freq, self.time_extend, self.freq_extend, chunk_ends, self.average_freq,
                self.flag_all_time_frac, self.flag_all_freq_frac, self.rho)
            out[:, :, bl] = bl_flags | nan_in[:, :, bl]
        return out
