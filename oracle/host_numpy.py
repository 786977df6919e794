"""numpy/pandas restatement of the reference CPU flagger.  TEST INFRASTRUCTURE ONLY.

This module restates, in float64 and with the same numpy/pandas primitives,
what ``katsdpsigproc.rfi.host`` computes (reference ``src/katsdpsigproc/rfi/
host.py``), plus the numpy expressions the reference's tests use as the oracle
for Percentile5, MaskedSum and Transpose.  It is the *reference-equivalent*
oracle: its outputs are byte-identical to the reference host classes on the
same inputs (checked in ``tests/test_oracle_golden.py`` against fixtures made
by importing the real reference, ``tests/golden/make_golden.py``).

Parity status: PINNED (reference known-answer tests + reference-generated
fixtures).

Third-party arithmetic at the boundary: numpy (``np.abs`` on complex64,
``np.median``, ``np.convolve``, ``np.percentile``) and pandas
(``DataFrame.rolling(...).median()``), both present in the image (numpy 2.3.5,
pandas 3.0.2); the reference only pins lower bounds (``setup.cfg:35-44``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this file.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

#: Ratio of the standard deviation to the median absolute deviation of a
#: normal distribution (reference ``src/katsdpsigproc/rfi/__init__.py:31``).
MAD_NORMAL = 1.4826


# --------------------------------------------------------------------------
# Stage 1: background (reference rfi/host.py:133-151)
# --------------------------------------------------------------------------
def background_median_filter(
    vis: np.ndarray,
    width: int,
    flags: Optional[np.ndarray] = None,
    amplitudes: bool = False,
) -> np.ndarray:
    """Deviation of each amplitude from a centred sliding median along channels.

    ``vis`` is (channels, baselines): complex64 visibilities, or amplitudes
    when ``amplitudes`` is true.  ``flags`` (any integer type; non-zero means
    "known bad") is either per channel (1-D) or per sample (2-D).  Flagged
    samples do not take part in the median and come out as 0.  Windows are
    clipped at the band edges (``min_periods=1``); a window holding an even
    number of usable samples yields the mean of the two middle ones.  The
    result is float64, as in the reference (host.py:147-151).
    """
    frame = pd.DataFrame(vis if amplitudes else np.abs(vis))
    if flags is not None:
        bad = np.asarray(flags).astype(np.bool_)
        if bad.ndim == 1:
            bad = bad[:, np.newaxis]
        frame = frame.mask(np.broadcast_to(bad, vis.shape))
    background = frame.rolling(width, center=True, min_periods=1).median()
    deviations = frame - background
    return deviations.fillna(0).values


# --------------------------------------------------------------------------
# Stage 2: noise estimate (reference rfi/host.py:157-163)
# --------------------------------------------------------------------------
def median_abs_nonzero(deviations: np.ndarray) -> np.ndarray:
    """Per-baseline median of the non-zero absolute deviations (before scaling)."""
    n_baselines = deviations.shape[1]
    med = np.empty(n_baselines)
    for bl in range(n_baselines):
        mag = np.abs(deviations[:, bl])
        med[bl] = np.median(mag[mag > 0])
    return med


def noise_est_mad(deviations: np.ndarray) -> np.ndarray:
    """Per-baseline noise sigma: 1.4826 x median(|deviation| over non-zero samples)."""
    return median_abs_nonzero(deviations) * MAD_NORMAL


# --------------------------------------------------------------------------
# Stage 3: thresholds (reference rfi/host.py:177-183 and :203-254)
# --------------------------------------------------------------------------
def threshold_simple(
    deviations: np.ndarray, noise: np.ndarray, n_sigma: float, flag_value: int = 1
) -> np.ndarray:
    """Flag every sample above ``n_sigma`` noise sigmas, independently."""
    hit = (deviations > n_sigma * noise).astype(np.uint8)
    return hit * flag_value


def sum_threshold_baseline(
    deviations: np.ndarray,
    threshold1,
    n_windows: int = 4,
    threshold_falloff: float = 1.2,
    near: Optional[Dict[str, int]] = None,
    rel_eps: float = 1e-6,
) -> np.ndarray:
    """Offringa SumThreshold on one baseline (reference host.py:218-246).

    Window sizes are 1, 2, 4, ... ``2**(n_windows-1)``; the per-sample
    threshold for window index ``w`` is ``float32(threshold1 * falloff**-w)``.
    Before each window size, samples already flagged are replaced by the
    current threshold; only windows lying fully inside the band are summed
    (``mode='valid'``); every member of a window whose sum exceeds
    ``threshold * window`` is flagged.

    If ``near`` is a dict, it receives counts used by the parity tests:
    ``near['band']`` = number of (window, position) pairs whose sum lies
    within ``rel_eps`` (relative) of the decision value, *excluding* windows
    whose members were all already flagged (those sum to exactly the decision
    value and can never fire: SURVEY.md R6).
    """
    data = np.array(deviations, copy=True)
    flagged = np.zeros(data.shape, dtype=np.bool_)
    for w in range(n_windows):
        window = 2**w
        level = np.float32(threshold1 * pow(threshold_falloff, -w))
        data[flagged] = level
        sums = np.convolve(data, np.ones(window), mode="valid")
        limit = level * window
        over = sums > limit
        if near is not None and sums.size:
            n_flagged = np.convolve(flagged.astype(np.float64), np.ones(window), mode="valid")
            close = np.abs(sums - np.float64(limit)) <= rel_eps * np.abs(np.float64(limit))
            close &= n_flagged < window
            near["band"] = near.get("band", 0) + int(close.sum())
        flagged |= np.convolve(over, np.ones(window, dtype=np.bool_))
    return flagged


def threshold_sum(
    deviations: np.ndarray,
    noise: np.ndarray,
    n_sigma: float,
    n_windows: int = 4,
    threshold_falloff: float = 1.2,
    flag_value: int = 1,
    near: Optional[Dict[str, int]] = None,
) -> np.ndarray:
    """SumThreshold over every baseline of a (channels, baselines) array."""
    out = np.empty(deviations.shape, dtype=np.uint8)
    for bl in range(deviations.shape[1]):
        hit = sum_threshold_baseline(
            deviations[:, bl], n_sigma * noise[bl], n_windows, threshold_falloff, near
        )
        out[:, bl] = hit * np.uint8(flag_value)
    return out


# --------------------------------------------------------------------------
# The composed flagger (reference rfi/host.py:257-273)
# --------------------------------------------------------------------------
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
    simple_threshold: bool = False,
    stages: Optional[dict] = None,
    near: Optional[Dict[str, int]] = None,
) -> np.ndarray:
    """background -> MAD noise -> threshold, exactly as ``FlaggerHost.__call__``.

    ``stages`` (a dict), when given, receives the intermediate ``deviations``
    and ``noise`` arrays (float64) for stage-by-stage parity checks.
    """
    deviations = background_median_filter(vis, width, input_flags, amplitudes)
    noise = noise_est_mad(deviations)
    if stages is not None:
        stages["deviations"] = deviations
        stages["noise"] = noise
    if simple_threshold:
        return threshold_simple(deviations, noise, n_sigma, flag_value)
    return threshold_sum(
        deviations, noise, n_sigma, n_windows, threshold_falloff, flag_value, near
    )


# --------------------------------------------------------------------------
# Helper operations: the numpy expressions the reference's tests compare with
# --------------------------------------------------------------------------
def percentile5(data: np.ndarray, column_range: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """[min, max, 25 %, 75 %, 50 %] "lower" percentiles of |data| along each row.

    Reference oracle expression: ``test/test_percentile.py:79-84``; rank
    formulas ``(n-1)/4, 3(n-1)/4, (n-1)/2`` in ``percentile.mako:131-133``.
    Returns float32 of shape (5, rows).
    """
    if column_range is None:
        column_range = (0, data.shape[1])
    sub = np.abs(data[:, column_range[0] : column_range[1]])
    out = np.percentile(sub, [0, 100, 25, 75, 50], axis=1, method="lower")
    return out.astype(np.float32)


def masked_sum(data: np.ndarray, mask: np.ndarray, use_amplitudes: bool = False) -> np.ndarray:
    """Per-column sum over rows of ``mask[row] * data[row, col]`` (or of ``|data|``).

    Reference oracle expression: ``test/test_maskedsum.py:62-67``.
    """
    terms = np.abs(data) if use_amplitudes else data
    return np.sum(terms * mask.reshape(data.shape[0], 1), axis=0)


def transpose(data: np.ndarray) -> np.ndarray:
    """Reference oracle expression: ``test/test_transpose.py:59`` (``ary.T``)."""
    return np.ascontiguousarray(data.T)


# --------------------------------------------------------------------------
# Synthetic visibilities (G1 noise + G2 spikes)
# --------------------------------------------------------------------------
def synthetic_vis(
    channels: int,
    baselines: int,
    seed: int = 1,
    spike_prob: float = 1.0 / 64.0,
    line_prob: float = 0.005,
    chunk: int = 1024,
) -> Tuple[np.ndarray, np.ndarray]:
    """Noise + injected RFI; returns ``(vis complex64, spikes uint8)``.

    Noise follows ``scripts/rfiflagtest.py:35-44`` (unit-variance complex
    normal, generated one channel row at a time from ``RandomState(seed)``);
    the interference follows ``test/rfi/test_flagger.py:45-50`` (amplitude
    U[50, 70), uniform phase) applied to isolated samples with probability
    ``spike_prob`` and to whole channels ("narrowband lines") with probability
    ``line_prob``.  Generation is chunked over channels to bound memory.
    """
    rs = np.random.RandomState(seed=seed)
    vis = np.empty((channels, baselines), np.complex64)
    spikes = np.empty((channels, baselines), np.uint8)
    for c0 in range(0, channels, chunk):
        n = min(chunk, channels - c0)
        shape = (n, baselines)
        block = rs.standard_normal(shape) + 1j * rs.standard_normal(shape)
        hit = rs.random_sample(shape) < spike_prob
        hit |= (rs.random_sample((n, 1)) < line_prob)
        amp = rs.random_sample(shape) * 20.0 + 50.0
        phase = rs.random_sample(shape) * (2j * np.pi)
        block += hit * (amp * np.exp(phase))
        vis[c0 : c0 + n] = block.astype(np.complex64)
        spikes[c0 : c0 + n] = hit
    return vis, spikes
