"""Golden vectors for the 2-D flagger from the UNMODIFIED reference (numba).

Run in the build container, where /root/reference exists::

    python tests/golden/make_golden_twodflag.py

Writes tests/golden/reference_twodflag.npz: inputs and the reference's outputs for every stage of
rfi/twodflag.py (``_average_freq`` ... ``_unaverage_freq``) and for ``SumThresholdFlagger.get_flags``
in several configurations.  tests/test_oracle_twodflag.py pins oracle/twodflag_numpy.py to them;
tests/test_gpu_twodflag.py compares the CUDA flagger with the same fixtures.
"""

import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
from katsdpsigproc.rfi import twodflag as tf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    "default": dict(),
    "avg2": dict(average_freq=2),
    "iter3": dict(background_iterations=3, freq_chunks=3),
    "onechunk": dict(freq_chunks=1, time_extend=5, freq_extend=1),
    "wide": dict(windows_time=[1, 2, 4, 8, 16], windows_freq=[1, 2, 4, 8, 16, 32], outlier_nsigma=3.5,
                 spike_width_time=4.0, spike_width_freq=6.0, rho=1.5),
}


def make_input(rs, shape, complex_=False):
    """Smooth background + noise + injected interference of several shapes (after
    test/rfi/test_twodflag.py:524-560)."""
    n_time, n_freq, n_bl = shape
    x = np.linspace(0.0, 1.0, n_freq)[None, :, None]
    y = np.linspace(0.0, 1.0, n_time)[:, None, None]
    data = (3.0 + np.sin(2 * x) + 0.3 * np.cos(3 * y) + rs.standard_normal(shape) * 0.1).astype(np.float32)
    data[n_time // 3:n_time // 3 + 3, n_freq // 4:n_freq // 2, :] += 0.6        # block
    data[:, n_freq // 2 + 5, :] += 0.4                                           # channel
    data[n_time // 2, :, n_bl - 1] += 0.5                                        # dump
    spikes = rs.random_sample(shape) < 0.004
    data[spikes] += 2.0
    data[1, 2, 0] = np.nan
    flags = rs.random_sample(shape) < 0.02
    if complex_:
        phase = rs.random_sample(shape) * 2 * np.pi
        data = (data * np.exp(1j * phase)).astype(np.complex64)
    return data, flags


def main():
    rs = np.random.RandomState(20)
    out = {}
    # stage fixtures
    d2 = (rs.standard_normal((40, 64)) * 2).astype(np.float32)
    d2[5:8, 10:30] += 6
    f2 = rs.random_sample(d2.shape) < 0.1
    ends = np.linspace(0, 64, 4).astype(np.int_)
    out["st_data"], out["st_flags"], out["st_ends"] = d2, f2, ends
    bg = tf._get_background2d(d2, f2, 2, np.array((3.0, 5.0)), 2.0, ends)
    out["st_background"] = bg
    w = np.array([1, 2, 4, 8])
    out["st_sum_time"] = tf._sum_threshold(d2, f2, 0, w, 3.0, 1.3)
    out["st_sum_freq"] = tf._sum_threshold(d2, f2, 1, w, 3.0, 1.3, ends)
    m = np.empty_like(d2)
    tf.masked_gaussian_filter(d2, f2, np.array((3.0, 5.0)), m)
    out["st_masked"] = m
    tm = tf._time_median(d2, f2)
    out["st_time_median"], out["st_time_median_flags"] = tm
    c3 = (rs.standard_normal((6, 17, 3)) + 1j * rs.standard_normal((6, 17, 3))).astype(np.complex64)
    f3 = rs.random_sample(c3.shape) < 0.3
    out["avg_in"], out["avg_in_flags"] = c3, f3
    for factor in (1, 2, 3):
        a, af = tf._average_freq(c3, f3, tf._as_min_dtype(factor))
        out[f"avg{factor}_data"], out[f"avg{factor}_flags"] = a, af
    # whole flagger
    for name, kw in CONFIGS.items():
        shape = (48, 130, 3) if name != "wide" else (64, 96, 2)
        data, flags = make_input(rs, shape)
        out[f"full_{name}_data"], out[f"full_{name}_flags"] = data, flags
        out[f"full_{name}_out"] = tf.SumThresholdFlagger(**kw).get_flags(data, flags)
    data, flags = make_input(rs, (32, 80, 2), complex_=True)
    out["full_complex_data"], out["full_complex_flags"] = data, flags
    out["full_complex_out"] = tf.SumThresholdFlagger().get_flags(data, flags)
    data, flags = make_input(rs, (20, 70, 2))
    out["full_allflagged_data"] = data
    out["full_allflagged_out"] = tf.SumThresholdFlagger().get_flags(data, np.ones(data.shape, np.bool_))
    np.savez_compressed(os.path.join(HERE, "reference_twodflag.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith("full") and k.endswith("out")})
    print({k: int(v.sum()) for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
