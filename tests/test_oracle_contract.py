"""Pin oracle/contract.c (float32 device contract) to oracle/host_numpy.py.

Rules (SURVEY.md section 8(c')): R1 amplitudes bit-exact with np.abs on this
host; R2/R3 deviations == float32(host float64 deviations); R4 pre-scale median
bit-exact, noise within 1 ulp; R5-R8 flags identical except where a window sum
lies within 1e-6 (relative) of its decision value -- those are counted.
"""

import numpy as np
import pytest

from oracle import contract
from oracle import host_numpy as hn


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_abs_mode_detected(abs_mode):
    assert abs_mode in (contract.ABS_NUMPY, contract.ABS_HYPOT)


def test_amplitude_rule(abs_mode):
    rs = np.random.RandomState(5)
    v = (rs.standard_normal(200000) + 1j * rs.standard_normal(200000)).astype(np.complex64)
    v[:1000] *= 1e-20
    v[1000:2000] *= 1e18
    v[2000:3000] = (rs.standard_normal(1000) * 1e-42).astype(np.float32)
    v[3000:3006] = [0, 1, 2j, np.inf, complex(np.nan, 1), complex(np.inf, np.nan)]
    np.testing.assert_array_equal(np.abs(v), contract.amplitude(v, abs_mode))


def test_amplitude_rule_matches_generating_host(golden):
    """The AVX-512 rule reproduces the amplitudes stored by the fixture generator."""
    for vis, amp in (("bg_in_vis", "bg_in_amp"), ("flg_in_vis", "flg_in_amp"),
                     ("cfg1_in_vis", "cfg1_in_amp"), ("pct_in_cplx", "pct_in_cplx_abs")):
        np.testing.assert_array_equal(golden[amp], contract.amplitude(golden[vis], contract.ABS_NUMPY))


@pytest.mark.parametrize("width", [3, 5, 13])
@pytest.mark.parametrize("flag_kind", ["none", "channel", "full"])
@pytest.mark.parametrize("amplitudes", [False, True])
def test_background(golden, abs_mode, width, flag_kind, amplitudes):
    vis, flags = golden["bg_in_vis"], golden["bg_in_flags"]
    if amplitudes:
        vis = np.abs(vis)
    fl = {"none": None, "channel": flags[:, 0], "full": flags}[flag_kind]
    host = hn.background_median_filter(vis, width, fl, amplitudes)
    dev = contract.background(vis, width, fl, amplitudes, abs_mode)
    np.testing.assert_array_equal(host.astype(np.float32), dev)


def test_background_golden_exact(golden):
    """Against the reference-generated fixtures directly (AVX-512 amplitude rule)."""
    vis, flags = golden["bg_in_vis"], golden["bg_in_flags"]
    for width in (5, 13):
        np.testing.assert_array_equal(
            golden[f"bg_w{width}_none"].astype(np.float32),
            contract.background(vis, width, None, False, contract.ABS_NUMPY))
        np.testing.assert_array_equal(
            golden[f"bg_w{width}_full"].astype(np.float32),
            contract.background(vis, width, flags, False, contract.ABS_NUMPY))
    np.testing.assert_array_equal(
        golden["cfg1_dev"].astype(np.float32),
        contract.background(golden["cfg1_in_vis"], 13, None, False, contract.ABS_NUMPY))


def test_background_edge_cases(abs_mode):
    # fewer channels than the window, single channel, NaN amplitude, all flagged
    rs = np.random.RandomState(2)
    for channels in (1, 2, 5, 12, 13, 14):
        vis = (rs.standard_normal((channels, 7)) + 1j * rs.standard_normal((channels, 7))).astype(np.complex64)
        host = hn.background_median_filter(vis, 13)
        np.testing.assert_array_equal(host.astype(np.float32), contract.background(vis, 13, None, False, abs_mode))
    amp = np.abs(rs.standard_normal((40, 5))).astype(np.float32)
    amp[7, 2] = np.nan
    amp[20:30, 1] = np.nan
    host = hn.background_median_filter(amp, 5, amplitudes=True)
    np.testing.assert_array_equal(host.astype(np.float32), contract.background(amp, 5, None, True))
    flags = np.ones((40, 5), np.uint8)
    np.testing.assert_array_equal(np.zeros((40, 5), np.float32), contract.background(amp, 5, flags, True))
    with pytest.raises(ValueError):
        contract.background(amp, 4, None, True)


def test_noise(golden):
    for key in ("noise_kat_in_dev", "noise_in_dev"):
        dev = golden[key]
        noise, med = contract.noise_mad(dev)
        host_med = hn.median_abs_nonzero(dev)
        np.testing.assert_array_equal(host_med.astype(np.float32), med)
        assert ulp_diff(hn.noise_est_mad(dev).astype(np.float32), noise).max() <= 1
        noise_t, med_t = contract.noise_mad(np.ascontiguousarray(dev.T), transposed=True)
        np.testing.assert_array_equal(noise, noise_t)
        np.testing.assert_array_equal(med, med_t)


def test_noise_all_zero_is_nan():
    dev = np.zeros((9, 3), np.float32)
    dev[:, 1] = [0, 1, 0, -2, 0, 0, 3, 0, 0]
    noise, med = contract.noise_mad(dev)
    assert np.isnan(noise[0]) and np.isnan(noise[2])
    assert med[1] == 2.0 and noise[1] == np.float32(2.0 * 1.4826)


def test_threshold_simple(golden):
    dev = golden["thr_in_dev"]
    ramp = np.linspace(0.0, 50.0, 273).astype(np.float32)
    np.testing.assert_array_equal(golden["thr_simple_ramp"], contract.threshold_simple(dev, ramp, 11.0))
    out_t = contract.threshold_simple(np.ascontiguousarray(dev.T), ramp, 11.0, transposed=True)
    np.testing.assert_array_equal(golden["thr_simple_ramp"], out_t.T)


def check_sum_threshold(dev, noise, n_sigma, n_windows, rho, flag_value=1):
    near = {}
    host = hn.threshold_sum(dev, noise, n_sigma, n_windows, rho, flag_value, near)
    out = contract.threshold_sum(dev, noise, n_sigma, n_windows, rho, flag_value)
    mismatches = int((host != out).sum())
    if near.get("band", 0) == 0:
        assert mismatches == 0
    out_t = contract.threshold_sum(np.ascontiguousarray(dev.T), noise, n_sigma, n_windows, rho,
                                   flag_value, transposed=True)
    np.testing.assert_array_equal(out, out_t.T)
    return mismatches, near.get("band", 0)


def test_threshold_sum_reference_cases(golden):
    dev = golden["thr_in_dev"]
    const = np.repeat(10.0, 273).astype(np.float32)
    ramp = np.linspace(0.0, 50.0, 273).astype(np.float32)
    assert check_sum_threshold(dev, const, 11.0, 4, 1.2) == (0, 0)
    # the ramp starts at noise == 0: threshold 0, decision value 0, band is degenerate there
    mism, _ = check_sum_threshold(dev, ramp, 11.0, 4, 1.2)
    assert mism == 0
    mism, _ = check_sum_threshold(dev, ramp, 11.0, 7, 1.5, 5)
    assert mism == 0


@pytest.mark.parametrize("rho", [1.2, 1.5, 2.5])
def test_threshold_sum_broad(golden, rho):
    """Weak broad interference: larger windows fire, incl. at the band edges (R7)."""
    mism, band = check_sum_threshold(golden["thr2_in_dev"], np.full(24, 1.0, np.float32), 3.0, 6, rho)
    assert mism == 0, (mism, band)


def test_threshold_sum_random_sweep():
    rs = np.random.RandomState(11)
    total_mismatch = total_band = 0
    for case in range(30):
        channels = int(rs.choice([1, 2, 3, 17, 64, 65, 300]))
        dev = rs.standard_normal((channels, 6)).astype(np.float32)
        if channels > 20:
            for bl in range(6):
                s = rs.randint(0, channels - 10)
                dev[s:s + rs.randint(1, 30), bl] += rs.uniform(1.0, 8.0)
        noise = rs.uniform(0.5, 2.0, 6).astype(np.float32)
        n_windows = int(rs.choice([1, 2, 4, 5, 7]))
        while 2 ** (n_windows - 1) > channels:
            n_windows -= 1
        mism, band = check_sum_threshold(dev, noise, float(rs.choice([2.5, 3.0, 4.5])), n_windows,
                                         float(rs.choice([1.2, 1.5, 2.5])))
        total_mismatch += mism
        total_band += band
    assert total_mismatch == 0, (total_mismatch, total_band)


def test_threshold_sum_nan_noise_and_ties():
    dev = np.ones((64, 2), np.float32)
    noise = np.array([np.nan, 1.0 / 11.0], np.float32)  # second baseline: thr_0 == 1 == every sample
    out = contract.threshold_sum(dev, noise, 11.0, 7, 1.2)
    assert out[:, 0].sum() == 0  # NaN threshold never fires
    host = hn.threshold_sum(dev, noise, 11.0, 7, 1.2)
    np.testing.assert_array_equal(host, out)


@pytest.mark.parametrize("case", ["flg", "cfg1"])
@pytest.mark.parametrize("n_windows", [0, 4, 7])
def test_flagger(golden, abs_mode, case, n_windows):
    vis = golden[f"{case}_in_vis"]
    in_flags = golden["flg_in_flags"] if case == "flg" else None
    for fl in ([None, in_flags[:, 0], in_flags] if in_flags is not None else [None]):
        stages, near = {}, {}
        host = hn.flagger(vis, fl, n_windows=max(n_windows, 1), simple_threshold=(n_windows == 0),
                          stages=stages, near=near)
        flags, dev, noise = contract.flagger(vis, fl, n_windows=n_windows, abs_mode=abs_mode)
        np.testing.assert_array_equal(stages["deviations"].astype(np.float32), dev)
        assert ulp_diff(stages["noise"].astype(np.float32), noise).max() <= 1
        mismatches = int((host != flags).sum())
        assert mismatches == 0 or mismatches <= near.get("band", 0), (mismatches, near)


def test_percentile5(golden, abs_mode):
    amp, cplx = golden["pct_in_amp"], golden["pct_in_cplx"]
    np.testing.assert_array_equal(hn.percentile5(amp), contract.percentile5(amp))
    np.testing.assert_array_equal(hn.percentile5(amp, (10, 290)), contract.percentile5(amp, (10, 290)))
    np.testing.assert_array_equal(hn.percentile5(cplx), contract.percentile5(cplx, None, abs_mode))
    one = amp[:, :1]
    np.testing.assert_array_equal(hn.percentile5(one), contract.percentile5(one))


def test_masked_sum(golden, abs_mode):
    """R10: float64-accumulated, rounded once.  numpy's float32 pairwise sum (the expression
    the reference test uses) agrees to 1e-6 of the summed magnitudes, not of the result."""
    data, mask = golden["msum_in_data"], golden["msum_in_mask"]
    exact = np.sum(data.astype(np.complex128) * mask.reshape(-1, 1), axis=0)
    out = contract.masked_sum(data, mask)
    np.testing.assert_array_equal(exact.astype(np.complex64), out)
    scale = np.sum(np.abs(data) * mask.reshape(-1, 1), axis=0)
    assert np.all(np.abs(hn.masked_sum(data, mask) - out) <= 1e-6 * scale)
    exact_amp = np.sum(np.abs(data).astype(np.float64) * mask.reshape(-1, 1), axis=0)
    out_amp = contract.masked_sum(data, mask, True, abs_mode)
    np.testing.assert_array_equal(exact_amp.astype(np.float32), out_amp)
    np.testing.assert_allclose(hn.masked_sum(data, mask, True), out_amp, rtol=1e-6)
