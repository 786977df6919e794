"""oracle/twodflag_numpy.py (the numpy restatement of the reference's numba 2-D flagger) pinned bit
for bit: against tests/golden/reference_twodflag.npz (outputs of the UNMODIFIED reference,
tests/golden/make_golden_twodflag.py) everywhere, and against the reference itself, live, where
oracle/_ref and numba are present.  Also the reference's own known-answer tests for the pieces
(test/rfi/test_twodflag.py:51-235)."""

import os

import numpy as np
import pytest

import oracle
from oracle import twodflag_numpy as tn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "reference_twodflag.npz")) as data:
        return {k: data[k] for k in data.files}


def same_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


CONFIGS = {
    "default": dict(),
    "avg2": dict(average_freq=2),
    "iter3": dict(background_iterations=3, freq_chunks=3),
    "onechunk": dict(freq_chunks=1, time_extend=5, freq_extend=1),
    "wide": dict(windows_time=[1, 2, 4, 8, 16], windows_freq=[1, 2, 4, 8, 16, 32], outlier_nsigma=3.5,
                 spike_width_time=4.0, spike_width_freq=6.0, rho=1.5),
    "complex": dict(),
}


def test_stages_against_golden(gold):
    d, f, ends = gold["st_data"], gold["st_flags"], gold["st_ends"]
    assert same_f32(gold["st_background"], tn.get_background2d(d, f, 2, np.array((3.0, 5.0)), 2.0, ends))
    w = np.array([1, 2, 4, 8])
    np.testing.assert_array_equal(gold["st_sum_time"], tn.sum_threshold(d, f, 0, w, 3.0, 1.3))
    np.testing.assert_array_equal(gold["st_sum_freq"], tn.sum_threshold(d, f, 1, w, 3.0, 1.3, ends))
    m = np.empty_like(d)
    tn.masked_gaussian_filter(d, f, (3.0, 5.0), m)
    assert same_f32(gold["st_masked"], m)
    med, med_flags = tn.time_median(d, f)
    assert same_f32(gold["st_time_median"], med)
    np.testing.assert_array_equal(gold["st_time_median_flags"], med_flags)
    for factor in (1, 2, 3):
        a, af = tn.average_freq(gold["avg_in"], gold["avg_in_flags"], factor)
        assert same_f32(gold[f"avg{factor}_data"], a)
        np.testing.assert_array_equal(gold[f"avg{factor}_flags"], af)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_flagger_against_golden(gold, name):
    flags = gold[f"full_{name}_flags"]
    out = tn.SumThresholdFlagger(**CONFIGS[name]).get_flags(gold[f"full_{name}_data"].copy(), flags.copy())
    np.testing.assert_array_equal(gold[f"full_{name}_out"], out)


def test_all_flagged_against_golden(gold):
    data = gold["full_allflagged_data"]
    out = tn.SumThresholdFlagger().get_flags(data.copy(), np.ones(data.shape, np.bool_))
    np.testing.assert_array_equal(gold["full_allflagged_out"], out)


def test_against_the_live_reference():
    ref = oracle.reference_twodflag()
    if ref is None:
        pytest.skip("oracle/_ref (or numba) is not present")
    rs = np.random.RandomState(7)
    shape = (30, 77, 2)
    data = (5 + rs.standard_normal(shape) * 0.1).astype(np.float32)
    data[10:13, 20:40] += 1.0
    flags = rs.random_sample(shape) < 0.03
    for kw in (dict(), dict(average_freq=3, freq_chunks=2), dict(background_iterations=2, time_extend=1)):
        want = ref.SumThresholdFlagger(**kw).get_flags(data, flags)
        np.testing.assert_array_equal(want, tn.SumThresholdFlagger(**kw).get_flags(data.copy(), flags.copy()))


# ---- the reference's known answers for the pieces (test/rfi/test_twodflag.py)
def test_average_freq_known_answers():
    data = np.arange(30, dtype=np.float32).reshape(5, 6, 1).repeat(2, axis=2)
    flags = np.zeros(data.shape, np.bool_)
    flags[0, 2, :] = True
    flags[2, 0:2, :] = True
    avg, avg_flags = tn.average_freq(data, flags, 2)
    want = np.array([[0.5, 3.0, 4.5], [6.5, 8.5, 10.5], [0.0, 14.5, 16.5], [18.5, 20.5, 22.5], [24.5, 26.5, 28.5]],
                    np.float32)
    np.testing.assert_array_equal(avg[0], want)
    assert avg_flags[0, 2, 0] and avg_flags.sum() == 2


def test_time_median_known_answers():
    data = np.array([[2.0, 1.0, 2.0, 5.0], [3.0, 1.0, 8.0, 6.0], [4.0, 1.0, 4.0, 7.0], [5.0, 1.0, 5.0, 6.5],
                     [1.5, 1.0, 1.5, 5.5]], np.float32)
    flags = np.array([[0, 1, 0, 0], [0, 1, 0, 1], [0, 1, 0, 0], [0, 1, 0, 1], [0, 1, 0, 0]], np.bool_)
    med, med_flags = tn.time_median(data, flags)
    np.testing.assert_array_equal(np.array([[3.0, 0.0, 4.0, 5.5]], np.float32), med)
    np.testing.assert_array_equal(np.array([[False, True, False, False]]), med_flags)


def test_interpolate_known_answers():
    data = np.array([np.nan, np.nan, 4.0, np.nan, np.nan, 10.0, np.nan, -2.0, np.nan, np.nan], np.float32)[None]
    tn.linearly_interpolate_nans(data)
    np.testing.assert_allclose([[4.0, 4.0, 4.0, 6.0, 8.0, 10.0, 4.0, -2.0, -2.0, -2.0]], data)
    empty = np.full((1, 5), np.nan, np.float32)
    tn.linearly_interpolate_nans(empty)
    np.testing.assert_array_equal(np.zeros((1, 5), np.float32), empty)


def test_box_filter_one_pass_is_a_box():
    data = np.array([50.0, 10.0, 60.0, -70.0, 30.0, 20.0, -15.0], np.float32)
    out = np.empty_like(data)
    # (one pass, radius 1: the centred sum of 3 over 3, zeros outside)
    import oracle.twodflag_numpy as m
    padded = np.concatenate([[0, 0], data, [0, 0]]).astype(np.float32)
    want = np.array([padded[i + 1:i + 4].sum() / 3 for i in range(7)], np.float32)
    saved = m.f32_pow_int
    m.box_gaussian_filter1d(data, 1, out, 1)
    assert m.f32_pow_int is saved
    np.testing.assert_allclose(want, out, rtol=1e-6)
