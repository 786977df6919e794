"""Parity of every CUDA kernel, called through the C ABI, against the oracle.

The checker is ``oracle.contract`` (plain C, float32 device contract) which is
itself pinned to the reference host classes by ``test_oracle_contract.py`` and
``test_oracle_golden.py``; where the golden fixtures hold reference outputs they
are compared directly as well.  Shapes follow the reference's own device tests
(test/rfi/test_background.py:78-104, test_noise_est.py:54-61,
test_threshold.py:60-93, test_flagger.py:74-132, test_percentile.py:37-90,
test_transpose.py:35-59, test_maskedsum.py:35-67).
"""

import numpy as np
import pytest

import cabi_util as cu
from oracle import contract

pytestmark = pytest.mark.gpu


def complex_normal(rs, shape):
    return (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)


def assert_same_f32(a, b):
    """Bit-for-bit equal, NaNs included."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape
    bad = a.view(np.uint32) != b.view(np.uint32)
    both_nan = np.isnan(a) & np.isnan(b)
    bad &= ~both_nan
    assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:5], a[bad][:5], b[bad][:5])


# ------------------------------------------------------------------ transpose
@pytest.mark.parametrize("shape", [(4, 5), (53, 7), (53, 81), (32, 64), (200, 333), (1, 1)])
@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.complex64, np.uint16, np.complex128])
@pytest.mark.parametrize("pad", [0, 3])
def test_transpose(shape, dtype, pad):
    rs = np.random.RandomState(1)
    a = (rs.uniform(0, 250, shape)).astype(dtype)
    out = cu.transpose(a, pad, pad)
    np.testing.assert_array_equal(a.T, out)


def test_transpose_bytes_vector_path():
    rs = np.random.RandomState(2)
    a = rs.randint(0, 256, (192, 320)).astype(np.uint8)
    np.testing.assert_array_equal(a.T, cu.transpose(a))
    np.testing.assert_array_equal(a.T, cu.transpose(a, 4, 8))


@pytest.mark.parametrize("shape,pads", [((256, 384), (0, 0)), ((300, 500), (0, 0)), ((128, 128), (16, 32)),
                                        ((391, 130), (9, 2)), ((1024, 640), (0, 16)), ((640, 1), (0, 0))])
def test_transpose_bytes_128bit_tiles(shape, pads):
    """uint8 arrays with 16-byte aligned rows take the 128 x 128-tile kernel for the whole tiles and
    the 64 x 64 one for the ragged strips; misaligned ones take the latter throughout."""
    rs = np.random.RandomState(3)
    a = rs.randint(0, 256, shape).astype(np.uint8)
    np.testing.assert_array_equal(a.T, cu.transpose(a, *pads))


@pytest.mark.parametrize("shape,pads", [((64, 64), (0, 0)), ((256, 384), (0, 0)), ((300, 500), (0, 0)),
                                        ((128, 192), (4, 8)), ((391, 132), (0, 1)), ((200, 333), (3, 0)),
                                        ((1024, 640), (0, 16)), ((640, 1), (0, 0)), ((65, 4160), (0, 0))])
def test_transpose_words_128bit_tiles(shape, pads):
    """float32 arrays with 16-byte aligned rows take the 64 x 64-tile kernel with 128-bit loads and
    stores (ragged edge tiles element by element); misaligned ones take the 32 x 32 one."""
    rs = np.random.RandomState(4)
    a = rs.standard_normal(shape).astype(np.float32)
    np.testing.assert_array_equal(a.T, cu.transpose(a, *pads))


def test_transpose_rejects_bad_element_size():
    from ctypes import c_void_p

    from katsdpsigproc_b200 import _capi
    lib = _capi.load()
    for size in (0, 3, 32):
        assert lib.ksp_transpose(None, c_void_p(256), c_void_p(512), 4, 4, 4, 4, size) == -1


# ------------------------------------------------------------------ background
def bg_inputs(channels, baselines, seed=1):
    rs = np.random.RandomState(seed)
    vis = complex_normal(rs, (channels, baselines))
    flags = (rs.random_sample((channels, baselines)) < 0.1).astype(np.uint8)
    flags[:, min(3, baselines - 1)] = 1            # a fully flagged baseline
    flags[40:60, :] = (flags[40:60, :] * 3) | 4    # fully flagged windows, values other than 0/1
    return vis, flags


@pytest.mark.parametrize("width", [5, 13, 1, 21])
@pytest.mark.parametrize("flag_kind", ["none", "channel", "full"])
@pytest.mark.parametrize("amplitudes", [False, True])
@pytest.mark.parametrize("transposed", [False, True])
def test_background(abs_mode, width, flag_kind, amplitudes, transposed):
    vis, flags = bg_inputs(417, 313)
    if amplitudes:
        vis = np.abs(vis)
    fl = {"none": None, "channel": np.ascontiguousarray(flags[:, 7]), "full": flags}[flag_kind]
    expect = contract.background(vis, width, fl, amplitudes, abs_mode)
    out = cu.background(vis, width, fl, amplitudes, abs_mode, transposed, pad=5)
    assert_same_f32(expect, out.T if transposed else out)


@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (12, 40), (13, 33), (31, 64), (1500, 70)])
@pytest.mark.parametrize("transposed", [False, True])
def test_background_small_and_long(abs_mode, shape, transposed):
    vis, flags = bg_inputs(*shape, seed=3)
    for fl in (None, flags):
        expect = contract.background(vis, 13, fl, False, abs_mode)
        out = cu.background(vis, 13, fl, False, abs_mode, transposed)
        assert_same_f32(expect, out.T if transposed else out)


@pytest.mark.parametrize("width", [3, 7, 9, 11, 15, 21, 31, 33, 63])
@pytest.mark.parametrize("transposed", [False, True])
def test_background_other_widths(abs_mode, width, transposed):
    """Odd widths up to 31 take the sliding-sorted-window kernel, wider ones the generic one;
    clean tiles, tiles with flags / NaN, band edges and a ragged last tile in both directions."""
    vis, flags = bg_inputs(700, 45, seed=width)
    vis[350:420, 11:] *= 3.0
    vis[500, 2] = np.nan
    for fl in (None, np.ascontiguousarray(flags[:, 3]), flags):
        expect = contract.background(vis, width, fl, False, abs_mode)
        out = cu.background(vis, width, fl, False, abs_mode, transposed, pad=3)
        assert_same_f32(expect, out.T if transposed else out)
    amp = np.abs(vis)
    amp[np.isnan(amp)] = 1.0
    expect = contract.background(amp, width, None, True, abs_mode)
    out = cu.background(amp, width, None, True, abs_mode, transposed)
    assert_same_f32(expect, out.T if transposed else out)


def test_background_special_values(abs_mode):
    rs = np.random.RandomState(4)
    vis = complex_normal(rs, (300, 40))
    vis[rs.random_sample(vis.shape) < 0.02] = np.nan
    vis[rs.random_sample(vis.shape) < 0.02] = 0
    vis[17, 3] = np.inf
    vis[100:140, 5] = 0
    vis[200, 7] = 1e-30 + 1e-41j
    for transposed in (False, True):
        expect = contract.background(vis, 13, None, False, abs_mode)
        out = cu.background(vis, 13, None, False, abs_mode, transposed)
        assert_same_f32(expect, out.T if transposed else out)


def test_background_golden(golden, abs_mode):
    """Reference-generated deviations (float64) rounded to float32."""
    if abs_mode != contract.ABS_NUMPY:
        pytest.skip("fixtures were generated on an AVX-512F host")
    vis, flags = golden["bg_in_vis"], golden["bg_in_flags"]
    for width in (5, 13):
        for kind, fl in (("none", None), ("channel", np.ascontiguousarray(flags[:, 0])),
                         ("full", flags)):
            out = cu.background(vis, width, fl, False, abs_mode)
            assert_same_f32(golden[f"bg_w{width}_{kind}"].astype(np.float32), out)
        out = cu.background(golden["bg_in_amp"], width, flags, True, abs_mode)
        assert_same_f32(golden[f"bg_w{width}_amp_full"].astype(np.float32), out)


# ------------------------------------------------------------------ noise
@pytest.mark.parametrize("transposed", [False, True])
def test_madnz(golden, transposed):
    dev = golden["noise_in_dev"]
    expect, _ = contract.noise_mad(dev)
    out = cu.madnz(np.ascontiguousarray(dev.T) if transposed else dev, transposed, pad=3)
    assert_same_f32(expect, out)
    ref = golden["noise_big"].astype(np.float32)
    assert np.max(np.abs(out.view(np.int32) - ref.view(np.int32))) <= 1   # R4: <= 1 ulp of host


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("channels", [1, 2, 6, 1000, 10000, 10001, 40000, 65535, 65536, 70001])
def test_madnz_sizes(transposed, channels):
    rs = np.random.RandomState(channels)
    baselines = 5
    dev = rs.standard_normal((channels, baselines)).astype(np.float32)
    dev[rs.random_sample(dev.shape) < 0.08] = 0
    dev[:, 2] = 0                                     # NaN result
    if channels > 5:
        dev[:, 3] = np.float32(0.5) * rs.randint(1, 4, channels)   # heavy ties
        dev[1::2, 4] = 0
    expect, _ = contract.noise_mad(dev)
    out = cu.madnz(np.ascontiguousarray(dev.T) if transposed else dev, transposed)
    assert_same_f32(expect, out)


# ------------------------------------------------------------------ thresholds
def test_threshold_simple(golden):
    dev = golden["thr_in_dev"]
    ramp = np.linspace(0.0, 50.0, 273).astype(np.float32)
    out = cu.threshold_simple(dev, ramp, 11.0, pad=2)
    np.testing.assert_array_equal(golden["thr_simple_ramp"], out)
    out_t = cu.threshold_simple(np.ascontiguousarray(dev.T), ramp, 11.0, 7, True, pad=2)
    np.testing.assert_array_equal(golden["thr_simple_ramp"] * 7, out_t.T)


def check_sum(dev, noise, n_sigma, n_windows, rho, flag_value=1, pad=0):
    expect = contract.threshold_sum(dev, noise, n_sigma, n_windows, rho, flag_value)
    out = cu.threshold_sum(np.ascontiguousarray(dev.T), noise, n_sigma, n_windows, rho, flag_value,
                           pad)
    np.testing.assert_array_equal(expect, out.T)
    return expect


def test_threshold_sum_reference_cases(golden):
    dev = golden["thr_in_dev"]
    const = np.repeat(10.0, 273).astype(np.float32)
    ramp = np.linspace(0.0, 50.0, 273).astype(np.float32)
    np.testing.assert_array_equal(golden["thr_sum_const"], check_sum(dev, const, 11.0, 4, 1.2))
    np.testing.assert_array_equal(golden["thr_sum_ramp"], check_sum(dev, ramp, 11.0, 4, 1.2, pad=7))
    np.testing.assert_array_equal(golden["thr_sum_ramp_w7_fv5"],
                                  check_sum(dev, ramp, 11.0, 7, 1.5, 5))


@pytest.mark.parametrize("rho", [1.2, 1.5, 2.5])
def test_threshold_sum_broad(golden, rho):
    out = check_sum(golden["thr2_in_dev"], np.full(24, 1.0, np.float32), 3.0, 6, rho)
    np.testing.assert_array_equal(golden[f"thr2_sum_w6_rho{rho}"], out)


def test_threshold_sum_random_sweep():
    rs = np.random.RandomState(11)
    for case in range(40):
        channels = int(rs.choice([1, 2, 3, 17, 31, 32, 33, 64, 65, 300, 1025, 4096]))
        dev = rs.standard_normal((channels, 6)).astype(np.float32)
        if channels > 40:
            for bl in range(6):
                for _ in range(1 + channels // 400):
                    s = rs.randint(0, channels - 10)
                    dev[s:s + rs.randint(1, 90), bl] += rs.uniform(1.0, 8.0)
        noise = rs.uniform(0.5, 2.0, 6).astype(np.float32)
        n_windows = int(rs.choice([1, 2, 4, 5, 7]))
        check_sum(dev, noise, float(rs.choice([2.5, 3.0, 4.5])), n_windows,
                  float(rs.choice([1.2, 1.5, 2.5])), pad=int(rs.choice([0, 1, 4])))


@pytest.mark.parametrize("channels", [32768, 32769, 40000, 70000])
def test_threshold_sum_long_rows(channels):
    """Rows longer than one block's span are processed in overlapping chunks."""
    rs = np.random.RandomState(5)
    dev = rs.standard_normal((channels, 3)).astype(np.float32)
    for bl in range(3):
        for _ in range(60):
            s = rs.randint(0, channels - 100)
            dev[s:s + rs.randint(1, 100), bl] += rs.uniform(1.0, 6.0)
    # interference straddling the chunk seams
    for seam in (32768 - 128, 32768 - 64, 2 * (32768 - 256)):
        if seam + 80 < channels:
            dev[seam - 40:seam + 40, 1] += 2.5
    noise = np.array([1.0, 0.9, 1.1], np.float32)
    check_sum(dev, noise, 3.0, 7, 1.2)


@pytest.mark.parametrize("channels, baselines", [(12288, 700), (4096, 1300), (8192 + 32, 450),
                                                 (1024, 9000), (96, 20000), (2048 + 7, 3000)])
def test_threshold_sum_many_tiles_per_block(channels, baselines):
    """More (row, span) tiles than resident blocks: every block of the persistent grid walks
    several tiles, alternating between its two span buffers."""
    rs = np.random.RandomState(channels + baselines)
    dev = rs.standard_normal((channels, baselines)).astype(np.float32)
    for _ in range(baselines * 3):                       # interference of all widths, everywhere
        wmax = min(100, channels // 2)
        bl, s = rs.randint(0, baselines), rs.randint(0, channels - wmax)
        dev[s:s + rs.randint(1, wmax), bl] += rs.uniform(1.0, 6.0)
    dev[rs.random_sample(dev.shape) < 1 / 64] += 40.0    # spikes
    noise = rs.uniform(0.8, 1.3, baselines).astype(np.float32)
    check_sum(dev, noise, 3.5, 7, 1.2)


@pytest.mark.parametrize("n_windows", [8, 9, 10, 11])
@pytest.mark.parametrize("channels", [700, 8192, 20000])
def test_threshold_sum_many_windows(n_windows, channels):
    """Window sizes beyond 64 (n_windows 8..11) take the general kernel: broad interference of
    every width, chunk seams, rows shorter than the largest window."""
    rs = np.random.RandomState(n_windows * 1000 + channels)
    baselines = 6
    dev = rs.standard_normal((channels, baselines)).astype(np.float32)
    for bl in range(baselines):
        for _ in range(4 + channels // 1500):
            width = int(rs.choice([1, 3, 20, 90, 200, 400, 900]))
            width = min(width, channels // 2)
            s = rs.randint(0, channels - width)
            dev[s:s + width, bl] += rs.uniform(0.4, 5.0)
    dev[rs.random_sample(dev.shape) < 1 / 128] += 30.0
    noise = rs.uniform(0.8, 1.2, baselines).astype(np.float32)
    noise[5] = np.nan
    check_sum(dev, noise, 3.0, n_windows, 1.2, flag_value=2)
    check_sum(dev, noise, 2.5, n_windows, 1.5, pad=4)


def test_flagger_fused_many_windows(abs_mode):
    rs = np.random.RandomState(12)
    channels, baselines = 9000, 40
    vis = complex_normal(rs, (channels, baselines))
    for _ in range(60):
        bl, width = rs.randint(0, baselines), int(rs.choice([5, 60, 300, 700]))
        s = rs.randint(0, channels - width)
        vis[s:s + width, bl] += rs.uniform(0.5, 3.0)
    flags, dev, noise = contract.flagger(vis, None, n_windows=10, n_sigma=4.0, abs_mode=abs_mode)
    out_flags, out_noise = cu.flagger(vis, None, n_windows=10, n_sigma=4.0, abs_mode=abs_mode)
    assert_same_f32(noise, out_noise)
    np.testing.assert_array_equal(flags, out_flags)
    assert flags.any()


def test_threshold_sum_deep_dips_next_to_marginal_wide_features():
    """The filters of the threshold kernel must stay sound when a run holds large NEGATIVE
    deviations (dropped samples on a bright band: dev = -median, far beyond the threshold in
    magnitude): the bound on the sum of the positive samples is computed from two rounded sums and
    its slack has to scale with sum |u|.  Wide features (16..64 channels) whose window sums sit
    within ~1e-5 of firing, placed right next to such dips, must be flagged exactly as the
    contract flags them."""
    rs = np.random.RandomState(31)
    channels, baselines = 4096, 48
    dev = (rs.standard_normal((channels, baselines)) * 0.05).astype(np.float32)
    noise = np.full(baselines, 1.0, np.float32)
    n_sigma, rho, n_windows = 4.0, 1.2, 7
    fired = 0
    for b in range(baselines):
        for k in range(12):
            w = int(rs.choice([16, 32, 64]))
            c0 = int(rs.randint(200, channels - 200))
            thr = np.float32(n_sigma * 1.0 * rho ** -int(np.log2(w)))
            # a plateau whose mean is the window threshold times (1 +- a few 1e-6 .. 1e-4)
            level = np.float32(thr * (1.0 + rs.choice([-1, 1]) * 10 ** rs.uniform(-5.5, -4.0)))
            dev[c0:c0 + w, b] = level
            # deep dips inside the same 32-channel runs, before and after the plateau
            dev[c0 - rs.randint(2, 12), b] = np.float32(-rs.uniform(200.0, 5000.0))
            dev[c0 + w + rs.randint(1, 10), b] = np.float32(-rs.uniform(200.0, 5000.0))
    expect = check_sum(dev, noise, n_sigma, n_windows, rho)
    fired = int(expect.sum())
    assert 0 < fired < expect.size // 2          # some of the marginal windows fire, some do not


def test_threshold_sum_nan_noise_and_ties():
    dev = np.ones((64, 3), np.float32)
    noise = np.array([np.nan, 1.0 / 11.0, -1.0], np.float32)
    check_sum(dev, noise, 11.0, 7, 1.2)


# ------------------------------------------------------------------ fused flagger
@pytest.mark.parametrize("case", ["flg", "cfg1"])
@pytest.mark.parametrize("n_windows", [1, 4, 7])
def test_flagger_fused(golden, abs_mode, case, n_windows):
    vis = golden[f"{case}_in_vis"]
    in_flags = golden["flg_in_flags"] if case == "flg" else None
    variants = [None] if in_flags is None else [None, np.ascontiguousarray(in_flags[:, 0]), in_flags]
    for fl in variants:
        flags, dev, noise = contract.flagger(vis, fl, n_windows=n_windows, abs_mode=abs_mode)
        out_flags, out_noise = cu.flagger(vis, fl, n_windows=n_windows, abs_mode=abs_mode, pad=3)
        assert_same_f32(noise, out_noise)
        np.testing.assert_array_equal(flags, out_flags)
        # several chunks, ragged last chunk
        out_flags, out_noise = cu.flagger(vis, fl, n_windows=n_windows, abs_mode=abs_mode,
                                          chunk_baselines=32)
        assert_same_f32(noise, out_noise)
        np.testing.assert_array_equal(flags, out_flags)


def test_flagger_fused_many_tiles(abs_mode):
    """Enough (baseline, span) tiles that the persistent threshold blocks loop (packed output)."""
    rs = np.random.RandomState(77)
    channels, baselines = 8192, 300
    vis = (rs.standard_normal((channels, baselines)) +
           1j * rs.standard_normal((channels, baselines))).astype(np.complex64)
    vis[rs.random_sample(vis.shape) < 1 / 64] += 60.0
    for _ in range(200):
        bl, s = rs.randint(0, baselines), rs.randint(0, channels - 70)
        vis[s:s + rs.randint(2, 64), bl] += rs.uniform(2.0, 6.0)
    flags, dev, noise = contract.flagger(vis, None, n_windows=7, n_sigma=4.0, abs_mode=abs_mode)
    out_flags, out_noise = cu.flagger(vis, None, n_windows=7, n_sigma=4.0, abs_mode=abs_mode)
    assert_same_f32(noise, out_noise)
    np.testing.assert_array_equal(flags, out_flags)


def test_flagger_fused_golden(golden, abs_mode):
    """Flags produced by the unmodified reference FlaggerHost."""
    if abs_mode != contract.ABS_NUMPY:
        pytest.skip("fixtures were generated on an AVX-512F host")
    out_flags, out_noise = cu.flagger(golden["cfg1_in_vis"], n_windows=7, abs_mode=abs_mode)
    np.testing.assert_array_equal(golden["cfg1_flags"], out_flags)
    ref = golden["cfg1_noise"].astype(np.float32)
    assert np.max(np.abs(out_noise.view(np.int32) - ref.view(np.int32))) <= 1
    for kind, fl in (("none", None), ("channel", np.ascontiguousarray(golden["flg_in_flags"][:, 0])),
                     ("full", golden["flg_in_flags"])):
        for nw in (4, 7):
            out_flags, _ = cu.flagger(golden["flg_in_vis"], fl, n_windows=nw, abs_mode=abs_mode)
            np.testing.assert_array_equal(golden[f"flg_sum{nw}_{kind}"], out_flags)


def test_flagger_fused_medium(abs_mode):
    """4096 x 200 with injected spikes and narrowband lines: fused == contract, stage by stage."""
    rs = np.random.RandomState(7)
    channels, baselines = 4096, 200
    vis = complex_normal(rs, (channels, baselines))
    spikes = rs.random_sample(vis.shape) < 1 / 64
    vis += (spikes * (rs.random_sample(vis.shape) * 20 + 50)
            * np.exp(2j * np.pi * rs.random_sample(vis.shape))).astype(np.complex64)
    vis[1000:1037, :] += 3.0
    flags, dev, noise = contract.flagger(vis, None, n_windows=7, abs_mode=abs_mode)
    out_flags, out_noise = cu.flagger(vis, None, n_windows=7, abs_mode=abs_mode, chunk_baselines=96)
    assert_same_f32(noise, out_noise)
    np.testing.assert_array_equal(flags, out_flags)


@pytest.mark.parametrize("flag_kind", ["none", "channel", "full"])
def test_flagger_fused_more_chunks_than_lanes(abs_mode, flag_kind):
    """With more chunks than lanes every chunk launches its own background filter (with up to 4
    chunks one launch covers them all); baselines not a multiple of 16 or 32, padded and
    unpadded flag rows (128-bit and byte-wise flag expansion)."""
    rs = np.random.RandomState(11)
    channels, baselines = 2048, 203
    vis = complex_normal(rs, (channels, baselines))
    vis[rs.random_sample(vis.shape) < 1 / 64] += 40.0
    vis[700:730, ::3] += 2.5
    fl = None
    if flag_kind == "channel":
        fl = (rs.random_sample(channels) < 0.05).astype(np.uint8)
    elif flag_kind == "full":
        fl = (rs.random_sample(vis.shape) < 0.05).astype(np.uint8)
    flags, dev, noise = contract.flagger(vis, fl, n_windows=7, abs_mode=abs_mode)
    for chunk, pad in ((32, 0), (32, 5), (64, 13), (0, 0)):
        out_flags, out_noise = cu.flagger(vis, fl, n_windows=7, abs_mode=abs_mode, chunk_baselines=chunk, pad=pad)
        assert_same_f32(noise, out_noise)
        np.testing.assert_array_equal(flags, out_flags)


# ------------------------------------------------------------------ helpers
@pytest.mark.parametrize("shape,column_range", [((4096, 1), None), ((4096, 4029), None),
                                                ((64, 300), (8, 280)), ((27, 301), (0, 301)),
                                                ((3, 50000), None), ((2, 70001), (5, 70000))])
@pytest.mark.parametrize("is_amplitude", [True, False])
def test_percentile5(abs_mode, shape, column_range, is_amplitude):
    rs = np.random.RandomState(1)
    rows, cols = shape
    rows = min(rows, 64)
    if is_amplitude:
        src = np.abs(rs.standard_normal((rows, cols))).astype(np.float32)
    else:
        src = complex_normal(rs, (rows, cols))
    expect = contract.percentile5(src, column_range, abs_mode)
    out = cu.percentile5(src, column_range, abs_mode, pad=3)
    assert_same_f32(expect, out)
    data = np.abs(src)
    if column_range:
        data = data[:, column_range[0]:column_range[1]]
    ref = np.percentile(data, [0, 100, 25, 75, 50], axis=1, method="lower").astype(np.float32)
    if is_amplitude or abs_mode == contract.detect_abs_mode():
        assert_same_f32(ref, out)


@pytest.mark.parametrize("cols, column_range, pad", [(8192, None, 0), (16384, (8, 16380), 4),
                                                     (32768, None, 0), (50000, (4, 49996), 0),
                                                     (65536, None, 0)])
@pytest.mark.parametrize("is_amplitude", [True, False])
def test_percentile5_fast_path(abs_mode, cols, column_range, pad, is_amplitude):
    """Rows readable with aligned 16-byte loads take the sampled-bracket kernel; the rows below
    include everything that must push single ranks or whole rows onto its fallbacks.  (Rows
    holding NaN are left out: their order statistics are unspecified, as in the reference.)"""
    rs = np.random.RandomState(cols)
    rows = 24
    if is_amplitude:
        src = np.abs(rs.standard_normal((rows, cols))).astype(np.float32)
    else:
        src = complex_normal(rs, (rows, cols))
    src[1, rs.random_sample(cols) < 0.07] = 0                       # exact zeros
    src[2] = np.round(src[2] * 4) / 4                               # heavy ties: overlapping brackets
    src[3] = 1.5                                                    # constant row
    src[4, 100] = 1e-38                                             # one denormal amplitude
    src[5, 7] = np.inf
    src[6] = src[6] * 1e-3 + 1.0                                    # narrow value range
    src[7, : cols // 2] = 0                                         # half zeros: the 25 % rank is 0
    src[8] = np.round(src[8] * 64) / 64                             # lighter ties: crowded bins
    src[9] *= 1e-30                                                 # tiny values (denormal amplitudes)
    src[10] *= 1e18                                                 # huge values
    expect = contract.percentile5(src, column_range, abs_mode)
    before = cu.selection_fallbacks(reset=True)
    out = cu.percentile5(src, column_range, abs_mode, pad=pad)
    fallbacks = cu.selection_fallbacks(reset=True)
    assert_same_f32(expect, out)
    # ordinary rows must not fall back: 13 of the 24 rows are plain noise
    assert fallbacks <= 3 * 11 + 2, (before, fallbacks)


def test_percentile5_golden(golden, abs_mode):
    assert_same_f32(golden["pct_amp_all"], cu.percentile5(golden["pct_in_amp"]))
    assert_same_f32(golden["pct_amp_range"], cu.percentile5(golden["pct_in_amp"], (10, 290)))
    if abs_mode == contract.ABS_NUMPY:
        assert_same_f32(golden["pct_cplx_all"], cu.percentile5(golden["pct_in_cplx"], None, abs_mode))


@pytest.mark.parametrize("cols", [2, 4029, 4030, 4031, 4032])
@pytest.mark.parametrize("use_amplitudes", [False, True])
def test_masked_sum(abs_mode, cols, use_amplitudes):
    rs = np.random.RandomState(1)
    rows = 4096
    src = complex_normal(rs, (rows, cols))
    mask = (rs.random_sample(rows) < 0.9).astype(np.float32)
    expect = contract.masked_sum(src, mask, use_amplitudes, abs_mode)
    out = cu.masked_sum(src, mask, use_amplitudes, abs_mode, pad=3)
    scale = np.sum(np.abs(src) * mask[:, None], axis=0)
    assert np.all(np.abs(expect - out) <= 1e-7 * scale)
    # the reference test's expression (test/test_maskedsum.py:62-67), evaluated in float64 so that
    # numpy's own float32 summation error (~2e-6 here) does not mask ours; tolerance as there
    data = (np.abs(src).astype(np.float64) if use_amplitudes else src.astype(np.complex128))
    ref = np.sum(data * mask[:, None], axis=0)
    assert np.all(np.abs(ref - out) <= 1e-6 * scale)


@pytest.mark.parametrize("rows,cols,pad", [(1, 1, 0), (3, 33, 1), (255, 31, 1), (256, 32, 0),
                                           (700, 1, 1), (5000, 97, 1), (5000, 97, 0), (2049, 640, 0),
                                           (300, 9000, 0)])
@pytest.mark.parametrize("use_amplitudes", [False, True])
def test_masked_sum_geometry(abs_mode, rows, cols, pad, use_amplitudes):
    """Row splits of 1 to 8 blocks per strip of columns (a thread block cluster), one or two
    columns per lane (odd last column, odd strides), fewer rows than row groups, no rows."""
    rs = np.random.RandomState(rows * 131 + cols)
    src = complex_normal(rs, (rows, cols))
    mask = rs.uniform(0, 2, rows).astype(np.float32)
    out = cu.masked_sum(src, mask, use_amplitudes, abs_mode, pad=pad)
    data = (contract.amplitude(src.ravel(), abs_mode).reshape(src.shape).astype(np.float64)
            if use_amplitudes else src.astype(np.complex128))
    ref = np.sum(data * mask.astype(np.float64)[:, None], axis=0)
    scale = np.sum(np.abs(src) * mask[:, None], axis=0)
    assert out.shape == (cols,)
    assert np.all(np.abs(ref - out) <= 1e-7 * scale + 1e-30)


# ------------------------------------------------------------------ amplitude rule (R1)
def test_amplitude_rule_wide_range(abs_mode):
    """A one-row MaskedSum of amplitudes with mask 1 returns the amplitudes themselves, so the
    kernel's complex absolute value can be compared bit for bit with the oracle's over the
    whole float range (normal data, huge / tiny exponents, denormals, zeros, inf, NaN)."""
    rs = np.random.RandomState(77)
    n = 1 << 20
    parts = [complex_normal(rs, n)]
    for lo, hi in ((-149, 128), (-70, -55), (55, 70), (-20, 20)):
        mag = np.ldexp(rs.uniform(1, 2, (2, n // 4)), rs.randint(lo, hi, (2, n // 4))).astype(np.float32)
        sign = rs.choice([-1.0, 1.0], (2, n // 4)).astype(np.float32)
        parts.append((mag[0] * sign[0] + 1j * (mag[1] * sign[1])).astype(np.complex64))
    special = np.array([0, 1e-45, 1e-38, 1.0, 3e38, np.inf, np.nan, -0.0, 2.0 ** -64, 2.0 ** 64,
                        2.0 ** -65, 2.0 ** 63, 1.17549435e-38], np.float32)
    grid = (special[:, None] + 1j * special[None, :]).astype(np.complex64).ravel()
    parts.append(grid)
    vis = np.concatenate(parts)[None, :]
    with np.errstate(all="ignore"):
        expect = contract.amplitude(vis[0], abs_mode)
    out = cu.masked_sum(vis, np.ones(1, np.float32), True, abs_mode)
    assert_same_f32(expect, out)
    if abs_mode == contract.detect_abs_mode():
        with np.errstate(all="ignore"):
            assert_same_f32(np.abs(vis[0]), out)
