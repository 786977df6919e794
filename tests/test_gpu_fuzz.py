"""Randomised parity: the fused flagger and the stage kernels against the C oracle over random
shapes, parameters and interference patterns, with low thresholds so that the rarely taken
paths (survivor evaluation, dilation + rebuild over several window sizes, masked medians,
radix-select fallbacks of the noise estimate) are exercised."""

import numpy as np
import pytest

import cabi_util as cu
from oracle import contract

pytestmark = pytest.mark.gpu


def random_vis(rs, channels, baselines):
    vis = (rs.standard_normal((channels, baselines))
           + 1j * rs.standard_normal((channels, baselines))).astype(np.complex64)
    scale = rs.choice([1.0, 1e-3, 250.0])
    vis *= np.complex64(scale)
    kind = rs.randint(0, 5)
    if kind >= 1:        # isolated spikes
        hit = rs.random_sample(vis.shape) < rs.choice([1 / 200, 1 / 30, 1 / 8])
        vis += (hit * scale * rs.uniform(3, 60)).astype(np.complex64)
    if kind >= 2 and channels > 8:        # narrow features of random width and strength
        for _ in range(rs.randint(1, 6)):
            c = rs.randint(0, channels - 1)
            w = rs.randint(1, min(channels - c, 12) + 1)
            b = rs.randint(0, baselines)
            vis[c:c + w, b:b + rs.randint(1, 8)] += np.complex64(scale * rs.uniform(1.0, 6.0))
    if kind == 3:        # quantised data: heavy ties
        vis = (np.round(vis.real / scale * 2) / 2 * scale
               + 1j * np.round(vis.imag / scale * 2) / 2 * scale).astype(np.complex64)
    if kind == 4:        # dead baselines and channels
        vis[:, rs.randint(0, baselines)] = 0
        vis[rs.randint(0, channels)] = 0
    return vis


@pytest.mark.parametrize("seed", range(40))
def test_fused_flagger_fuzz(abs_mode, seed):
    rs = np.random.RandomState(1000 + seed)
    channels = int(rs.choice([1, 7, 13, 14, 40, 127, 128, 129, 700, 2049, 4096 + 17]))
    baselines = int(rs.choice([1, 3, 31, 32, 33, 70]))
    vis = random_vis(rs, channels, baselines)
    n_windows = int(rs.randint(1, 8))
    n_sigma = float(rs.choice([2.0, 3.0, 4.5, 11.0]))
    falloff = float(rs.choice([1.2, 1.5, 2.5]))
    flag_kind = rs.randint(0, 3)
    fl = None
    if flag_kind == 1:
        fl = (rs.random_sample(channels) < 0.1).astype(np.uint8)
    elif flag_kind == 2:
        fl = (rs.random_sample(vis.shape) < rs.choice([0.02, 0.3])).astype(np.uint8) * 7
    with np.errstate(all="ignore"):
        want_flags, _, want_noise = contract.flagger(
            vis, fl, n_windows=n_windows, n_sigma=n_sigma, threshold_falloff=falloff,
            flag_value=3, abs_mode=abs_mode)
    flags, noise = cu.flagger(vis, fl, n_windows=n_windows, n_sigma=n_sigma, falloff=falloff,
                              flag_value=3, abs_mode=abs_mode,
                              chunk_baselines=int(rs.choice([0, 32])), pad=int(rs.choice([0, 5])))
    same = (noise.view(np.uint32) == want_noise.view(np.uint32)) | (np.isnan(noise) & np.isnan(want_noise))
    assert same.all(), (seed, noise[~same][:4], want_noise[~same][:4])
    np.testing.assert_array_equal(want_flags, flags, err_msg=f"seed {seed}")


@pytest.mark.parametrize("seed", range(30))
def test_fused_flagger_fuzz_other_parameters(abs_mode, seed):
    """The same with median widths other than 13 (sliding-window and generic kernels) and up to
    11 window sizes (general sum-threshold kernel), amplitude input now and then."""
    rs = np.random.RandomState(5000 + seed)
    channels = int(rs.choice([1, 9, 40, 129, 700, 2049, 4096 + 17, 9000]))
    baselines = int(rs.choice([1, 3, 31, 33, 70]))
    vis = random_vis(rs, channels, baselines)
    width = int(rs.choice([1, 3, 5, 7, 9, 11, 15, 17, 21, 31, 33, 45]))
    n_windows = int(rs.randint(1, 12))
    n_sigma = float(rs.choice([2.0, 3.0, 4.5]))
    falloff = float(rs.choice([1.2, 1.5]))
    amplitudes = bool(seed % 5 == 0)
    data = np.abs(vis) if amplitudes else vis
    flag_kind = rs.randint(0, 3)
    fl = None
    if flag_kind == 1:
        fl = (rs.random_sample(channels) < 0.1).astype(np.uint8)
    elif flag_kind == 2:
        fl = (rs.random_sample(vis.shape) < 0.05).astype(np.uint8) * 7
    with np.errstate(all="ignore"):
        want_flags, _, want_noise = contract.flagger(
            data, fl, width=width, n_windows=n_windows, n_sigma=n_sigma, threshold_falloff=falloff,
            flag_value=1, amplitudes=amplitudes, abs_mode=abs_mode)
    flags, noise = cu.flagger(data, fl, width=width, n_windows=n_windows, n_sigma=n_sigma,
                              falloff=falloff, flag_value=1, amplitudes=amplitudes,
                              abs_mode=abs_mode, pad=int(rs.choice([0, 3])))
    same = (noise.view(np.uint32) == want_noise.view(np.uint32)) | (np.isnan(noise) & np.isnan(want_noise))
    assert same.all(), (seed, noise[~same][:4], want_noise[~same][:4])
    np.testing.assert_array_equal(want_flags, flags, err_msg=f"seed {seed}")


@pytest.mark.parametrize("seed", range(25))
def test_threshold_sum_fuzz(seed):
    """Deviations with broad and clustered excesses close to the thresholds."""
    rs = np.random.RandomState(2000 + seed)
    channels = int(rs.choice([64, 100, 513, 4096, 4500, 9000]))
    baselines = 6
    dev = rs.standard_normal((channels, baselines)).astype(np.float32)
    for bl in range(baselines):
        for _ in range(rs.randint(0, 12)):
            c = rs.randint(0, channels - 2)
            w = rs.randint(1, min(channels - c, 150))
            dev[c:c + w, bl] += np.float32(rs.uniform(0.3, 4.0))
    if seed % 5 == 0:
        dev[rs.randint(0, channels), 0] = np.nan
        dev[rs.randint(0, channels), 1] = np.inf
        dev[rs.randint(0, channels), 2] = -np.inf
    noise = rs.uniform(0.3, 1.5, baselines).astype(np.float32)
    if seed % 7 == 0:
        noise[0] = 0.0
        noise[1] = -0.5
    n_windows = int(rs.randint(2, 8))
    n_sigma = float(rs.choice([1.5, 2.5, 3.5]))
    falloff = float(rs.choice([1.2, 1.5, 2.5]))
    with np.errstate(all="ignore"):
        want = contract.threshold_sum(dev, noise, n_sigma, n_windows, falloff)
    got = cu.threshold_sum(np.ascontiguousarray(dev.T), noise, n_sigma, n_windows, falloff)
    np.testing.assert_array_equal(want, got.T, err_msg=f"seed {seed}")


@pytest.mark.parametrize("seed", range(15))
def test_noise_fuzz(seed):
    """Rows with ties, tiny and huge values, few usable samples, denormals."""
    rs = np.random.RandomState(3000 + seed)
    channels = int(rs.choice([1, 2, 5, 33, 1000, 5000, 32768]))
    baselines = 9
    dev = rs.standard_normal((baselines, channels)).astype(np.float32)
    dev[0] = np.round(dev[0] * 2) / 2                      # ties
    dev[1] *= np.float32(1e-40)                            # denormals
    dev[2] *= np.float32(1e30)
    dev[3, rs.random_sample(channels) < 0.95] = 0          # nearly all zero
    dev[4] = 0
    dev[5] = np.float32(3.25)                              # all equal
    dev[6, ::2] = np.nan
    dev[7] = np.abs(dev[7]) ** 8                           # heavy tail
    with np.errstate(all="ignore"):
        want, _ = contract.noise_mad(dev, transposed=True)
    got = cu.madnz(dev, True)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (seed, got, want)
    got_cm = cu.madnz(np.ascontiguousarray(dev.T), False)
    same = (got_cm.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got_cm) & np.isnan(want))
    assert same.all(), (seed, got_cm, want)
