"""Parity at BASELINE.json's full channel counts, through checks that do not need the
whole oracle: a deterministic subset of baselines against the C oracle (baselines are
independent), plus size-independent properties -- sharding invariance, chunk / lane
invariance, injected interference recovered, input flags never set in the output."""

import os

import numpy as np
import pytest

import cabi_util as cu
from katsdpsigproc_b200 import sharding
from katsdpsigproc_b200.rfi import device as rfi
from oracle import contract
from oracle import host_numpy as hn

pytestmark = pytest.mark.gpu


def subset(baselines):
    """First 16, every 65th, last 16 (parity check at scale)."""
    idx = set(range(min(16, baselines))) | set(range(0, baselines, 65)) | \
        set(range(max(0, baselines - 16), baselines))
    return np.array(sorted(idx))


def run_flagger(context, queue, vis, abs_mode, input_flags=None, n_windows=7, **kw):
    use = rfi.BackgroundFlags.NONE
    if input_flags is not None:
        use = rfi.BackgroundFlags.CHANNEL if input_flags.ndim == 1 else rfi.BackgroundFlags.FULL
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, use_flags=use, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 1 << 20),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=n_windows), **kw)
    fn = template.instantiate(queue, *vis.shape, threshold_args={"n_sigma": 11.0})
    fn.ensure_all_bound()
    fn.buffer("vis").set(queue, vis)
    if input_flags is not None:
        fn.buffer("input_flags").set(queue, input_flags)
    fn()
    return np.array(fn.buffer("flags").get(queue)), np.array(fn.buffer("noise").get(queue))


@pytest.fixture(scope="module")
def meerkat_dump():
    """32768 channels x 416 baselines (1/20 of the MeerKAT dump), noise + spikes + lines."""
    vis, spikes = hn.synthetic_vis(32768, 416, seed=3)
    vis[20000:20004, 100:140] += np.complex64(6.0)        # a weak 4-channel feature: larger windows
    return vis, spikes


def test_cfg2_subset_against_oracle(context, command_queue, abs_mode, meerkat_dump):
    vis, spikes = meerkat_dump
    flags, noise = run_flagger(context, command_queue, vis, abs_mode)
    pick = np.union1d(subset(vis.shape[1]), np.arange(100, 140, 3))
    want_flags, _, want_noise = contract.flagger(np.ascontiguousarray(vis[:, pick]), None,
                                                 n_windows=7, abs_mode=abs_mode)
    assert np.array_equal(noise[pick].view(np.uint32), want_noise.view(np.uint32))
    np.testing.assert_array_equal(want_flags, flags[:, pick])
    # every injected sample is found and little else is flagged
    assert np.all(flags[spikes != 0] == 1)
    assert flags.mean() < spikes.mean() + 0.001


def test_sharding_invariance(context, command_queue, abs_mode, meerkat_dump):
    """Flagging a rank's column block alone gives that block of the full result."""
    vis, _ = meerkat_dump
    flags, noise = run_flagger(context, command_queue, vis, abs_mode)
    for world in (2, 8):
        for rank in (0, world - 1):
            start, stop = sharding.baseline_range(vis.shape[1], rank, world)
            part = np.ascontiguousarray(vis[:, start:stop])
            f, n = run_flagger(context, command_queue, part, abs_mode)
            np.testing.assert_array_equal(flags[:, start:stop], f)
            assert np.array_equal(noise[start:stop].view(np.uint32), n.view(np.uint32))


def test_chunk_and_lane_invariance(abs_mode, meerkat_dump):
    vis = np.ascontiguousarray(meerkat_dump[0][:8192, :200])
    ref_flags, ref_noise = cu.flagger(vis, n_windows=7, abs_mode=abs_mode)
    old = os.environ.get("KSP_LANES")
    try:
        for lanes, chunk in ((1, 32), (2, 64), (4, 32), (3, 96), (4, 0)):
            os.environ["KSP_LANES"] = str(lanes)
            flags, noise = cu.flagger(vis, n_windows=7, abs_mode=abs_mode, chunk_baselines=chunk)
            np.testing.assert_array_equal(ref_flags, flags)
            assert np.array_equal(ref_noise.view(np.uint32), noise.view(np.uint32))
    finally:
        if old is None:
            os.environ.pop("KSP_LANES", None)
        else:
            os.environ["KSP_LANES"] = old


def test_fused_equals_sequence_at_full_channels(context, command_queue, abs_mode, meerkat_dump):
    vis = np.ascontiguousarray(meerkat_dump[0][:, :96])
    f1, n1 = run_flagger(context, command_queue, vis, abs_mode, fused=True)
    f2, n2 = run_flagger(context, command_queue, vis, abs_mode, fused=False)
    np.testing.assert_array_equal(f1, f2)
    assert np.array_equal(n1.view(np.uint32), n2.view(np.uint32))


def test_input_flags_are_never_set(context, command_queue, abs_mode, meerkat_dump):
    vis = np.ascontiguousarray(meerkat_dump[0][:, :64])
    first, _ = run_flagger(context, command_queue, vis, abs_mode)
    again, _ = run_flagger(context, command_queue, vis, abs_mode, input_flags=first)
    assert not np.any(again[first != 0])
    want, _, _ = contract.flagger(vis, first, n_windows=7, abs_mode=abs_mode)
    np.testing.assert_array_equal(want, again)


def test_cfg5_shard_shape(context, command_queue, abs_mode):
    """An 8-GPU shard of the 80-antenna array: 1620 baselines (not a multiple of 32)."""
    vis, spikes = hn.synthetic_vis(4096, 1620, seed=11)
    flags, noise = run_flagger(context, command_queue, vis, abs_mode)
    pick = subset(1620)
    want_flags, _, want_noise = contract.flagger(np.ascontiguousarray(vis[:, pick]), None,
                                                 n_windows=7, abs_mode=abs_mode)
    np.testing.assert_array_equal(want_flags, flags[:, pick])
    assert np.array_equal(noise[pick].view(np.uint32), want_noise.view(np.uint32))
    assert np.all(flags[spikes != 0] == 1)


def fast_dump(channels, baselines, seed):
    """Noise + isolated spikes + narrowband lines like hn.synthetic_vis, from numpy's fast float32
    generator (a 32768 x 8320 dump in seconds rather than minutes)."""
    rng = np.random.default_rng(seed)
    vis = np.empty((channels, baselines), np.complex64)
    spikes = np.empty((channels, baselines), bool)
    for c0 in range(0, channels, 2048):
        n = min(2048, channels - c0)
        block = vis[c0:c0 + n].view(np.float32).reshape(n, baselines, 2)
        block[...] = rng.standard_normal((n, baselines, 2), dtype=np.float32)
        hit = rng.random((n, baselines), dtype=np.float32) < np.float32(1 / 64)
        hit |= rng.random((n, 1), dtype=np.float32) < np.float32(0.005)
        amp = rng.random((n, baselines), dtype=np.float32) * np.float32(20) + np.float32(50)
        phase = rng.random((n, baselines), dtype=np.float32) * np.float32(2 * np.pi)
        block[..., 0] += hit * amp * np.cos(phase)
        block[..., 1] += hit * amp * np.sin(phase)
        spikes[c0:c0 + n] = hit
    return vis, spikes


@pytest.mark.parametrize("baselines", [8320, 1620])
def test_benchmark_shape_on_the_default_path(context, command_queue, abs_mode, baselines):
    """The shapes bench.py times, run the way bench.py runs them (FlaggerDeviceTemplate defaults:
    8320 baselines = chunks of 3 x 2368 + 1216 on 4 lanes; 1620, the 8-GPU shard of the 12960-baseline
    dump = two launches per stage), with the parity check of SURVEY.md 8(d): a deterministic subset
    of baselines at the full channel count against the oracle, plus injected RFI recovered."""
    vis, spikes = fast_dump(32768, baselines, seed=baselines)
    flags, noise = run_flagger(context, command_queue, vis, abs_mode)
    idx = set(range(64)) | set(range(0, baselines, 65)) | set(range(baselines - 64, baselines))
    pick = np.array(sorted(idx))
    want_flags, _, want_noise = contract.flagger(np.ascontiguousarray(vis[:, pick]), None,
                                                 n_windows=7, abs_mode=abs_mode)
    assert np.array_equal(noise[pick].view(np.uint32), want_noise.view(np.uint32))
    np.testing.assert_array_equal(want_flags, flags[:, pick])
    assert np.all(flags[spikes] == 1)
    assert flags.mean() < spikes.mean() + 0.001
    assert np.all(np.isfinite(noise)) and noise.min() > 0.5 and noise.max() < 2.0


@pytest.mark.parametrize("channels", [256, 4096, 65536])
def test_cfg3_noise_and_percentile_sweep(abs_mode, channels):
    """Percentile5 / MAD sweep of BASELINE.json configs[2] on a slice of rows: bit-exact."""
    rs = np.random.RandomState(channels)
    rows = 48
    dev = rs.standard_normal((rows, channels)).astype(np.float32)
    dev[rs.random_sample(dev.shape) < 0.07] = 0
    np.testing.assert_array_equal(contract.noise_mad(dev, transposed=True)[0].view(np.uint32),
                                  cu.madnz(dev, True).view(np.uint32))
    amp = np.abs(dev)
    ref = np.percentile(amp, [0, 100, 25, 75, 50], axis=1, method="lower").astype(np.float32)
    np.testing.assert_array_equal(ref, cu.percentile5(amp))
    cplx = (dev + 1j * rs.standard_normal(dev.shape)).astype(np.complex64)
    np.testing.assert_array_equal(contract.percentile5(cplx, None, abs_mode),
                                  cu.percentile5(cplx, None, abs_mode))
