"""The dataflow form of the fused flagger (one persistent kernel per dump, csrc/dataflow.cu):
bit-exact against the C oracle over shapes that exercise every part of its schedule - ragged
last strips and groups, rings that wrap, one and several threshold spans per row, every input
and flag mode - and equal to the chunked four-kernel form on large inputs.

Reference behaviour: rfi/device.py:1111-1166 (FlaggerDevice), test/rfi/test_flagger.py:74-132.
"""

import numpy as np
import pytest

import cabi_util as cu
from oracle import contract
from oracle import host_numpy as hn

pytestmark = pytest.mark.gpu

DATAFLOW = -1      # chunk_baselines < 0: the dataflow form or KSP_EINVAL


def make_vis(rs, channels, baselines, features=True):
    vis = (rs.standard_normal((channels, baselines))
           + 1j * rs.standard_normal((channels, baselines))).astype(np.complex64)
    spikes = rs.random_sample(vis.shape) < 1 / 64
    vis += (spikes * (rs.random_sample(vis.shape) * 20 + 50)
            * np.exp(2j * np.pi * rs.random_sample(vis.shape))).astype(np.complex64)
    if features and channels >= 256:
        for _ in range(max(4, baselines // 4)):               # weak broad features: larger windows
            bl, s = rs.randint(0, baselines), rs.randint(0, channels - 70)
            vis[s:s + rs.randint(2, 64), bl] += np.complex64(rs.uniform(2.0, 6.0))
    return vis


def check(vis, input_flags=None, *, n_windows=7, n_sigma=11.0, abs_mode, amplitudes=False, pad=0,
          flag_value=1):
    want_flags, _, want_noise = contract.flagger(vis, input_flags, n_windows=n_windows,
                                                 n_sigma=n_sigma, abs_mode=abs_mode,
                                                 amplitudes=amplitudes, flag_value=flag_value)
    flags, noise = cu.flagger(vis, input_flags, n_windows=n_windows, n_sigma=n_sigma,
                              abs_mode=abs_mode, amplitudes=amplitudes, chunk_baselines=DATAFLOW,
                              pad=pad, flag_value=flag_value)
    st = dict(cu.LAST_FLAGGER_STATS)
    assert st["dataflow"] == 1 and st["error"] == 0
    a, b = np.ascontiguousarray(noise), np.ascontiguousarray(want_noise)
    bad = (a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))
    assert not bad.any(), (int(bad.sum()), np.argwhere(bad)[:5])
    np.testing.assert_array_equal(want_flags, flags)
    return st


@pytest.mark.parametrize("channels,baselines", [
    (32, 1), (64, 5), (256, 33), (512, 128), (2048, 96), (4096, 130), (8192, 31),
    (5152, 40),        # just over one span: two threshold spans per row
    (8224, 40),
    (10240, 64), (32768, 70),
    (2048, 1000),      # 32 strips: both rings wrap several times
    (1024, 3000),
])
def test_shapes(abs_mode, channels, baselines):
    rs = np.random.RandomState(channels + baselines)
    st = check(make_vis(rs, channels, baselines), abs_mode=abs_mode, n_sigma=5.0)
    strips = -(-baselines // 32)
    tiles = -(-channels // 256)
    # background tiles, noise rows, threshold rows, expansion tiles
    assert st["items"] == strips * tiles + 2 * baselines + -(-strips // 4) * tiles


@pytest.mark.parametrize("n_windows", [1, 2, 4, 6, 7])
def test_window_counts(abs_mode, n_windows):
    rs = np.random.RandomState(n_windows)
    check(make_vis(rs, 12288, 75), n_windows=n_windows, n_sigma=4.0, abs_mode=abs_mode, pad=5)


@pytest.mark.parametrize("kind", ["channel", "full"])
def test_input_flags(abs_mode, kind):
    rs = np.random.RandomState(11)
    channels, baselines = 4096, 100
    vis = make_vis(rs, channels, baselines)
    if kind == "channel":
        fl = (rs.random_sample(channels) < 0.02).astype(np.uint8) * 3
        fl[100:140] = 1                                   # whole windows flagged
    else:
        fl = (rs.random_sample((channels, baselines)) < 0.02).astype(np.uint8) * 5
        fl[200:230, 10:20] = 1
    check(vis, fl, abs_mode=abs_mode, n_sigma=6.0, pad=3)


def test_amplitude_input_and_flag_value(abs_mode):
    rs = np.random.RandomState(5)
    amp = np.abs(make_vis(rs, 4096, 64)).astype(np.float32)
    check(amp, abs_mode=abs_mode, amplitudes=True, flag_value=7)


def test_other_abs_mode():
    rs = np.random.RandomState(6)
    vis = make_vis(rs, 2048, 64)
    for mode in (contract.ABS_NUMPY, contract.ABS_HYPOT):
        check(vis, abs_mode=mode)


def test_nan_and_zero_rows(abs_mode):
    rs = np.random.RandomState(8)
    vis = make_vis(rs, 2048, 64)
    vis[:, 3] = 0                                          # no usable deviation: noise NaN
    vis[100:110, 5] = np.nan
    vis[:, 7] = np.complex64(1 + 1j)                       # constant: all deviations zero
    check(vis, abs_mode=abs_mode)


def test_equals_chunked_form_at_full_channel_count(abs_mode):
    """32768 x 416: dataflow == chunked, and a subset of baselines == oracle."""
    vis, _ = hn.synthetic_vis(32768, 416, seed=4)
    vis[20000:20004, 100:140] += np.complex64(6.0)
    f1, n1 = cu.flagger(vis, n_windows=7, abs_mode=abs_mode, chunk_baselines=DATAFLOW)
    assert cu.LAST_FLAGGER_STATS["dataflow"] == 1 and cu.LAST_FLAGGER_STATS["error"] == 0
    f2, n2 = cu.flagger(vis, n_windows=7, abs_mode=abs_mode, chunk_baselines=128)
    assert cu.LAST_FLAGGER_STATS["dataflow"] == 0
    assert np.array_equal(n1.view(np.uint32), n2.view(np.uint32))
    np.testing.assert_array_equal(f1, f2)
    pick = np.r_[0:8, 100:140:3, 408:416]
    want_flags, _, want_noise = contract.flagger(np.ascontiguousarray(vis[:, pick]), None,
                                                 n_windows=7, abs_mode=abs_mode)
    assert np.array_equal(n1[pick].view(np.uint32), want_noise.view(np.uint32))
    np.testing.assert_array_equal(want_flags, f1[:, pick])


def test_repeated_launches_same_scratch(context, command_queue, abs_mode):
    """The Operation reuses its scratch (ring, counters) from call to call."""
    from katsdpsigproc_b200.rfi import device as rfi
    rs = np.random.RandomState(9)
    channels, baselines = 4096, 200
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 1 << 20),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=7))
    fn = rfi.FusedFlaggerDevice(template.background, template.threshold, command_queue, channels,
                                baselines, 11.0, chunk_baselines=DATAFLOW)
    fn.ensure_all_bound()
    assert fn.parameters()["dataflow"]
    for _ in range(3):
        vis = make_vis(rs, channels, baselines)
        fn.buffer("vis").set(command_queue, vis)
        fn()
        flags = np.array(fn.buffer("out_flags").get(command_queue))
        noise = np.array(fn.buffer("noise").get(command_queue))
        st = fn.stats()
        assert st["error"] == 0 and st["items"] > 0
        want_flags, _, want_noise = contract.flagger(vis, None, n_windows=7, abs_mode=abs_mode)
        assert np.array_equal(noise.view(np.uint32), want_noise.view(np.uint32))
        np.testing.assert_array_equal(want_flags, flags)


def test_not_legal_is_an_error():
    vis = np.zeros((48, 4), np.complex64)                  # channels not a multiple of 32
    with pytest.raises(Exception):
        cu.flagger(vis, chunk_baselines=DATAFLOW)
