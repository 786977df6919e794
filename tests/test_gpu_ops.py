"""Parity through the public Template / Operation API on a real device.

These read like the reference's own device tests (``test/rfi/test_background.py``,
``test_noise_est.py``, ``test_threshold.py``, ``test_flagger.py``, ``test_percentile.py``,
``test_transpose.py``, ``test_maskedsum.py``, ``test_accel.py:95-380``): fixed-seed numpy
input, run through a ``*HostFromDevice`` wrapper or directly bound slots with
non-trivial padding, compare with the host oracle.
"""

import numpy as np
import pytest

from katsdpsigproc_b200 import accel, maskedsum, percentile, transpose
from katsdpsigproc_b200.accel import DeviceArray, HostArray
from katsdpsigproc_b200.rfi import device as rfi
from oracle import contract
from oracle import host_numpy as hn

pytestmark = pytest.mark.gpu


def complex_normal(rs, shape):
    return (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32))
                                              | (np.isnan(a) & np.isnan(b))))


# ----------------------------------------------------------------------------- DeviceArray
class TestDeviceArray:
    @pytest.fixture(autouse=True)
    def setup(self, context, command_queue):
        self.context, self.queue = context, command_queue
        self.shape, self.padded = (17, 13), (32, 16)
        self.array = DeviceArray(context, self.shape, np.int32, self.padded)

    def test_set_get_roundtrip(self):
        ary = np.random.RandomState(1).randint(0, 100, self.shape).astype(np.int32)
        self.array.set(self.queue, ary)
        np.testing.assert_array_equal(ary, self.array.get(self.queue))
        host = self.array.empty_like()
        host[...] = ary + 1
        self.array.set_async(self.queue, host)
        out = self.array.get_async(self.queue)
        self.queue.finish()
        np.testing.assert_array_equal(ary + 1, out)
        target = self.array.empty_like()
        assert self.array.get(self.queue, target) is target

    def test_set_rejects_other_dtypes(self):
        with pytest.raises(TypeError):
            self.array.set(self.queue, np.zeros(self.shape, np.float32))

    def test_zero(self):
        self.array.set(self.queue, np.ones(self.shape, np.int32))
        self.array.zero(self.queue)
        assert not self.array.get(self.queue).any()

    def test_regions(self):
        rs = np.random.RandomState(2)
        src = rs.randint(0, 1000, self.shape).astype(np.int32)
        self.array.set(self.queue, src)
        other = DeviceArray(self.context, (9, 20), np.int32, (10, 24))
        other.zero(self.queue)
        self.array.copy_region(self.queue, other, np.s_[3:8, 2:12:2], np.s_[1:6, 10:15])
        expect = np.zeros((9, 20), np.int32)
        expect[1:6, 10:15] = src[3:8, 2:12:2]
        np.testing.assert_array_equal(expect, other.get(self.queue))

        host = HostArray((6, 13), np.int32, (8, 16), context=self.context)
        host[...] = -1
        self.array.get_region(self.queue, host, np.s_[10:14, 1:], np.s_[2:6, :12])
        expect_h = np.full((6, 13), -1, np.int32)
        expect_h[2:6, :12] = src[10:14, 1:]
        np.testing.assert_array_equal(expect_h, host)
        with pytest.raises(ValueError):
            self.array.get_region(self.queue, np.zeros((6, 13), np.int32), np.s_[:6], np.s_[:])

        patch = rs.randint(0, 1000, (4, 5)).astype(np.int32)       # plain array: staged
        self.array.set_region(self.queue, patch, np.s_[0:4, 8:13], np.s_[:, :])
        src[0:4, 8:13] = patch
        np.testing.assert_array_equal(src, self.array.get(self.queue))
        self.array.set_region(self.queue, host, np.s_[16, :], np.s_[5, :])
        src[16, :] = host[5, :]
        np.testing.assert_array_equal(src, self.array.get(self.queue))

    def test_events_time_a_marker_pair(self):
        start = self.queue.enqueue_marker()
        self.array.zero(self.queue)
        end = self.queue.enqueue_marker()
        assert end.time_since(start) >= 0.0 and start.time_till(end) >= 0.0
        other = self.context.create_command_queue()
        other.enqueue_wait_for_events([end])
        other.finish()


# ----------------------------------------------------------------------------- stages
@pytest.mark.parametrize("width", [5, 13])
@pytest.mark.parametrize("use_flags", list(rfi.BackgroundFlags))
@pytest.mark.parametrize("amplitudes", [False, True])
def test_background(context, command_queue, abs_mode, width, use_flags, amplitudes):
    rs = np.random.RandomState(1)
    vis = complex_normal(rs, (417, 313))
    flags = (rs.random_sample(vis.shape) < 0.1).astype(np.uint8) * 3
    flags[100:110, :] = 1
    if amplitudes:
        vis = np.abs(vis)
    fl = {rfi.BackgroundFlags.NONE: None, rfi.BackgroundFlags.CHANNEL: flags[:, 0].copy(),
          rfi.BackgroundFlags.FULL: flags}[use_flags]
    template = rfi.BackgroundMedianFilterDeviceTemplate(context, width, amplitudes, use_flags,
                                                        abs_mode=abs_mode)
    out = rfi.BackgroundHostFromDevice(template, command_queue)(vis, fl)
    assert same_bits(contract.background(vis, width, fl, amplitudes, abs_mode), out)
    # the reference's own check: device vs host class to atol 1e-6
    host = hn.background_median_filter(vis, width, fl, amplitudes)
    np.testing.assert_allclose(host, out, atol=1e-6)
    if abs_mode == contract.detect_abs_mode():
        assert same_bits(host.astype(np.float32), out)


@pytest.mark.parametrize("transposed", [False, True])
def test_noise_est(context, command_queue, transposed):
    rs = np.random.RandomState(1)
    deviations = rs.standard_normal((117, 273)).astype(np.float32)
    template = (rfi.NoiseEstMADTDeviceTemplate(context, 10240) if transposed
                else rfi.NoiseEstMADDeviceTemplate(context))
    out = rfi.NoiseEstHostFromDevice(template, command_queue)(deviations)
    assert same_bits(contract.noise_mad(deviations)[0], out)
    np.testing.assert_allclose(hn.noise_est_mad(deviations), out, rtol=2e-7)


@pytest.mark.parametrize("shape", [(2048, 37), (4100, 70), (2047, 33)])
def test_noise_est_channel_major_long_rows(context, command_queue, shape):
    """From NoiseEstMADDevice.TRANSPOSE_FROM channels on, the channel-major operation transposes
    into its scratch_t slot and runs the baseline-major kernel; shorter rows (2047) use the
    channel-major kernel.  Zeros and NaN are skipped either way; an all-zero baseline gives NaN."""
    rs = np.random.RandomState(5)
    deviations = rs.standard_normal(shape).astype(np.float32)
    deviations[rs.random_sample(shape) < 0.05] = 0.0
    deviations[rs.random_sample(shape) < 0.01] = np.nan
    deviations[:, 3] = 0.0
    template = rfi.NoiseEstMADDeviceTemplate(context)
    fn = template.instantiate(command_queue, *shape)
    assert ("scratch_t" in fn.slots) == (shape[0] >= fn.TRANSPOSE_FROM)
    out = rfi.NoiseEstHostFromDevice(template, command_queue)(deviations)
    expect = contract.noise_mad(deviations)[0]
    assert same_bits(expect, out)
    assert np.isnan(out[3])


def threshold_case():
    rs = np.random.RandomState(1)
    deviations = rs.standard_normal((117, 273)).astype(np.float32) * 10.0
    spikes = rs.random_sample(deviations.shape) < 0.25
    deviations[spikes] += 200.0
    noise = np.linspace(0.0, 50.0, 273).astype(np.float32)
    return deviations, noise


@pytest.mark.parametrize("transposed", [False, True])
def test_threshold_simple(context, command_queue, transposed):
    deviations, noise = threshold_case()
    template = rfi.ThresholdSimpleDeviceTemplate(context, transposed, flag_value=3)
    out = rfi.ThresholdHostFromDevice(template, command_queue, 11.0)(deviations, noise)
    np.testing.assert_array_equal(hn.threshold_simple(deviations, noise, 11.0, 3), out)


@pytest.mark.parametrize("n_windows", [1, 4, 7])
def test_threshold_sum(context, command_queue, n_windows):
    deviations, noise = threshold_case()
    template = rfi.ThresholdSumDeviceTemplate(context, n_windows=n_windows)
    out = rfi.ThresholdHostFromDevice(template, command_queue, 11.0, threshold_falloff=1.5)(
        deviations, noise)
    near = {}
    host = hn.threshold_sum(deviations, noise, 11.0, n_windows, 1.5, 1, near)
    np.testing.assert_array_equal(
        contract.threshold_sum(deviations, noise, 11.0, n_windows, 1.5), out)
    assert int((host != out).sum()) <= near.get("band", 0)


# ----------------------------------------------------------------------------- host_class
# The reference's own device tests, written as they are written there: the expected value comes
# from ``template.host_class(...)`` (test/rfi/test_background.py:78-104, test_noise_est.py:54-61,
# test_threshold.py:60-70).  host_class resolves to the reference's katsdpsigproc.rfi.host, which
# oracle/_ref holds in the build container and on the GPU box (oracle/make_ref.py).
@pytest.fixture
def reference_on_path():
    import oracle
    if oracle.reference_host() is None:
        pytest.skip("oracle/_ref (the reference's rfi/host.py) is not present")


@pytest.mark.parametrize("amplitudes", [False, True])
@pytest.mark.parametrize("use_flags", list(rfi.BackgroundFlags))
def test_reference_background_test(context, command_queue, abs_mode, reference_on_path, amplitudes, use_flags):
    width = 5
    rs = np.random.RandomState(seed=1)                      # as _make_vis / _make_flags there
    vis = (rs.standard_normal((417, 313)) + rs.standard_normal((417, 313)) * 1j).astype(np.complex64)
    flags = (rs.random_sample((417, 313)) < 0.1).astype(np.uint8)
    flags[:, 3] = 1
    flags[100:110, :] = 1
    flags *= rs.randint(1, 256, flags.shape).astype(np.uint8)
    if amplitudes:
        vis = np.abs(vis)
    if use_flags is rfi.BackgroundFlags.CHANNEL:
        flags = flags[:, 0].copy()
    elif use_flags is rfi.BackgroundFlags.NONE:
        flags = None
    template = rfi.BackgroundMedianFilterDeviceTemplate(context, width, amplitudes, use_flags,
                                                        abs_mode=abs_mode)
    bg_host = template.host_class(width, amplitudes)
    bg_device = rfi.BackgroundHostFromDevice(template, command_queue)
    full = None if flags is None else (flags if flags.ndim == 2 else np.repeat(flags[:, None], 313, 1))
    expected = bg_host(vis, full)
    np.testing.assert_allclose(expected, bg_device(vis, flags), atol=1e-6)


@pytest.mark.parametrize("transposed", [False, True])
def test_reference_noise_est_test(context, command_queue, reference_on_path, transposed):
    rs = np.random.RandomState(seed=1)
    deviations = rs.standard_normal((117, 273)).astype(np.float32)
    template = (rfi.NoiseEstMADTDeviceTemplate(context, 10240) if transposed
                else rfi.NoiseEstMADDeviceTemplate(context))
    ne_host = template.host_class()
    ne_device = rfi.NoiseEstHostFromDevice(template, command_queue)
    np.testing.assert_allclose(ne_host(deviations), ne_device(deviations), rtol=2e-7)


@pytest.mark.parametrize("kind", ["simple", "simple_t", "sum"])
def test_reference_threshold_test(context, command_queue, reference_on_path, kind):
    deviations, noise = threshold_case()
    template = {"simple": lambda: rfi.ThresholdSimpleDeviceTemplate(context, False),
                "simple_t": lambda: rfi.ThresholdSimpleDeviceTemplate(context, True),
                "sum": lambda: rfi.ThresholdSumDeviceTemplate(context)}[kind]()
    th_host = template.host_class(11.0)
    th_device = rfi.ThresholdHostFromDevice(template, command_queue, 11.0)
    np.testing.assert_equal(th_host(deviations, noise), th_device(deviations, noise))


# ----------------------------------------------------------------------------- flagger
def flagger_case(channels=117, baselines=131):
    rs = np.random.RandomState(1)
    vis = complex_normal(rs, (channels, baselines))
    spikes = rs.random_sample(vis.shape) < 1 / 16
    vis += (spikes * (rs.random_sample(vis.shape) * 20 + 50)
            * np.exp(rs.random_sample(vis.shape) * 2j * np.pi)).astype(np.complex64)
    input_flags = (rs.random_sample(vis.shape) < 0.05).astype(np.uint8)
    return vis, spikes, input_flags


@pytest.mark.parametrize("noise_t", [False, True])
@pytest.mark.parametrize("threshold", ["simple", "simple_t", "sum"])
@pytest.mark.parametrize("use_flags", list(rfi.BackgroundFlags))
def test_flagger_sequence(context, command_queue, abs_mode, noise_t, threshold, use_flags):
    """The reference's test matrix (test/rfi/test_flagger.py:74-132): flags == injected spikes."""
    vis, spikes, input_flags = flagger_case()
    fl = {rfi.BackgroundFlags.NONE: None, rfi.BackgroundFlags.CHANNEL: input_flags[:, 0].copy(),
          rfi.BackgroundFlags.FULL: input_flags}[use_flags]
    background = rfi.BackgroundMedianFilterDeviceTemplate(context, 13, use_flags=use_flags,
                                                          abs_mode=abs_mode)
    noise = (rfi.NoiseEstMADTDeviceTemplate(context, 10240) if noise_t
             else rfi.NoiseEstMADDeviceTemplate(context))
    thr = {"simple": lambda: rfi.ThresholdSimpleDeviceTemplate(context, False),
           "simple_t": lambda: rfi.ThresholdSimpleDeviceTemplate(context, True),
           "sum": lambda: rfi.ThresholdSumDeviceTemplate(context, n_windows=4)}[threshold]()
    template = rfi.FlaggerDeviceTemplate(background, noise, thr, fused=False)
    out = rfi.FlaggerHostFromDevice(template, command_queue,
                                    threshold_args={"n_sigma": 11.0})(vis, fl)
    expect = spikes.astype(np.uint8)
    if fl is not None:
        mask = fl.reshape(-1, 1) if fl.ndim == 1 else fl
        expect = np.where(mask, 0, expect).astype(np.uint8)
    if threshold != "sum":
        np.testing.assert_array_equal(expect, out)
    want, _, _ = contract.flagger(vis, fl, n_windows=4 if threshold == "sum" else 0,
                                  abs_mode=abs_mode)
    np.testing.assert_array_equal(want, out)


@pytest.mark.parametrize("use_flags", list(rfi.BackgroundFlags))
@pytest.mark.parametrize("n_windows", [4, 7])
def test_flagger_fused_equals_sequence_and_host(context, command_queue, abs_mode, use_flags,
                                                n_windows):
    vis, spikes, input_flags = flagger_case(1000, 77)
    vis[300:330, :] += 2.0          # a broad weak feature for the larger windows
    fl = {rfi.BackgroundFlags.NONE: None, rfi.BackgroundFlags.CHANNEL: input_flags[:, 0].copy(),
          rfi.BackgroundFlags.FULL: input_flags}[use_flags]

    def run(fused):
        template = rfi.FlaggerDeviceTemplate(
            rfi.BackgroundMedianFilterDeviceTemplate(context, 13, use_flags=use_flags,
                                                     abs_mode=abs_mode),
            rfi.NoiseEstMADTDeviceTemplate(context, 10240),
            rfi.ThresholdSumDeviceTemplate(context, n_windows=n_windows, flag_value=2),
            fused=fused)
        fn = template.instantiate(command_queue, *vis.shape, threshold_args={"n_sigma": 7.0})
        fn.ensure_all_bound()
        fn.buffer("vis").set(command_queue, vis)
        if fl is not None:
            fn.buffer("input_flags").set(command_queue, fl)
        fn()
        return fn.buffer("flags").get(command_queue), fn.buffer("noise").get(command_queue), fn

    flags_f, noise_f, fn_f = run(True)
    flags_s, noise_s, fn_s = run(False)
    assert fn_f.fused_op is not None and fn_s.fused_op is None
    np.testing.assert_array_equal(flags_s, flags_f)
    assert same_bits(noise_s, noise_f)
    want_flags, want_dev, want_noise = contract.flagger(vis, fl, n_windows=n_windows, n_sigma=7.0,
                                                        flag_value=2, abs_mode=abs_mode)
    np.testing.assert_array_equal(want_flags, flags_f)
    assert same_bits(want_noise, noise_f)
    assert same_bits(want_dev, fn_s.buffer("deviations").get(command_queue))
    # and the reference host flagger, with the near-threshold allowance (R6)
    if abs_mode == contract.detect_abs_mode():
        near = {}
        host = hn.flagger(vis, fl, n_sigma=7.0, n_windows=n_windows, flag_value=2, near=near)
        assert int((host != flags_f).sum()) <= near.get("band", 0)


def test_flagger_rebinding_user_buffers(context, command_queue, abs_mode):
    """Bind caller-owned buffers (allocated to the slots' padded shapes) and run twice."""
    vis, spikes, _ = flagger_case(256, 40)
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 10240),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=4))
    fn = template.instantiate(command_queue, 256, 40, threshold_args={"n_sigma": 11.0})
    slot_v, slot_f = fn.slots["vis"], fn.slots["flags"]
    mine_v = DeviceArray(context, slot_v.shape, slot_v.dtype, slot_v.required_padded_shape())
    mine_f = DeviceArray(context, slot_f.shape, slot_f.dtype, slot_f.required_padded_shape())
    mine_v.set(command_queue, vis)
    fn(vis=mine_v, flags=mine_f)
    np.testing.assert_array_equal(spikes.astype(np.uint8), mine_f.get(command_queue))
    with pytest.raises(ValueError):
        fn.bind(vis=DeviceArray(context, slot_v.shape, slot_v.dtype, (256, 48)))
    with pytest.raises(TypeError):
        fn.bind(vis=DeviceArray(context, slot_v.shape, np.float32, slot_v.required_padded_shape()))


# ----------------------------------------------------------------------------- helpers
@pytest.mark.parametrize("shape", [(4, 5), (53, 7), (53, 81), (32, 64)])
@pytest.mark.parametrize("dtype,ctype", [(np.float32, "float"), (np.uint8, "unsigned char"),
                                         (np.complex64, "float2")])
def test_transpose(context, command_queue, shape, dtype, ctype):
    template = transpose.TransposeTemplate(context, dtype, ctype)
    fn = template.instantiate(command_queue, shape)
    # force non-trivial padding, as the reference's test does (test/test_transpose.py:46-51)
    fn.slots["src"].dimensions[0].link(accel.Dimension(shape[0], min_padded_round=5))
    fn.slots["src"].dimensions[1].link(accel.Dimension(shape[1], min_padded_round=7))
    fn.slots["dest"].dimensions[1].link(accel.Dimension(shape[0], min_padded_round=3))
    fn.ensure_all_bound()
    ary = np.random.RandomState(1).uniform(0, 200, shape).astype(dtype)
    fn.buffer("src").set(command_queue, ary)
    fn()
    np.testing.assert_array_equal(ary.T, fn.buffer("dest").get(command_queue))


@pytest.mark.parametrize("shape,column_range", [((4096, 1), None), ((4096, 4029), None),
                                                ((4096, 4030), (8, 4000)), ((27, 301), (0, 301))])
@pytest.mark.parametrize("is_amplitude", [True, False])
def test_percentile5(context, command_queue, abs_mode, shape, column_range, is_amplitude):
    rs = np.random.RandomState(1)
    rows, cols = min(shape[0], 128), shape[1]
    src = (np.abs(rs.standard_normal((rows, cols))).astype(np.float32) if is_amplitude
           else complex_normal(rs, (rows, cols)))
    template = percentile.Percentile5Template(context, 5000, is_amplitude, abs_mode=abs_mode)
    fn = template.instantiate(command_queue, (rows, cols), column_range)
    fn.slots["src"].dimensions[1].link(accel.Dimension(cols, min_padded_round=13))
    fn.ensure_all_bound()
    fn.buffer("src").set(command_queue, src)
    fn()
    out = fn.buffer("dest").get(command_queue)
    expected = hn.percentile5(src, column_range)
    if is_amplitude or abs_mode == contract.detect_abs_mode():
        assert same_bits(expected, out)
    else:
        np.testing.assert_allclose(expected, out, rtol=1e-6)


@pytest.mark.parametrize("cols", [2, 4029, 4032])
@pytest.mark.parametrize("use_amplitudes", [False, True])
def test_maskedsum(context, command_queue, abs_mode, cols, use_amplitudes):
    rs = np.random.RandomState(1)
    rows = 4096
    src = complex_normal(rs, (rows, cols))
    mask = np.ones(rows, np.float32)
    mask[rs.random_sample(rows) < 0.1] = 0
    template = maskedsum.MaskedSumTemplate(context, use_amplitudes, abs_mode=abs_mode)
    fn = template.instantiate(command_queue, (rows, cols))
    fn.ensure_all_bound()
    fn.buffer("src").set(command_queue, src)
    fn.buffer("mask").set(command_queue, mask)
    fn()
    out = fn.buffer("dest").get(command_queue)
    data = np.abs(src).astype(np.float64) if use_amplitudes else src.astype(np.complex128)
    expected = np.sum(data * mask[:, None], axis=0)
    scale = np.sum(np.abs(src) * mask[:, None], axis=0)
    assert np.all(np.abs(expected - out) <= 1e-6 * scale)


# ----------------------------------------------------------------------------- streaming ingest
@pytest.mark.parametrize("depth", [1, 2, 3])
@pytest.mark.parametrize("use_flags", [rfi.BackgroundFlags.NONE, rfi.BackgroundFlags.CHANNEL])
def test_streaming_flagger(context, abs_mode, depth, use_flags):
    """Dumps go through upload / compute / download queues with `depth` in flight; every
    result must equal the one-dump-at-a-time oracle, in order."""
    from katsdpsigproc_b200 import streaming

    channels, baselines = 512, 70
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, use_flags=use_flags, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 10240),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=7))
    stream = streaming.StreamingFlagger(template, channels, baselines, depth=depth,
                                        threshold_args={"n_sigma": 9.0})
    rs = np.random.RandomState(5)
    dumps, chan_flags, results = [], [], []
    for i in range(7):
        vis = complex_normal(rs, (channels, baselines))
        spikes = rs.random_sample(vis.shape) < 1 / 32
        vis += (spikes * 60.0).astype(np.complex64)
        fl = (rs.random_sample(channels) < 0.05).astype(np.uint8) if use_flags else None
        dumps.append(vis)
        chan_flags.append(fl)
        out = stream.submit(vis, fl)
        if out is not None:
            results.append(out.copy())
    results.extend(out.copy() for out in stream.drain())
    assert len(results) == len(dumps)
    for vis, fl, got in zip(dumps, chan_flags, results):
        want, _, _ = contract.flagger(vis, fl, n_windows=7, n_sigma=9.0, abs_mode=abs_mode)
        np.testing.assert_array_equal(want, got)
    with pytest.raises(TypeError):
        stream.submit(dumps[0], None if use_flags else np.zeros(channels, np.uint8))


@pytest.mark.parametrize("depth", [1, 2])
def test_streaming_flagger_zero_copy_staging(context, abs_mode, depth):
    """The zero-copy pattern: the producer writes each dump straight into host_vis() and calls
    submit(None).  host_vis() must not hand the staging array out while the upload of the dump
    it still holds is in flight (large dumps, so that an upload takes a while)."""
    from katsdpsigproc_b200 import streaming

    channels, baselines = 4096, 1024
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 10240),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=7))
    stream = streaming.StreamingFlagger(template, channels, baselines, depth=depth,
                                        threshold_args={"n_sigma": 11.0})
    rs = np.random.RandomState(6)
    base = complex_normal(rs, (channels, baselines))
    spike_sets, results = [], []
    for i in range(6):
        spikes = rs.random_sample(base.shape) < 1 / 256
        spike_sets.append(spikes)
        staging = stream.host_vis()
        staging[...] = 0                              # a different dump every time
        np.copyto(staging, base + (spikes * 80.0).astype(np.complex64))
        out = stream.submit(None)
        if out is not None:
            results.append(out.copy())
    results.extend(out.copy() for out in stream.drain())
    assert len(results) == len(spike_sets)
    for spikes, got in zip(spike_sets, results):
        np.testing.assert_array_equal(spikes, got != 0)


@pytest.mark.parametrize("depth", [1, 3])
def test_async_streaming_flagger(context, abs_mode, depth):
    """The asyncio pipeline (Resource / JobQueue / async_wait_for_events): every task resolves to
    the flags of its own dump while the producer keeps `depth` dumps in flight."""
    import asyncio

    from katsdpsigproc_b200 import resource, streaming

    channels, baselines = 384, 45
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 10240),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=5))
    rs = np.random.RandomState(11)
    dumps = []
    for i in range(6):
        vis = complex_normal(rs, (channels, baselines))
        vis += ((rs.random_sample(vis.shape) < 1 / 32) * 60.0).astype(np.complex64)
        dumps.append(vis)

    async def main():
        stream = streaming.AsyncStreamingFlagger(template, channels, baselines, depth=depth,
                                                 threshold_args={"n_sigma": 9.0})
        assert stream.depth == depth
        jobs = resource.JobQueue()
        results = [None] * len(dumps)

        async def consume(index, task):            # tasks of different buffer sets may finish in any order
            results[index] = (await task).copy()
        for index, vis in enumerate(dumps):
            jobs.add(consume(index, stream.submit(vis)))
            await jobs.finish(max_remaining=depth - 1)
        await jobs.finish()
        with pytest.raises(TypeError):
            await stream.submit(dumps[0], np.zeros(channels, np.uint8))
        # the failed submission released its buffer set: the stream still works
        again = (await stream.submit(dumps[0])).copy()
        return results, again

    results, again = asyncio.run(main())
    assert len(results) == len(dumps)
    for vis, got in zip(dumps, results):
        want, _, _ = contract.flagger(vis, None, n_windows=5, n_sigma=9.0, abs_mode=abs_mode)
        np.testing.assert_array_equal(want, got)
    np.testing.assert_array_equal(results[0], again)


def test_wrap_external_device_memory(context, command_queue, abs_mode):
    """Zero-copy hand-over: bind a torch tensor's memory as the flagger's vis buffer."""
    torch = pytest.importorskip("torch")
    vis, spikes, _ = flagger_case(256, 64)
    t = torch.from_numpy(vis.view(np.float32).reshape(256, 64, 2)).cuda()
    template = rfi.FlaggerDeviceTemplate(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 13, abs_mode=abs_mode),
        rfi.NoiseEstMADTDeviceTemplate(context, 10240),
        rfi.ThresholdSumDeviceTemplate(context, n_windows=4))
    fn = template.instantiate(command_queue, 256, 64, threshold_args={"n_sigma": 11.0})
    slot = fn.slots["vis"]
    assert slot.required_padded_shape() == (256, 64)
    torch.cuda.synchronize()
    wrapped = DeviceArray.wrap(context, t.data_ptr(), (256, 64), np.complex64, owner=t)
    fn(vis=wrapped)
    np.testing.assert_array_equal(spikes.astype(np.uint8), fn.buffer("flags").get(command_queue))


# ----------------------------------------------------------------------------- Fill / HReduce
@pytest.mark.parametrize("dtype, ctype, value", [
    (np.uint8, "unsigned char", 0xA5),
    (np.int16, "short", -1234),
    (np.float32, "float", 1.5e-3),
    (np.uint32, "unsigned int", 0xDEADBEEF),
    (np.complex64, "float2", 2.5 - 1.25j),
    (np.complex128, "double2", -3.0 + 7.0j),
])
@pytest.mark.parametrize("shape", [(75, 63), (1,), (1000003,)])
def test_fill(context, command_queue, dtype, ctype, value, shape):
    """Reference test/test_fill.py:35-56: every element INCLUDING the padding takes the value."""
    from katsdpsigproc_b200 import fill

    template = fill.FillTemplate(context, dtype, ctype)
    fn = template.instantiate(command_queue, shape)
    for dim, extra in zip(fn.slots["data"].dimensions, (5, 10)):
        accel.Dimension(dim.size, min_padded_size=dim.size + extra).link(dim)
    fn.ensure_all_bound()
    data = fn.buffer("data")
    assert all(p >= s + 5 for p, s in zip(data.padded_shape, shape))
    poison = data.empty_like()                  # so that untouched bytes are noticed
    HostArray.padded_view(poison).view(np.uint8)[...] = 0x3C
    data.set(command_queue, poison)
    fn.set_value(value)
    fn()
    ret = data.get(command_queue)
    want = np.dtype(dtype).type(value)
    whole = HostArray.padded_view(ret)
    assert whole.shape == data.padded_shape
    assert np.all(whole == want)
    assert fn.parameters()["value"] == want
    # the default value is the dtype's zero
    assert template.instantiate(command_queue, shape).value == np.dtype(dtype).type()
    fn.set_value(0)
    fn()
    assert not np.any(HostArray.padded_view(data.get(command_queue)).view(np.uint8))


@pytest.mark.parametrize("rows, columns, column_range", [
    (129, 173, (67, 128)),
    (7, 8, (1, 7)),            # narrower than a warp
    (64, 5000, None),
    (1, 1, None),
])
def test_hreduce_sum_uint32(context, command_queue, rows, columns, column_range):
    """Reference test/test_reduce.py:86-107 (uint32 'a + b', identity '0') plus wider shapes."""
    from katsdpsigproc_b200 import reduce

    template = reduce.HReduceTemplate(context, np.uint32, "unsigned int", "a + b", "0")
    fn = template.instantiate(command_queue, (rows, columns), column_range)
    fn.ensure_all_bound()
    src = fn.buffer("src").empty_like()
    rs = np.random.RandomState(1)
    src[:] = rs.randint(0, 100000, (rows, columns))
    fn.buffer("src").set(command_queue, src)
    fn()
    dest = fn.buffer("dest").get(command_queue)
    lo, hi = column_range if column_range else (0, columns)
    np.testing.assert_equal(np.sum(src[:, lo:hi], axis=1, dtype=np.uint32), dest)
    assert fn.parameters()["column_range"] == (lo, hi)


def test_hreduce_other_operators(context, command_queue):
    from katsdpsigproc_b200 import reduce

    rs = np.random.RandomState(2)
    rows, columns = 37, 1234
    data = rs.standard_normal((rows, columns)).astype(np.float32)
    # max of floats
    t = reduce.HReduceTemplate(context, np.float32, "float", "fmaxf(a, b)", "-INFINITY")
    fn = t.instantiate(command_queue, (rows, columns), (3, 1200))
    fn.ensure_all_bound()
    fn.buffer("src").set(command_queue, data)
    fn()
    np.testing.assert_array_equal(data[:, 3:1200].max(axis=1), fn.buffer("dest").get(command_queue))
    # a two-field element with helper code: (min, max) pairs, 8 bytes per element
    pairs = np.stack([data, data], axis=-1).copy().view(np.complex64)[..., 0]
    extra = ("struct mm { float lo, hi; };\n"
             "__device__ mm mm_join(mm a, mm b) { mm r; r.lo = fminf(a.lo, b.lo); "
             "r.hi = fmaxf(a.hi, b.hi); return r; }\n"
             "__device__ mm mm_id() { mm r; r.lo = INFINITY; r.hi = -INFINITY; return r; }\n")
    t2 = reduce.HReduceTemplate(context, np.complex64, "mm", "mm_join(a, b)", "mm_id()", extra)
    fn2 = t2.instantiate(command_queue, (rows, columns))
    fn2.ensure_all_bound()
    fn2.buffer("src").set(command_queue, pairs)
    fn2()
    out = fn2.buffer("dest").get(command_queue)
    np.testing.assert_array_equal(data.min(axis=1), out.real)
    np.testing.assert_array_equal(data.max(axis=1), out.imag)
    # float64 sum: tree order differs from numpy's, compare to rounding error
    d64 = rs.standard_normal((rows, columns))
    t3 = reduce.HReduceTemplate(context, np.float64, "double", "a + b", "0.0")
    fn3 = t3.instantiate(command_queue, (rows, columns))
    fn3.ensure_all_bound()
    fn3.buffer("src").set(command_queue, d64)
    fn3()
    np.testing.assert_allclose(d64.sum(axis=1), fn3.buffer("dest").get(command_queue),
                               rtol=0, atol=1e-11)


def test_hreduce_errors(context, command_queue):
    from katsdpsigproc_b200 import reduce

    with pytest.raises(RuntimeError, match="could not be built"):
        reduce.HReduceTemplate(context, np.float32, "float", "a +* b", "0")
    with pytest.raises(RuntimeError, match="sizes differ"):
        reduce.HReduceTemplate(context, np.float32, "double", "a + b", "0")
    t = reduce.HReduceTemplate(context, np.int32, "int", "a + b", "0")
    with pytest.raises(ValueError):
        t.instantiate(command_queue, (4, 5, 6))
    with pytest.raises(ValueError):
        t.instantiate(command_queue, (4, 5), (0, 6))
    with pytest.raises(ValueError):
        t.instantiate(command_queue, (4, 5), (3, 3))
