"""Template / operation surface of the flagging path, checked without a GPU.

A fake context stands in for the device, so these tests see exactly what a
caller of the reference API sees: constructor validation, slot names, shapes,
dtypes, shared strides and the wiring of ``FlaggerDevice`` (reference
``rfi/device.py:177-209,302-309,587-592,960-966,1139-1166``; ``transpose.py:143-144``;
``percentile.py:175-191``; ``maskedsum.py:132-139``).
"""

import numpy as np
import pytest

from katsdpsigproc_b200 import accel, maskedsum, percentile, transpose
from katsdpsigproc_b200.rfi import device as rfi
from test_accel import FakeContext, FakeQueue


@pytest.fixture
def context():
    return FakeContext()


@pytest.fixture
def queue(context):
    return FakeQueue(context)


def slot_table(op):
    return {name: (slot.shape, slot.dtype) for name, slot in op.slots.items()}


def test_background_template_validation(context):
    T = rfi.BackgroundMedianFilterDeviceTemplate
    assert T(context, 13, use_flags=True).use_flags is rfi.BackgroundFlags.CHANNEL
    assert T(context, 13, use_flags=False).use_flags is rfi.BackgroundFlags.NONE
    assert not rfi.BackgroundFlags.NONE and rfi.BackgroundFlags.FULL
    with pytest.raises(TypeError):
        T(context, 13, use_flags=1)
    with pytest.raises(ValueError):
        T(context, 12)
    with pytest.raises(ValueError):
        T(context, 65)
    assert T(context, 13, tuning={"wgs": 64, "csplit": 2}).tuning == {"wgs": 64, "csplit": 2}
    assert set(T.autotune(context, 13, False, rfi.BackgroundFlags.NONE)) == {"wgs", "csplit"}


@pytest.mark.parametrize("use_flags,flag_shape", [(rfi.BackgroundFlags.NONE, None),
                                                  (rfi.BackgroundFlags.CHANNEL, (417,)),
                                                  (rfi.BackgroundFlags.FULL, (417, 313))])
@pytest.mark.parametrize("is_amplitude", [False, True])
def test_background_slots(context, queue, use_flags, flag_shape, is_amplitude):
    template = rfi.BackgroundMedianFilterDeviceTemplate(context, 5, is_amplitude, use_flags)
    op = template.instantiate(queue, 417, 313)
    table = slot_table(op)
    assert table["vis"] == ((417, 313), np.float32 if is_amplitude else np.complex64)
    assert table["deviations"] == ((417, 313), np.float32)
    assert ("flags" in table) == bool(use_flags)
    if flag_shape:
        assert table["flags"] == (flag_shape, np.uint8)
    # one baseline Dimension: every 2-D slot ends up with the same row stride (in elements)
    strides = {slot.required_padded_shape()[1] for slot in op.slots.values() if len(slot.shape) == 2}
    assert len(strides) == 1 and strides.pop() >= 313
    assert op.parameters() == {"width": 5, "use_flags": use_flags.name, "channels": 417,
                               "baselines": 313}


def test_noise_templates(context, queue):
    madt = rfi.NoiseEstMADTDeviceTemplate(context, 10240)
    assert madt.transposed and not rfi.NoiseEstMADDeviceTemplate(context).transposed
    op = madt.instantiate(queue, 117, 273)
    assert slot_table(op) == {"noise": ((273,), np.float32), "deviations": ((273, 117), np.float32)}
    with pytest.raises(ValueError, match="channels exceeds max_channels"):
        madt.instantiate(queue, 10241, 4)
    op = rfi.NoiseEstMADDeviceTemplate(context).instantiate(queue, 117, 273)
    assert slot_table(op) == {"noise": ((273,), np.float32), "deviations": ((117, 273), np.float32)}


def test_noise_channel_major_long_rows_slot(context, queue):
    """From TRANSPOSE_FROM channels on, the channel-major operation owns a transposed scratch
    (it runs the baseline-major kernel behind one transposition); inside an unfused flagger the
    slot is exported under the child's name and allocated with the rest."""
    T = rfi.NoiseEstMADDeviceTemplate(context)
    short = T.instantiate(queue, rfi.NoiseEstMADDevice.TRANSPOSE_FROM - 1, 40)
    assert "scratch_t" not in short.slots
    op = T.instantiate(queue, 4096, 40)
    assert slot_table(op) == {"noise": ((40,), np.float32), "deviations": ((4096, 40), np.float32),
                              "scratch_t": ((40, 4096), np.float32)}
    template = make_flagger(context, False, False, fused=False)
    fn = template.instantiate(queue, 4096, 40, threshold_args={"n_sigma": 11.0})
    assert set(fn.slots) == {"vis", "deviations", "noise", "flags", "noise_est:scratch_t"}
    fn.ensure_all_bound()
    assert fn.noise_est.buffer("scratch_t").shape == (40, 4096)


def test_threshold_templates(context, queue):
    with pytest.raises(ValueError):
        rfi.ThresholdSumDeviceTemplate(context, n_windows=12)
    with pytest.raises(ValueError):
        rfi.ThresholdSumDeviceTemplate(context, n_windows=0)
    assert rfi.ThresholdSumDeviceTemplate(context, n_windows=11).n_windows == 11
    template = rfi.ThresholdSumDeviceTemplate(context, n_windows=3, flag_value=5)
    assert template.transposed
    op = template.instantiate(queue, 117, 273, 11.0, threshold_falloff=1.5)
    assert slot_table(op) == {"deviations": ((273, 117), np.float32), "noise": ((273,), np.float32),
                              "flags": ((273, 117), np.uint8)}
    assert op.slots["deviations"].dimensions[1] is op.slots["flags"].dimensions[1]
    assert op.slots["deviations"].required_padded_shape() == op.slots["flags"].required_padded_shape()
    params = op.parameters()
    assert params["flag_value"] == 5
    assert params["n_sigma"] == [np.float32(11.0), np.float32(11.0 / 1.5), np.float32(11.0 / 2.25)]
    with pytest.raises(TypeError):
        template.instantiate(queue, 117, 273)          # n_sigma has no default

    for transposed, shape in ((False, (117, 273)), (True, (273, 117))):
        op = rfi.ThresholdSimpleDeviceTemplate(context, transposed).instantiate(queue, 117, 273, 11.0)
        assert slot_table(op) == {"deviations": (shape, np.float32), "noise": ((273,), np.float32),
                                  "flags": (shape, np.uint8)}
        assert op.transposed == transposed


def make_flagger(context, noise_t, threshold_sum, use_flags=rfi.BackgroundFlags.NONE, fused=None):
    background = rfi.BackgroundMedianFilterDeviceTemplate(context, 13, use_flags=use_flags)
    noise = (rfi.NoiseEstMADTDeviceTemplate(context, 10240) if noise_t
             else rfi.NoiseEstMADDeviceTemplate(context))
    threshold = (rfi.ThresholdSumDeviceTemplate(context, n_windows=4) if threshold_sum
                 else rfi.ThresholdSimpleDeviceTemplate(context, transposed=False))
    return rfi.FlaggerDeviceTemplate(background, noise, threshold, fused=fused)


@pytest.mark.parametrize("noise_t", [False, True])
@pytest.mark.parametrize("threshold_sum", [False, True])
def test_flagger_sequence_slots(context, queue, noise_t, threshold_sum):
    template = make_flagger(context, noise_t, threshold_sum, rfi.BackgroundFlags.FULL, fused=False)
    fn = template.instantiate(queue, 117, 131, threshold_args={"n_sigma": 11.0})
    expected = {"vis", "input_flags", "deviations", "noise", "flags"}
    if noise_t or threshold_sum:
        expected.add("deviations_t")
    if threshold_sum:
        expected.add("flags_t")
    assert set(fn.slots) == expected
    names = list(fn.operations)
    assert names[0] == "background" and "noise_est" in names and "threshold" in names
    assert ("transpose_deviations" in names) == (noise_t or threshold_sum)
    assert ("transpose_flags" in names) == threshold_sum
    assert fn.slots["vis"].dtype == np.complex64 and fn.slots["flags"].shape == (117, 131)
    if threshold_sum:
        assert fn.slots["flags_t"].shape == (131, 117)
        # deviations_t and flags_t share the channel Dimension of the threshold operation
        assert (fn.slots["deviations_t"].required_padded_shape()
                == fn.slots["flags_t"].required_padded_shape())
    fn.ensure_all_bound()
    assert fn.background.buffer("deviations") is fn.buffer("deviations")
    assert fn.noise_est.buffer("noise") is fn.threshold.buffer("noise") is fn.buffer("noise")
    assert fn.background.buffer("flags") is fn.buffer("input_flags")


def test_flagger_fused_slots(context, queue):
    template = make_flagger(context, True, True, rfi.BackgroundFlags.CHANNEL)
    assert template.fused
    fn = template.instantiate(queue, 1024, 96, threshold_args={"n_sigma": 11.0})
    assert set(fn.slots) == {"vis", "input_flags", "noise", "flags", "scratch"}
    assert list(fn.operations) == ["fused"]
    assert fn.slots["input_flags"].shape == (1024,)
    assert fn.slots["flags"].shape == (1024, 96) and fn.slots["flags"].dtype == np.uint8
    assert fn.slots["scratch"].shape[0] >= 96 * 1024 * 4
    assert fn.parameters()["fused"]
    assert fn.fused_op.parameters()["n_windows"] == 4
    # stage operations stay reachable for introspection
    assert fn.threshold.parameters()["channels"] == 1024
    with pytest.raises(ValueError):
        make_flagger(context, True, False, fused=True)
    assert not make_flagger(context, True, False).fused


def test_flagger_requires_one_context(context):
    background = rfi.BackgroundMedianFilterDeviceTemplate(context, 13)
    noise = rfi.NoiseEstMADTDeviceTemplate(FakeContext(), 10240)
    threshold = rfi.ThresholdSumDeviceTemplate(context)
    with pytest.raises(AssertionError):
        rfi.FlaggerDeviceTemplate(background, noise, threshold)


def test_host_wrappers_check_flag_arguments(context, queue):
    vis = np.zeros((8, 4), np.complex64)
    wrapper = rfi.BackgroundHostFromDevice(rfi.BackgroundMedianFilterDeviceTemplate(context, 5), queue)
    with pytest.raises(TypeError, match="flags were provided"):
        wrapper(vis, np.zeros(8, np.uint8))
    wrapper = rfi.BackgroundHostFromDevice(
        rfi.BackgroundMedianFilterDeviceTemplate(context, 5, use_flags=True), queue)
    with pytest.raises(TypeError, match="flags were expected"):
        wrapper(vis)
    flagger = rfi.FlaggerHostFromDevice(make_flagger(context, True, True), queue,
                                        threshold_args={"n_sigma": 11.0})
    with pytest.raises(TypeError, match="channel flags were provided"):
        flagger(vis, np.zeros(8, np.uint8))


def test_transpose_template(context, queue):
    template = transpose.TransposeTemplate(context, np.float32, "float")
    op = template.instantiate(queue, (53, 81))
    assert slot_table(op) == {"src": ((53, 81), np.float32), "dest": ((81, 53), np.float32)}
    assert op.parameters()["shape"] == (53, 81)
    with pytest.raises(ValueError):
        transpose.TransposeTemplate(context, np.dtype([("a", np.uint8, 3)]), "rgb")


def test_percentile_template(context, queue):
    template = percentile.Percentile5Template(context, 5000, is_amplitude=False)
    op = template.instantiate(queue, (27, 301), column_range=(8, 280))
    assert slot_table(op) == {"src": ((27, 301), np.complex64), "dest": ((5, 27), np.float32)}
    assert op.parameters()["column_range"] == (8, 280)
    with pytest.raises(ValueError, match="empty"):
        template.instantiate(queue, (27, 301), column_range=(8, 8))
    with pytest.raises(IndexError):
        template.instantiate(queue, (27, 301), column_range=(-1, 8))
    with pytest.raises(IndexError):
        template.instantiate(queue, (27, 301), column_range=(0, 302))
    with pytest.raises(ValueError, match="max_columns"):
        percentile.Percentile5Template(context, 100).instantiate(queue, (27, 301))


def test_maskedsum_template(context, queue):
    op = maskedsum.MaskedSumTemplate(context).instantiate(queue, (4096, 4029))
    assert slot_table(op) == {"src": ((4096, 4029), np.complex64), "mask": ((4096,), np.float32),
                              "dest": ((4029,), np.complex64)}
    op = maskedsum.MaskedSumTemplate(context, use_amplitudes=True).instantiate(queue, (16, 2))
    assert op.slots["dest"].dtype == np.float32


def test_missing_library_is_loud(monkeypatch):
    from katsdpsigproc_b200 import _capi

    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libksp_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _capi.load()


def test_backend_implements_the_abstract_interfaces():
    from katsdpsigproc_b200 import abc, cuda

    for concrete, interface in ((cuda.Event, abc.AbstractEvent), (cuda.Device, abc.AbstractDevice),
                                (cuda.Context, abc.AbstractContext),
                                (cuda.CommandQueue, abc.AbstractCommandQueue)):
        assert issubclass(concrete, interface)
        assert not getattr(concrete, "__abstractmethods__", None), concrete.__abstractmethods__


def test_host_class_resolves_to_the_reference(context):
    """``Template.host_class`` (reference rfi/device.py:174,380,504,679,834): the reference's own
    host classes when katsdpsigproc.rfi.host is importable, ImportError (not a silent CPU
    substitute) when it is not."""
    import sys

    import oracle
    templates = [
        (rfi.BackgroundMedianFilterDeviceTemplate(context, 13), "BackgroundMedianFilterHost"),
        (rfi.NoiseEstMADDeviceTemplate(context), "NoiseEstMADHost"),
        (rfi.NoiseEstMADTDeviceTemplate(context, 4096), "NoiseEstMADHost"),
        (rfi.ThresholdSimpleDeviceTemplate(context, False), "ThresholdSimpleHost"),
        (rfi.ThresholdSumDeviceTemplate(context), "ThresholdSumHost"),
    ]
    ref = oracle.reference_host()
    if ref is not None:
        for template, name in templates:
            assert template.host_class is getattr(ref, name)
            assert type(template).host_class is getattr(ref, name)
        assert templates[0][0].host_class(5, True)(np.ones((7, 2), np.float32)).shape == (7, 2)
    else:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.startswith("katsdpsigproc.")
                 or k == "katsdpsigproc"}
        try:
            with pytest.raises(ImportError):
                templates[0][0].host_class
        finally:
            sys.modules.update(saved)
