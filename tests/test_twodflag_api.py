"""Host side of the 2-D flagger (no GPU): the constructor and the parameter conditioning of
katsdpsigproc_b200.rfi.twodflag.SumThresholdFlagger follow the reference's
(rfi/twodflag.py:951-1026), and what it hands to the C ABI is what the reference hands to numba."""

import numpy as np
import pytest

from katsdpsigproc_b200 import _capi
from katsdpsigproc_b200.rfi import twodflag


def test_constructor_conditions_parameters_like_the_reference():
    f = twodflag.SumThresholdFlagger(windows_freq=[1, 2, 4, 8], average_freq=2, spike_width_freq=10.0,
                                     time_extend=3, freq_extend=5)
    # windows scaled by the averaging, rounded up, duplicates dropped; spike width scaled
    np.testing.assert_array_equal(f.windows_freq, [1, 2, 4])
    assert f.spike_width_freq == 5.0
    assert f.time_extend.dtype == np.uint8 and int(f.freq_extend) == 5 and int(f.average_freq) == 2
    big = twodflag.SumThresholdFlagger(time_extend=300)
    assert big.time_extend.dtype == np.uint16
    with pytest.raises(ValueError):
        twodflag.SumThresholdFlagger(time_extend=-1)


def test_abi_parameters():
    f = twodflag.SumThresholdFlagger(background_iterations=3, freq_chunks=4, average_freq=3, rho=1.5,
                                     windows_time=[1, 2, 4, 8, 200], windows_freq=[1, 2, 4, 8, 64])
    p = f._params((10, 100, 7), True)
    assert (p.n_time, p.n_freq, p.n_bl, p.is_complex, p.average_freq) == (10, 100, 7, 1, 3)
    averaged = 34
    np.testing.assert_array_equal(list(p.chunk_ends)[:5], np.linspace(0, averaged, 5).astype(np.int_))
    # time windows are clipped against the channel count (as in the reference), frequency windows
    # against the averaged channel count
    assert list(p.windows_time)[:p.n_windows_time] == [1, 2, 4, 8]
    assert list(p.windows_freq)[:p.n_windows_freq] == [1, 2, 3, 22]
    for i in range(p.n_windows_freq):
        assert p.tf_freq[i] == pow(1.5, np.log2(p.windows_freq[i]))
    for e in (1, 2, 3):
        sigma = e * np.array((12.5, 10.0 / 3))
        r = (0.5 * np.sqrt(12.0 * sigma**2 / 4 + 1)).astype(np.int_)
        assert (p.r_time[e], p.r_freq[e]) == (r[0], r[1])
    lib = _capi.load()
    from ctypes import byref
    one = lib.ksp_twodflag_scratch_bytes(byref(p), 1)
    assert one > 0 and lib.ksp_twodflag_scratch_bytes(byref(p), 5) == 5 * one


def test_abi_rejects_bad_parameters():
    from ctypes import byref
    lib = _capi.load()
    f = twodflag.SumThresholdFlagger()
    p = f._params((10, 100, 7), False)
    assert lib.ksp_twodflag(None, byref(p), None, None, None, None, 0, 1) == -1       # null pointers
    p.chunk_ends[p.n_chunks] = 99                                                      # must end at the channels
    assert lib.ksp_twodflag_scratch_bytes(byref(p), 1) == 0
    p = f._params((10, 100, 7), False)
    p.windows_freq[0] = 65
    assert lib.ksp_twodflag_scratch_bytes(byref(p), 1) == 0
    with pytest.raises(ValueError):
        twodflag.SumThresholdFlagger(freq_chunks=100)._params((4, 400, 1), False)
    with pytest.raises(ValueError):
        twodflag.SumThresholdFlagger(windows_time=[128])._params((4, 400, 1), False)
