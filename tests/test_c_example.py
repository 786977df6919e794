"""The C ABI from plain C: tests/c/flagger_smoke.c is compiled against include/ksp_b200.h and
linked with the library (CPU), then run on a device and compared with the oracle (GPU)."""

import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "katsdpsigproc_b200", "_lib")
SRC = os.path.join(ROOT, "tests", "c", "flagger_smoke.c")


def build(tmp_path):
    exe = str(tmp_path / "flagger_smoke")
    subprocess.run(
        ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
         "-L", LIBDIR, "-lksp_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"],
        check=True, capture_output=True, text=True)
    return exe


def test_c_example_builds_and_links(tmp_path):
    assert os.path.exists(os.path.join(LIBDIR, "libksp_b200.so")), "run __graft_entry__.build() first"
    exe = build(tmp_path)
    # every ksp_* symbol the program needs was resolved at link time; without a device it must
    # fail cleanly (exit code 3: "no CUDA device"), not crash
    import ctypes
    lib = ctypes.CDLL(os.path.join(LIBDIR, "libksp_b200.so"))
    count = ctypes.c_int(-1)
    assert lib.ksp_device_count(ctypes.byref(count)) == 0
    if count.value == 0:
        done = subprocess.run([exe, "64", "4"], capture_output=True, text=True)
        assert done.returncode == 3, done.stderr


def xorshift_dump(channels, baselines):
    """The dump flagger_smoke.c generates (32-bit xorshift, three draws per visibility)."""
    n = channels * baselines
    state = 12345
    out = np.empty((n, 2), np.float32)
    mask = 0xFFFFFFFF
    scale = np.float32(1.0 / 16777216.0)

    def draw():
        nonlocal state
        state ^= (state << 13) & mask
        state ^= state >> 17
        state ^= (state << 5) & mask
        return np.float32(state >> 8) * scale
    half = np.float32(0.5)
    for i in range(n):
        re = draw() - half
        im = draw() - half
        if draw() < np.float32(1.0 / 64.0):
            re = np.float32(re + np.float32(20.0))
        out[i, 0] = re
        out[i, 1] = im
    return out.view(np.complex64).reshape(channels, baselines)


@pytest.mark.gpu
def test_c_example_matches_oracle(tmp_path):
    from oracle import contract

    channels, baselines = 1024, 40
    exe = build(tmp_path)
    done = subprocess.run([exe, str(channels), str(baselines)], capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    words = done.stdout.split()
    flagged, digest, noise_sum = int(words[1]), int(words[3]), float(words[5])

    vis = xorshift_dump(channels, baselines)
    flags, _, noise = contract.flagger(vis, None, n_windows=7, n_sigma=11.0,
                                       abs_mode=contract.ABS_NUMPY)
    want = 1469598103934665603
    for v in flags.ravel().tolist():
        want = ((want ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert flagged == int(flags.sum()) and flagged > 0
    assert digest == want
    assert abs(noise_sum - float(noise.astype(np.float64).sum())) < 1e-5
