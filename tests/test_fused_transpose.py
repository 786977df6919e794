"""include/ksp_transpose_base.cuh, the fusable transpose tools (reference
transpose_base.mako:34-137): a downstream kernel written with them (tests/c/fused_transpose.cu -
complex amplitude x row weight, written transposed) compiles on its own with nvcc for sm_100a
(CPU) and gives numpy's answer on a device (GPU), ragged edges and padded strides included."""

import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "fused_transpose.cu")


def build(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "fused_transpose")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "--extended-lambda",
                    "-I", os.path.join(ROOT, "include"), SRC, "-o", exe],
                   check=True, capture_output=True, text=True)
    return exe


def test_header_compiles_in_a_downstream_kernel(tmp_path):
    build(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(53, 81), (32, 64), (4, 5), (200, 333)])
def test_fused_kernel_matches_numpy(tmp_path, rows, cols):
    exe = build(tmp_path)
    done = subprocess.run([exe, str(rows), str(cols)], capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    got = np.array(done.stdout.split(), np.float32).reshape(cols, rows)
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    re = ((r * 31 + c * 17) % 23 - 11).astype(np.float32)
    im = ((r * 13 + c * 7) % 19 - 9).astype(np.float32)
    weight = np.where(np.arange(rows) % 7 == 3, 0.0, 1.0 + 0.25 * (np.arange(rows) % 3)).astype(np.float32)
    want = (weight[:, None] * np.sqrt(re * re + im * im)).T
    np.testing.assert_allclose(got, want, rtol=1e-6)
