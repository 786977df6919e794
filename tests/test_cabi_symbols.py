"""The C-ABI library loads without a GPU and exports every symbol the header declares."""

import ctypes
import os
import re

import pytest

from katsdpsigproc_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ksp_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ksp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_functions()
    for needed in ("ksp_flagger", "ksp_background_median_filter", "ksp_madnz_t",
                   "ksp_threshold_sum", "ksp_transpose", "ksp_percentile5", "ksp_maskedsum"):
        assert needed in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [name for name in declared_functions() if not hasattr(lib, name)]
    assert not missing, missing


def test_python_prototypes_cover_the_header():
    assert sorted(_capi.PROTOTYPES) == declared_functions()


def test_abi_version_and_error_strings():
    lib = _capi.load()
    assert lib.ksp_abi_version() == 1
    assert _capi.error_string(0) == "success"
    assert "aligned" in _capi.error_string(-2)
    assert _capi.kernel_launch_count() == 0 or _capi.kernel_launch_count() > 0


def test_argument_errors_need_no_gpu():
    """Validation happens before any CUDA call, so these are safe on a CPU-only box."""
    lib = _capi.load()
    assert lib.ksp_transpose(None, None, None, -1, 4, 4, 4, 4) == -1
    assert lib.ksp_transpose(None, None, None, 4, 4, 4, 4, 3) in (-1,)
    assert lib.ksp_madnz_t(None, None, None, 8, 8, 4) == -1          # stride < channels
    assert lib.ksp_threshold_sum(None, None, None, None, 8, 8, 8, 8, 9, 1.0, None, 1) == -1
    with pytest.raises(_capi.KspError):
        _capi.call("ksp_percentile5", None, None, None, 4, 8, 4, 0, 0, 1, 0)


def test_struct_layout_matches_header():
    # int64 x5, int x6, double, double[KSP_MAX_WINDOWS], int64
    assert ctypes.sizeof(_capi.FlaggerParams) == 5 * 8 + 6 * 4 + 8 + _capi.MAX_WINDOWS * 8 + 8
    header = open(HEADER).read()
    assert f"#define KSP_MAX_WINDOWS {_capi.MAX_WINDOWS} " in header
