"""The pytest plugin (reference ``pytest_plugin.py:30-131``): a tiny suite written the way a
downstream project writes its tests is run in a sub-process with the plugin loaded."""

import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SUITE = textwrap.dedent('''
    import pytest
    from katsdpsigproc_b200 import tune


    class T:
        @classmethod
        @tune.autotuner(test={"wgs": 7})
        def autotune(cls, context):
            return {"wgs": 99}


    def test_stubbed(patch_autotune):
        assert T.autotune(object()) == {"wgs": 7}


    @pytest.mark.force_autotune
    def test_forced(patch_autotune):
        assert T.autotune(object()) == {"wgs": 99}


    def test_on_a_device(context, command_queue):
        assert command_queue.context is context
        assert context.device.is_cuda and context.device.platform_name


    @pytest.mark.opencl_only
    def test_opencl_only(device):
        raise AssertionError("there is no OpenCL device")


    @pytest.mark.cuda_only(min_compute_capability=(99, 0))
    def test_too_new(device):
        raise AssertionError("no such device")
''')


def run(tmp_path, *args):
    (tmp_path / "test_downstream.py").write_text(SUITE)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, "-m", "pytest", "-p", "katsdpsigproc_b200.pytest_plugin",
                           "-q", "-rA", "-p", "no:cacheprovider", str(tmp_path), *args],
                          capture_output=True, text=True, env=env, cwd=str(tmp_path))


def test_fixtures_markers_and_option(tmp_path):
    out = run(tmp_path)
    text = out.stdout + out.stderr
    assert "PASSED test_downstream.py::test_stubbed" in text, text
    assert "PASSED test_downstream.py::test_forced" in text, text
    # with a GPU the device test runs; without one it is reported as xfail and not run
    assert ("PASSED test_downstream.py::test_on_a_device" in text
            or "XFAIL test_downstream.py::test_on_a_device" in text), text
    assert "XFAIL test_downstream.py::test_opencl_only" in text, text
    assert "XFAIL test_downstream.py::test_too_new" in text, text
    assert out.returncode == 0, text
    out = run(tmp_path, "--devices=none")
    text = out.stdout + out.stderr
    assert "SKIPPED" in text and "--devices=none passed on command line" in text, text
    assert out.returncode == 0, text
