"""The autotuning cache: same behaviour and same database layout as the reference's ``tune.py``
(cases of reference ``test/test_tune.py:30-250``: parallel search, failing candidates, the cache
key, exact / nearest matching, the stub and force replacements), plus the layout itself and the
environment variables.  No GPU needed: devices are stand-ins, the database is in memory."""

import sqlite3
import threading
from types import SimpleNamespace
from unittest import mock

import numpy as np
import pytest

from katsdpsigproc_b200 import tune


def make_context(name="dev", platform="plat", version="1.0"):
    return SimpleNamespace(device=SimpleNamespace(name=name, platform_name=platform,
                                                  driver_version=version))


# ------------------------------------------------------------------ autotune(): the search
def test_search_tries_every_combination_and_keeps_the_best():
    seen, lock = [], threading.Lock()

    def generate(a, b):
        with lock:
            seen.append((a, b))
        return lambda iters: a * b

    best = tune.autotune(generate, time_limit=0.001, a=[1, 2], b=[7, 3])
    assert sorted(seen) == [(1, 3), (1, 7), (2, 3), (2, 7)]
    assert best == {"a": 1, "b": 3}


def test_search_of_nothing_is_an_error():
    with pytest.raises(ValueError):
        tune.autotune(lambda x, y: lambda iters: 0, x=[1, 2], y=[])


class Boom(RuntimeError):
    pass


def test_search_skips_candidates_that_fail_or_decline():
    def generate(x):
        if x == 1:
            raise Boom("x = 1")
        if x == 4:
            return None                                   # "not suitable", without the noise

        def measure(iters):
            if x == 3:
                raise Boom("x = 3")
            return -x
        return measure

    assert tune.autotune(generate, x=[0, 1, 2, 3, 4]) == {"x": 2}


def always_fails(x):
    raise Boom(f"x = {x}")


def test_search_reraises_the_last_failure_when_all_fail():
    with pytest.raises(Boom, match=r"^x = 3$") as info:
        tune.autotune(always_fails, x=[1, 2, 3])
    assert info.traceback[-1].name == "always_fails"      # raised where it happened


def test_make_measure_averages_over_the_tuning_queue():
    queue = mock.Mock()
    queue.stop_tuning.return_value = 6.0
    calls = []
    measure = tune.make_measure(queue, lambda: calls.append(1))
    assert measure(3) == 2.0 and len(calls) == 3
    queue.start_tuning.assert_called_once_with()


# ------------------------------------------------------------------ @autotuner: the cache
@pytest.fixture(autouse=True)
def real_autotuner(monkeypatch):
    """conftest stubs autotuner_impl for every test; these tests are about the real one."""
    monkeypatch.setattr(tune, "autotuner_impl", REAL_IMPL)


REAL_IMPL = tune.autotuner_impl


@pytest.fixture
def conn(monkeypatch):
    """An in-memory database that stays open across calls."""
    c = sqlite3.connect(":memory:")
    monkeypatch.setattr(tune, "_open_db", lambda: c)
    monkeypatch.setattr(tune, "_close_db", lambda conn: None)
    yield c
    c.close()


class Tuned:
    autotune_version = 3
    work = mock.Mock()

    @classmethod
    @tune.autotuner(test={"a": 3, "b": -1})
    def autotune(cls, context, param):
        return cls.work(context, param)

    @classmethod
    @tune.autotuner(test={"a": 3, "b": -1})
    def autotune_no_args(cls, context):
        return cls.work(context)

    @classmethod
    @tune.autotuner(test={"x": 0})
    def autotune_typed(cls, context, dtype, kind, flag=False):
        return cls.work(context, dtype, kind, flag)


@pytest.fixture(autouse=True)
def fresh_work():
    Tuned.work.reset_mock(return_value=True, side_effect=True)


def test_cache_hit_miss_and_matching(conn, monkeypatch):
    monkeypatch.setattr(tune, "KATSDPSIGPROC_TUNE_MATCH", "exact")
    one, two, four = {"a": 1, "b": 2}, {"a": 3, "b": 4}, {"a": 7, "b": 8}
    ctx = make_context("mock device", "mock platform", "mock version")
    Tuned.work.return_value = one
    assert Tuned.autotune(ctx, "xyz") == one
    assert Tuned.autotune(ctx, "xyz") == one
    Tuned.work.assert_called_once_with(ctx, "xyz")          # the second call came from the cache
    other = make_context("another device", "another platform", "another version")
    Tuned.work.return_value = two
    assert Tuned.autotune(other, "xyz") == two              # different device: tuned again

    monkeypatch.setattr(tune, "KATSDPSIGPROC_TUNE_MATCH", "nearest")
    Tuned.work.return_value = four
    assert Tuned.autotune(ctx, "zzz") == four               # different argument: tuned again
    Tuned.work.side_effect = RuntimeError                   # from here on everything is cached
    assert Tuned.autotune(make_context("mock device", "mock platform", "abc"), "xyz") == one
    assert Tuned.autotune(make_context("another device", "abc", "abc"), "xyz") == two
    assert Tuned.autotune(ctx, "xyz") == one
    assert Tuned.autotune(make_context("z", "y", "x"), "xyz") in (one, two)   # any record of that key


def test_exact_matching_does_not_fall_back(conn, monkeypatch):
    monkeypatch.setattr(tune, "KATSDPSIGPROC_TUNE_MATCH", "exact")
    Tuned.work.return_value = {"a": 1, "b": 2}
    Tuned.autotune(make_context(version="1"), "p")
    Tuned.work.return_value = {"a": 5, "b": 6}
    assert Tuned.autotune(make_context(version="2"), "p") == {"a": 5, "b": 6}
    assert Tuned.work.call_count == 2


def test_function_without_arguments(conn):
    Tuned.work.return_value = {"a": 1, "b": 2}
    ctx = make_context()
    assert Tuned.autotune_no_args(ctx) == {"a": 1, "b": 2}
    assert Tuned.autotune_no_args(ctx) == {"a": 1, "b": 2}
    Tuned.work.assert_called_once_with(ctx)


def test_database_layout_is_the_references(conn):
    """Table ``<module>_<class>_<function>__<version>``; key columns ``arg_*`` and ``device_*``
    (types and dtypes by repr, enums by name); result columns ``value_*``; the keys are the
    primary key with ON CONFLICT REPLACE."""
    import enum

    class Kind(enum.Enum):
        FULL = 2

    Tuned.work.return_value = {"x": 17}
    Tuned.autotune_typed(make_context("B200", "CUDA", "12.9"), np.dtype(np.float32), Kind.FULL)
    table = "test_tune_Tuned_autotune_typed__3"
    cols = [r[1] for r in conn.execute(f"PRAGMA table_info({table})")]
    assert cols == ["arg_dtype", "arg_kind", "arg_flag", "device_name", "device_platform",
                    "device_version", "value_x"]
    row = conn.execute(f"SELECT * FROM {table}").fetchone()
    assert tuple(row) == ("dtype('float32')", "FULL", 0, "B200", "CUDA", "12.9", 17)
    sql = conn.execute("SELECT sql FROM sqlite_master WHERE name=?", (table,)).fetchone()[0]
    assert "PRIMARY KEY (arg_dtype, arg_kind, arg_flag, device_name, device_platform, device_version)" in sql
    assert "ON CONFLICT REPLACE" in sql
    # a second result for the same key replaces the first
    keys = {"arg_dtype": "dtype('float32')", "arg_kind": "FULL", "arg_flag": 0, "device_name": "B200",
            "device_platform": "CUDA", "device_version": "12.9"}
    tune._save(conn, table, keys, {"value_x": 18})
    assert [tuple(r) for r in conn.execute(f"SELECT value_x FROM {table}")] == [(18,)]


def test_stub_and_force_replacements(monkeypatch):
    ctx = make_context()
    Tuned.work.return_value = {"a": 9, "b": 9}
    monkeypatch.setattr(tune, "autotuner_impl", tune.stub_autotuner)
    assert Tuned.autotune(ctx, "p") == {"a": 3, "b": -1}       # the decorator's test= value
    Tuned.work.assert_not_called()
    monkeypatch.setattr(tune, "autotuner_impl", tune.force_autotuner)
    assert Tuned.autotune(ctx, "p") == {"a": 9, "b": 9}
    assert Tuned.autotune(ctx, "p") == {"a": 9, "b": 9}
    assert Tuned.work.call_count == 2                          # no caching


def test_database_location(monkeypatch, tmp_path):
    path = tmp_path / "sub" / "tuning.db"
    path.parent.mkdir()
    monkeypatch.setenv("KATSDPSIGPROC_TUNE_DB", str(path))
    Tuned.work.return_value = {"a": 1, "b": 2}
    Tuned.autotune(make_context(), "p")
    assert path.exists()
    with sqlite3.connect(str(path)) as c:
        assert c.execute("SELECT value_a, value_b FROM test_tune_Tuned_autotune__3").fetchall() == [(1, 2)]
    monkeypatch.delenv("KATSDPSIGPROC_TUNE_DB")
    monkeypatch.setenv("XDG_CACHE_HOME", str(tmp_path / "cache"))
    Tuned.autotune(make_context(), "p")
    assert (tmp_path / "cache" / "katsdpsigproc" / "tuning.db").exists()


def test_unusable_database_is_not_fatal(monkeypatch, tmp_path):
    monkeypatch.setenv("KATSDPSIGPROC_TUNE_DB", str(tmp_path / "no" / "such" / "dir" / "t.db"))
    Tuned.work.return_value = {"a": 4, "b": 5}
    assert Tuned.autotune(make_context(), "p") == {"a": 4, "b": 5}


def test_templates_tune_through_the_cache(conn):
    """The package's own templates: reference signatures, answers cached per device."""
    from katsdpsigproc_b200 import transpose
    from katsdpsigproc_b200.rfi import device as rfi

    ctx = make_context("NVIDIA B200", "CUDA", "12090")
    first = rfi.BackgroundMedianFilterDeviceTemplate.autotune(ctx, 13, False, rfi.BackgroundFlags.NONE)
    assert set(first) == {"wgs", "csplit"}
    tables = {r[0] for r in conn.execute("SELECT name FROM sqlite_master WHERE type='table'")}
    assert "katsdpsigproc_b200_rfi_device_BackgroundMedianFilterDeviceTemplate_autotune__1" in tables
    with mock.patch.object(rfi.BackgroundMedianFilterDeviceTemplate, "_TUNING", {"wgs": 1, "csplit": 1}):
        again = rfi.BackgroundMedianFilterDeviceTemplate.autotune(ctx, 13, False, rfi.BackgroundFlags.NONE)
        fresh = rfi.BackgroundMedianFilterDeviceTemplate.autotune(ctx, 15, False, rfi.BackgroundFlags.NONE)
    assert again == first                                   # from the database
    assert fresh == {"wgs": 1, "csplit": 1}                  # a new key is computed
    assert set(transpose.TransposeTemplate.autotune(ctx, np.float32, "float")) == {"block", "vtx", "vty"}
