"""Autotuning cache with the reference's interface and database layout.

Mirrors reference ``tune.py:75-81,92-129,132-241,254-334,337-442`` from the caller's side:

* ``@autotuner(test={...})`` on a ``classmethod autotune(cls, context, <named args>)`` caches its
  result in an sqlite database - ``$KATSDPSIGPROC_TUNE_DB`` or ``<user cache>/katsdpsigproc/
  tuning.db`` - in a table named ``<module>_<class>_<function>__<autotune_version>`` whose key
  columns are ``arg_<name>`` for every argument after ``cls, context`` plus ``device_name``,
  ``device_platform`` and ``device_version``, and whose result columns are ``value_<key>``:
  the layout the reference writes, so one database file serves both packages side by side;
* ``KATSDPSIGPROC_TUNE_MATCH=nearest`` retries a miss ignoring, in turn, the driver version, the
  platform and the device name;
* :func:`autotuner_impl` is the patch point; :func:`stub_autotuner` (return the ``test``
  value, no database) and :func:`force_autotuner` (always run the function) are the two
  replacements the reference's pytest plugin installs (``pytest_plugin.py:30-35``);
* :func:`autotune` runs the Cartesian product of candidate values through a ``generate``
  callback and returns the best-scoring combination; :func:`make_measure` times a callable on
  a tuning command queue.

The kernels of this package fix their launch geometry for sm_100a inside the library, so the
``autotune`` classmethods here have nothing to search: they return at once and the cache
stores that answer.  The machinery is complete all the same - a downstream operation built on
this runtime can tune its own parameters with it exactly as it would with the reference's.

Unlike the reference, a database that cannot be opened or written (read-only home, locked
file) is not fatal: the tuning function is simply run.
"""

from __future__ import annotations

import concurrent.futures
import enum
import functools
import inspect
import itertools
import logging
import multiprocessing
import os
import sqlite3
import time
from typing import Any, Callable, Dict, Mapping, Optional, Sequence

import numpy as np

_logger = logging.getLogger(__name__)

KATSDPSIGPROC_TUNE_MATCH = os.getenv("KATSDPSIGPROC_TUNE_MATCH", "exact")
if KATSDPSIGPROC_TUNE_MATCH not in ("exact", "nearest"):
    _logger.debug("KATSDPSIGPROC_TUNE_MATCH is neither 'exact' nor 'nearest': using 'exact'")
    KATSDPSIGPROC_TUNE_MATCH = "exact"

_DEVICE_KEYS = ("device_version", "device_platform", "device_name")   # order in which "nearest" drops them


def adapt_value(value: Any) -> Any:
    """A lookup-key value in a form sqlite can store: types and dtypes by ``repr``, enums by name."""
    if isinstance(value, (type, np.dtype)):
        return repr(value)
    if isinstance(value, enum.Enum):
        return value.name
    return value


def _db_keys(fn: Callable[..., Any], args: Sequence[Any], kwargs: Mapping[str, Any]) -> Dict[str, Any]:
    """Database key of one call of a tuning function: its arguments by name (all but ``cls`` and
    ``context``) and the device the context is on."""
    bound = inspect.signature(fn).bind(*args, **kwargs)
    bound.apply_defaults()
    keys = {"arg_" + name: adapt_value(value) for name, value in list(bound.arguments.items())[2:]}
    device = args[1].device
    keys["device_name"] = device.name
    keys["device_platform"] = device.platform_name
    keys["device_version"] = device.driver_version
    return keys


def _query(conn: sqlite3.Connection, tablename: str, keys: Mapping[str, Any]) -> Optional[sqlite3.Row]:
    where = " AND ".join(f"{key}=?" for key in keys)
    sql = f"SELECT * FROM {tablename}" + (f" WHERE {where}" if where else "")
    return conn.cursor().execute(sql, list(keys.values())).fetchone()


def _fetch(conn: sqlite3.Connection, tablename: str, keys: Mapping[str, Any]
           ) -> Optional[Mapping[str, Any]]:
    """The cached record for ``keys``, or - with ``KATSDPSIGPROC_TUNE_MATCH=nearest`` - the nearest
    one; ``None`` if there is none (or no such table yet)."""
    try:
        row = _query(conn, tablename, keys)
        if row is None and KATSDPSIGPROC_TUNE_MATCH == "nearest":
            partial = dict(keys)
            for dropped in _DEVICE_KEYS:
                _logger.debug("Retrying the tuning query without %s", dropped)
                partial.pop(dropped, None)
                try:
                    row = _query(conn, tablename, partial)
                except sqlite3.Error:
                    _logger.debug("Query failed", exc_info=True)
                if row is not None:
                    break
        if row is None:
            return None
        return {name[len("value_"):]: row[name] for name in row.keys() if name.startswith("value_")}
    except sqlite3.Error:
        _logger.debug("Query failed", exc_info=True)       # e.g. the table does not exist yet
        return None


def _create_table(conn: sqlite3.Connection, tablename: str, keys: Mapping[str, Any],
                  values: Mapping[str, Any]) -> None:
    columns = ", ".join(f"{name} NOT NULL" for name in itertools.chain(keys, values))
    conn.execute(f"CREATE TABLE IF NOT EXISTS {tablename} ({columns}, "
                 f"PRIMARY KEY ({', '.join(keys)}) ON CONFLICT REPLACE)")


def _save(conn: sqlite3.Connection, tablename: str, keys: Mapping[str, Any],
          values: Mapping[str, Any]) -> None:
    """Store one result, creating the table on first use."""
    _create_table(conn, tablename, keys, values)
    entries = {**keys, **values}
    marks = ", ".join("?" for _ in entries)
    with conn:                                             # one transaction
        conn.execute(f"INSERT OR REPLACE INTO {tablename}({', '.join(entries)}) VALUES ({marks})",
                     list(entries.values()))


def _user_cache_dir() -> str:
    base = os.getenv("XDG_CACHE_HOME") or os.path.join(os.path.expanduser("~"), ".cache")
    return os.path.join(base, "katsdpsigproc")


def _open_db() -> sqlite3.Connection:
    cache_file = os.getenv("KATSDPSIGPROC_TUNE_DB")
    if cache_file is None:
        cache_dir = _user_cache_dir()
        os.makedirs(cache_dir, exist_ok=True)
        cache_file = os.path.join(cache_dir, "tuning.db")
    return sqlite3.connect(cache_file)


def _close_db(conn: sqlite3.Connection) -> None:
    """Separate so that tests can keep an in-memory database open across calls."""
    conn.close()


def autotuner_impl(test: Mapping[str, Any], fn: Callable[..., Mapping[str, Any]], *args: Any,
                   **kwargs: Any) -> Mapping[str, Any]:
    """What :func:`autotuner` does (a function of its own so that it can be patched): look the
    call up in the database, run ``fn`` and store its answer on a miss."""
    cls = args[0]
    classname = f"{cls.__module__}.{cls.__name__}.{fn.__name__}"
    tablename = classname.replace(".", "_") + "__" + str(getattr(cls, "autotune_version", 0))
    keys = _db_keys(fn, args, kwargs)
    try:
        conn = _open_db()
    except (sqlite3.Error, OSError):
        _logger.warning("Tuning database unavailable; tuning %s without it", classname, exc_info=True)
        return fn(*args, **kwargs)
    conn.row_factory = sqlite3.Row
    try:
        ans = _fetch(conn, tablename, keys)
        if ans is None:
            _logger.info("Performing autotuning for %s with key %s", classname, keys)
            ans = fn(*args, **kwargs)
            try:
                _save(conn, tablename, keys, {"value_" + key: value for key, value in ans.items()})
            except (sqlite3.Error, OSError):
                _logger.warning("Could not store the tuning result of %s", classname, exc_info=True)
        else:
            _logger.debug("Autotuning cache hit for %s with key %s", classname, keys)
    finally:
        _close_db(conn)
    return ans


def autotuner(test: Mapping[str, Any]) -> Callable[[Callable[..., Any]], Callable[..., Any]]:
    r"""Decorator: make ``fn(cls, context, <named arguments>)`` a cached tuning function.

    The arguments after ``context`` form the cache key together with the device and the
    function's qualified name, so each must have a name (no ``*args``).  ``test`` is what
    :func:`stub_autotuner` returns in its place.
    """
    def decorate(fn: Callable[..., Any]) -> Callable[..., Any]:
        @functools.wraps(fn)
        def wrapper(*args: Any, **kwargs: Any) -> Mapping[str, Any]:
            # looked up at call time: tests patch the module attribute
            return autotuner_impl(test, fn, *args, **kwargs)
        return wrapper
    return decorate


def force_autotuner(test: Mapping[str, Any], fn: Callable[..., Mapping[str, Any]], *args: Any,
                    **kwargs: Any) -> Mapping[str, Any]:
    """Replacement for :func:`autotuner_impl`: always run the tuning function, no database."""
    return fn(*args, **kwargs)


def stub_autotuner(test: Mapping[str, Any], fn: Callable[..., Mapping[str, Any]], *args: Any,
                   **kwargs: Any) -> Mapping[str, Any]:
    """Replacement for :func:`autotuner_impl`: return the ``test`` value, tune nothing."""
    return test


def make_measure(queue: Any, function: Callable[[], None]) -> Callable[[int], float]:
    """Scoring function for :func:`autotune`: mean seconds per call of ``function``, measured
    by the tuning command queue (``start_tuning`` / ``stop_tuning``)."""
    def measure(iters: int) -> float:
        queue.start_tuning()
        for _ in range(iters):
            function()
        return queue.stop_tuning() / iters
    return measure


def autotune(generate: Callable[..., Optional[Callable[[int], float]]], time_limit: float = 0.1,
             threads: Optional[int] = None, **kwargs: Any) -> Mapping[str, Any]:
    """Try every combination of the iterables in ``kwargs`` and return the best one.

    ``generate(**combination)`` returns a scoring function ``score(iterations) -> float`` (lower
    is better) or ``None`` to skip the combination; exceptions from either are swallowed unless
    every combination fails, in which case the last one is re-raised.  Each candidate gets a
    warm-up call, a timing call to size the run, then about ``time_limit`` seconds.  Raises
    ``ValueError`` if there is nothing to try.
    """
    names = list(kwargs)
    combos = itertools.product(*kwargs.values())
    if threads is None:
        try:
            threads = multiprocessing.cpu_count()
        except NotImplementedError:
            threads = 1
    best: Optional[Dict[str, Any]] = None
    best_score: Optional[float] = None
    last_error: Optional[Exception] = None
    with concurrent.futures.ThreadPoolExecutor(threads) as pool:
        while True:
            batch = [dict(zip(names, combo)) for combo in itertools.islice(combos, threads)]
            if not batch:
                break
            futures = [pool.submit(generate, **keywords) for keywords in batch]
            concurrent.futures.wait(futures)
            for keywords, future in zip(batch, futures):
                try:
                    measure = future.result()
                    if measure is None:
                        continue
                    measure(1)                                   # warm-up
                    start = time.time()
                    measure(1)
                    elapsed = max(time.time() - start, 1e-4)
                    iters = max(3, int(time_limit / elapsed))
                    score = measure(iters)
                    _logger.debug("Configuration %s scored %f in %d iterations", keywords, score, iters)
                    if best_score is None or score < best_score:
                        best, best_score = keywords, score
                except Exception as exc:                         # noqa: BLE001 - as the reference
                    last_error = exc
                    _logger.debug("Exception while testing configuration %s", keywords, exc_info=True)
    if best is None:
        if last_error is not None:
            raise last_error
        raise ValueError("No options to test")
    return best
