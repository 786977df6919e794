"""Small helpers shared by the operation classes: argument marshalling for the C ABI."""

from __future__ import annotations

from ctypes import c_void_p
from typing import Any, Mapping, Optional

from . import _capi


def ptr(array: Any) -> c_void_p:
    """Device pointer of a ``DeviceArray`` as ``void *``."""
    return c_void_p(array.buffer.ptr)


def stream(command_queue: Any) -> c_void_p:
    return c_void_p(command_queue.stream)


def launch(command_queue: Any, name: str, *args: Any) -> None:
    """Make the queue's device current and enqueue one C-ABI kernel on its stream."""
    command_queue.context._make_current()
    _capi.call(name, stream(command_queue), *args)


class FixedTuning:
    """Mixin for templates: the reference's ``tuning=`` / ``autotune`` surface.

    The reference searches work-group shapes at run time and caches them in sqlite
    (``tune.py:254-334``).  These kernels target one chip (sm_100a) and pick their launch
    geometry inside the library from the problem size and the SM count, so every template's
    ``autotune`` classmethod - same name, same arguments as the reference's, decorated with
    :func:`katsdpsigproc_b200.tune.autotuner` and therefore cached in the same database layout -
    returns at once with a fixed dictionary, and a user-supplied ``tuning`` mapping is accepted
    and kept (``template.tuning``) but has no effect on the launch.
    """

    autotune_version = 1
    _TUNING: Mapping[str, Any] = {}

    def _init_tuning(self, context: Any, tuning: Optional[Mapping[str, Any]], *key: Any) -> None:
        self.tuning = dict(self.autotune(context, *key) if tuning is None else tuning)
