"""Hand-over of shared buffers between asyncio tasks that drive the device.

Same public surface and semantics as the reference's ``katsdpsigproc/resource.py``
(``wait_until`` :31-56, ``async_wait_for_events`` :59-81, ``ResourceAllocation`` :84-166,
``Resource`` :169-214, ``JobQueue`` :217-244): a pipeline of coroutines (receive a dump,
upload it, flag it, download the flags) passes each device buffer from stage to stage
*in acquisition order*, and what is handed over is not "the buffer is free" but "the
buffer is free once these device events have fired", so the next stage can either block
on the host (:meth:`ResourceAllocation.wait_events`) or make its command queue wait
(``queue.enqueue_wait_for_events(await alloc.wait())``) without stalling the event loop.

Nothing here touches CUDA; events only need the :class:`~katsdpsigproc_b200.abc.AbstractEvent`
interface.  :class:`katsdpsigproc_b200.streaming.StreamingFlagger` is the synchronous
packaging of the same idea for the flagger.
"""

from __future__ import annotations

import asyncio
import collections
import logging
from types import TracebackType
from typing import Awaitable, Deque, Generic, Iterable, List, Optional, Type, TypeVar

from .abc import AbstractEvent

_T = TypeVar("_T")
_logger = logging.getLogger(__name__)
_EventList = List[AbstractEvent]


async def wait_until(future: Awaitable[_T], when: float,
                     loop: Optional[asyncio.AbstractEventLoop] = None) -> _T:
    """Await `future` with a deadline on the loop's clock (``loop.time()``).

    Returns the result (or raises the exception) of `future` if it completes by `when`;
    otherwise cancels it and raises :exc:`asyncio.TimeoutError`.  Wrap the awaitable in
    :func:`asyncio.shield` to keep it alive past the deadline.
    """
    if loop is None:
        loop = asyncio.get_event_loop()
    task = asyncio.ensure_future(future, loop=loop)
    if not task.done():
        await asyncio.wait([task], timeout=max(0.0, when - loop.time()))
    if task.done():
        return task.result()
    task.cancel()
    raise asyncio.TimeoutError()


def _block_on(events: _EventList) -> None:
    """Worker-thread half of :func:`async_wait_for_events`."""
    while events:
        events[0].wait()
        # drop our reference here, in the worker, BEFORE the awaiting coroutine resumes:
        # the caller may release its own references as soon as it wakes up, and the last
        # reference to a device event should not die at an arbitrary later time on this thread
        del events[0]


async def async_wait_for_events(events: Iterable[AbstractEvent],
                                loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
    """Wait for device events without blocking the event loop (the blocking
    ``event.wait()`` calls run in the loop's default executor)."""
    if loop is None:
        loop = asyncio.get_event_loop()
    pending = list(events)
    if pending:
        await loop.run_in_executor(None, _block_on, pending)


class ResourceAllocation(Generic[_T]):
    """One turn at a :class:`Resource`; obtained from :meth:`Resource.acquire` only.

    The turn begins when the future returned by :meth:`wait` resolves - to the list of device
    events the previous holder passed to :meth:`ready` - and ends when this holder calls
    :meth:`ready` with the events of its own device work.  Use it as a context manager
    (``with alloc as buffer:``) so that an exception inside the block is forwarded to the
    holders queued behind it instead of leaving them waiting for ever.
    """

    def __init__(self, start: "asyncio.Future[_EventList]", end: "asyncio.Future[_EventList]",
                 value: _T, loop: asyncio.AbstractEventLoop) -> None:
        self._start = start
        self._end = end
        self._loop = loop
        self.value = value

    def wait(self) -> "asyncio.Future[_EventList]":
        """Future for the start of the turn; its result is the events to wait for before use."""
        return self._start

    async def wait_events(self) -> None:
        """Start of the turn AND completion of the previous holder's device work, on the host."""
        await async_wait_for_events(await self._start, loop=self._loop)

    def ready(self, events: Optional[_EventList] = None) -> None:
        """End the turn.  `events` mark the device work the next holder has to wait for.
        Must not be called before the turn has begun, even if the resource ends up unused."""
        self._end.set_result([] if events is None else events)

    def __enter__(self) -> _T:
        return self.value

    def __exit__(self, exc_type: Optional[Type[BaseException]],
                 exc_value: Optional[BaseException], exc_tb: Optional[TracebackType]) -> None:
        if self._end.done():
            return
        if exc_value is None:
            _logger.warning("Resource allocation was not explicitly made ready")
            self.ready()
        else:
            self._end.set_exception(exc_value)
            self._end.exception()      # mark it retrieved: it also propagates out of the with block


class Resource(Generic[_T]):
    """A value (typically a device buffer or a whole operation) used by one task at a time,
    strictly in the order of the :meth:`acquire` calls.

    ``acquire`` never blocks: it returns a :class:`ResourceAllocation` whose start is chained
    to the end of the allocation handed out before it.
    """

    def __init__(self, value: _T, loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        if loop is None:
            loop = asyncio.get_event_loop()
        self._loop = loop
        self.value = value
        self._tail: "asyncio.Future[_EventList]" = loop.create_future()
        self._tail.set_result([])      # nobody before the first holder

    def acquire(self) -> ResourceAllocation[_T]:
        start = self._tail
        self._tail = self._loop.create_future()
        return ResourceAllocation(start, self._tail, self.value, self._loop)


class JobQueue:
    """FIFO of in-flight jobs (futures or coroutines), to bound how far a producer runs ahead."""

    def __init__(self) -> None:
        self._jobs: Deque["asyncio.Future"] = collections.deque()

    def add(self, job: Awaitable) -> None:
        """Append a job; a coroutine is scheduled as a task."""
        self._jobs.append(asyncio.ensure_future(job))

    def clean(self) -> None:
        """Drop finished jobs from the front, re-raising the exception of any that failed."""
        while self._jobs and self._jobs[0].done():
            self._jobs.popleft().result()

    async def finish(self, max_remaining: int = 0) -> None:
        """Await jobs from the front until at most `max_remaining` are left."""
        while len(self._jobs) > max_remaining:
            await self._jobs.popleft()

    def __len__(self) -> int:
        return len(self._jobs)

    def __bool__(self) -> bool:
        return len(self._jobs) > 0

    def __contains__(self, item: object) -> bool:
        return item in self._jobs


__all__ = ["wait_until", "async_wait_for_events", "Resource", "ResourceAllocation", "JobQueue"]
