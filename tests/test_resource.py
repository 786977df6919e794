"""katsdpsigproc_b200.resource: the cases of the reference's test/test_resource.py
(wait_until :31-75, Resource :104-157, JobQueue :160-211), driven with asyncio.run
because pytest-asyncio is not in this image."""

import asyncio
import logging
import queue
import threading
import time

import pytest

from katsdpsigproc_b200 import resource
from katsdpsigproc_b200.abc import AbstractEvent

FAR = 1e6


def run(coro):
    return asyncio.run(coro)


# ------------------------------------------------------------------ wait_until
def test_wait_until_result():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        loop.call_later(0.05, future.set_result, 42)
        return await resource.wait_until(future, loop.time() + FAR)
    assert run(main()) == 42


def test_wait_until_already_done():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        future.set_result(7)
        return await resource.wait_until(future, loop.time() + FAR)
    assert run(main()) == 7


def test_wait_until_exception():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        loop.call_later(0.05, future.set_exception, ValueError("boom"))
        await resource.wait_until(future, loop.time() + FAR)
    with pytest.raises(ValueError):
        run(main())


def test_wait_until_timeout_cancels():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        with pytest.raises(asyncio.TimeoutError):
            await resource.wait_until(future, loop.time() + 0.01)
        return future.cancelled()
    assert run(main())


def test_wait_until_deadline_in_the_past():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        with pytest.raises(asyncio.TimeoutError):
            await resource.wait_until(future, loop.time() - 1.0)
        return future.cancelled()
    assert run(main())


def test_wait_until_shield_keeps_inner_alive():
    async def main():
        loop = asyncio.get_running_loop()
        future = loop.create_future()
        with pytest.raises(asyncio.TimeoutError):
            await resource.wait_until(asyncio.shield(future), loop.time() + 0.01)
        return future.cancelled()
    assert not run(main())


def test_wait_until_accepts_coroutine():
    async def answer():
        await asyncio.sleep(0.01)
        return "done"

    async def main():
        return await resource.wait_until(answer(), asyncio.get_running_loop().time() + FAR)
    assert run(main()) == "done"


# ------------------------------------------------------------------ events
class SlowEvent(AbstractEvent):
    """wait() sleeps briefly, then logs itself into a queue (once)."""

    def __init__(self, log: "queue.Queue") -> None:
        self.log = log
        self.fired = False
        self.thread = None

    def wait(self) -> None:
        if not self.fired:
            time.sleep(0.05)
            self.thread = threading.get_ident()
            self.log.put(self)
            self.fired = True

    def time_since(self, prior_event):
        return 0.0

    def time_till(self, next_event):
        return 0.0


def drain(q):
    out = []
    while True:
        try:
            out.append(q.get_nowait())
        except queue.Empty:
            return out


def test_async_wait_for_events_runs_off_loop():
    log = queue.Queue()
    events = [SlowEvent(log), SlowEvent(log)]

    async def main():
        ticks = 0

        async def ticker():
            nonlocal ticks
            while True:
                await asyncio.sleep(0.005)
                ticks += 1
        t = asyncio.ensure_future(ticker())
        await resource.async_wait_for_events(events)
        t.cancel()
        return ticks
    ticks = run(main())
    assert drain(log) == events
    assert all(e.thread != threading.get_ident() for e in events)
    assert ticks >= 3          # the loop kept running during the 0.1 s of blocking waits


def test_async_wait_for_events_empty():
    run(resource.async_wait_for_events([]))
    run(resource.async_wait_for_events(iter(())))


# ------------------------------------------------------------------ Resource
def test_resource_order_and_events():
    log = queue.Queue()

    async def frame(alloc, event):
        with alloc as value:
            assert value == 42
            await alloc.wait_events()
            alloc.ready([event])
            log.put(alloc)

    async def main():
        r = resource.Resource(42)
        a0, a1 = r.acquire(), r.acquire()
        e0, e1 = SlowEvent(log), SlowEvent(log)
        second = asyncio.ensure_future(frame(a1, e1))      # started first, must still run second
        first = asyncio.ensure_future(frame(a0, e0))
        await first
        await second
        return a0, e0, a1
    a0, e0, a1 = run(main())
    # a0 completes; a1 then waits for a0's event e0 before completing; e1 is never waited for
    assert drain(log) == [a0, e0, a1]


def test_resource_wait_returns_previous_events():
    async def main():
        r = resource.Resource("buf")
        a0, a1 = r.acquire(), r.acquire()
        assert r.value == "buf" and a0.value == "buf"
        assert await a0.wait() == []
        assert not a1.wait().done()
        marker = object()
        a0.ready([marker])
        assert await a1.wait() == [marker]
        a1.ready()
        assert await r.acquire().wait() == []
    run(main())


def test_resource_exception_is_forwarded():
    async def main():
        r = resource.Resource(None)
        a0, a1 = r.acquire(), r.acquire()
        with pytest.raises(RuntimeError):
            with a0:
                await a0.wait_events()
                raise RuntimeError("stage failed")
        with pytest.raises(RuntimeError):
            with a1:
                await a1.wait_events()
                a1.ready()
    run(main())


def test_resource_missing_ready_warns(caplog):
    async def main():
        r = resource.Resource(None)
        a0, a1 = r.acquire(), r.acquire()
        with a0:
            pass
        assert await a1.wait() == []       # a0 was made ready on exit
        a1.ready()
    with caplog.at_level(logging.WARNING, logger="katsdpsigproc_b200.resource"):
        run(main())
    assert caplog.record_tuples == [
        ("katsdpsigproc_b200.resource", logging.WARNING,
         "Resource allocation was not explicitly made ready")]


# ------------------------------------------------------------------ JobQueue
def _futures(loop, n, done):
    out = [loop.create_future() for _ in range(n)]
    if done:
        for i, f in enumerate(out):
            f.set_result(i)
    return out


def test_jobqueue_clean_len_bool_contains():
    async def main():
        loop = asyncio.get_running_loop()
        fin, unf = _futures(loop, 3, True), _futures(loop, 3, False)
        jobs = resource.JobQueue()
        assert not jobs and len(jobs) == 0 and fin[0] not in jobs
        jobs.add(fin[0])
        jobs.add(unf[0])
        jobs.add(fin[1])
        assert jobs and len(jobs) == 3 and fin[0] in jobs and fin[2] not in jobs
        jobs.clean()                        # only the finished job at the FRONT goes
        assert len(jobs) == 2 and fin[0] not in jobs and fin[1] in jobs
        for f in unf:
            f.cancel()
    run(main())


def test_jobqueue_clean_rethrows():
    async def main():
        loop = asyncio.get_running_loop()
        bad = loop.create_future()
        bad.set_exception(KeyError("x"))
        jobs = resource.JobQueue()
        jobs.add(bad)
        with pytest.raises(KeyError):
            jobs.clean()
        assert len(jobs) == 0
    run(main())


def test_jobqueue_finish_leaves_max_remaining():
    async def main():
        loop = asyncio.get_running_loop()
        fin, unf = _futures(loop, 1, True), _futures(loop, 3, False)
        jobs = resource.JobQueue()
        for f in (fin[0], unf[0], unf[1], unf[2]):
            jobs.add(f)

        async def finisher():
            for i, f in enumerate(unf):
                await asyncio.sleep(0.01)
                f.set_result(i)
        t = asyncio.ensure_future(finisher())
        await jobs.finish(max_remaining=1)
        state = (unf[0].done(), unf[1].done(), unf[2].done(), len(jobs))
        await t
        return state
    assert run(main()) == (True, True, False, 1)


def test_jobqueue_wraps_coroutines():
    async def main():
        async def work(x):
            await asyncio.sleep(0.001)
            return x
        jobs = resource.JobQueue()
        jobs.add(work(1))
        jobs.add(work(2))
        await jobs.finish()
        return len(jobs)
    assert run(main()) == 0
