"""Streaming ingest around :class:`~katsdpsigproc_b200.rfi.device.FlaggerDevice`.

The flagger itself needs ~1.2 ms per MeerKAT dump on a B200, moving the dump over
PCIe takes ~40 ms, so an ingest loop is transfer-bound and must overlap the three
legs: host -> device copy of dump *i+1*, flagging of dump *i*, device -> host copy
of the flags of dump *i-1*.  This is the pattern the reference documents for its
users (``doc/user/sync.rst``, ``doc/user/resource.rst``: one command queue per
activity, ordered with ``enqueue_marker`` / ``enqueue_wait_for_events``), packaged
for the flagger: ``depth`` complete sets of device buffers, three in-order queues
(upload, compute, download), events between them, pinned host buffers.

Nothing here is specific to one GPU: with baseline sharding each rank owns one
``StreamingFlagger`` for its baseline range.
"""

from __future__ import annotations

import asyncio
from typing import Any, List, Mapping, Optional

import numpy as np

from . import accel, resource
from .rfi import device as rfi_device


class _Slot:
    """One set of buffers in flight."""

    def __init__(self, flagger: rfi_device.FlaggerDevice, context: Any) -> None:
        self.flagger = flagger
        flagger.ensure_all_bound()
        self.vis = flagger.buffer("vis")
        self.flags = flagger.buffer("flags")
        self.input_flags = flagger.buffer("input_flags") if "input_flags" in flagger.slots else None
        self.host_vis = self.vis.empty_like()           # pinned, same padding as the device buffer
        # two result buffers, used alternately: the one handed out by submit() stays valid
        # while the slot's next dump downloads into the other
        self.host_flags_pair = [self.flags.empty_like(), self.flags.empty_like()]
        self.host_flags = self.host_flags_pair[0]
        self.uses = 0
        self.host_input_flags = self.input_flags.empty_like() if self.input_flags is not None else None
        self.uploaded = None      # events of the most recent use of this slot
        self.computed = None
        self.downloaded = None
        self.busy = False


class StreamingFlagger:
    """Flag a stream of dumps with upload, compute and download overlapped.

    Parameters
    ----------
    template
        :class:`~katsdpsigproc_b200.rfi.device.FlaggerDeviceTemplate`
    channels, baselines
        Shape of every dump
    depth
        Dumps in flight (2 is enough to hide compute behind the transfers)
    background_args, noise_est_args, threshold_args
        As for ``FlaggerDeviceTemplate.instantiate``

    Usage::

        stream = StreamingFlagger(template, channels, baselines, threshold_args={"n_sigma": 11.0})
        for vis in dumps:
            done = stream.submit(vis)          # returns flags of an EARLIER dump, or None
            if done is not None:
                consume(done)
        for flags in stream.drain():
            consume(flags)

    The arrays returned are pinned host buffers owned by the object; each stays valid for
    the next ``depth`` submissions (every buffer set has two result buffers, used alternately).
    """

    def __init__(self, template: rfi_device.FlaggerDeviceTemplate, channels: int, baselines: int,
                 depth: int = 2, background_args: Mapping[str, Any] = {},
                 noise_est_args: Mapping[str, Any] = {}, threshold_args: Mapping[str, Any] = {}
                 ) -> None:
        if depth < 1:
            raise ValueError("depth must be at least 1")
        context = template.context
        self.context = context
        self.upload_queue = context.create_command_queue()
        self.compute_queue = context.create_command_queue()
        self.download_queue = context.create_command_queue()
        self._slots: List[_Slot] = []
        for _ in range(depth):
            flagger = template.instantiate(self.compute_queue, channels, baselines, background_args,
                                           noise_est_args, threshold_args)
            self._slots.append(_Slot(flagger, context))
        self._next = 0
        self._pending: List[_Slot] = []     # submitted, flags not yet handed out (oldest first)

    @property
    def depth(self) -> int:
        return len(self._slots)

    def host_vis(self, index: Optional[int] = None) -> accel.HostArray:
        """The pinned staging array the NEXT (or ``index``-th) submission copies from; a
        producer can write straight into it and call :meth:`submit` with ``None``.

        If that buffer set is still in flight, this first waits until its host -> device copy
        has completed: the array may be overwritten as soon as it is returned.
        """
        slot = self._slots[self._next if index is None else index]
        if slot.uploaded is not None:
            slot.uploaded.wait()          # the asynchronous upload still reads the staging array
        return slot.host_vis

    def submit(self, vis: Optional[np.ndarray], input_flags: Optional[np.ndarray] = None
               ) -> Optional[np.ndarray]:
        """Enqueue one dump.  If all buffer sets are in flight, first waits for the oldest
        dump and returns its flags; otherwise returns ``None``."""
        slot = self._slots[self._next]
        result = None
        if slot.busy:
            result = self._collect(slot)
        if vis is not None:
            if vis is not slot.host_vis:
                np.copyto(slot.host_vis, vis, casting="no")
        if slot.input_flags is not None:
            if input_flags is None:
                raise TypeError("input flags were expected but not provided")
            np.copyto(slot.host_input_flags, input_flags, casting="no")
        elif input_flags is not None:
            raise TypeError("input flags were provided but not included in the template")

        self._enqueue(slot)
        slot.busy = True
        self._pending.append(slot)
        self._next = (self._next + 1) % len(self._slots)
        return result

    def _enqueue(self, slot: _Slot) -> None:
        """Upload, flag and download the dump staged in ``slot``'s pinned arrays: three queues,
        ordered by events, nothing blocks."""
        # upload: the device vis buffer is free once the previous flagging of this slot is done
        if slot.computed is not None:
            self.upload_queue.enqueue_wait_for_events([slot.computed])
        slot.vis.set_async(self.upload_queue, slot.host_vis)
        if slot.input_flags is not None:
            slot.input_flags.set_async(self.upload_queue, slot.host_input_flags)
        slot.uploaded = self.upload_queue.enqueue_marker()
        # compute: after the upload; the flags buffer is free once its last download is done
        waits = [slot.uploaded]
        if slot.downloaded is not None:
            waits.append(slot.downloaded)
        self.compute_queue.enqueue_wait_for_events(waits)
        slot.flagger()
        slot.computed = self.compute_queue.enqueue_marker()
        # download
        self.download_queue.enqueue_wait_for_events([slot.computed])
        slot.uses += 1
        slot.host_flags = slot.host_flags_pair[slot.uses & 1]
        slot.flags.get_async(self.download_queue, slot.host_flags)
        slot.downloaded = self.download_queue.enqueue_marker()

    def _collect(self, slot: _Slot) -> np.ndarray:
        assert self._pending and self._pending[0] is slot
        slot.downloaded.wait()
        slot.busy = False
        self._pending.pop(0)
        return slot.host_flags_pair[slot.uses & 1]

    def drain(self) -> List[np.ndarray]:
        """Wait for everything in flight; returns the remaining flag arrays, oldest first."""
        out = []
        while self._pending:
            out.append(self._collect(self._pending[0]))
        return out


class AsyncStreamingFlagger:
    """The same pipeline for asyncio applications, built from the reference's own ordering
    primitives (:class:`~katsdpsigproc_b200.resource.Resource`, ``async_wait_for_events``;
    the pattern of the reference's ``doc/user/resource.rst``).

    Every buffer set is a :class:`Resource`; :meth:`submit` takes the next one in round-robin
    order and returns a task that resolves to the flags of that dump.  Tasks complete in
    submission order per buffer set, and the event loop is never blocked: host-side waits
    and the staging copy run in the default executor.

    ::

        stream = AsyncStreamingFlagger(template, channels, baselines, threshold_args={...})
        jobs = resource.JobQueue()
        async for vis in receiver:
            jobs.add(consume(stream.submit(vis)))
            await jobs.finish(max_remaining=stream.depth - 1)
        await jobs.finish()

    The array a task returns is a pinned buffer owned by the object; it stays valid until
    ``depth`` further dumps have been submitted AND awaited.
    """

    def __init__(self, template: rfi_device.FlaggerDeviceTemplate, channels: int, baselines: int,
                 depth: int = 2, background_args: Mapping[str, Any] = {},
                 noise_est_args: Mapping[str, Any] = {}, threshold_args: Mapping[str, Any] = {},
                 loop: Optional[asyncio.AbstractEventLoop] = None) -> None:
        self._engine = StreamingFlagger(template, channels, baselines, depth, background_args,
                                        noise_est_args, threshold_args)
        if loop is None:
            loop = asyncio.get_event_loop()
        self._loop = loop
        self._resources = [resource.Resource(slot, loop=loop) for slot in self._engine._slots]
        self._next = 0

    @property
    def depth(self) -> int:
        return len(self._resources)

    def submit(self, vis: np.ndarray, input_flags: Optional[np.ndarray] = None
               ) -> "asyncio.Future[np.ndarray]":
        """Queue one dump; returns a task whose result is the ``uint8`` flag array."""
        alloc = self._resources[self._next].acquire()      # order is fixed here, synchronously
        self._next = (self._next + 1) % len(self._resources)
        return asyncio.ensure_future(self._run(alloc, vis, input_flags), loop=self._loop)

    async def _run(self, alloc: "resource.ResourceAllocation[_Slot]", vis: np.ndarray,
                   input_flags: Optional[np.ndarray]) -> np.ndarray:
        with alloc as slot:
            await alloc.wait_events()                       # earlier dump in this buffer set is out
            if (slot.input_flags is None) != (input_flags is None):
                alloc.ready()
                raise TypeError("input flags were provided but not included in the template"
                                if input_flags is not None else
                                "input flags were expected but not provided")

            def stage() -> None:
                np.copyto(slot.host_vis, vis, casting="no")
                if input_flags is not None:
                    np.copyto(slot.host_input_flags, input_flags, casting="no")
            await self._loop.run_in_executor(None, stage)
            with self._engine.context:
                self._engine._enqueue(slot)
            flags = slot.host_flags
            downloaded = slot.downloaded
            await resource.async_wait_for_events([downloaded], loop=self._loop)
            alloc.ready()                                   # everything of this dump has left the device
            return flags
