"""Baseline sharding of the flagger across the GPUs of one box.

The reference has no multi-GPU code.  Every stage of the flagger works on one
baseline at a time along the channel axis (reference
``rfi/background_median_filter.mako:208-219``, ``rfi/madnz_t.mako:79-86``,
``rfi/threshold_sum.mako:72-73``), so baselines are independent units: each
rank (one process per GPU) flags a contiguous range of baselines of every dump
and no data-path collective is needed.  The only optional exchange is a gather
of the per-rank flag blocks into a full ``channels x baselines`` array
(:func:`gather_flags`), which runs over NCCL/NVLink on GPUs and over gloo in the
CPU tests.
"""

from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple


def baseline_ranges(baselines: int, world_size: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous ``[start, stop)`` baseline ranges, one per rank.

    Range boundaries fall on multiples of ``align`` (rows of a shard then start on
    32-baseline = 256-byte boundaries of the complex64 input) and sizes differ by at
    most ``align``; with fewer than ``world_size`` aligned blocks the last ranks get
    empty ranges.
    """
    if baselines < 0 or world_size < 1 or align < 1:
        raise ValueError("baselines >= 0, world_size >= 1 and align >= 1 are required")
    blocks = -(-baselines // align)
    base, extra = divmod(blocks, world_size)
    ranges = []
    start_block = 0
    for rank in range(world_size):
        n = base + (1 if rank < extra else 0)
        start = min(start_block * align, baselines)
        stop = min((start_block + n) * align, baselines)
        ranges.append((start, stop))
        start_block += n
    return ranges


def baseline_range(baselines: int, rank: int, world_size: int, align: int = 32) -> Tuple[int, int]:
    return baseline_ranges(baselines, world_size, align)[rank]


def shard_columns(array: Any, rank: int, world_size: int, align: int = 32) -> Any:
    """The column block ``array[:, start:stop]`` of this rank (numpy or torch; a view)."""
    start, stop = baseline_range(array.shape[1], rank, world_size, align)
    return array[:, start:stop]


def gather_flags(local_flags: Any, baselines: int, group: Optional[Any] = None, align: int = 32
                 ) -> Any:
    """All-gather the per-rank flag blocks (``channels x local_baselines`` uint8 torch tensors)
    into the full ``channels x baselines`` array on every rank.

    Blocks are padded to the largest shard so that one fixed-size ``all_gather`` does
    the exchange (1 byte per visibility; off the critical path of the flagger).
    """
    import torch
    import torch.distributed as dist

    world_size = dist.get_world_size(group)
    ranges = baseline_ranges(baselines, world_size, align)
    channels = local_flags.shape[0]
    widest = max(stop - start for start, stop in ranges)
    mine = torch.zeros((channels, widest), dtype=local_flags.dtype, device=local_flags.device)
    mine[:, : local_flags.shape[1]] = local_flags
    blocks = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(blocks, mine, group=group)
    full = torch.empty((channels, baselines), dtype=local_flags.dtype, device=local_flags.device)
    for (start, stop), block in zip(ranges, blocks):
        full[:, start:stop] = block[:, : stop - start]
    return full


def shard_sizes(baselines: int, world_size: int, align: int = 32) -> Sequence[int]:
    return [stop - start for start, stop in baseline_ranges(baselines, world_size, align)]


def bind_to_device_locality(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs that are closest to CUDA device ``device_index``.

    With one process per GPU, the pinned staging buffers of
    :class:`~katsdpsigproc_b200.streaming.StreamingFlagger` are then allocated (first touch)
    on the NUMA node the GPU hangs off, and the PCIe copies of all ranks stop crossing the
    socket interconnect - on an 8-GPU box that is the difference between the ranks sharing one
    memory controller and each using its own.  Uses NVML (``nvidia-ml-py``); returns the CPU
    list it bound to, or ``None`` if NVML or the affinity information is not available (the
    process is then left as it was).
    """
    import os

    try:
        import pynvml

        import ctypes

        from . import _capi

        # NVML numbers the GPUs of the machine, CUDA those this process may see
        # (CUDA_VISIBLE_DEVICES): name the device by its PCI bus id, which both agree on
        bus_id = ctypes.create_string_buffer(32)
        _capi.call("ksp_device_pci_bus_id", int(device_index), bus_id, len(bus_id))
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.value)
            words = -(-(os.cpu_count() or 1) // 64)
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + bit for w, word in enumerate(mask) for bit in range(64) if (word >> bit) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
