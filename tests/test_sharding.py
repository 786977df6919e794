"""Baseline sharding across ranks: range arithmetic, and the optional flag gather over a
world_size-2 gloo group on CPU (the NCCL path runs the same code on GPUs)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from katsdpsigproc_b200 import sharding


@pytest.mark.parametrize("baselines,world,align", [(8320, 8, 32), (12960, 8, 32), (12960, 4, 32),
                                                   (12960, 2, 128), (100, 8, 32), (0, 2, 32),
                                                   (33, 2, 1), (8320, 1, 32)])
def test_ranges_partition_the_baselines(baselines, world, align):
    ranges = sharding.baseline_ranges(baselines, world, align)
    assert len(ranges) == world
    assert ranges[0][0] == 0 and ranges[-1][1] == baselines
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in ranges]
    assert sum(sizes) == baselines
    assert all(a % align == 0 for a, _ in ranges if a < baselines)
    assert max(sizes) - min(sizes) <= align or baselines < world * align
    assert sharding.shard_sizes(baselines, world, align) == sizes


def test_meerkat_shards():
    assert sharding.shard_sizes(8320, 8) == [1056, 1056, 1056, 1056, 1024, 1024, 1024, 1024]
    assert sharding.baseline_range(12960, 7, 8) == (11360, 12960)


def test_shard_columns_is_a_view():
    a = np.arange(12 * 100).reshape(12, 100)
    block = sharding.shard_columns(a, 1, 2, align=32)
    assert block.shape == (12, 36) and block.base is not None
    np.testing.assert_array_equal(block, a[:, 64:])


def test_bad_arguments():
    with pytest.raises(ValueError):
        sharding.baseline_ranges(10, 0)
    with pytest.raises(ValueError):
        sharding.baseline_ranges(-1, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, channels, baselines, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = (np.arange(channels * baselines).reshape(channels, baselines) % 251).astype(np.uint8)
        mine = sharding.shard_columns(full, rank, world)
        local = torch.from_numpy(np.ascontiguousarray(mine))
        gathered = sharding.gather_flags(local, baselines)
        ok = bool(np.array_equal(gathered.numpy(), full))
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_flags_gloo_world2(tmp_path):
    world, channels, baselines = 2, 64, 100      # ragged: 64 + 36 baselines
    mp.spawn(_gather_worker, args=(world, _free_port(), channels, baselines, str(tmp_path)),
             nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / f"rank{rank}.txt").read_text() == "ok"
