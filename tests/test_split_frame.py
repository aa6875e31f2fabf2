"""CPU: the sort-first band split and the composite step, world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kazan_b200.split_frame import SplitFrame, band_rows, shard_invocations


@pytest.mark.parametrize("height,world", [(2160, 1), (2160, 2), (2160, 8), (4320, 8), (1080, 4), (1081, 3), (7, 2), (8, 3), (540, 7)])
def test_bands_partition_the_frame(height, world):
    bands = [band_rows(height, world, r) for r in range(world)]
    assert bands[0][0] == 0 and bands[-1][1] == height
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(y0 % 16 == 0 for y0, _ in bands)  # no raster brick is shared between ranks
    heights = [b - a for a, b in bands if b > a]
    assert max(heights) - min(heights) <= 16 or len(heights) < world


@pytest.mark.parametrize("total,world", [(1 << 24, 8), (1000, 3), (255, 2), (0, 4), (257, 2)])
def test_dispatch_shards_cover_every_invocation_once(total, world):
    shards = [shard_invocations(total, world, r) for r in range(world)]
    assert shards[0][0] == 0 and sum(c for _, c in shards) == total
    for (f0, c0), (f1, _) in zip(shards, shards[1:]):
        assert f0 + c0 == f1
    assert all(f % 256 == 0 for f, _ in shards)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, width, height, presenter, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sf = SplitFrame(width, height, world, rank, presenter)
        frame = torch.full((width * height * 4,), 0xEE, dtype=torch.uint8)
        a, b = sf.byte_range(rank)
        frame[a:b] = torch.arange(a, b, dtype=torch.int64).mul_(2654435761 + rank).bitwise_and_(0xFF).to(torch.uint8)
        sf.composite(frame, dist)
        if rank == presenter:
            expect = torch.empty_like(frame)
            for r in range(world):
                a, b = sf.byte_range(r)
                expect[a:b] = torch.arange(a, b, dtype=torch.int64).mul_(2654435761 + r).bitwise_and_(0xFF).to(torch.uint8)
            result.put(bool(torch.equal(frame, expect)))
            assert sf.composite_bytes() == width * height * 4 - (sf.byte_range(presenter)[1] - sf.byte_range(presenter)[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,presenter,height", [(2, 0, 64), (3, 1, 100)])
def test_composite_gathers_every_band_on_the_presenter(world, presenter, height):
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 48, height, presenter, result)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert result.get(timeout=10) is True
