"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the band partition, the uniform
band gather and the frame reassembly exactly as bench.py / a multi-GPU caller uses them (NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rtb200
from rtb200 import tiling


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pixel(y, x):  # a deterministic stand-in for "the pixel rank r computed"
    return (y * 7919 + x * 104729 + 12345) & 0x7FFFFFFF


def _worker(rank, world, port, h, w, band_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the scene "blob": rank 0 owns it, everyone must end with identical bytes
        blob = torch.arange(1000, dtype=torch.int64) * 3 if rank == 0 else torch.zeros(1000, dtype=torch.int64)
        dist.broadcast(blob, src=0)
        assert torch.equal(blob, torch.arange(1000, dtype=torch.int64) * 3)
        # every rank fills ONLY its own rows of a frame
        frame = torch.full((h, w), -1, dtype=torch.int32)
        rows = tiling.owned_rows(rank, world, h, band_rows)
        ys, xs = np.meshgrid(rows, np.arange(w), indexing="ij")
        if len(rows):
            frame[torch.as_tensor(rows)] = torch.as_tensor(_pixel(ys, xs).astype(np.int32))
        band = frame.index_select(0, tiling.band_index(rank, world, h, band_rows))
        gathered = tiling.gather_bands(dist, band)
        full = tiling.assemble_frame(gathered, h, band_rows)
        ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        ok = torch.equal(full, torch.as_tensor(_pixel(ys, xs).astype(np.int32)))
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            q.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w,band_rows", [(2, 64, 96, 16), (2, 70, 40, 8), (3, 52, 24, 4)])
def test_band_gather_roundtrip_gloo(world, h, w, band_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, band_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


def test_partition_is_exact_cover():
    for world, h, band in [(1, 1080, 16), (2, 1528, 16), (4, 2160, 16), (8, 3056, 16), (3, 50, 4)]:
        seen = np.concatenate([tiling.owned_rows(r, world, h, band) for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(h))
        sizes = [len(tiling.owned_rows(r, world, h, band)) for r in range(world)]
        assert max(sizes) - min(sizes) <= band
        assert tiling.rows_per_rank(world, h, band) == max(sizes)
    with pytest.raises(ValueError):
        tiling.owned_rows(0, 2, 64, 6)
