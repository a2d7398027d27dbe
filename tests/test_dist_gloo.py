"""CPU suite part 3: the N>1 host logic (tile sharding + the single all-gather) under gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsplitting_b200.parallel import chunk_ranges, gather_tiles, shard_chunks


def test_chunk_tables_cover_every_tile_once():
    for total, chunk, world in ((490, 16, 8), (490, 16, 3), (45, 8, 2), (5, 8, 4), (490, 1, 8)):
        chunks = chunk_ranges(total, chunk)
        assert sum(n for _, n in chunks) == total and chunks[0][0] == 0
        seen = []
        for r in range(world):
            seen += list(shard_chunks(len(chunks), r, world))
        assert seen == list(range(len(chunks)))


def _worker(rank, world, port, total, chunk, q):
    try:
        _worker_body(rank, world, port, total, chunk, q)
    except Exception as e:      # surface the failure instead of a queue timeout
        q.put((rank, repr(e)))


def _worker_body(rank, world, port, total, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunks = chunk_ranges(total, chunk)
    mine = shard_chunks(len(chunks), rank, world)
    tiles = [torch.full((n, 2, 4, 4), 0.0) + torch.arange(f, f + n).float().view(n, 1, 1, 1) for f, n in (chunks[c] for c in mine)]
    local = torch.cat(tiles) if tiles else torch.zeros((0, 2, 4, 4))
    counts = [sum(chunks[c][1] for c in shard_chunks(len(chunks), r, world)) for r in range(world)]
    full = gather_tiles(local, counts)
    ok = full.shape[0] == total and bool((full[:, 0, 0, 0] == torch.arange(total).float()).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_tiles_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 45, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
