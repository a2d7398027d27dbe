"""CPU suite part 3: the N>1 host logic (round-robin tile sharding, packed-box layout, the single gather, stitch from the
gathered boxes) under gloo, world_size 2.  The two CUDA kernels of that path (`ds_pack_tile_regions`, `ds_stitch_packed`)
are replaced here - in the test only - by numpy stand-ins built on the oracle's copy-box arithmetic; the kernels
themselves are checked bit for bit on the GPU (tests/test_gpu_parity.py::test_packed_region_exchange_*)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsplitting_b200 import parallel as PAR
from diffsplitting_b200.data import TileIndexManager, TilingMode
from diffsplitting_b200.parallel import PackedLayout, chunk_ranges, rank_tile_ids, region_table, shard_chunks
from oracle import tiling_ref as TR

DATA, GRID, PATCH = (3, 100, 130), (1, 16, 16), (1, 32, 32)


def test_chunk_tables_cover_every_tile_once_and_balance():
    for total, chunk, world in ((490, 16, 8), (490, 16, 3), (45, 8, 2), (5, 8, 4), (490, 1, 8), (490, 8, 8)):
        chunks = chunk_ranges(total, chunk)
        assert sum(n for _, n in chunks) == total and chunks[0][0] == 0
        per = [list(shard_chunks(len(chunks), r, world)) for r in range(world)]
        assert sorted(c for p in per for c in p) == list(range(len(chunks)))
        assert max(len(p) for p in per) - min(len(p) for p in per) <= 1          # round-robin: floor or ceil of chunks / world
        ids = np.concatenate([rank_tile_ids(total, chunk, r, world) for r in range(world)])
        assert sorted(ids.tolist()) == list(range(total))


def test_region_table_matches_the_oracle_copy_boxes():
    for mode in (TilingMode.ShiftBoundary, TilingMode.TrimBoundary):
        mng = TileIndexManager(DATA, GRID, PATCH, mode)
        tg = TR.TileGrid(DATA, GRID, PATCH, int(mode))
        reg = region_table(mng)
        assert reg.shape == (mng.total_grid_count(), 5)
        for i in range(reg.shape[0]):
            vs, ve, _ = tg.copy_boxes(i)
            assert reg[i].tolist() == [vs[0], vs[1], ve[1], vs[2], ve[2]]
    # BASELINE frame set: the packed boxes are exactly the stitched frames (335 MB), not the 1.03 GB of whole tiles
    mng = TileIndexManager((10, 2048, 2048), (1, 256, 256), (1, 512, 512), TilingMode.ShiftBoundary)
    lay = PackedLayout(mng, 2, 8, 8)
    assert lay.total == 490 and lay.payload_bytes == 10 * 2048 * 2048 * 2 * 4
    assert max(lay.rank_len) <= 1.2 * min(lay.rank_len)


def _np_pack(tiles, mng, tile_ids, offsets, out):
    tg = TR.TileGrid(tuple(mng.data_shape), tuple(mng.grid_shape), tuple(mng.patch_shape), int(mng.tiling_mode))
    o = out.numpy()
    for t, (i, off) in enumerate(zip(tile_ids, offsets)):
        vs, ve, rs = tg.copy_boxes(int(i))
        hh, ww = ve[1] - vs[1], ve[2] - vs[2]
        box = tiles[t, :, rs[1]:rs[1] + hh, rs[2]:rs[2] + ww].numpy()
        o[off:off + box.size] = box.reshape(-1)
    return out


def _np_stitch(packed, mng, global_off, channels):
    tg = TR.TileGrid(tuple(mng.data_shape), tuple(mng.grid_shape), tuple(mng.patch_shape), int(mng.tiling_mode))
    out = np.zeros(tuple(mng.data_shape) + (channels,), dtype=np.float32)
    p = packed.numpy()
    for i in range(tg.total):
        vs, ve, _ = tg.copy_boxes(i)
        hh, ww = ve[1] - vs[1], ve[2] - vs[2]
        box = p[global_off[i]:global_off[i] + channels * hh * ww].reshape(channels, hh, ww)
        out[vs[0]:ve[0], vs[1]:ve[1], vs[2]:ve[2], :] = box.transpose(1, 2, 0)[None]
    return torch.from_numpy(out)


class _Frames:
    """CPU stand-in for TiledFrames (the product class needs a GPU): serves oracle-cropped tiles."""

    def __init__(self, frames):
        self.tile_manager = TileIndexManager(DATA, GRID, PATCH, TilingMode.ShiftBoundary)
        self.tg = TR.TileGrid(DATA, GRID, PATCH, int(TilingMode.ShiftBoundary))
        self.np_frames = frames
        self.frames = torch.zeros(1)
        self.patch_size = PATCH[1]

    def __len__(self):
        return self.tg.total

    def batch(self, first, n):
        t = torch.from_numpy(TR.crop_tiles(self.np_frames, self.tg, range(first, first + n)))
        return t, t


def _worker(rank, world, port, root, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        PAR.pack_tile_regions, PAR.stitch_packed = _np_pack, _np_stitch
        frames = np.arange(2 * np.prod(DATA), dtype=np.float32).reshape((2,) + DATA)
        tf = _Frames(frames)
        out = PAR.tiled_predict_and_stitch(lambda x: x, tf, chunk=5, out_channels=2, root=root)
        if out is None:
            q.put((rank, root is not None and rank != root))
        else:
            q.put((rank, bool(np.array_equal(out.numpy(), frames.transpose(1, 2, 3, 0)))))
        dist.destroy_process_group()
    except Exception as e:      # surface the failure instead of a queue timeout
        q.put((rank, repr(e)))


def _run(root):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, root, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_shard_pack_gather_stitch_world2_gloo_all_ranks():
    _run(None)


def test_shard_pack_gather_stitch_world2_gloo_root_only():
    _run(0)
