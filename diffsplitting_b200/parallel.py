"""Multi-GPU tiled prediction: one process per GPU, tile chunks dealt round-robin, ONE collective (gather of the packed
destination boxes of the predicted tiles) before a single stitch.

The reference predicts its 490 tiles serially at batch 1 on one GPU (notebooks/EvaluateJointIndi.ipynb cell 23;
split.py:59-70) and has no process-group code.  Tiles are independent sampling problems, so rank r of R takes the chunks
c = r, r + R, r + 2R, ... of the global chunk table (round-robin: every rank gets floor or ceil of chunks / R, whatever the
chunk count) and runs its own sampling loops with no inter-GPU traffic.  Noise is made independent of R by seeding every
chunk's Philox offset from its GLOBAL chunk index.

Only what the stitcher reads travels: ``stitch_predictions`` (data/tile_stitcher.py:28-57) copies from each tile its inner
grid cell, widened to the frame edge where the patch touches it.  Every rank packs exactly those boxes
(``ds_pack_tile_regions``), the packed buffers are gathered (NCCL over NVLink: to the root, or to every rank), and
``ds_stitch_packed`` writes the (F,H,W,C) frames once.  For 10 x 2048^2 frames, 512^2 tiles, 2 channels that is 335 MB
instead of the 1.03 GB of whole tiles.
"""
import ctypes as C
from typing import Callable, List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def chunk_ranges(total: int, chunk: int) -> List[Tuple[int, int]]:
    """Global chunk table [(first, n)], identical on every rank."""
    return [(s, min(chunk, total - s)) for s in range(0, total, chunk)]


def shard_chunks(n_chunks: int, rank: int, world: int) -> range:
    """Chunk indices of `rank`: round-robin (c % world == rank)."""
    return range(rank, n_chunks, world)


def rank_tile_ids(total: int, chunk: int, rank: int, world: int) -> np.ndarray:
    """Global tile indices predicted by `rank`, in the order it predicts them."""
    chunks = chunk_ranges(total, chunk)
    ids = [np.arange(chunks[c][0], chunks[c][0] + chunks[c][1], dtype=np.int64) for c in shard_chunks(len(chunks), rank, world)]
    return np.concatenate(ids) if ids else np.zeros((0,), dtype=np.int64)


def region_table(mng) -> np.ndarray:
    """(total, 5) int32: (frame, y_lo, y_hi, x_lo, x_hi) of the destination box of every tile of the manager."""
    total = mng.total_grid_count()
    out = np.zeros((total, 5), dtype=np.int32)
    d, g, p = mng._c_shapes()
    _lib.check(_lib.lib().ds_tile_regions(d, g, p, int(mng.tiling_mode), 0, total,
                                          out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


class PackedLayout:
    """Where every tile's packed box lives: per-rank send buffers (tiles in the rank's own order) laid side by side, each
    padded to the longest.  Pure host arithmetic, identical on every rank."""

    def __init__(self, mng, channels: int, chunk: int, world: int):
        reg = region_table(mng)
        self.total = reg.shape[0]
        self.channels, self.world = channels, world
        size = (reg[:, 2] - reg[:, 1]).astype(np.int64) * (reg[:, 4] - reg[:, 3]).astype(np.int64) * channels
        self.ids = [rank_tile_ids(self.total, chunk, r, world) for r in range(world)]
        self.local_off = []
        lens = []
        for r in range(world):
            sz = size[self.ids[r]]
            off = np.concatenate([[0], np.cumsum(sz)[:-1]]).astype(np.int64) if len(sz) else np.zeros((0,), dtype=np.int64)
            self.local_off.append(off)
            lens.append(int(sz.sum()))
        self.rank_len = lens
        self.slot = max(max(lens), 1)                                   # elements per rank in the gathered buffer
        self.global_off = np.zeros((self.total,), dtype=np.int64)
        for r in range(world):
            self.global_off[self.ids[r]] = r * self.slot + self.local_off[r]
        self.payload_bytes = int(size.sum()) * 4


def pack_tile_regions(tiles: torch.Tensor, mng, tile_ids: np.ndarray, offsets: np.ndarray, out: torch.Tensor):
    """tiles (n,C,P,P) CUDA fp32 with global indices `tile_ids` -> their destination boxes at `out[offsets[i]:]`."""
    _lib.require_cuda(tiles, "tiles")
    n = int(tiles.shape[0])
    if n == 0:
        return out
    dev = tiles.device
    ids = torch.from_numpy(np.ascontiguousarray(tile_ids, dtype=np.int64)).to(dev)
    off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).to(dev)
    d, g, p = mng._c_shapes()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ds_pack_tile_regions(tiles.contiguous().data_ptr(), int(tiles.shape[1]), d, g, p, int(mng.tiling_mode),
                                                   ids.data_ptr(), off.data_ptr(), n, out.data_ptr(), _lib.stream_ptr()))
    return out


def stitch_packed(packed: torch.Tensor, mng, global_off: np.ndarray, channels: int) -> torch.Tensor:
    """Gathered packed boxes of ALL tiles -> (F,H,W,C) frames; same bits as ``stitch_predictions`` on the whole tiles."""
    _lib.require_cuda(packed, "packed tiles")
    dev = packed.device
    off = torch.from_numpy(np.ascontiguousarray(global_off, dtype=np.int64)).to(dev)
    out = torch.empty(tuple(mng.data_shape) + (channels,), dtype=torch.float32, device=dev)
    d, g, p = mng._c_shapes()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ds_stitch_packed(packed.data_ptr(), off.data_ptr(), channels, d, g, p, int(mng.tiling_mode),
                                               out.data_ptr(), _lib.stream_ptr()))
    return out


def gather_packed(local: torch.Tensor, world: int, root=None, group=None):
    """One collective: every rank's `slot`-long send buffer -> (world * slot,) on the root (``root`` = its rank; the others
    get None) or on every rank (``root=None``).  NCCL on GPUs; gloo on CPU tensors for the host-logic tests."""
    if world == 1:
        return local
    rank = dist.get_rank(group)
    if root is None:
        out = torch.empty((world * local.numel(),), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    if rank == root:
        out = torch.empty((world * local.numel(),), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.chunk(world)), dst=root, group=group)
        return out
    dist.gather(local, None, dst=root, group=group)
    return None


def predict_tiles(infer: Callable[[torch.Tensor], torch.Tensor], tiled, chunk: int, rank: int = 0, world: int = 1,
                  seed_base: int = None, offset_stride: int = 0):
    """Run ``infer(input_batch) -> (n, C, P, P)`` over this rank's chunks of ``tiled`` (a ``TiledFrames``).
    Returns (local predictions or None, global tile indices of this rank).  If ``offset_stride`` > 0 the CUDA generator
    offset is set to ``global_chunk_index * offset_stride`` before every chunk: the noise does not depend on `world`."""
    total = len(tiled)
    chunks = chunk_ranges(total, chunk)
    outs = []
    gen = None
    if offset_stride:
        dev = tiled.frames.device
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        if seed_base is not None:
            gen.manual_seed(seed_base)
    for ci in shard_chunks(len(chunks), rank, world):
        first, n = chunks[ci]
        inp, _ = tiled.batch(first, n)
        if gen is not None:
            gen.set_offset(ci * offset_stride)
        outs.append(infer(inp))
    local = torch.cat(outs, dim=0) if outs else None
    return local, rank_tile_ids(total, chunk, rank, world)


def tiled_predict_and_stitch(infer, tiled, chunk: int, out_channels: int, group=None, offset_stride: int = 0,
                             seed_base: int = None, root=None):
    """Shard (round-robin chunks) -> predict -> pack the destination boxes -> ONE gather -> stitch once.
    ``root=None``: every rank returns the stitched (F,H,W,C) frames; ``root=r``: rank r does, the others return None."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mng = tiled.tile_manager
    lay = PackedLayout(mng, out_channels, chunk, world)
    local, ids = predict_tiles(infer, tiled, chunk, rank, world, seed_base, offset_stride)
    dev = tiled.frames.device
    send = torch.zeros((lay.slot,), dtype=torch.float32, device=dev)
    if local is not None:
        pack_tile_regions(local, mng, ids, lay.local_off[rank], send)
    packed = gather_packed(send, world, root, group)
    if packed is None:
        return None
    return stitch_packed(packed, mng, lay.global_off, out_channels)
