"""Multi-GPU tiled prediction: one process per GPU, tiles sharded by index, ONE collective (all-gather of the
predicted tiles) before stitching.

The reference predicts its 490 tiles serially at batch 1 on one GPU (notebooks/EvaluateJointIndi.ipynb cell 23;
split.py:59-70) and has no process-group code.  Tiles are independent sampling problems, so rank r of R takes
a contiguous block of tile *chunks* and runs its own sampling loops with no inter-GPU traffic.  Noise is made
independent of R by seeding every chunk's Philox offset from its GLOBAL chunk index.
"""
from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def chunk_ranges(total: int, chunk: int) -> List[Tuple[int, int]]:
    """Global chunk table [(first, n)], identical on every rank."""
    return [(s, min(chunk, total - s)) for s in range(0, total, chunk)]


def shard_chunks(n_chunks: int, rank: int, world: int) -> range:
    """Contiguous block of chunk indices for `rank` (keeps a rank's tiles inside few frames)."""
    per = -(-n_chunks // world)
    lo = min(rank * per, n_chunks)
    return range(lo, min(lo + per, n_chunks))


def gather_tiles(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """All-gather variable-length per-rank tile blocks (n_r, C, P, P) into (sum n_r, C, P, P) on every rank.
    NCCL over NVLink on GPUs; works with gloo on CPU tensors for the host-logic tests."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    assert len(counts) == world
    nmax = max(counts)
    shape = (nmax,) + tuple(local.shape[1:])
    padded = local
    if local.shape[0] != nmax:
        padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * nmax,) + shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    out = out.view((world,) + shape)
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def predict_tiles(infer: Callable[[torch.Tensor], torch.Tensor], tiled, chunk: int, rank: int = 0, world: int = 1,
                  seed_base: int = None, offset_stride: int = 0):
    """Run ``infer(input_batch) -> (n, C, P, P)`` over this rank's chunks of ``tiled`` (a ``TiledFrames``).
    Returns (local predictions, per-rank tile counts).  If ``offset_stride`` > 0 the CUDA generator offset is set
    to ``global_chunk_index * offset_stride`` before every chunk, which makes the noise independent of `world`."""
    total = len(tiled)
    chunks = chunk_ranges(total, chunk)
    mine = shard_chunks(len(chunks), rank, world)
    outs = []
    gen = None
    if offset_stride:
        dev = tiled.frames.device
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        if seed_base is not None:
            gen.manual_seed(seed_base)
    for ci in mine:
        first, n = chunks[ci]
        inp, _ = tiled.batch(first, n)
        if gen is not None:
            gen.set_offset(ci * offset_stride)
        outs.append(infer(inp))
    counts = [sum(chunks[c][1] for c in shard_chunks(len(chunks), r, world)) for r in range(world)]
    if outs:
        local = torch.cat(outs, dim=0)
    else:
        local = None
    return local, counts


def tiled_predict_and_stitch(infer, tiled, chunk: int, out_channels: int, group=None, offset_stride: int = 0,
                             seed_base: int = None):
    """Shard -> predict -> all-gather -> stitch.  Every rank returns the stitched (F, H, W, C) frames."""
    from .data.tile_stitcher import stitch_predictions
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local, counts = predict_tiles(infer, tiled, chunk, rank, world, seed_base, offset_stride)
    P = tiled.patch_size
    if local is None:
        local = torch.zeros((0, out_channels, P, P), dtype=torch.float32, device=tiled.frames.device)
    tiles = gather_tiles(local.contiguous(), counts, group)
    return stitch_predictions(tiles, tiled.tile_manager)
