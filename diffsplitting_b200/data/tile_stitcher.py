"""``stitch_predictions(predictions, idx_manager)`` with the reference's signature (data/tile_stitcher.py:10-81),
executed by the ``ds_stitch_tiles`` kernel.  (N,C,P,P) tiles -> (F,H,W,C) frames, bit-exact copy; where
destination boxes overlap, the tile the reference's sequential loop writes last wins."""
import numpy as np
import torch

from .. import _lib


def stitch_predictions(predictions, idx_manager, device=None):
    mng = idx_manager
    if len(mng.data_shape) != 3:
        raise NotImplementedError("stitching is implemented for (F, H, W) data with (1, P, P) patches")
    as_numpy = isinstance(predictions, np.ndarray)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("diffsplit_b200: stitch_predictions needs a CUDA device (no CPU fallback)")
        orig_dtype = predictions.dtype
        if predictions.dtype != np.float32:
            # the copy kernel moves 32-bit words; other dtypes would need their own instantiation
            raise TypeError(f"stitch_predictions: float32 tiles expected, got {orig_dtype}")
        tiles = torch.from_numpy(np.ascontiguousarray(predictions)).to(device or "cuda")
    else:
        tiles = predictions
        _lib.require_cuda(tiles, "tiles")
        if tiles.dtype != torch.float32:
            raise TypeError(f"stitch_predictions: float32 tiles expected, got {tiles.dtype}")
        tiles = tiles.contiguous()
    N, Cc, P1, P2 = tiles.shape
    total = mng.total_grid_count()
    if N != total:
        raise ValueError(f"stitch_predictions needs all {total} tiles, got {N}")
    if (P1, P2) != tuple(mng.patch_shape[1:]):
        raise ValueError(f"tile size {(P1, P2)} != patch shape {tuple(mng.patch_shape[1:])}")
    out = torch.empty(tuple(mng.data_shape) + (Cc,), dtype=torch.float32, device=tiles.device)
    d, g, p = mng._c_shapes()
    with torch.cuda.device(tiles.device):
        _lib.check(_lib.lib().ds_stitch_tiles(tiles.data_ptr(), Cc, d, g, p, int(mng.tiling_mode), out.data_ptr(),
                                              _lib.stream_ptr()))
    return out.cpu().numpy() if as_numpy else out
