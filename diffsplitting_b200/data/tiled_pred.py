"""Tiled-prediction data path on the device.

  * ``TiledFrames``         - frames resident in HBM; tiles gathered by ``ds_crop_tiles`` in dataset order
                              (the crop of SplitDataset.__getitem__, data/split_dataset.py:239-249, with
                              SplitDatasetTiledPred.patch_location, data/split_dataset_tiledpred.py:27-32),
                              normalised and mixed like the reference (:198-204, 262-272).
  * ``get_tile_manager`` / ``get_tiling_dataset`` - the two ``predtiler`` entry points ``split.py:14,59-62``
    imports (the package itself is an un-vendored, un-pinned dependency; its observable behaviour - 490 tiles
    for (10,2048,2048)/256/512 - equals the in-repo ShiftBoundary manager).
"""
import numpy as np
import torch

from .. import _lib
from .tiling_manager import TileIndexManager, TilingMode


def get_tile_manager(data_shape, grid_shape, patch_shape, tiling_mode=TilingMode.ShiftBoundary):
    return TileIndexManager(tuple(data_shape), tuple(grid_shape), tuple(patch_shape), tiling_mode)


def get_tiling_dataset(dataset_class, tile_manager):
    """Subclass ``dataset_class`` so that it enumerates the manager's tiles (predtiler-shaped shim)."""

    class TilingDataset(dataset_class):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.tile_manager = tile_manager

        def __len__(self):
            return self.tile_manager.total_grid_count()

        def patch_location(self, index):
            return self.tile_manager.get_patch_location_from_dataset_idx(index)

    return TilingDataset


class TiledFrames:
    """Two-channel frame stack on the GPU, served as overlapping tiles.

    ``frames``: array/tensor (2, F, H, W), float32 or uint16.  ``normalization_dict`` uses the reference keys
    (``mean_input, std_input, mean_target, std_target``).  ``batch(first, n)`` returns CUDA fp32 tensors
    ``input`` (n,1,P,P) and ``target`` (n,2,P,P) identical to stacking ``SplitDatasetTiledPred[i]`` items.
    """

    def __init__(self, frames, patch_size, grid_size=None, normalization_dict=None, channel_weights=(1, 1),
                 input_from_normalized_target=False, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("diffsplit_b200: TiledFrames needs a CUDA device (no CPU fallback)")
        if isinstance(frames, np.ndarray):
            if frames.dtype == np.uint16:
                frames = torch.from_numpy(np.ascontiguousarray(frames)).to(device)
            else:
                frames = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.float32)).to(device)
        _lib.require_cuda(frames, "frames")
        self.frames = frames.contiguous()
        self.elem_size = self.frames.element_size()
        if self.elem_size not in (2, 4):
            raise TypeError("frames must be float32 or uint16")
        Cc, F, H, W = self.frames.shape
        self.C = Cc
        grid_size = grid_size or patch_size // 2
        self.patch_size = patch_size
        self.tile_manager = TileIndexManager((F, H, W), (1, grid_size, grid_size), (1, patch_size, patch_size),
                                             TilingMode.ShiftBoundary)
        nd = normalization_dict or {"mean_input": 0.0, "std_input": 1.0, "mean_target": np.zeros(Cc),
                                    "std_target": np.ones(Cc)}
        self.norm = nd
        self.weights = channel_weights
        self.input_from_normalized_target = input_from_normalized_target

    def __len__(self):
        return self.tile_manager.total_grid_count()

    def patch_location(self, index):
        return self.tile_manager.get_patch_location_from_dataset_idx(index)

    def raw_tiles(self, first, n):
        tiles = torch.empty((n, self.C, self.patch_size, self.patch_size), dtype=torch.float32, device=self.frames.device)
        d, g, p = self.tile_manager._c_shapes()
        with torch.cuda.device(self.frames.device):
            _lib.check(_lib.lib().ds_crop_tiles(self.frames.data_ptr(), self.elem_size, self.C, d, g, p,
                                                int(self.tile_manager.tiling_mode), first, n, tiles.data_ptr(),
                                                _lib.stream_ptr()))
        return tiles

    def _tile_norm(self):
        nd = self.norm
        mt = np.asarray(nd["mean_target"], dtype=np.float64).reshape(-1)
        st = np.asarray(nd["std_target"], dtype=np.float64).reshape(-1)
        if mt.size != 2 or st.size != 2:
            raise ValueError("TiledFrames.batch: mean_target / std_target must have one entry per channel (2)")
        w0, w1 = self.weights
        return _lib.TileNorm((_lib.C.c_double * 2)(*mt), (_lib.C.c_double * 2)(*st), float(nd["mean_input"]),
                             float(nd["std_input"]), float(w0), float(w1), int(bool(self.input_from_normalized_target)))

    def batch(self, first, n):
        """Tiles [first, first+n) as the stacked ``SplitDatasetTiledPred[i]`` items: one fused kernel (crop, float64
        normalisation with one rounding to fp32, channel mix) straight from the resident frames."""
        if self.C != 2:
            raise ValueError("TiledFrames.batch: the splitting data path has two channels")
        P = self.patch_size
        dev = self.frames.device
        inp = torch.empty((n, 1, P, P), dtype=torch.float32, device=dev)
        target = torch.empty((n, 2, P, P), dtype=torch.float32, device=dev)
        d, g, p = self.tile_manager._c_shapes()
        nm = self._tile_norm()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ds_tile_batch(self.frames.data_ptr(), self.elem_size, d, g, p,
                                                int(self.tile_manager.tiling_mode), first, n, _lib.C.byref(nm),
                                                inp.data_ptr(), target.data_ptr(), _lib.stream_ptr()))
        return inp, target

    def __getitem__(self, index):
        inp, target = self.batch(int(index), 1)
        return {"input": inp[0].cpu().numpy(), "target": target[0].cpu().numpy()}
