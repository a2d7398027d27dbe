"""``TileIndexManager`` with the reference's interface (data/tiling_manager.py:6-191), exact integer arithmetic.

The reference computes grid counts with ``np.ceil`` / ``np.floor`` of float quotients and returns numpy floats
from ``get_grid_index``; here everything is integer ``//``.  Rank-3 ``(F, H, W)`` managers (the only rank the
datasets create, data/split_dataset_tiledpred.py:17-22) additionally expose bulk tables computed by the native
library (``ds_tile_counts`` / ``ds_tile_patch_locations``), which is what the device gather / stitch kernels use.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib


class TilingMode:
    TrimBoundary = 0
    PadBoundary = 1
    ShiftBoundary = 2


def _cdiv(a, b):
    return -((-a) // b)


@dataclass
class TileIndexManager:
    data_shape: tuple
    grid_shape: tuple
    patch_shape: tuple
    tiling_mode: int

    def __post_init__(self):
        assert len(self.data_shape) == len(self.grid_shape), \
            f"Data shape:{self.data_shape} and grid size:{self.grid_shape} must have the same dimension"
        assert len(self.data_shape) == len(self.patch_shape), \
            f"Data shape:{self.data_shape} and patch shape:{self.patch_shape} must have the same dimension"
        for dim, (p, g) in enumerate(zip(self.patch_shape, self.grid_shape)):
            if p - g < 0:
                raise ValueError(f"Patch shape:{self.patch_shape} must be greater than or equal to grid shape:{self.grid_shape} in dimension {dim}")
            if (p - g) % 2 != 0:
                raise ValueError(f"Patch shape:{self.patch_shape} must have even padding in dimension {dim}")

    # ---- scalar interface (any rank) -------------------------------------------------------------
    def _check_dim(self, dim):
        assert dim < len(self.data_shape), f"Dimension {dim} is out of bounds for data shape {self.data_shape}"
        assert dim >= 0, "Dimension must be greater than or equal to 0"

    def patch_offset(self):
        return (np.array(self.patch_shape) - np.array(self.grid_shape)) // 2

    def get_individual_dim_grid_count(self, dim: int):
        self._check_dim(dim)
        D, G, P = int(self.data_shape[dim]), int(self.grid_shape[dim]), int(self.patch_shape[dim])
        if G == 1 and P == 1:
            return D
        if self.tiling_mode == TilingMode.PadBoundary:
            return _cdiv(D, G)
        if self.tiling_mode == TilingMode.ShiftBoundary:
            return _cdiv(D - (P - G), G)
        return (D - (P - G)) // G

    def grid_count(self, dim: int):
        self._check_dim(dim)
        n = 1
        for d in range(dim + 1, len(self.data_shape)):
            n *= self.get_individual_dim_grid_count(d)
        return n

    def total_grid_count(self):
        return self.grid_count(0) * self.get_individual_dim_grid_count(0)

    def get_grid_index(self, dim: int, coordinate: int):
        self._check_dim(dim)
        assert coordinate < self.data_shape[dim], f"Coordinate {coordinate} is out of bounds for data shape {self.data_shape}"
        G, P = int(self.grid_shape[dim]), int(self.patch_shape[dim])
        if G == 1 and P == 1:
            return coordinate
        if self.tiling_mode == TilingMode.PadBoundary:
            return coordinate // G
        excess = (P - G) // 2
        if self.tiling_mode == TilingMode.ShiftBoundary and coordinate + G + excess == self.data_shape[dim]:
            return self.get_individual_dim_grid_count(dim) - 1
        if self.tiling_mode in (TilingMode.TrimBoundary, TilingMode.ShiftBoundary):
            return max(0, (coordinate - excess) // G)
        raise ValueError(f"Unsupported tiling mode {self.tiling_mode}")

    def dataset_idx_from_grid_idx(self, grid_idx: tuple):
        assert len(grid_idx) == len(self.data_shape), \
            f"Dimension indices {grid_idx} must have the same dimension as data shape {self.data_shape}"
        return sum(int(grid_idx[d]) * self.grid_count(d) for d in range(len(grid_idx)))

    def get_gridstart_location_from_dim_index(self, dim: int, dim_index: int):
        self._check_dim(dim)
        n = self.get_individual_dim_grid_count(dim)
        assert dim_index < n, f"Dimension index {dim_index} is out of bounds for data shape {self.data_shape}"
        G, P = int(self.grid_shape[dim]), int(self.patch_shape[dim])
        if G == 1 and P == 1:
            return dim_index
        if self.tiling_mode == TilingMode.PadBoundary:
            return dim_index * G
        excess = (P - G) // 2
        if self.tiling_mode == TilingMode.TrimBoundary or dim_index < n - 1:
            return dim_index * G + excess
        if self.tiling_mode == TilingMode.ShiftBoundary:
            return int(self.data_shape[dim]) - G - excess
        raise ValueError(f"Unsupported tiling mode {self.tiling_mode}")

    def get_location_from_dataset_idx(self, dataset_idx: int):
        loc = []
        for dim in range(len(self.data_shape)):
            gc = self.grid_count(dim)
            loc.append(self.get_gridstart_location_from_dim_index(dim, dataset_idx // gc))
            dataset_idx = dataset_idx % gc
        return tuple(loc)

    def get_patch_location_from_dataset_idx(self, dataset_idx: int):
        grid_location = self.get_location_from_dataset_idx(dataset_idx)
        return tuple(np.array(grid_location) - np.array(self.patch_offset()))

    def get_dataset_idx_from_grid_location(self, location: tuple):
        assert len(location) == len(self.data_shape), \
            f"Location {location} must have the same dimension as data shape {self.data_shape}"
        return self.dataset_idx_from_grid_idx(tuple(self.get_grid_index(d, location[d]) for d in range(len(location))))

    def on_boundary(self, dataset_idx: int, dim: int, only_end: bool = False):
        self._check_dim(dim)
        if dim > 0:
            dataset_idx = dataset_idx % self.grid_count(dim - 1)
        dim_index = dataset_idx // self.grid_count(dim)
        last = self.get_individual_dim_grid_count(dim) - 1
        return dim_index == last if only_end else (dim_index == 0 or dim_index == last)

    def next_grid_along_dim(self, dataset_idx: int, dim: int):
        self._check_dim(dim)
        new_idx = dataset_idx + self.grid_count(dim)
        return None if new_idx >= self.total_grid_count() else new_idx

    def prev_grid_along_dim(self, dataset_idx: int, dim: int):
        self._check_dim(dim)
        new_idx = dataset_idx - self.grid_count(dim)
        return None if new_idx < 0 else None      # the reference falls off the end of this method: always None

    # ---- bulk native interface (rank 3) ---------------------------------------------------------------
    def _c_shapes(self):
        if len(self.data_shape) != 3:
            raise ValueError("native tiling supports (F, H, W) data only")
        return _lib.i3(self.data_shape), _lib.i3(self.grid_shape), _lib.i3(self.patch_shape)

    def native_counts(self):
        d, g, p = self._c_shapes()
        counts, total = (C.c_int32 * 3)(), C.c_int64()
        _lib.check(_lib.lib().ds_tile_counts(d, g, p, int(self.tiling_mode), counts, C.byref(total)))
        return tuple(counts), total.value

    def patch_locations(self, first=0, n=None) -> np.ndarray:
        """int32 [n,3] patch origins (f, h, w) of tiles [first, first+n) from the native library."""
        d, g, p = self._c_shapes()
        if n is None:
            n = self.native_counts()[1] - first
        out = np.empty((n, 3), dtype=np.int32)
        _lib.check(_lib.lib().ds_tile_patch_locations(d, g, p, int(self.tiling_mode), first, n,
                                                      out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out
