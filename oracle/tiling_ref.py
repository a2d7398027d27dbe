"""Integer restatement of the reference tiling path (oracle; test infrastructure).

  * tile index arithmetic   data/tiling_manager.py:14-191
  * tiled dataset contract  data/split_dataset_tiledpred.py:9-32, data/split_dataset.py:237-278
  * stitching               data/tile_stitcher.py:10-81

The reference evaluates ``np.ceil``/``np.floor`` of float quotients; here every
count is exact integer arithmetic.  ``oracle/make_golden.py`` checks equality
with the reference over the full index range of many shapes.
"""
from typing import Sequence, Tuple

import numpy as np

TRIM, PAD, SHIFT = 0, 1, 2       # TilingMode (tiling_manager.py:6-12)


def _cdiv(a: int, b: int) -> int:
    return -((-a) // b)


class TileGrid:
    """Pure-integer equivalent of TileIndexManager."""

    def __init__(self, data_shape: Sequence[int], grid_shape: Sequence[int],
                 patch_shape: Sequence[int], mode: int = SHIFT):
        self.data = tuple(int(v) for v in data_shape)
        self.grid = tuple(int(v) for v in grid_shape)
        self.patch = tuple(int(v) for v in patch_shape)
        self.mode = mode
        if not (len(self.data) == len(self.grid) == len(self.patch)):
            raise AssertionError("rank mismatch")
        for d, (p, g) in enumerate(zip(self.patch, self.grid)):
            if p < g:
                raise ValueError(f"patch < grid in dim {d}")
            if (p - g) % 2:
                raise ValueError(f"odd padding in dim {d}")
        self.nd = len(self.data)
        self.offset = tuple((p - g) // 2 for p, g in zip(self.patch, self.grid))   # :31-32
        self.counts = tuple(self._count(d) for d in range(self.nd))              # :34-50
        st = [1] * self.nd                                                        # :58-67
        for d in range(self.nd - 2, -1, -1):
            st[d] = st[d + 1] * self.counts[d + 1]
        self.strides = tuple(st)
        self.total = self.strides[0] * self.counts[0]                             # :52-56

    def _count(self, d: int) -> int:
        D, G, P = self.data[d], self.grid[d], self.patch[d]
        if G == 1 and P == 1:
            return D
        if self.mode == PAD:
            return _cdiv(D, G)
        if self.mode == SHIFT:
            return _cdiv(D - (P - G), G)
        return (D - (P - G)) // G

    def _grid_start(self, d: int, k: int) -> int:                                 # :120-143
        G, P = self.grid[d], self.patch[d]
        if G == 1 and P == 1:
            return k
        if self.mode == PAD:
            return k * G
        ex = (P - G) // 2
        if self.mode == TRIM or k < self.counts[d] - 1:
            return k * G + ex
        return self.data[d] - G - ex

    def grid_index(self, idx: int) -> Tuple[int, ...]:                            # :149-152
        out = []
        for d in range(self.nd):
            out.append(idx // self.strides[d])
            idx %= self.strides[d]
        return tuple(out)

    def grid_location(self, idx: int) -> Tuple[int, ...]:                         # :145-154
        return tuple(self._grid_start(d, k) for d, k in enumerate(self.grid_index(idx)))

    def patch_location(self, idx: int) -> Tuple[int, ...]:                        # :106-112
        return tuple(g - o for g, o in zip(self.grid_location(idx), self.offset))

    def patch_table(self) -> np.ndarray:
        return np.array([self.patch_location(i) for i in range(self.total)], dtype=np.int64).reshape(self.total, self.nd)

    def copy_boxes(self, idx: int):
        """(dst_start, dst_end, src_start) of the region tile ``idx`` contributes (tile_stitcher.py:28-57)."""
        gs = self.grid_location(idx)
        ps = self.patch_location(idx)
        vs, ve = list(gs), [g + s for g, s in zip(gs, self.grid)]
        if self.mode == SHIFT:
            for d in range(self.nd):
                if ps[d] == 0:
                    vs[d] = 0
                if ps[d] + self.patch[d] == self.data[d]:
                    ve[d] = self.data[d]
        rs = [v - p for v, p in zip(vs, ps)]
        return vs, ve, rs


def stitch(predictions: np.ndarray, tg: TileGrid) -> np.ndarray:
    """(N,C,P,P) tiles -> (F,H,W,C) frames; rank-3 data only (the path the reference tests)."""
    assert tg.nd == 3
    out = np.zeros(tg.data + (predictions.shape[1],), dtype=predictions.dtype)
    for i in range(predictions.shape[0]):
        vs, ve, rs = tg.copy_boxes(i)
        hh, ww = ve[1] - vs[1], ve[2] - vs[2]
        for c in range(predictions.shape[1]):
            out[vs[0]:ve[0], vs[1]:ve[1], vs[2]:ve[2], c] = \
                predictions[i, c, rs[1]:rs[1] + hh, rs[2]:rs[2] + ww]
    return out


def crop_tiles(frames: np.ndarray, tg: TileGrid, idxs=None) -> np.ndarray:
    """frames (C,F,H,W) -> tiles (N,C,P,P) float32 as SplitDataset.__getitem__ crops them
    (split_dataset.py:239-249), before normalisation."""
    idxs = range(tg.total) if idxs is None else idxs
    P = tg.patch[1]
    out = []
    for i in idxs:
        f, h, w = tg.patch_location(i)
        out.append(frames[:, f, h:h + P, w:w + P].astype(np.float32))
    return np.stack(out)


def normalise_and_mix(tiles: np.ndarray, mean_t, std_t, mean_in, std_in, weights=(1, 1), input_from_normalized_target=False):
    """target = (tile-mean_t)/std_t ; input = (w0*ch0 + w1*ch1 - mean_in)/std_in, or w0*target0 + w1*target1 when
    ``input_from_normalized_target`` (split_dataset.py:195-201, 262-272), float32 results.  The normalisation constants are
    numpy float64 in the reference (compute_normalization_dict :28-75: np.quantile(...)/2, np.array([...])), so the
    float32 patches promote to float64 and are rounded once by ``astype``; the python-scalar channel weights do not
    promote (float32 products and sum)."""
    mt = np.asarray(mean_t, dtype=np.float64).reshape(1, -1, 1, 1)
    st = np.asarray(std_t, dtype=np.float64).reshape(1, -1, 1, 1)
    tiles = tiles.astype(np.float32)
    target = ((tiles - mt) / st).astype(np.float32)
    w0, w1 = (np.float32(w) for w in weights)
    if input_from_normalized_target:
        inp = w0 * target[:, 0:1] + w1 * target[:, 1:2]
    else:
        inp = w0 * tiles[:, 0:1] + w1 * tiles[:, 1:2]
        inp = ((inp - np.float64(mean_in)) / np.float64(std_in)).astype(np.float32)
    return inp, target
