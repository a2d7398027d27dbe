"""Tiling side of the hot path: tile index manager, device tile gather, bit-exact stitching."""
from .tiling_manager import TileIndexManager, TilingMode  # noqa: F401
from .tile_stitcher import stitch_predictions  # noqa: F401
from .tiled_pred import TiledFrames, get_tile_manager, get_tiling_dataset  # noqa: F401
