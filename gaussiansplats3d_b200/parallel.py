"""Multi-GPU plumbing for the hot path: one process per GPU (torch.distributed), the rasteriser sharded by screen tile
row, the frame assembled with one all-gather of finished tile strips (the only collective on the data path).

Ownership is row-interleaved (tile row r belongs to rank r % world) so that near and far rows of the picture spread
evenly; every rank holds the whole scene (0.7 GB for 16M SH0 splats -- replicas only) and sorts the full depth list
(the order restricted to a rank's tiles is the global order), so no splat data ever crosses NVLink.
"""
from __future__ import annotations

import numpy as np

TILE = 16


def owned_tile_rows(tiles_y: int, rank: int, world: int) -> range:
    return range(rank, tiles_y, world)


def strip_rows(height: int, rank: int, world: int) -> int:
    """pixel rows in the compact strip buffer rank `rank` produces (whole tiles)."""
    return len(owned_tile_rows((height + TILE - 1) // TILE, rank, world)) * TILE


def assemble_frame(strips: list, width: int, height: int, flip_y: bool = True):
    """strips[r]: array/tensor [strip_rows(r), width, 4] (GL row order inside each tile row) -> full frame.
    Works on numpy arrays and torch tensors alike (uses only slicing / assignment)."""
    world = len(strips)
    tiles_y = (height + TILE - 1) // TILE
    first = strips[0]
    out = first.new_empty((tiles_y * TILE, width, 4)) if hasattr(first, "new_empty") else np.empty((tiles_y * TILE, width, 4), first.dtype)
    v = out.reshape(tiles_y, TILE, width, 4) if not hasattr(out, "view") or isinstance(out, np.ndarray) else out.view(tiles_y, TILE, width, 4)
    for r, s in enumerate(strips):
        rows = len(owned_tile_rows(tiles_y, r, world))
        sv = s.reshape(rows, TILE, width, 4)
        v[r::world] = sv[:rows]
    out = out[:height]
    if flip_y:
        out = out.flip(0) if hasattr(out, "flip") and not isinstance(out, np.ndarray) else out[::-1]
    return out
