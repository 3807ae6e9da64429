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
        sv = s[: rows * TILE].reshape(rows, TILE, width, 4)   # strips may be padded to the longest rank's length
        v[r::world] = sv
    out = out[:height]
    if flip_y:
        out = out.flip(0) if hasattr(out, "flip") and not isinstance(out, np.ndarray) else out[::-1]
    return out


class _DevicePointer:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class TileGather:
    """The final tile gather: every rank's finished strips -> the whole frame on every rank, one NCCL all-gather enqueued on
    the ENGINE's stream (torch.cuda.ExternalStream), so it is ordered after the blend kernel without a host sync."""

    def __init__(self, engine, width: int, height: int, rank: int, world: int, fmt: int):
        import torch
        import torch.distributed as dist
        from . import _native as N
        self.torch, self.dist = torch, dist
        self.engine, self.width, self.height, self.rank, self.world = engine, width, height, rank, world
        self.bpp = 4 if fmt == N.GS_FRAME_RGBA8 else 16
        self.tiles_y = (height + TILE - 1) // TILE
        self.padded_rows = ((self.tiles_y + world - 1) // world) * TILE
        ptr, nbytes = engine.buffer_dev(N.GS_BUF_FRAME)
        need = self.padded_rows * width * self.bpp
        dev = torch.device("cuda", torch.cuda.current_device())
        raw = torch.as_tensor(_DevicePointer(ptr, max(need, 1)), device=dev)
        self.strip = raw[:need]
        self.gathered = torch.empty((world, need), dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.ExternalStream(engine.stream(), device=dev)

    def all_gather(self) -> None:
        with self.torch.cuda.stream(self.stream):
            self.dist.all_gather_into_tensor(self.gathered.view(-1), self.strip)

    def sync_to_torch(self) -> None:
        self.torch.cuda.current_stream().wait_stream(self.stream)

    def assemble(self, flip_y: bool = True):
        """[world, strips] -> [height, width, bpp] image tensor on the device (tile row r + k*world lives at [r, k])."""
        t = self.gathered.view(self.world, self.padded_rows // TILE, TILE, self.width, self.bpp)
        img = t.permute(1, 0, 2, 3, 4).reshape(-1, self.width, self.bpp)[: self.height]
        return img.flip(0) if flip_y else img
