"""Multi-GPU plumbing for the hot path: one process per GPU (torch.distributed), the rasteriser sharded by screen tile, the frame
assembled with ONE collective (the only one on the data path).

Ownership: coarse tile (cx, cy) (128x64 px) belongs to rank (cx + cy) % world -- a diagonal interleave that spreads the dense centre
and the empty border over all ranks.  Every rank holds the whole scene (0.7 GB for 16M SH0 splats: replicas only) and sorts the full
depth list (the order restricted to a rank's tiles is the global order), so no splat data ever crosses NVLink.  Each
rank writes its tiles into a full-size frame that is zero elsewhere; the frames are summed with one NCCL all-reduce on the ENGINE's
stream ("the final tile gather").
"""
from __future__ import annotations

import numpy as np

TILE = 16
COARSE_W, COARSE_H = 128, 64   # pixels (8 x 4 fine tiles), csrc/raster_kernels.cuh kCoarseW/kCoarseH


def owner_of_coarse(cx: int, cy: int, world: int) -> int:
    return (cx + cy) % world if world > 1 else 0


def ownership_map(width: int, height: int, world: int) -> np.ndarray:
    """[height, width] array of the rank that rasterises each pixel (GL row order)."""
    ys, xs = np.mgrid[0:height, 0:width]
    return ((xs // COARSE_W + ys // COARSE_H) % max(world, 1)).astype(np.int32)


def combine_frames(frames: list):
    """Sum of the per-rank frames (what the all-reduce computes); numpy arrays or torch tensors."""
    out = frames[0].copy() if isinstance(frames[0], np.ndarray) else frames[0].clone()
    for f in frames[1:]:
        out += f
    return out


class _DevicePointer:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class TileGather:
    """The final tile gather: every rank's finished tiles -> the whole frame on every rank, one NCCL all-reduce (SUM of frames that
    are zero outside their own tiles) enqueued on the ENGINE's stream (torch.cuda.ExternalStream), so it is ordered after the blend
    kernel without a host sync."""

    def __init__(self, engine, width: int, height: int, rank: int, world: int, fmt: int):
        import torch
        import torch.distributed as dist
        from . import _native as N
        self.torch, self.dist = torch, dist
        self.engine, self.width, self.height, self.rank, self.world = engine, width, height, rank, world
        self.rgba8 = fmt == N.GS_FRAME_RGBA8
        ptr, _ = engine.buffer_dev(N.GS_BUF_FRAME)
        nbytes = width * height * (4 if self.rgba8 else 16)
        dev = torch.device("cuda", torch.cuda.current_device())
        raw = torch.as_tensor(_DevicePointer(ptr, nbytes), device=dev)
        self.frame = raw if self.rgba8 else raw.view(torch.float32)      # summed in place: afterwards every rank holds the picture
        self.stream = torch.cuda.ExternalStream(engine.stream(), device=dev)

    def all_reduce(self) -> None:
        with self.torch.cuda.stream(self.stream):
            self.dist.all_reduce(self.frame, op=self.dist.ReduceOp.SUM)

    all_gather = all_reduce   # historical name used by bench.py

    def sync_to_torch(self) -> None:
        self.torch.cuda.current_stream().wait_stream(self.stream)

    def image(self):
        """[height, width, 4] view of the assembled frame (uint8 or float32) on the device."""
        return self.frame.view(self.height, self.width, -1) if self.rgba8 else self.frame.view(self.height, self.width, 4)


class PeerGather:
    """The final tile gather fused into the blend kernel: ranks > 0 store their finished pixels directly into rank 0's frame buffer
    over NVLink (CUDA IPC mapping, device-side release/arrive handshake); no collective call at all on the data path.  The IPC
    handles travel once through torch.distributed at set-up.  After a frame, only rank 0 holds the picture."""

    def __init__(self, engine, rank: int, world: int):
        import torch.distributed as dist
        self.engine, self.rank, self.world = engine, rank, world
        payload = [None]
        if rank == 0:
            payload = [engine.peer_export()]
        dist.broadcast_object_list(payload, src=0)
        if rank != 0:
            engine.peer_attach(*payload[0])
        dist.barrier()

    def all_gather(self) -> None:   # nothing to do: the gather happened inside the blend kernel
        pass

    all_reduce = all_gather
