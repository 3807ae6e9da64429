"""Tile-row sharding of the rasteriser: strips rendered by N engines (rank r owns tile rows r, r+N, ...) must reassemble into
exactly the single-engine frame.  GPU part: N engines on one device (no NCCL needed to check the arithmetic); CPU part: the
gather + assembly plumbing over torch.distributed/gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, height, width, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gaussiansplats3d_b200.parallel import TILE, assemble_frame, owned_tile_rows, strip_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles_y = (height + TILE - 1) // TILE
    # the "frame": pixel value encodes its (GL) row and column, so any misplacement is visible
    full = np.zeros((tiles_y * TILE, width, 4), np.float32)
    full[..., 0] = np.arange(tiles_y * TILE)[:, None]
    full[..., 1] = np.arange(width)[None, :]
    rows = [r for t in owned_tile_rows(tiles_y, rank, world) for r in range(t * TILE, (t + 1) * TILE)]
    padded = ((tiles_y + world - 1) // world) * TILE
    strip = np.zeros((padded, width, 4), np.float32)
    strip[: len(rows)] = full[rows]
    assert strip_rows(height, rank, world) == len(rows)
    out = [torch.zeros(padded, width, 4) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(strip))                       # the final tile gather (NCCL on the GPUs)
    img = assemble_frame([o.numpy() for o in out], width, height, flip_y=True)
    ok = np.array_equal(np.asarray(img), full[:height][::-1])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [1080, 270, 33])
def test_gather_and_assemble_gloo_world2(height):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, height, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_strips_equal_single_engine_frame(gs, world):
    from gaussiansplats3d_b200.parallel import assemble_frame
    from gaussiansplats3d_b200.scenes import CAMERAS, synthetic_scene
    from gaussiansplats3d_b200.viewer import Viewer
    n, w, h = 120_000, 801, 455
    raw = synthetic_scene(n, seed=8, kind="bonsai", sh_degree=1)
    c = CAMERAS["bonsai"]
    opts = dict(cameraUp=c["up"], initialCameraPosition=c["position"], initialCameraLookAt=c["look_at"], width=w, height=h, sphericalHarmonicsDegree=1)
    v1 = Viewer(opts)
    v1.addSplatScene(raw)
    want = v1.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=True)
    v1.dispose()
    strips = []
    for r in range(world):
        v = Viewer(dict(opts, rank=r, world_size=world))
        v.addSplatScene(raw)
        strips.append(v.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=True).copy())
        v.dispose()
    padded = max(s.shape[0] for s in strips)
    strips = [np.concatenate([s, np.zeros((padded - s.shape[0], w, 4), s.dtype)]) for s in strips]
    got = np.asarray(assemble_frame(strips, w, h, flip_y=True))
    assert np.array_equal(got, want), "sharded strips do not reassemble into the single-GPU frame bit for bit"
