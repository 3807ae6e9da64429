"""Coarse-tile sharding of the rasteriser: rank r renders the 128x64-px tiles with (cx + cy) % N == r into a full-size frame that is
zero elsewhere; the ranks' frames must SUM to exactly the single-engine frame.  GPU part: N engines on one device (no NCCL needed to
check the arithmetic); CPU part: the reduction plumbing over torch.distributed/gloo with world_size 2 and the ownership map."""
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
    from gaussiansplats3d_b200.parallel import ownership_map
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the "frame": pixel value encodes its row and column, so any misplacement is visible
    full = np.zeros((height, width, 4), np.int32)
    full[..., 0] = np.arange(height)[:, None] % 251
    full[..., 1] = np.arange(width)[None, :] % 251
    full[..., 3] = 255
    mine = ownership_map(width, height, world) == rank
    part = np.where(mine[..., None], full, 0)
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                            # the final tile gather (NCCL on the GPUs)
    q.put((rank, bool(np.array_equal(t.numpy(), full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [1080, 270, 33])
def test_all_reduce_assembles_frame_gloo_world2(height):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, height, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_ownership_map_is_a_partition():
    from gaussiansplats3d_b200.parallel import COARSE_H, COARSE_W, owner_of_coarse, ownership_map
    for world in (1, 2, 3, 4, 8):
        m = ownership_map(1920, 1080, world)
        assert m.min() == 0 and m.max() == world - 1
        assert m[0, 0] == owner_of_coarse(0, 0, world) and m[1079, 1919] == owner_of_coarse(1919 // COARSE_W, 1079 // COARSE_H, world)
        share = np.bincount(m.ravel(), minlength=world) / m.size
        assert share.max() - share.min() < 0.08, "diagonal interleave should balance the screen area"


@pytest.mark.gpu
@pytest.mark.parametrize("subset_sort", [False, True])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_frames_sum_to_single_engine_frame(gs, world, subset_sort, monkeypatch):
    # both sort strategies of a sharded frame: replicated full sort (small scenes) and per-rank subset sort (default from 3M splats)
    monkeypatch.setenv("GS_SUBSET_MIN", "1" if subset_sort else "4000000000")
    from gaussiansplats3d_b200.parallel import combine_frames, ownership_map
    from gaussiansplats3d_b200.scenes import CAMERAS, synthetic_scene
    from gaussiansplats3d_b200.viewer import Viewer
    n, w, h = 120_000, 801, 455
    raw = synthetic_scene(n, seed=8, kind="bonsai", sh_degree=1)
    c = CAMERAS["bonsai"]
    opts = dict(cameraUp=c["up"], initialCameraPosition=c["position"], initialCameraLookAt=c["look_at"], width=w, height=h, sphericalHarmonicsDegree=1)
    v1 = Viewer(opts)
    v1.addSplatScene(raw)
    want = v1.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=False)
    v1.dispose()
    own = ownership_map(w, h, world)
    frames = []
    for r in range(world):
        v = Viewer(dict(opts, rank=r, world_size=world))
        v.addSplatScene(raw)
        f = v.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=False).copy()
        assert not f[own != r].any(), "a rank wrote pixels outside its own coarse tiles"
        frames.append(f)
        v.dispose()
    got = combine_frames(frames)
    assert np.array_equal(got, want), "sharded frames do not sum to the single-GPU frame bit for bit"


def _peer_worker(rank, world, port, q):
    """One process per GPU: rank 0 exports its frame through CUDA IPC, the others blend into it (fused tile gather)."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gaussiansplats3d_b200 as gs
    from gaussiansplats3d_b200.parallel import PeerGather
    from gaussiansplats3d_b200.scenes import CAMERAS, synthetic_scene
    from gaussiansplats3d_b200.viewer import Viewer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n, w, h = 100_000, 801, 455
    raw = synthetic_scene(n, seed=8, kind="bonsai", sh_degree=0)
    c = CAMERAS["bonsai"]
    opts = dict(cameraUp=c["up"], initialCameraPosition=c["position"], initialCameraLookAt=c["look_at"], width=w, height=h, device=rank)
    want = None
    if rank == 0:
        v1 = Viewer(opts)
        v1.addSplatScene(raw)
        want = v1.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=True).copy()
        v1.dispose()
    v = Viewer(dict(opts, rank=rank, world_size=world))
    v.addSplatScene(raw)
    PeerGather(v.engine, rank, world)
    ok = True
    for it in range(3):     # several frames: the release/arrive handshake must hold across frames (graph replay included)
        got = v.frame(frame_format=gs._native.GS_FRAME_RGBA8, flip_y=True)
        if rank == 0:
            ok = ok and bool(np.array_equal(got, want))
    dist.barrier()
    q.put((rank, ok))
    v.dispose()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_fused_peer_gather_two_gpus(gs):
    """Needs 2 GPUs (skipped on a 1-GPU box): the picture assembled in rank 0's memory by peer stores equals the single-GPU frame."""
    if gs._native.load().gs_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
