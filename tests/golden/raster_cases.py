"""Seeded raster cases shared by tests/golden/make_raster_golden.py (writes the fixture) and the tests that read it.

The raster stage has no reference-produced vectors (the reference needs a browser + WebGL: "parity unpinned"), so this
fixture pins the CPU restatement (oracle/raster_oracle.c) against silent change and gives the GPU tests a committed target; the draw
order inside it comes from the sort oracle, which IS pinned by the compiled reference."""
import numpy as np

CASES = {
    # name: (splats, seed, kind, sh_degree, width, height, camera)
    "bonsai-sh0-160x100": (4000, 7, "bonsai", 0, 160, 100, "bonsai"),
    "bonsai-sh2-128x96": (3000, 8, "bonsai", 2, 128, 96, "bonsai"),
    "garden-sh1-200x120": (5000, 9, "garden", 1, 200, 120, "garden"),
}


class _NoEngine:
    """Stands in for the CUDA engine so the Viewer's host code (camera, uniforms, packing) can run where there is no GPU."""

    def __init__(self, *a, **k):
        pass

    def upload_splat_data(self, *a, **k):
        pass

    def upload_centers(self, *a, **k):
        pass


def host_viewer(name):
    """(viewer without a device engine, raw scene) for a case: the same Viewer code path the GPU tests drive."""
    import gaussiansplats3d_b200.viewer as V
    from gaussiansplats3d_b200.scenes import CAMERAS, synthetic_scene
    n, seed, kind, sh, w, h, cam = CASES[name]
    raw = synthetic_scene(n, seed=seed, kind=kind, sh_degree=sh)
    c = CAMERAS[cam]
    real = V.Engine
    V.Engine = _NoEngine
    try:
        v = V.Viewer(dict(cameraUp=c["up"], initialCameraPosition=c["position"], initialCameraLookAt=c["look_at"], width=w, height=h, sphericalHarmonicsDegree=sh))
        v.addSplatScene(raw)
    finally:
        V.Engine = real
    v.camera.update()
    v.updateSplatMesh()
    return v, raw


def oracle_outputs(name, oracle):
    """(draw order, float frame bottom-up, projected records) from the CPU restatements for a case."""
    v, raw = host_viewer(name)
    n = raw.count
    p = v.splatMesh.packed
    mvp = v.mvp_matrix().astype(np.float32)
    order = oracle.port_sort_indexes(np.arange(n, dtype=np.uint32), p.int_centers, None, mvp, None, None, 1 << 16, n, n, n, False, True, False)
    frame, ps = oracle.render(v.uniforms(), p.centers_colors, p.covariances, order, v.renderWidth, v.renderHeight, sh=p.sh, sh_degree=p.sh_degree)
    return order, frame, ps
