"""CPU: host-side mirror (three.js math, packing, message protocol shapes) and the raster oracle's self-consistency."""
import numpy as np
import pytest

import cases


def test_three_math_invert_and_perspective():
    from gaussiansplats3d_b200 import three_math as TM
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(0, 1, (4, 4)) + 3 * np.eye(4)
        e = TM.to_elements(m)
        assert np.allclose(TM.to_mat(TM.invert(e)), np.linalg.inv(m), rtol=1e-10, atol=1e-12)
        assert np.allclose(TM.to_mat(TM.multiply(e, TM.invert(e))), np.eye(4), atol=1e-10)
    p = TM.make_perspective(50, 16 / 9, 0.1, 1000)
    assert np.allclose(p, cases.perspective()), "matches the independent restatement in tests/cases.py"
    assert np.isclose(p[5], 1 / np.tan(np.deg2rad(25)))
    w = TM.camera_world_matrix((0, 10, 15), (0, 0, 0), (0, 1, 0))
    assert np.allclose(w, cases.look_at_world((0, 10, 15), (0, 0, 0), (0, 1, 0)))
    # camera looks down -Z of its own frame
    fwd = -TM.to_mat(w)[:3, 2]
    assert np.allclose(fwd, np.array([0, -10, -15]) / np.hypot(10, 15))


def test_compose_matches_rotation_then_scale():
    from gaussiansplats3d_b200 import three_math as TM
    q = np.array([0.1, 0.2, 0.3, 0.9]); q /= np.linalg.norm(q)
    m = TM.to_mat(TM.compose((1, 2, 3), q, (2, 3, 4)))
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    assert np.allclose(m[:3, :3], R @ np.diag([2, 3, 4])) and np.allclose(m[:3, 3], [1, 2, 3])


def test_packing_matches_reference_layout(oracle_mod):
    from gaussiansplats3d_b200.scenes import synthetic_scene, pack_scene, integer_centers
    raw = synthetic_scene(5000, seed=1, sh_degree=2)
    p = pack_scene(raw)
    assert p.centers_colors.dtype == np.uint32 and p.centers_colors.shape == (5000, 4)
    assert np.array_equal(p.centers_colors[:, 1:].view(np.float32), raw.centers)
    rgba = p.centers_colors[:, 0]
    assert np.array_equal(rgba & 255, raw.colors[:, 0]) and np.array_equal((rgba >> 16) & 255, raw.colors[:, 2])
    a = raw.colors[:, 3].astype(np.uint32)
    assert np.array_equal(rgba >> 24, np.where(a >= 1, a, 0))
    # covariance = R S S R^T, symmetric PSD, 6 unique entries
    c = p.covariances.astype(np.float64)
    full = np.stack([np.stack([c[:, 0], c[:, 1], c[:, 2]], 1), np.stack([c[:, 1], c[:, 3], c[:, 4]], 1), np.stack([c[:, 2], c[:, 4], c[:, 5]], 1)], 1)
    ev = np.linalg.eigvalsh(full)
    assert np.allclose(np.sort(ev, 1), np.sort(raw.scales.astype(np.float64) ** 2, 1), rtol=2e-2, atol=1e-8)
    assert p.sh.dtype == np.float16 and p.sh.shape == (5000, 24)
    # integer centres agree with the oracle's restatement of getIntegerCenters
    assert np.array_equal(integer_centers(raw.centers), oracle_mod.integer_centers(raw.centers))


def test_raster_oracle_single_splat_analytic(oracle_mod):
    """One isotropic splat at the optical axis: the oracle's alpha profile must be exp(-r^2 / (2 sigma_px^2)) * a."""
    from gaussiansplats3d_b200.engine import Uniforms
    from gaussiansplats3d_b200 import three_math as TM
    from gaussiansplats3d_b200.scenes import pack_centers_colors
    W, H = 200, 120
    cam = TM.PerspectiveCamera(50, W / H, 0.1, 1000)
    cam.position = np.array([0, 0, 5.0]); cam.look_at((0, 0, 0))
    sigma = 0.05
    cc = pack_centers_colors(np.zeros((1, 3), np.float32), np.array([[255, 128, 0, 255]], np.uint8))
    cov = np.array([[sigma**2, 0, 0, sigma**2, 0, sigma**2]], np.float32)
    fx = cam.projectionMatrix[0] * 0.5 * W; fy = cam.projectionMatrix[5] * 0.5 * H
    u = Uniforms(model_view=cam.matrixWorldInverse.astype(np.float32), projection=cam.projectionMatrix.astype(np.float32), camera_position=cam.position.astype(np.float32),
                 focal=(fx, fy), viewport=(W, H))
    frame, ps = oracle_mod.render(u, cc, cov, np.zeros(1, np.uint32), W, H)
    assert ps["valid"][0] == 1 and abs(ps["cx"][0] - W / 2) < 1e-3 and abs(ps["cy"][0] - H / 2) < 1e-3
    var_px = (sigma * fx / 5.0) ** 2 + 0.3
    ys, xs = np.mgrid[0:H, 0:W]
    r2 = (xs + 0.5 - W / 2) ** 2 + (ys + 0.5 - H / 2) ** 2
    want_a = np.where(r2 / var_px <= 8.0 * (1 + 1e-5), np.exp(-0.5 * r2 / var_px), 0.0)
    # the eigen split clamps term2 >= sqrt(0.1): an isotropic splat gets lambda = var +- 0.316 -> compare loosely
    inside = r2 / var_px < 4.0
    assert np.abs(frame[..., 3] - want_a)[inside].max() < 0.12
    assert frame[..., 3].max() > 0.85 and frame[0, 0, 3] == 0.0
    assert np.allclose(frame[..., 0][inside], frame[..., 3][inside] * 1.0, atol=1e-5)       # premultiplied red = alpha * 1.0
    assert np.allclose(frame[..., 1][inside], frame[..., 3][inside] * (128 / 255), atol=1e-5)


def test_blend_order_dependence(oracle_mod):
    """'over' compositing: the later splat in the draw order wins; alpha accumulates as 1 - prod(1 - a)."""
    from gaussiansplats3d_b200 import _native as N
    ps = np.zeros(2, N.PROJECTED_DTYPE)
    for i, (r, g) in enumerate(((1.0, 0.0), (0.0, 1.0))):
        ps[i] = (10.5, 10.5, 6.0, 0.0, 0.0, 6.0, r, g, 0.0, 0.8, 0.0, 1)
    f01 = oracle_mod.blend(ps, np.array([0, 1], np.uint32), 21, 21)
    f10 = oracle_mod.blend(ps, np.array([1, 0], np.uint32), 21, 21)
    c = f01[10, 10]
    assert np.isclose(c[3], 1 - 0.2 * 0.2, atol=1e-6) and np.isclose(c[1], 0.8, atol=1e-6) and np.isclose(c[0], 0.8 * 0.2, atol=1e-6)
    assert np.isclose(f10[10, 10][0], 0.8, atol=1e-6)
    assert f01[0, 0, 3] == 0.0  # outside the quad's inscribed disc


# ---- static scene transform baked at load (scenes.transform_scene / pack_scene(transform16=...)) ------------------------------------
def _demo_transform():
    from gaussiansplats3d_b200 import three_math as TM
    q = np.array([0.31, -0.52, 0.18, 0.77])
    q /= np.linalg.norm(q)
    return TM.compose((0.7, -1.1, 0.4), q, (1.6, 1.6, 1.6)), q


def test_baked_transform_matches_scalar_restatement():
    """Vectorised packing vs the per-splat restatement of SplatBuffer.js (oracle/pack_oracle.py): centres, covariances, SH bands 1 and 2."""
    from oracle import pack_oracle as PO
    from gaussiansplats3d_b200.scenes import compute_covariances, rotation_of_transform, synthetic_scene, transform_scene
    raw = synthetic_scene(300, seed=5, sh_degree=2)
    T, _ = _demo_transform()
    baked, t3 = transform_scene(raw, T)
    cov = compute_covariances(baked.scales, baked.rotations, t3)
    R = rotation_of_transform(T)
    for i in range(raw.count):
        c, cov6, sh = PO.bake_one([float(v) for v in raw.centers[i]], [float(v) for v in raw.scales[i]], [float(v) for v in raw.rotations[i]],
                                  [[float(v) for v in t] for t in raw.sh[i]], 2, [float(v) for v in T], R.tolist())
        assert np.allclose(baked.centers[i], c, rtol=1e-6, atol=1e-6)
        assert np.allclose(cov[i], cov6, rtol=2e-6, atol=1e-9)
        assert np.allclose(baked.sh[i], np.array(sh), rtol=1e-5, atol=1e-6)


def test_sh_rotation_composes_and_inverts():
    """Rotating by A then B equals rotating by B*A; the identity rotates nothing; band matrices are orthogonal."""
    from gaussiansplats3d_b200.scenes import sh_rotation_matrices
    rng = np.random.default_rng(3)
    qa, qb = rng.normal(size=4), rng.normal(size=4)

    def rot(q):
        x, y, z, w = q / np.linalg.norm(q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    A, B = rot(qa), rot(qb)
    a1, a2 = sh_rotation_matrices(A)
    b1, b2 = sh_rotation_matrices(B)
    c1, c2 = sh_rotation_matrices(B @ A)
    assert np.allclose(b1 @ a1, c1, atol=1e-12) and np.allclose(b2 @ a2, c2, atol=1e-12)
    assert np.allclose(a2 @ a2.T, np.eye(5), atol=1e-12)
    i1, i2 = sh_rotation_matrices(np.eye(3))
    assert np.allclose(i1, np.eye(3)) and np.allclose(i2, np.eye(5))


@pytest.mark.parametrize("degree", [0, 1, 2])
def test_baked_transform_renders_like_the_same_transform_as_model_matrix(oracle_mod, degree):
    """The rendering invariant that pins the baking (incl. the direction of the SH rotation): a scene with transform T baked in, seen
    by camera V, must give the picture of the untouched scene drawn with model matrix T (modelView = V*T, camera position T^-1 * eye)
    -- which is also exactly what the reference's dynamic mode does in the shader."""
    from gaussiansplats3d_b200 import three_math as TM
    from gaussiansplats3d_b200.engine import Uniforms
    from gaussiansplats3d_b200.scenes import pack_scene, synthetic_scene
    W, H = 320, 200
    raw = synthetic_scene(3000, seed=9, kind="bonsai", sh_degree=degree)
    T, _ = _demo_transform()
    cam = TM.PerspectiveCamera(50, W / H, 0.1, 1000)
    cam.position = np.array([2.0, 3.5, -11.0]); cam.look_at((0.5, -1.0, 0.5))
    fx, fy = cam.projectionMatrix[0] * 0.5 * W, cam.projectionMatrix[5] * 0.5 * H
    base = dict(projection=cam.projectionMatrix.astype(np.float32), focal=(fx, fy), viewport=(W, H), sh_degree=degree)
    baked = pack_scene(raw, sh_format="f32", transform16=T)
    plain = pack_scene(raw, sh_format="f32")
    # one draw order for both pictures: back to front by view-space depth of the transformed centres
    V = TM.to_mat(cam.matrixWorldInverse)
    centres_t = baked.centers_colors[:, 1:].view(np.float32).astype(np.float64)
    order = np.argsort((centres_t @ V[:3, :3].T + V[:3, 3])[:, 2], kind="stable").astype(np.uint32)
    ua = Uniforms(model_view=cam.matrixWorldInverse.astype(np.float32), camera_position=cam.position.astype(np.float32), **base)
    eye_model = TM.to_mat(TM.invert(T)) @ np.append(cam.position, 1.0)
    ub = Uniforms(model_view=TM.multiply(cam.matrixWorldInverse, T).astype(np.float32), camera_position=eye_model[:3].astype(np.float32), **base)
    fa, pa = oracle_mod.render(ua, baked.centers_colors, baked.covariances, order, W, H, sh=baked.sh, sh_degree=baked.sh_degree)
    fb, pb = oracle_mod.render(ub, plain.centers_colors, plain.covariances, order, W, H, sh=plain.sh, sh_degree=plain.sh_degree)
    assert pa["valid"].sum() > 1000 and np.array_equal(pa["valid"], pb["valid"])
    assert np.abs(fa - fb).max() < 4e-3 and np.abs(fa - fb).mean() < 2e-5
    if degree > 0:   # the test has teeth: without the SH rotation the colours differ visibly
        unrot = pack_scene(raw, sh_format="f32")
        fc, _ = oracle_mod.render(ua, baked.centers_colors, baked.covariances, order, W, H, sh=unrot.sh, sh_degree=degree)
        assert np.abs(fa - fc).max() > 2e-2


def test_splat_mesh_bakes_transform_only_when_static():
    """SplatMesh.js:1872-1883: a static mesh bakes the scene transform (the sorter's centres included); a dynamic one does not."""
    from gaussiansplats3d_b200 import three_math as TM
    from gaussiansplats3d_b200.scenes import synthetic_scene
    from gaussiansplats3d_b200.viewer import SplatMesh
    raw = synthetic_scene(200, seed=2, sh_degree=1)
    T = TM.compose((1.0, 2.0, 3.0), (0.0, 0.0, 0.0, 1.0), (2.0, 2.0, 2.0))
    static = SplatMesh(sphericalHarmonicsDegree=1)
    static.build(raw, transform16=T)
    want = raw.centers * 2.0 + np.array([1.0, 2.0, 3.0], np.float32)
    assert np.allclose(static.getFloatCenters(0, 199)[:, :3], want, atol=1e-5)
    assert np.array_equal(static.getIntegerCenters(0, 199, True)[:, :3], np.floor(static.raw.centers.astype(np.float64) * 1000.0 + 0.5).astype(np.int32))
    assert np.allclose(static.packed.centers_colors[:, 1:].view(np.float32), want, atol=1e-5)
    assert np.allclose(static.packed.covariances, 4.0 * _plain_cov(raw), rtol=1e-5)      # T3 = 2 I: Sigma scales by 4
    dynamic = SplatMesh(dynamicMode=True, sphericalHarmonicsDegree=1)
    dynamic.build(raw, transform16=T)
    assert np.array_equal(dynamic.getFloatCenters(0, 199)[:, :3], raw.centers)


def _plain_cov(raw):
    from gaussiansplats3d_b200.scenes import compute_covariances
    return compute_covariances(raw.scales, raw.rotations)


def test_ellipse_block_test_never_drops_a_covered_block(tmp_path):
    """csrc/ellipse_mask.h on the host, used the way the blend kernel uses it: brute force over pixel centres for 150 K random splats --
    an 8x8-px block holding a covered pixel always passes the exact minimum-of-q test, and the test prunes most AABB corners."""
    import json
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "ellipse_mask_check"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", str(exe), str(root / "oracle" / "ellipse_mask_check.cpp")], check=True)
    out = subprocess.run([str(exe), "150000"], capture_output=True, text=True)
    stats = json.loads(out.stdout)
    assert out.returncode == 0 and stats["dropped_hits"] == 0
    assert stats["exact"] <= stats["kept"] < stats["blocks_in_aabb"]          # conservative, yet it prunes
    assert stats["kept"] < 1.25 * stats["exact"]                              # ... nearly all of what can be pruned
