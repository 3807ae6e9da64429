"""CPU: pin oracle/raster_oracle.c (the line-by-line GLSL restatement the GPU tests compare against) with a SECOND formulation that
shares none of its code or algebra (oracle/raster_independent.py): projective map + finite-difference Jacobian, INRIA conic
falloff, SciPy real spherical harmonics, a software triangle rasteriser for the quad's two triangles, f64 'over' chain.

Reference semantics being pinned: SplatMaterial.js:156-166 (transform, cull), :173-341 (SH colour), SplatMaterial3D.js:111-149
(J W Sigma W^T J^T + kernel), :174-213 (eigen basis, quad), :234-252 (fragment), SplatGeometry.js:14-23 (two triangles)."""
import numpy as np
import pytest

import cases
from oracle import raster_independent as RI


def _scene_and_uniforms(n, seed, kind, sh_degree, w, h, eye=(1.54163, 2.68515, -6.37228), target=(0.45622, 1.95338, 1.51278), up=(0.01933, -0.7583, -0.65161)):
    from gaussiansplats3d_b200.engine import Uniforms           # struct marshalling only
    from gaussiansplats3d_b200.scenes import pack_scene, synthetic_scene
    raw = synthetic_scene(n, seed=seed, kind=kind, sh_degree=sh_degree)
    p = pack_scene(raw)
    _, view, proj = cases.camera_mvp(eye=eye, target=target, up=up, aspect=w / h)
    u = Uniforms(model_view=view.astype(np.float32), projection=proj.astype(np.float32), camera_position=np.asarray(eye, np.float32),
                 focal=(proj[0] * 0.5 * w, proj[5] * 0.5 * h), viewport=(w, h), sh_degree=sh_degree)
    return raw, p, u, view, proj, np.asarray(eye, np.float64)


def _independent_projection(p, u, view, proj, eye, w, h, sh_degree):
    cc = p.centers_colors
    centers = cc[:, 1:].copy().view(np.float32).astype(np.float64)
    rgba = np.stack([(cc[:, 0] >> (8 * k)) & 255 for k in range(4)], 1)
    sh = None if p.sh is None else p.sh.astype(np.float64).reshape(p.count, -1, 3)
    # same f32-rounded matrices the shaders receive
    return RI.project(view.astype(np.float32), proj.astype(np.float32), eye.astype(np.float32), (w, h), centers, rgba, p.covariances.astype(np.float64),
                      sh=sh, sh_degree=sh_degree, kernel2d=u.kernel_2d_size)


@pytest.mark.parametrize("kind,sh_degree,seed", [("bonsai", 0, 4), ("bonsai", 2, 5), ("garden", 1, 6)])
def test_projection_agrees_with_conic_formulation(oracle_mod, kind, sh_degree, seed):
    n, w, h = 20_000, 640, 360
    raw, p, u, view, proj, eye = _scene_and_uniforms(n, seed, kind, sh_degree, w, h)
    got = oracle_mod.project(u, p.centers_colors, p.covariances, p.sh, p.sh_degree)
    ind = _independent_projection(p, u, view, proj, eye, w, h, sh_degree)
    B1, B2, clamped, positive = RI.eigen_basis(ind["sigma2"])
    valid_ind = ind["valid"] & positive
    assert (got["valid"].astype(bool) != valid_ind).mean() < 2e-4          # f32 vs f64 at the 1.2 w cull boundary
    m = got["valid"].astype(bool) & valid_ind
    assert m.sum() > n // 4
    # screen position through the projective map
    assert np.abs(got["cx"][m] - ind["mean"][m, 0]).max() < 2e-3 and np.abs(got["cy"][m] - ind["mean"][m, 1]).max() < 2e-3
    assert np.abs(got["ndc_z"][m] - ind["ndc_z"][m]).max() < 1e-5
    # the quad basis spans 8 Sigma' exactly where the reference's clamps are inactive: B1 B1^T + B2 B2^T = 8 (J Sv J^T + k I),
    # with J from FINITE DIFFERENCES of the projective map (no focal / z expressions on this side)
    Q = np.stack([got["b1x"] * got["b1x"] + got["b2x"] * got["b2x"], got["b1x"] * got["b1y"] + got["b2x"] * got["b2y"],
                  got["b1y"] * got["b1y"] + got["b2y"] * got["b2y"]], 1).astype(np.float64)
    S = np.stack([ind["sigma2"][:, 0, 0], ind["sigma2"][:, 0, 1], ind["sigma2"][:, 1, 1]], 1) * 8.0
    un = m & ~clamped
    assert un.sum() > 1000 and (m & clamped).sum() > 10      # both populations are exercised
    rel = np.abs(Q[un] - S[un]).max(1) / np.abs(S[un]).max(1)
    assert np.quantile(rel, 0.999) < 5e-4 and rel.max() < 2e-2, (np.quantile(rel, 0.999), rel.max())
    # clamped splats: restated clamp lines applied to the independent covariance
    Qc = np.stack([B1[:, 0] ** 2 + B2[:, 0] ** 2, B1[:, 0] * B1[:, 1] + B2[:, 0] * B2[:, 1], B1[:, 1] ** 2 + B2[:, 1] ** 2], 1)
    cl = m & clamped
    relc = np.abs(Q[cl] - Qc[cl]).max(1) / np.abs(Qc[cl]).max(1)
    assert np.quantile(relc, 0.999) < 5e-4 and relc.max() < 2e-2
    # colour: SciPy's real spherical harmonics against the shader's hand-written polynomials
    for k, ch in enumerate("rgb"):
        assert np.abs(got[ch][m] - ind["rgb"][m, k]).max() < 3e-6 * (1 + 50 * (sh_degree > 0)), ch
    assert np.abs(got["a"][m] - ind["opacity"][m]).max() < 1e-6


def test_two_triangle_rasterisation_equals_inverse_map_shortcut():
    """The reference draws each splat as 2 triangles with interpolated vPosition and discards A > 8.  Both the CPU restatement and
    the CUDA blend use the closed form instead (A from the inverse affine map, no triangles).  Check that closed form against an
    actual triangle rasteriser: identical coverage after the A <= 8 discard, identical A."""
    rng = np.random.default_rng(11)
    w, h = 96, 64
    for trial in range(300):
        ang = rng.uniform(0, np.pi)
        e1 = np.array([np.cos(ang), np.sin(ang)])
        e2 = np.array([e1[1], -e1[0]])
        l1, l2 = np.exp(rng.uniform(-1, 3.5)), np.exp(rng.uniform(-1, 3.5))
        B1, B2 = e1 * l1, e2 * l2
        mean = np.array([rng.uniform(-10, w + 10), rng.uniform(-10, h + 10)])
        covered, A_tri = RI.rasterise_quad_triangles(mean, B1, B2, w, h)
        ys, xs = np.mgrid[0:h, 0:w]
        alpha, A_short = RI.quad_alpha_shortcut(mean, B1, B2, 1.0, xs, ys)
        drawn_tri = covered & (A_tri <= 8.0)
        drawn_short = A_short <= 8.0
        edge = np.abs(A_short - 8.0) < 1e-9
        assert np.array_equal(drawn_tri | edge, drawn_short | edge), trial
        both = drawn_tri & drawn_short
        if both.any():
            assert np.abs(A_tri[both] - A_short[both]).max() < 1e-9
        # every pixel with A <= 8 lies inside the quad (the unit disc is inscribed): the triangles never clip the splat
        assert not (drawn_short & ~covered & ~edge).any()


def test_conic_equals_eigen_quad_where_unclamped():
    """exp(-1/2 d^T Sigma'^-1 d) with the Mahalanobis^2 <= 8 cut == the eigen-basis quad evaluation (SplatMaterial3D.js:154-171's own
    claim), for covariances on which neither clamp acts."""
    rng = np.random.default_rng(3)
    ys, xs = np.mgrid[0:40, 0:40]
    for _ in range(200):
        a = rng.normal(size=(2, 2)) * rng.uniform(0.5, 6)
        S2 = a @ a.T + 0.3 * np.eye(2)
        B1, B2, clamped, pos = RI.eigen_basis(S2[None])
        if clamped[0] or not pos[0]:
            continue
        mean = rng.uniform(10, 30, 2)
        ca, m2 = RI.conic_alpha(mean, S2, 0.9, xs, ys)
        qa, A = RI.quad_alpha_shortcut(mean, B1[0], B2[0], 0.9, xs, ys)
        near_cut = np.abs(m2 - 8.0) < 1e-6
        assert np.abs(ca - qa)[~near_cut].max() < 1e-9
        assert np.abs(m2 - A)[~near_cut].max() < 1e-6


@pytest.mark.parametrize("kind,sh_degree,n,w,h,seed", [("bonsai", 0, 6000, 160, 100, 1), ("bonsai", 2, 5000, 128, 96, 2), ("garden", 1, 6000, 200, 120, 3)])
def test_frame_agrees_with_independent_renderer(oracle_mod, kind, sh_degree, n, w, h, seed):
    """Whole frames: the GLSL restatement (f32, eigen quads) vs the independent renderer (f64, conics for every unclamped splat)."""
    raw, p, u, view, proj, eye = _scene_and_uniforms(n, seed, kind, sh_degree, w, h)
    rng = np.random.default_rng(seed)
    order = rng.permutation(n).astype(np.uint32)        # any draw order: both sides must honour it
    want, _ = oracle_mod.render(u, p.centers_colors, p.covariances, order, w, h, sh=p.sh, sh_degree=p.sh_degree)
    ind = _independent_projection(p, u, view, proj, eye, w, h, sh_degree)
    got, clamped = RI.render(ind, order, w, h)
    assert (~clamped).sum() > n // 10
    err = np.abs(got - want.astype(np.float64))
    # differences: f32 vs f64 and pixels whose A sits within rounding of the A = 8 cut (one splat's edge appears / disappears:
    # at most opacity * exp(-4) = 1.8 % of a colour step chain)
    assert (err <= 0.25 / 255).mean() >= 0.999, (err <= 0.25 / 255).mean()
    assert err.max() <= 6.0 / 255, err.max() * 255
    assert want[..., 3].max() > 0.5
