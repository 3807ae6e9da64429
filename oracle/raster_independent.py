"""oracle/raster_independent.py -- TEST INFRASTRUCTURE ONLY.

A SECOND, structurally different statement of what the reference's splat shaders draw, used to pin oracle/raster_oracle.c (which
follows the GLSL line by line and therefore shares its formulation with the CUDA projection kernel).  Nothing here is derived from
raster_oracle.c or from the product: it is the textbook EWA-splatting description the reference's own comments appeal to
(/root/reference/src/splatmesh/SplatMaterial3D.js:117-134 "Jacobian of the affine approximation of the projection",
:154-171 "full gaussian: exp(-0.5 (X - mean) conic (X - mean))"), evaluated in float64 with NumPy / SciPy:

  * screen position   : the full projective map  world -> clip -> NDC -> pixels  (SplatMaterial.js:156-166)
  * 2D covariance     : J Sigma_view J^T where J is the FINITE-DIFFERENCE Jacobian of that projective map at the splat centre
                        (no closed-form focal/z expressions), + kernel2DSize on the diagonal (SplatMaterial3D.js:111-149)
  * falloff           : INRIA conic form  alpha = a exp(-1/2 d^T Sigma'^-1 d), cut at d^T Sigma'^-1 d > 8
                        (equals the eigen-basis quad of SplatMaterial3D.js:174-213, :234-252 wherever its clamps are inactive)
  * colour            : real spherical harmonics (Condon-Shortley phase) from scipy.special.sph_harm_y dotted with the coefficients
                        (SplatMaterial.js:173-341 spells the same polynomials out by hand)
  * coverage          : a software rasteriser for the reference's actual geometry -- 2 triangles (0,1,2),(0,2,3) over the 4 corners
                        (-1,-1),(-1,1),(1,1),(1,-1) (SplatGeometry.js:14-23), pixel-centre sampling, barycentric interpolation of
                        vPosition -- against which the "A <= 8 from the inverse affine map" shortcut both other implementations use
                        is checked
  * blend             : back-to-front "over" in draw order (SplatMaterial3D.js:65-75 NormalBlending)

The two reference-specific clamps (term2 >= sqrt(0.1) in the eigen split, basis length <= maxScreenSpaceSplatSize) have no
textbook counterpart; `eigen_basis` restates just those two lines so that whole frames can be compared, and reports which
splats they touched so tests can also compare the untouched majority against the pure conic form.
"""
from __future__ import annotations

import numpy as np

SQRT8 = np.sqrt(8.0)


def _mat(colmajor16) -> np.ndarray:
    return np.asarray(colmajor16, np.float64).reshape(4, 4).T


def pixel_of_view_point(P: np.ndarray, v: np.ndarray, viewport) -> np.ndarray:
    """View-space points [n,3] -> pixel coordinates [n,2] through the projective map (GL window coordinates, y up)."""
    h = np.concatenate([v, np.ones((v.shape[0], 1))], 1) @ P.T
    ndc = h[:, :2] / h[:, 3:4]
    return (ndc + 1.0) * 0.5 * np.asarray(viewport, np.float64)[None, :]


def numeric_jacobian(P: np.ndarray, v: np.ndarray, viewport, orthographic=False) -> np.ndarray:
    """d(pixel)/d(view xyz) at each view-space point by central differences, [n,2,3]."""
    n = v.shape[0]
    J = np.empty((n, 2, 3))
    step = 1e-5 * np.maximum(np.abs(v[:, 2]), 1e-3)
    for k in range(3):
        d = np.zeros_like(v)
        d[:, k] = step
        J[:, :, k] = (pixel_of_view_point(P, v + d, viewport) - pixel_of_view_point(P, v - d, viewport)) / (2.0 * step)[:, None]
    del orthographic
    return J


def real_sh_colour(direction: np.ndarray, coeffs: np.ndarray, degree: int) -> np.ndarray:
    """Sum over bands 1..degree of Y_lm(direction) * coeff_lm per colour channel.  coeffs: [n, ncoef, 3] in the reference's GPU-side
    order (l=1: m=-1,0,1; l=2: m=-2..2).  Y_lm = real SH with the Condon-Shortley phase (what the 3DGS constants encode)."""
    from scipy.special import sph_harm_y
    x, y, z = direction[:, 0], direction[:, 1], direction[:, 2]
    theta = np.arccos(np.clip(z, -1.0, 1.0))
    phi = np.arctan2(y, x)
    out = np.zeros((direction.shape[0], 3))
    k = 0
    for l in range(1, degree + 1):
        for m in range(-l, l + 1):
            Y = sph_harm_y(l, abs(m), theta, phi)
            if m < 0:
                real = np.sqrt(2.0) * Y.imag          # the Condon-Shortley sign of Y_l^|m| is kept (3DGS convention)
            elif m == 0:
                real = Y.real
            else:
                real = np.sqrt(2.0) * Y.real
            out += real[:, None] * coeffs[:, k, :]
            k += 1
    return out


def project(model_view16, projection16, camera_position, viewport, centers, rgba8, cov6, *, sh=None, sh_degree=0, kernel2d=0.3,
            focal_adjust_inv=1.0) -> dict:
    """Per-splat screen-space Gaussian: mean [n,2] px, Sigma2 [n,2,2] px^2 (kernel included), colour [n,3], opacity [n], valid [n]."""
    MV, P = _mat(model_view16), _mat(projection16)
    c = np.asarray(centers, np.float64)
    n = c.shape[0]
    v = c @ MV[:3, :3].T + MV[:3, 3]
    clip = np.concatenate([v, np.ones((n, 1))], 1) @ P.T
    lim = 1.2 * clip[:, 3]
    culled = (clip[:, 2] < -lim) | (np.abs(clip[:, 0]) > lim) | (np.abs(clip[:, 1]) > lim)        # SplatMaterial.js:160-164
    ndc_z = clip[:, 2] / clip[:, 3]
    mean = pixel_of_view_point(P, v, viewport)
    J = numeric_jacobian(P, v, viewport)
    S = np.asarray(cov6, np.float64)
    Sigma = np.stack([np.stack([S[:, 0], S[:, 1], S[:, 2]], 1), np.stack([S[:, 1], S[:, 3], S[:, 4]], 1), np.stack([S[:, 2], S[:, 4], S[:, 5]], 1)], 1)
    Sv = MV[:3, :3][None] @ Sigma @ MV[:3, :3].T[None]                   # covariance in view space
    S2 = J @ Sv @ np.transpose(J, (0, 2, 1))
    S2[:, 0, 0] += kernel2d
    S2[:, 1, 1] += kernel2d
    col = np.asarray(rgba8, np.float64) / 255.0
    rgb = col[:, :3].copy()
    if sh is not None and sh_degree >= 1:
        d = c - np.asarray(camera_position, np.float64)[None, :]
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rgb = np.clip(rgb + real_sh_colour(d, np.asarray(sh, np.float64), sh_degree), 0.0, 1.0)
    valid = (~culled) & (ndc_z >= -1.0) & (ndc_z <= 1.0)
    del focal_adjust_inv
    return dict(mean=mean, sigma2=S2, rgb=rgb, opacity=col[:, 3].copy(), valid=valid, culled=culled, ndc_z=ndc_z, view=v)


def eigen_basis(S2: np.ndarray, splat_scale=1.0, max_size=1024.0):
    """The quad basis of SplatMaterial3D.js:174-196 from a 2x2 covariance: B1, B2 [n,2] px, plus `clamped` [n] = one of the two
    reference-specific clamps changed the result (then B1 B1^T + B2 B2^T != 8 Sigma')."""
    a, b, d = S2[:, 0, 0], S2[:, 0, 1], S2[:, 1, 1]
    half_tr = 0.5 * (a + d)
    disc = half_tr * half_tr - (a * d - b * b)
    term2 = np.sqrt(np.maximum(0.1, disc))
    l1, l2 = half_tr + term2, half_tr - term2
    e1 = np.stack([b, l1 - a], 1)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.stack([e1[:, 1], -e1[:, 0]], 1)
    r1, r2 = SQRT8 * np.sqrt(np.maximum(l1, 0)), SQRT8 * np.sqrt(np.maximum(l2, 0))
    clamped = (disc < 0.1) | (r1 > max_size) | (r2 > max_size)
    B1 = e1 * (splat_scale * np.minimum(r1, max_size))[:, None]
    B2 = e2 * (splat_scale * np.minimum(r2, max_size))[:, None]
    return B1, B2, clamped, l2 > 0


def conic_alpha(mean, S2, opacity, px, py):
    """INRIA form at pixel centres (px+0.5, py+0.5): opacity * exp(-1/2 d^T S2^-1 d), 0 where the Mahalanobis^2 exceeds 8."""
    det = S2[0, 0] * S2[1, 1] - S2[0, 1] * S2[0, 1]
    ca, cb, cc = S2[1, 1] / det, -S2[0, 1] / det, S2[0, 0] / det
    dx, dy = (px + 0.5) - mean[0], (py + 0.5) - mean[1]
    m2 = ca * dx * dx + 2.0 * cb * dx * dy + cc * dy * dy
    return np.where(m2 <= 8.0, opacity * np.exp(-0.5 * m2), 0.0), m2


def quad_alpha_shortcut(mean, B1, B2, opacity, px, py):
    """A from the inverse of the affine quad map (what raster_oracle.c and the CUDA blend evaluate)."""
    dx, dy = (px + 0.5) - mean[0], (py + 0.5) - mean[1]
    u = (dx * B1[0] + dy * B1[1]) / (B1 @ B1)
    w = (dx * B2[0] + dy * B2[1]) / (B2 @ B2)
    A = 8.0 * (u * u + w * w)
    return np.where(A <= 8.0, opacity * np.exp(-0.5 * A), 0.0), A


def rasterise_quad_triangles(mean, B1, B2, width, height):
    """The reference's real geometry: corners mean + qx B1 + qy B2 for q = (-1,-1), (-1,1), (1,1), (1,-1); triangles (0,1,2), (0,2,3)
    (SplatGeometry.js:14-23); varying vPosition = q * sqrt(8) interpolated with barycentric weights at pixel centres (w = 1, so
    perspective-correct == linear).  Returns (covered mask [h,w], A [h,w]) with A = dot(vPosition, vPosition) where covered."""
    q = np.array([[-1.0, -1.0], [-1.0, 1.0], [1.0, 1.0], [1.0, -1.0]])
    verts = mean[None, :] + q[:, :1] * B1[None, :] + q[:, 1:] * B2[None, :]
    vpos = q * SQRT8
    ys, xs = np.mgrid[0:height, 0:width]
    sx, sy = xs + 0.5, ys + 0.5
    covered = np.zeros((height, width), bool)
    A = np.full((height, width), np.inf)
    for tri in ((0, 1, 2), (0, 2, 3)):
        p0, p1, p2 = verts[tri[0]], verts[tri[1]], verts[tri[2]]
        area = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0])
        if area == 0:
            continue
        w0 = ((p1[0] - sx) * (p2[1] - sy) - (p1[1] - sy) * (p2[0] - sx)) / area
        w1 = ((p2[0] - sx) * (p0[1] - sy) - (p2[1] - sy) * (p0[0] - sx)) / area
        w2 = 1.0 - w0 - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        vp = w0[..., None] * vpos[tri[0]] + w1[..., None] * vpos[tri[1]] + w2[..., None] * vpos[tri[2]]
        a = (vp * vp).sum(-1)
        new = inside & ~covered
        A[new] = a[new]
        covered |= inside
    return covered, A


def render(proj: dict, order, width, height, *, splat_scale=1.0, max_size=1024.0, use_conic_when_unclamped=True):
    """Back-to-front 'over' chain in draw order with f64 accumulators; frame rows bottom-up (GL window coordinates).
    Unclamped splats are drawn from the CONIC (no eigen decomposition at all); clamped ones from the restated basis."""
    B1, B2, clamped, positive = eigen_basis(proj["sigma2"], splat_scale, max_size)
    frame = np.zeros((height, width, 4))
    ok = proj["valid"] & positive
    for s in np.asarray(order, np.int64):
        if not ok[s]:
            continue
        mean = proj["mean"][s]
        ext = np.abs(B1[s]) + np.abs(B2[s])
        x0, x1 = int(max(np.floor(mean[0] - ext[0] - 1), 0)), int(min(np.ceil(mean[0] + ext[0] + 1), width - 1))
        y0, y1 = int(max(np.floor(mean[1] - ext[1] - 1), 0)), int(min(np.ceil(mean[1] + ext[1] + 1), height - 1))
        if x0 > x1 or y0 > y1:
            continue
        ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1]
        if use_conic_when_unclamped and not clamped[s] and splat_scale == 1.0:
            alpha, _ = conic_alpha(mean, proj["sigma2"][s], proj["opacity"][s], xs, ys)
        else:
            alpha, _ = quad_alpha_shortcut(mean, B1[s], B2[s], proj["opacity"][s], xs, ys)
        blk = frame[y0:y1 + 1, x0:x1 + 1]
        om = (1.0 - alpha)[..., None]
        blk[..., :3] = proj["rgb"][s][None, None, :] * alpha[..., None] + blk[..., :3] * om
        blk[..., 3:] = alpha[..., None] + blk[..., 3:] * om
    return frame, clamped
