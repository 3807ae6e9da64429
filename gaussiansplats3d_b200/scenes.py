"""Seeded synthetic splat scenes standing in for the reference's demo data (bonsai / garden .ksplat files are an
external download that is not on disk: README.md:121) and the packing of raw splat attributes into the arrays the
reference's SplatMesh uploads (the load-time CPU loops of SplatBuffer.js / SplatMesh.js, done here with NumPy).

Sizes, SH degrees and cameras follow BASELINE.json `configs` / SURVEY.md 8(d)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class RawScene:
    """Uncompressed per-splat attributes, the content of a level-0 SplatBuffer (SplatBuffer.js:108-163)."""
    centers: np.ndarray      # f32 [n,3]
    scales: np.ndarray       # f32 [n,3] linear
    rotations: np.ndarray    # f32 [n,4] x,y,z,w (unit)
    colors: np.ndarray       # u8  [n,4] rgba
    sh: np.ndarray | None    # f32 [n, ncoef, 3] coefficient-major RGB triples (the GPU-side order), or None
    sh_degree: int

    @property
    def count(self) -> int:
        return self.centers.shape[0]


def synthetic_scene(n: int, seed: int, kind: str = "bonsai", sh_degree: int = 0) -> RawScene:
    """bonsai: clustered N(0,1.5^2) + 10% shell r=6.  garden: same + ground-plane disc r=12.  uniform: N(0,3^2)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        centers = rng.normal(0.0, 3.0, (n, 3))
    else:
        n_shell = n // 10
        n_disc = n // 4 if kind == "garden" else 0
        n_core = n - n_shell - n_disc
        core = rng.normal(0.0, 1.5, (n_core, 3))
        d = rng.normal(0, 1, (n_shell, 3))
        shell = 6.0 * d / np.linalg.norm(d, axis=1, keepdims=True)
        parts = [core, shell]
        if n_disc:
            r = 12.0 * np.sqrt(rng.uniform(0, 1, n_disc))
            th = rng.uniform(0, 2 * np.pi, n_disc)
            parts.append(np.stack([r * np.cos(th), rng.normal(-1.5, 0.05, n_disc), r * np.sin(th)], 1))
        centers = np.concatenate(parts, 0)
        centers = centers[rng.permutation(n)]
    log_scales = np.clip(rng.normal(-4.5, 1.2, (n, 3)), -8.0, -1.0)
    scales = np.exp(log_scales)
    q = rng.normal(0, 1, (n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1.0
    colors = np.empty((n, 4), np.uint8)
    colors[:, :3] = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    colors[:, 3] = (255.0 * (1.0 - rng.uniform(0, 1, n) ** 3)).astype(np.uint8)  # skewed high
    sh = None
    if sh_degree > 0:
        ncoef = 3 if sh_degree == 1 else 8
        sh = rng.normal(0.0, 0.15, (n, ncoef, 3)).astype(np.float32)
    return RawScene(centers.astype(np.float32), scales.astype(np.float32), q.astype(np.float32), colors, sh, sh_degree)


def compute_covariances(scales: np.ndarray, rotations_xyzw: np.ndarray, transform3x3: np.ndarray | None = None) -> np.ndarray:
    """SplatBuffer.computeCovariance (SplatBuffer.js:440-486): Sigma = (R S)(R S)^T [then T Sigma T^T], evaluated in
    float64, stored as float32 [m00, m01, m02, m11, m12, m22]."""
    s = scales.astype(np.float64)
    x, y, z, w = (rotations_xyzw[:, k].astype(np.float64) for k in range(4))
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    R = np.empty((scales.shape[0], 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = 1 - (yy + zz), xy - wz, xz + wy
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = xy + wz, 1 - (xx + zz), yz - wx
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = xz - wy, yz + wx, 1 - (xx + yy)
    M = R * s[:, None, :]
    cov = M @ np.transpose(M, (0, 2, 1))
    if transform3x3 is not None:
        T = np.asarray(transform3x3, np.float64)
        cov = T @ cov @ T.T
    out = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], 1)
    return out.astype(np.float32)


def pack_centers_colors(centers: np.ndarray, colors: np.ndarray, minimum_alpha: int = 1) -> np.ndarray:
    """SplatMesh.updateCenterColorsPaddedData (SplatMesh.js:1143-1153) + fillSplatColorArray's alpha floor
    (SplatBuffer.js:541-542): uvec4 {r | g<<8 | b<<16 | a<<24, bits(x), bits(y), bits(z)}."""
    n = centers.shape[0]
    c = colors.astype(np.uint32)
    a = np.where(c[:, 3] >= minimum_alpha, c[:, 3], 0)
    out = np.empty((n, 4), np.uint32)
    out[:, 0] = c[:, 0] | (c[:, 1] << 8) | (c[:, 2] << 16) | (a << 24)
    out[:, 1:] = np.ascontiguousarray(centers, dtype=np.float32).view(np.uint32)
    return out


def integer_centers(centers: np.ndarray) -> np.ndarray:
    """SplatMesh.getIntegerCenters(padFour=true) (SplatMesh.js:1912-1926): Math.round(f32 * 1000.0) in f64, w = 1000."""
    n = centers.shape[0]
    out = np.empty((n, 4), np.int32)
    out[:, :3] = np.floor(centers.astype(np.float32).astype(np.float64) * 1000.0 + 0.5).astype(np.int32)
    out[:, 3] = 1000
    return out


def float_centers(centers: np.ndarray) -> np.ndarray:
    """SplatMesh.getFloatCenters(padFour=true) (SplatMesh.js:1935-1948): w = 1."""
    out = np.ones((centers.shape[0], 4), np.float32)
    out[:, :3] = centers
    return out


@dataclass
class PackedScene:
    """What SplatMesh keeps on the GPU (setupDataTextures, SplatMesh.js:637-898) + what it sends the sorter."""
    centers_colors: np.ndarray   # u32 [n,4]
    covariances: np.ndarray      # f32 or f16 [n,6]
    sh: np.ndarray | None        # f16 / u8 / f32 [n, ncomp]
    sh_degree: int
    int_centers: np.ndarray      # i32 [n,4]
    count: int


def pack_scene(raw: RawScene, *, half_covariances: bool = False, sh_format: str = "f16", minimum_alpha: int = 1,
               sh8_range: tuple[float, float] = (-1.5, 1.5)) -> PackedScene:
    cov = compute_covariances(raw.scales, raw.rotations)
    if half_covariances:  # halfPrecisionCovariancesOnGPU
        cov = cov.astype(np.float16)
    sh = None
    if raw.sh is not None and raw.sh_degree > 0:
        flat = raw.sh.reshape(raw.count, -1)
        if sh_format == "f16":  # getTargetSphericalHarmonicsCompressionLevel >= 1 (SplatMesh.js:1064-1066)
            sh = flat.astype(np.float16)
        elif sh_format == "u8":  # SplatBuffer.js:22-26
            lo, hi = sh8_range
            sh = np.floor((np.clip(flat, lo, hi) - lo) / (hi - lo) * 255.0).astype(np.uint8)
        else:
            sh = flat.astype(np.float32)
    return PackedScene(pack_centers_colors(raw.centers, raw.colors, minimum_alpha), cov, sh, raw.sh_degree if sh is not None else 0,
                       integer_centers(raw.centers), raw.count)


# Cameras of the reference's demo pages (demo/bonsai.html:38-41, demo/garden.html:38-41) and the Viewer default
# (Viewer.js:51-55)
CAMERAS = {
    "bonsai": dict(up=(0.01933, -0.75830, -0.65161), position=(1.54163, 2.68515, -6.37228), look_at=(0.45622, 1.95338, 1.51278)),
    "garden": dict(up=(0.0, -1.0, -0.54), position=(-3.15634, -0.16946, -0.51552), look_at=(1.52976, 2.27776, 1.65898)),
    "default": dict(up=(0.0, 1.0, 0.0), position=(0.0, 10.0, 15.0), look_at=(0.0, 0.0, 0.0)),
}
