"""oracle/ksplat_oracle.py -- TEST INFRASTRUCTURE ONLY.

NumPy restatement of how the reference turns a `.ksplat` buffer into the arrays its renderer and sorter consume
(static scene, identity scene transform), following /root/reference/src/loaders/SplatBuffer.js:
  centres   fillSplatCenterArray      :307-347  (level >= 1: (u16 - range) * (halfBlock / range) + bucketCentre, evaluated in f64, stored f32)
  bucket    getBucketIndex            :199-219
  cov       fillSplatCovarianceArray  :488-520 -> computeCovariance :440-486 (f64, stored f32 / f16)
  colour    fillSplatColorArray       :522-549  (alpha < minimumAlpha -> 0)
  SH        fillSphericalHarmonicsArray :551-734 (channel-major file order -> coefficient-major RGB triples; stored at
            max(1, level): f32 -> f16 for level-0 files, f16 and u8 pass through)  SplatMesh.js:1064-1066
and src/splatmesh/SplatMesh.js:1143-1153 (centres+colours texel), :1912-1948 (sorter centres).
PARITY UNPINNED against the JavaScript itself (no JS engine here): pinned only by the reference's own writer/reader
being inverse to each other, which tests/test_ksplat.py checks through gaussiansplats3d_b200.ksplat.write.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gaussiansplats3d_b200 import ksplat as K  # noqa: E402  (byte layout only)


def _rotation_matrices(q_xyzw: np.ndarray) -> np.ndarray:
    x, y, z, w = (q_xyzw[:, k].astype(np.float64) for k in range(4))
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    R = np.empty((q_xyzw.shape[0], 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = 1 - (yy + zz), xy - wz, xz + wy
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = xy + wz, 1 - (xx + zz), yz - wx
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = xz - wy, yz + wx, 1 - (xx + yy)
    return R


def decode(data: bytes, *, minimum_alpha: int = 1, half_covariances: bool = False) -> dict:
    h = K.parse(data)
    buf = np.frombuffer(data, np.uint8)
    level = h.compression_level
    L = K.LEVELS[level]
    out = {k: [] for k in ("centers", "scales", "rot", "colors", "sh")}
    sh_degree = min(s.sh_degree for s in h.sections) if h.sections else 0
    ncomp_out = K.SH_COMPONENTS[sh_degree]
    for s in h.sections:
        n = s.max_splat_count
        rec = buf[s.data_base: s.data_base + n * s.bytes_per_splat].reshape(n, s.bytes_per_splat)
        if level == 0:
            c = rec[:, 0:12].copy().view(np.float32).reshape(n, 3)
            sc = rec[:, 12:24].copy().view(np.float32).reshape(n, 3)
            r = rec[:, 24:40].copy().view(np.float32).reshape(n, 4)
            col = rec[:, 40:44].copy()
            sh_raw = rec[:, 44:].copy().view(np.float32).reshape(n, -1) if s.sh_degree else None
        else:
            cu = rec[:, 0:6].copy().view(np.uint16).reshape(n, 3).astype(np.float64)
            lengths = np.concatenate([np.full(s.full_bucket_count, s.bucket_size, np.int64),
                                      buf[s.base: s.base + 4 * s.partially_filled_bucket_count].copy().view(np.uint32).astype(np.int64)])
            bidx = np.repeat(np.arange(lengths.size), lengths)[:n]
            centres = buf[s.buckets_base: s.buckets_base + 12 * s.bucket_count].copy().view(np.float32).reshape(-1, 3).astype(np.float64)
            sf = (s.bucket_block_size / 2.0) / s.compression_scale_range
            c = ((cu - s.compression_scale_range) * sf + centres[bidx]).astype(np.float32)
            sc = rec[:, 6:12].copy().view(np.float16).reshape(n, 3).astype(np.float32)
            r = rec[:, 12:20].copy().view(np.float16).reshape(n, 4).astype(np.float32)
            col = rec[:, 20:24].copy()
            if s.sh_degree:
                sh_raw = rec[:, 24:].copy().view(np.float16).reshape(n, -1) if level == 1 else rec[:, 24:].copy()
            else:
                sh_raw = None
        out["centers"].append(c); out["scales"].append(sc)
        out["rot"].append(np.stack([r[:, 1], r[:, 2], r[:, 3], r[:, 0]], 1))     # file w,x,y,z -> x,y,z,w (:398-401)
        out["colors"].append(col)
        if ncomp_out:
            ncoef1 = sh_raw[:, :9].reshape(n, 3, 3).transpose(0, 2, 1).reshape(n, 9)          # [ch][coef] -> [coef][ch]
            parts = [ncoef1]
            if sh_degree >= 2:
                parts.append(sh_raw[:, 9:24].reshape(n, 3, 5).transpose(0, 2, 1).reshape(n, 15))
            out["sh"].append(np.concatenate(parts, 1))
    centers = np.concatenate(out["centers"]); scales = np.concatenate(out["scales"]); rot = np.concatenate(out["rot"]); colors = np.concatenate(out["colors"])
    n = centers.shape[0]
    sh = None
    if ncomp_out:
        sh = np.concatenate(out["sh"])
        if level == 0:
            sh = K.to_half_three(sh).view(np.float16)      # stored at compression level max(1, 0) = 1 on the GPU (THREE toHalfFloat)
    M = _rotation_matrices(rot) * scales.astype(np.float64)[:, None, :]
    def dot(r0, r1):  # Matrix3.multiplyMatrices: a1*b1 + a2*b2 + a3*b3, left to right, unfused
        return (M[:, r0, 0] * M[:, r1, 0] + M[:, r0, 1] * M[:, r1, 1]) + M[:, r0, 2] * M[:, r1, 2]
    cov6 = np.stack([dot(0, 0), dot(0, 1), dot(0, 2), dot(1, 1), dot(1, 2), dot(2, 2)], 1)
    cov6 = K.to_half_three(cov6.astype(np.float32)).view(np.float16) if half_covariances else cov6.astype(np.float32)
    a = colors[:, 3].astype(np.uint32)
    a = np.where(a >= minimum_alpha, a, 0)
    cc = np.empty((n, 4), np.uint32)
    cc[:, 0] = colors[:, 0].astype(np.uint32) | (colors[:, 1].astype(np.uint32) << 8) | (colors[:, 2].astype(np.uint32) << 16) | (a << 24)
    cc[:, 1:] = centers.view(np.uint32)
    ic = np.empty((n, 4), np.int32)
    ic[:, :3] = np.floor(centers.astype(np.float64) * 1000.0 + 0.5).astype(np.int32)
    ic[:, 3] = 1000
    fc = np.ones((n, 4), np.float32)
    fc[:, :3] = centers
    return dict(header=h, count=n, sh_degree=sh_degree, centers=centers, scales=scales, rotations=rot, colors=colors, sh=sh, covariances=cov6,
                centers_colors=cc, int_centers=ic, float_centers=fc)
