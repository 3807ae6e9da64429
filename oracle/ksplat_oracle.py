"""oracle/ksplat_oracle.py -- TEST INFRASTRUCTURE ONLY.  Imports nothing from the product package.

NumPy restatement of how the reference turns a `.ksplat` buffer into the arrays its renderer and sorter consume
(static scene; `transform16` = the SplatScene transform baked at load, None = identity), following /root/reference/src/loaders/SplatBuffer.js:
  header    parseHeader               :819-848, parseSectionHeaders :877-941 (own struct-based parser below)
  centres   fillSplatCenterArray      :307-347  (level >= 1: (u16 - range) * (halfBlock / range) + bucketCentre, evaluated in f64, stored f32)
  bucket    getBucketIndex            :199-219
  cov       fillSplatCovarianceArray  :488-520 -> computeCovariance :440-486 (f64, stored f32 / f16)
  colour    fillSplatColorArray       :522-549  (alpha < minimumAlpha -> 0)
  SH        fillSphericalHarmonicsArray :551-734 (channel-major file order -> coefficient-major RGB triples; stored at
            max(1, level): f32 -> f16 for level-0 files, f16 and u8 pass through)  SplatMesh.js:1064-1066
and src/splatmesh/SplatMesh.js:1143-1153 (centres+colours texel), :1912-1948 (sorter centres).
Pinning: the JavaScript itself cannot run here (no JS engine).  This decoder is pinned by hand-assembled byte fixtures
(tests/golden/ksplat_handmade.py builds them field by field from the offsets in SURVEY.md Appendix A and carries per-splat expected
values computed with scalar arithmetic), not by the product's writer; the product's parser / writer / GPU decode are then checked
against this file and against the same fixtures.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from . import pack_oracle as PO

HEADER_BYTES, SECTION_HEADER_BYTES = 4096, 1024                    # SplatBuffer.js:167-168
# SplatBuffer.CompressionLevels (:108-163): bytes of centre / scale / rotation / colour, bytes per SH component, default scale range
_LEVELS = {0: (12, 12, 16, 4, 4, 1), 1: (6, 6, 8, 4, 2, 32767), 2: (6, 6, 8, 4, 1, 32767)}
_SH_COMPONENTS = {0: 0, 1: 9, 2: 24}


@dataclass
class Section:
    max_splat_count: int
    bucket_size: int
    bucket_count: int
    bucket_block_size: float
    bucket_storage_bytes: int
    compression_scale_range: int
    full_bucket_count: int
    partially_filled_bucket_count: int
    sh_degree: int
    bytes_per_splat: int
    base: int
    buckets_base: int
    data_base: int


@dataclass
class Header:
    version: tuple
    max_section_count: int
    section_count: int
    max_splat_count: int
    splat_count: int
    compression_level: int
    scene_center: tuple
    min_sh: float
    max_sh: float
    sections: list = field(default_factory=list)


def parse(data: bytes) -> Header:
    """parseHeader (:819-848) + parseSectionHeaders (:877-941, secLoadedCountsToMax)."""
    data = bytes(data)
    major, minor = struct.unpack_from("<BB", data, 0)
    max_sections, sections, max_splats, splats = struct.unpack_from("<4I", data, 4)
    (level,) = struct.unpack_from("<H", data, 20)
    cx, cy, cz, lo, hi = struct.unpack_from("<5f", data, 24)
    h = Header((major, minor), max_sections, sections, max_splats, splats, level, (cx, cy, cz), lo or -1.5, hi or 1.5)
    cb, sb, rb, colb, shb, default_range = _LEVELS[level]
    base = HEADER_BYTES + max_sections * SECTION_HEADER_BYTES
    for i in range(max_sections):
        o = HEADER_BYTES + i * SECTION_HEADER_BYTES
        max_count, bucket_size, bucket_count = struct.unpack_from("<3I", data, o + 4)
        (block,) = struct.unpack_from("<f", data, o + 16)
        (storage,) = struct.unpack_from("<H", data, o + 20)
        (scale_range,) = struct.unpack_from("<I", data, o + 24)
        full, partial = struct.unpack_from("<2I", data, o + 32)
        (deg,) = struct.unpack_from("<H", data, o + 40)
        bps = cb + sb + rb + colb + shb * _SH_COMPONENTS[deg]
        meta = 4 * partial
        buckets_bytes = storage * bucket_count + meta
        h.sections.append(Section(max_count, bucket_size, bucket_count, block, storage, scale_range or default_range, full, partial, deg, bps,
                                  base, base + meta, base + buckets_bytes))
        base += bps * max_count + buckets_bytes
    return h


def _three_half_tables():
    """THREE.DataUtils._generateTables (three r160, not vendored: restated from its published source): base / shift per 9-bit
    sign+exponent."""
    base = np.zeros(512, np.uint32)
    shift = np.zeros(512, np.uint32)
    for i in range(256):
        e = i - 127
        if e < -27:
            b, sft = 0x0000, 24
        elif e < -14:
            b, sft = 0x0400 >> (-e - 14), -e - 1
        elif e <= 15:
            b, sft = (e + 15) << 10, 13
        elif e < 128:
            b, sft = 0x7C00, 24
        else:
            b, sft = 0x7C00, 13
        base[i], base[i | 0x100] = b, b | 0x8000
        shift[i] = shift[i | 0x100] = sft
    return base, shift


_HALF_BASE, _HALF_SHIFT = _three_half_tables()


def to_half_three(a: np.ndarray) -> np.ndarray:
    """THREE.DataUtils.toHalfFloat: clamp to +-65504, then baseTable[e] + (mantissa >> shiftTable[e]) -- the mantissa is TRUNCATED."""
    f = np.clip(np.asarray(a, np.float32), -65504.0, 65504.0).astype(np.float32).view(np.uint32)
    e = (f >> 23) & 0x1FF
    return (_HALF_BASE[e] + ((f & 0x007FFFFF) >> _HALF_SHIFT[e])).astype(np.uint16)




def rotation_of_transform(transform16):
    """Matrix4.decompose -> quaternion -> normalize -> makeRotationFromQuaternion (three r160 operation order), scalar Python.
    Returns rotation[row][col]."""
    import math
    e = [float(v) for v in np.asarray(transform16, np.float64).reshape(16)]
    sx = math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    sy = math.sqrt(e[4] * e[4] + e[5] * e[5] + e[6] * e[6])
    sz = math.sqrt(e[8] * e[8] + e[9] * e[9] + e[10] * e[10])
    det = (e[0] * (e[5] * e[10] - e[9] * e[6]) - e[4] * (e[1] * e[10] - e[9] * e[2]) + e[8] * (e[1] * e[6] - e[5] * e[2]))
    if det < 0:
        sx = -sx
    ix, iy, iz = 1.0 / sx, 1.0 / sy, 1.0 / sz
    m11, m21, m31 = e[0] * ix, e[1] * ix, e[2] * ix
    m12, m22, m32 = e[4] * iy, e[5] * iy, e[6] * iy
    m13, m23, m33 = e[8] * iz, e[9] * iz, e[10] * iz
    tr = m11 + m22 + m33
    if tr > 0:
        s = 0.5 / math.sqrt(tr + 1.0)
        w, x, y, z = 0.25 / s, (m32 - m23) * s, (m13 - m31) * s, (m21 - m12) * s
    elif m11 > m22 and m11 > m33:
        s = 2.0 * math.sqrt(1.0 + m11 - m22 - m33)
        w, x, y, z = (m32 - m23) / s, 0.25 * s, (m12 + m21) / s, (m13 + m31) / s
    elif m22 > m33:
        s = 2.0 * math.sqrt(1.0 + m22 - m11 - m33)
        w, x, y, z = (m13 - m31) / s, (m12 + m21) / s, 0.25 * s, (m23 + m32) / s
    else:
        s = 2.0 * math.sqrt(1.0 + m33 - m11 - m22)
        w, x, y, z = (m21 - m12) / s, (m13 + m31) / s, (m23 + m32) / s, 0.25 * s
    ln = math.sqrt(x * x + y * y + z * z + w * w)
    if ln == 0:
        x, y, z, w = 0.0, 0.0, 0.0, 1.0
    else:
        ln = 1.0 / ln
        x, y, z, w = x * ln, y * ln, z * ln, w * ln
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    return [[1 - (yy + zz), xy - wz, xz + wy], [xy + wz, 1 - (xx + zz), yz - wx], [xz - wy, yz + wx, 1 - (xx + yy)]]


def sh_rotation_matrices(rot):
    t11, t12, t13 = PO.band1_rows(rot)
    return np.array([t11, t12, t13]), np.array(PO.band2_rows(t11, t12, t13))


def _rotation_matrices(q_xyzw: np.ndarray) -> np.ndarray:
    x, y, z, w = (q_xyzw[:, k].astype(np.float64) for k in range(4))
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    R = np.empty((q_xyzw.shape[0], 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = 1 - (yy + zz), xy - wz, xz + wy
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = xy + wz, 1 - (xx + zz), yz - wx
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = xz - wy, yz + wx, 1 - (xx + yy)
    return R


def _ordered3(a0, b0, a1, b1, a2, b2):
    """a0*b0 + a1*b1 + a2*b2 evaluated left to right (Matrix3.multiplyMatrices / Vector3.applyMatrix4 order), f64."""
    return (a0 * b0 + a1 * b1) + a2 * b2


def decode(data: bytes, *, minimum_alpha: int = 1, half_covariances: bool = False, transform16=None) -> dict:
    """`transform16` (column-major 4x4, f64): baked like fillSplatDataArrays does for a static mesh (SplatMesh.js:1872-1897):
    centre.applyMatrix4 (:340-342), T3 (M M^T) T3^T (:461-466), SH decoded to floats, rotated (:684-716) and re-encoded at the
    GPU-side level (toHalfFloat / toUint8, :663-676)."""
    h = parse(data)
    buf = np.frombuffer(data, np.uint8)
    level = h.compression_level
    out = {k: [] for k in ("centers", "scales", "rot", "colors", "sh")}
    sh_degree = min(s.sh_degree for s in h.sections) if h.sections else 0
    ncomp_out = _SH_COMPONENTS[sh_degree]
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
    T = None if transform16 is None else np.asarray(transform16, np.float64).reshape(16)
    if T is not None:     # Vector3.applyMatrix4 on the decoded (f32) centre, JS doubles, stored back into a Float32Array
        x, y, z = (centers[:, k].astype(np.float64) for k in range(3))
        w = 1.0 / (((T[3] * x + T[7] * y) + T[11] * z) + T[15])
        centers = np.stack([(((T[0] * x + T[4] * y) + T[8] * z) + T[12]) * w, (((T[1] * x + T[5] * y) + T[9] * z) + T[13]) * w,
                            (((T[2] * x + T[6] * y) + T[10] * z) + T[14]) * w], 1).astype(np.float32)
    sh = None
    if ncomp_out:
        sh = np.concatenate(out["sh"])
        if T is None:
            if level == 0:
                sh = to_half_three(sh).view(np.float16)      # stored at compression level max(1, 0) = 1 on the GPU (THREE toHalfFloat)
        else:
            lo = h.min_sh if h.min_sh != 0 else -1.5
            hi = h.max_sh if h.max_sh != 0 else 1.5
            if level == 2:        # fromUint8 (:27-30) with the file's range, JS doubles
                f = sh.astype(np.float64) / 255 * (float(hi) - float(lo)) + float(lo)
            else:                 # f32 as is / fromHalfFloat (exact)
                f = sh.astype(np.float64)
            m1, m2 = sh_rotation_matrices(rotation_of_transform(T))
            tri = f.reshape(n, -1, 3)
            shr = np.empty_like(tri)
            for l in range(3):    # dot3: ((0 + in1 t0) + in2 t1) + in3 t2, per colour channel (:736-746)
                shr[:, l] = (tri[:, 0] * m1[l, 0] + tri[:, 1] * m1[l, 1]) + tri[:, 2] * m1[l, 2]
            if sh_degree >= 2:
                for l in range(5):    # dot5 (:752-772)
                    shr[:, 3 + l] = ((((tri[:, 3] * m2[l, 0] + tri[:, 4] * m2[l, 1]) + tri[:, 5] * m2[l, 2]) + tri[:, 6] * m2[l, 3]) + tri[:, 7] * m2[l, 4])
            flat = shr.reshape(n, -1)
            if level == 2:        # toUint8 (:21-25)
                v = np.clip(flat, float(lo), float(hi))
                sh = np.clip(np.floor((v - float(lo)) / (float(hi) - float(lo)) * 255), 0, 255).astype(np.uint8)
            else:                 # toHalfFloat of the f32-rounded value
                sh = to_half_three(flat.astype(np.float32)).view(np.float16)
    M = _rotation_matrices(rot) * scales.astype(np.float64)[:, None, :]
    def dot(r0, r1):  # Matrix3.multiplyMatrices: a1*b1 + a2*b2 + a3*b3, left to right, unfused
        return (M[:, r0, 0] * M[:, r1, 0] + M[:, r0, 1] * M[:, r1, 1]) + M[:, r0, 2] * M[:, r1, 2]
    cov6 = np.stack([dot(0, 0), dot(0, 1), dot(0, 2), dot(1, 1), dot(1, 2), dot(2, 2)], 1)
    if T is not None:     # transformedCovariance.multiply(T3^T).premultiply(T3): X = S T3^T, then Y = T3 X, ordered sums (:461-466)
        t3 = [[T[0], T[4], T[8]], [T[1], T[5], T[9]], [T[2], T[6], T[10]]]        # t3[row][col]
        S = {(0, 0): cov6[:, 0], (0, 1): cov6[:, 1], (0, 2): cov6[:, 2], (1, 1): cov6[:, 3], (1, 2): cov6[:, 4], (2, 2): cov6[:, 5]}
        sym = lambda i, j: S[(i, j)] if i <= j else S[(j, i)]
        X = [[_ordered3(sym(i, 0), t3[j][0], sym(i, 1), t3[j][1], sym(i, 2), t3[j][2]) for j in range(3)] for i in range(3)]
        Y = lambda i, j: _ordered3(t3[i][0], X[0][j], t3[i][1], X[1][j], t3[i][2], X[2][j])
        # elements [0], [3], [6], [4], [7], [8] of the column-major result = (0,0), (0,1), (0,2), (1,1), (1,2), (2,2)
        cov6 = np.stack([Y(0, 0), Y(0, 1), Y(0, 2), Y(1, 1), Y(1, 2), Y(2, 2)], 1)
    cov6 = to_half_three(cov6.astype(np.float32)).view(np.float16) if half_covariances else cov6.astype(np.float32)
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
