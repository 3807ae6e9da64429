"""Hand-assembled `.ksplat` fixtures + per-splat expected values.  TEST INFRASTRUCTURE ONLY.

Every byte is placed with struct.pack at the offsets of SURVEY.md Appendix A (= /root/reference/src/loaders/SplatBuffer.js:
writeHeaderToBuffer :856-875, writeSectionHeaderToBuffer :944-961, CompressionLevels :108-163, section body layout :926-927), and
every expected value is computed with scalar Python arithmetic straight from the reference's decode semantics
(fillSplatCenterArray :307-347, getBucketIndex :199-219, computeCovariance :440-486 with three's Matrix3 operation order,
fillSplatColorArray :522-549, fillSphericalHarmonicsArray :551-734, SplatMesh.getIntegerCenters SplatMesh.js:1912-1926).
Imports neither the product package nor oracle/ -- it is the third party both are checked against.

`python tests/golden/ksplat_handmade.py` rewrites tests/golden/ksplat_handmade_*.ksplat (committed, a few KB each)."""
from __future__ import annotations

import math
import random
import struct
from pathlib import Path

HERE = Path(__file__).resolve().parent

# bytes: centre, scale, rotation, colour, per SH component; default compressionScaleRange        SplatBuffer.js:108-163
LEVEL = {0: (12, 12, 16, 4, 4, 1), 1: (6, 6, 8, 4, 2, 32767), 2: (6, 6, 8, 4, 1, 32767)}
NCOMP = {0: 0, 1: 9, 2: 24}


def f32(x: float) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def half_bits_rne(x: float) -> int:
    return struct.unpack("<H", struct.pack("<e", x))[0]


def half_value(bits: int) -> float:
    return struct.unpack("<e", struct.pack("<H", bits))[0]


def half_bits_truncated(x: float) -> int:
    """THREE.DataUtils.toHalfFloat truncates the mantissa.  Derived here from IEEE rounding instead of three's tables: take the
    nearest half and step one ulp toward zero if it overshot in magnitude (normal range only -- the fixtures stay inside it)."""
    x = max(-65504.0, min(65504.0, f32(x)))
    b = half_bits_rne(x)
    if abs(half_value(b)) > abs(x):
        b -= 1
    return b


def rotation_matrix(x, y, z, w):
    """Matrix4.makeRotationFromQuaternion = compose(0, q, 1) (three r160), rows of the 3x3."""
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    return [[1 - (yy + zz), xy - wz, xz + wy], [xy + wz, 1 - (xx + zz), yz - wx], [xz - wy, yz + wx, 1 - (xx + yy)]]


def covariance6(scale, quat_xyzw):
    """SplatBuffer.computeCovariance without transform: M = R S, M M^T with Matrix3.multiplyMatrices' left-to-right sums; returns the
    f32-rounded [m00 m01 m02 m11 m12 m22]."""
    r = rotation_matrix(*quat_xyzw)
    m = [[(r[i][0] * (scale[0] if j == 0 else 0.0) + r[i][1] * (scale[1] if j == 1 else 0.0)) + r[i][2] * (scale[2] if j == 2 else 0.0) for j in range(3)] for i in range(3)]
    c = lambda i, j: (m[i][0] * m[j][0] + m[i][1] * m[j][1]) + m[i][2] * m[j][2]
    return [f32(c(0, 0)), f32(c(0, 1)), f32(c(0, 2)), f32(c(1, 1)), f32(c(1, 2)), f32(c(2, 2))]


def build(level: int, sh_degree: int, sections: list[dict], seed: int, *, scene_center=(0.25, -1.5, 3.0), sh_range=(-1.25, 1.75), minimum_alpha=1):
    """sections: [{'n': splats, 'bucket_size': k, 'block': f}] -> (file bytes, expected per-splat values)."""
    rng = random.Random(seed)
    cb, sb, rb, colb, shb, default_range = LEVEL[level]
    ncomp = NCOMP[sh_degree]
    bps = cb + sb + rb + colb + shb * ncomp
    lo, hi = sh_range
    header = bytearray(4096)
    struct.pack_into("<BB", header, 0, 0, 1)                               # version 0.1
    total = sum(s["n"] for s in sections)
    struct.pack_into("<4I", header, 4, len(sections), len(sections), total, total)
    struct.pack_into("<H", header, 20, level)
    struct.pack_into("<3f", header, 24, *scene_center)
    struct.pack_into("<2f", header, 36, lo, hi)
    sec_headers, bodies = bytearray(), bytearray()
    exp = dict(centers=[], rgba=[], cov=[], sh=[], int_centers=[], scales=[], rot_xyzw=[])
    for s in sections:
        n, bucket_size, block = s["n"], s.get("bucket_size", 4), s.get("block", 5.0)
        sr = default_range
        if level >= 1:
            full = n // bucket_size if s.get("all_partial") is None else 0
            rest = n - full * bucket_size
            partial = []
            while rest > 0:                                                   # several partially filled buckets of differing length
                take = min(rest, max(1, bucket_size - 1 - len(partial) % 2))
                partial.append(take)
                rest -= take
            lengths = [bucket_size] * full + partial
            bucket_centres = [[f32(rng.uniform(-8, 8)) for _ in range(3)] for _ in lengths]
        else:
            full, partial, lengths, bucket_centres = 0, [], [], []
        body = bytearray()
        for L in partial:
            body += struct.pack("<I", L)
        for bc in bucket_centres:
            body += struct.pack("<3f", *bc)
        bucket_of = [b for b, L in enumerate(lengths) for _ in range(L)]
        for i in range(n):
            rec = bytearray(bps)
            # ---- centre -------------------------------------------------------------------------------------------------------
            if level == 0:
                c = [f32(rng.uniform(-6, 6)) for _ in range(3)]
                struct.pack_into("<3f", rec, 0, *c)
            else:
                q = [rng.randrange(0, 2 * sr + 1) for _ in range(3)]
                if i == 0:
                    q = [0, sr, 2 * sr]                                        # range ends and the exact middle
                struct.pack_into("<3H", rec, 0, *q)
                block32 = f32(block)
                sf = (block32 / 2.0) / sr                                      # compressionScaleFactor (JS doubles)
                bc = bucket_centres[bucket_of[i]]
                c = [f32((q[k] - sr) * sf + bc[k]) for k in range(3)]          # stored into a Float32Array
            # ---- scale, rotation (file order w, x, y, z) ------------------------------------------------------------------------
            sc = [math.exp(rng.uniform(-6, -1)) for _ in range(3)]
            qv = [rng.gauss(0, 1) for _ in range(4)]
            nq = math.sqrt(sum(v * v for v in qv))
            wxyz = [v / nq for v in qv]
            if level == 0:
                sc = [f32(v) for v in sc]
                wxyz = [f32(v) for v in wxyz]
                struct.pack_into("<3f", rec, cb, *sc)
                struct.pack_into("<4f", rec, cb + sb, *wxyz)
            else:
                hb = [half_bits_rne(v) for v in sc]
                struct.pack_into("<3H", rec, cb, *hb)
                sc = [half_value(b) for b in hb]
                hq = [half_bits_rne(v) for v in wxyz]
                struct.pack_into("<4H", rec, cb + sb, *hq)
                wxyz = [half_value(b) for b in hq]
            xyzw = [wxyz[1], wxyz[2], wxyz[3], wxyz[0]]
            # ---- colour ---------------------------------------------------------------------------------------------------------
            rgba = [rng.randrange(256) for _ in range(4)]
            if i == 1:
                rgba[3] = 0                                                    # below minimumAlpha: rendered with alpha 0
            struct.pack_into("<4B", rec, cb + sb + rb, *rgba)
            # ---- spherical harmonics: file = channel-major per band; GPU array = coefficient-major RGB triples ------------------------------
            sh_out = []
            if ncomp:
                triples = [[rng.uniform(lo * 0.9, hi * 0.9) for _ in range(3)] for _ in range(ncomp // 3)]     # [coef][channel]
                file_vals = [triples[k][ch] for ch in range(3) for k in range(3)]
                if sh_degree == 2:
                    file_vals += [triples[3 + k][ch] for ch in range(3) for k in range(5)]
                at = cb + sb + rb + colb
                if level == 0:
                    stored = [f32(v) for v in file_vals]
                    struct.pack_into(f"<{ncomp}f", rec, at, *stored)
                    decode = lambda v: half_bits_truncated(v)                  # kept at half precision on the GPU (SplatMesh.js:1064-1066)
                elif level == 1:
                    hb = [half_bits_rne(v) for v in file_vals]
                    struct.pack_into(f"<{ncomp}H", rec, at, *hb)
                    stored = hb
                    decode = lambda b: b
                else:
                    u8 = [max(0, min(255, math.floor((max(lo, min(hi, v)) - lo) / (hi - lo) * 255))) for v in file_vals]
                    struct.pack_into(f"<{ncomp}B", rec, at, *u8)
                    stored = u8
                    decode = lambda b: b
                # re-order: output index 3*coef + channel
                for k in range(3):
                    for ch in range(3):
                        sh_out.append(decode(stored[ch * 3 + k]))
                if sh_degree == 2:
                    for k in range(5):
                        for ch in range(3):
                            sh_out.append(decode(stored[9 + ch * 5 + k]))
            body += rec
            exp["centers"].append(c)
            a = rgba[3] if rgba[3] >= minimum_alpha else 0
            exp["rgba"].append(rgba[0] | (rgba[1] << 8) | (rgba[2] << 16) | (a << 24))
            exp["cov"].append(covariance6(sc, xyzw))
            exp["sh"].append(sh_out)
            exp["int_centers"].append([math.floor(v * 1000.0 + 0.5) for v in c] + [1000])
            exp["scales"].append(sc)
            exp["rot_xyzw"].append(xyzw)
        sh_ = bytearray(1024)
        struct.pack_into("<2I", sh_, 0, n, n)
        if level >= 1:
            struct.pack_into("<2I", sh_, 8, bucket_size, len(lengths))
            struct.pack_into("<f", sh_, 16, block)
            struct.pack_into("<H", sh_, 20, 12)
            struct.pack_into("<I", sh_, 24, 0 if s.get("default_range") else sr)       # 0 -> the level's default range (:904-905)
            struct.pack_into("<2I", sh_, 32, full, len(partial))
        struct.pack_into("<I", sh_, 28, len(body))
        struct.pack_into("<H", sh_, 40, sh_degree)
        sec_headers += sh_
        bodies += body
    exp.update(level=level, sh_degree=sh_degree, count=total, sh_range=(f32(lo), f32(hi)), scene_center=tuple(f32(v) for v in scene_center),
               bytes_per_splat=bps, sections=len(sections))
    return bytes(header) + bytes(sec_headers) + bytes(bodies), exp


FIXTURES = {
    # name: (level, SH degree, sections, seed)
    "l0_sh2": (0, 2, [dict(n=6)], 101),
    "l0_sh0": (0, 0, [dict(n=5)], 102),
    "l1_sh1": (1, 1, [dict(n=11, bucket_size=4)], 103),                       # 2 full buckets + partial ones
    "l2_sh2_two_sections": (2, 2, [dict(n=5, bucket_size=4, block=5.0), dict(n=3, bucket_size=4, block=2.5, all_partial=True, default_range=True)], 104),
    "l1_sh0_many_buckets": (1, 0, [dict(n=37, bucket_size=3, block=1.25)], 105),
}


def fixture(name: str):
    level, deg, sections, seed = FIXTURES[name]
    return build(level, deg, sections, seed)


if __name__ == "__main__":
    for name in FIXTURES:
        data, exp = fixture(name)
        (HERE / f"ksplat_handmade_{name}.ksplat").write_bytes(data)
        print(name, len(data), "bytes,", exp["count"], "splats")
