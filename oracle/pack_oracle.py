"""oracle/pack_oracle.py -- TEST INFRASTRUCTURE ONLY.

Scalar restatement of how the reference bakes a static scene transform into one splat at load time, written per splat with
plain Python floats (f64, like the JS numbers) so that the vectorised product code (gaussiansplats3d_b200/scenes.py) has an
independent checker.  Follows /root/reference/src/loaders/SplatBuffer.js:
  centre       applyMatrix4                                   :340-342
  covariance   T3 * (M M^T) * T3^T with M = R(q) * diag(s)     :440-486
  SH band 1    three dot products with rows built from the rotation matrix   :628-634, :736-746, :774-778
  SH band 2    five dot products with rows built from the band-1 rows        :752-772, :780-816
parity unpinned by reference vectors (the reference ships no tests); pinned instead by the rendering invariant in
tests/test_host_logic.py (a baked transform must equal the same transform applied as the model matrix)."""
import math


def mat3_from_quaternion(q):
    x, y, z, w = q
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]


def band1_rows(r):
    """r[row][col] of the rotation; the three weight rows of the first band."""
    return ([r[1][1], -r[1][2], r[1][0]], [-r[2][1], r[2][2], -r[2][0]], [r[0][1], -r[0][2], r[0][0]])


def band2_rows(t11, t12, t13):
    k14, k34, k13, k43, k112 = math.sqrt(1 / 4), math.sqrt(3 / 4), math.sqrt(1 / 3), math.sqrt(4 / 3), math.sqrt(1 / 12)
    t21 = [k14 * ((t13[2] * t11[0] + t13[0] * t11[2]) + (t11[2] * t13[0] + t11[0] * t13[2])),
           t13[1] * t11[0] + t11[1] * t13[0],
           k34 * (t13[1] * t11[1] + t11[1] * t13[1]),
           t13[1] * t11[2] + t11[1] * t13[2],
           k14 * ((t13[2] * t11[2] - t13[0] * t11[0]) + (t11[2] * t13[2] - t11[0] * t13[0]))]
    t22 = [k14 * ((t12[2] * t11[0] + t12[0] * t11[2]) + (t11[2] * t12[0] + t11[0] * t12[2])),
           t12[1] * t11[0] + t11[1] * t12[0],
           k34 * (t12[1] * t11[1] + t11[1] * t12[1]),
           t12[1] * t11[2] + t11[1] * t12[2],
           k14 * ((t12[2] * t11[2] - t12[0] * t11[0]) + (t11[2] * t12[2] - t11[0] * t12[0]))]
    t23 = [k13 * (t12[2] * t12[0] + t12[0] * t12[2]) - k112 * ((t13[2] * t13[0] + t13[0] * t13[2]) + (t11[2] * t11[0] + t11[0] * t11[2])),
           k43 * t12[1] * t12[0] - k13 * (t13[1] * t13[0] + t11[1] * t11[0]),
           t12[1] * t12[1] - k14 * (t13[1] * t13[1] + t11[1] * t11[1]),
           k43 * t12[1] * t12[2] - k13 * (t13[1] * t13[2] + t11[1] * t11[2]),
           k13 * (t12[2] * t12[2] - t12[0] * t12[0]) - k112 * ((t13[2] * t13[2] - t13[0] * t13[0]) + (t11[2] * t11[2] - t11[0] * t11[0]))]
    t24 = [k14 * ((t12[2] * t13[0] + t12[0] * t13[2]) + (t13[2] * t12[0] + t13[0] * t12[2])),
           t12[1] * t13[0] + t13[1] * t12[0],
           k34 * (t12[1] * t13[1] + t13[1] * t12[1]),
           t12[1] * t13[2] + t13[1] * t12[2],
           k14 * ((t12[2] * t13[2] - t12[0] * t13[0]) + (t13[2] * t12[2] - t13[0] * t12[0]))]
    t25 = [k14 * ((t13[2] * t13[0] + t13[0] * t13[2]) - (t11[2] * t11[0] + t11[0] * t11[2])),
           t13[1] * t13[0] - t11[1] * t11[0],
           k34 * (t13[1] * t13[1] - t11[1] * t11[1]),
           t13[1] * t13[2] - t11[1] * t11[2],
           k14 * ((t13[2] * t13[2] - t13[0] * t13[0]) - (t11[2] * t11[2] - t11[0] * t11[0]))]
    return t21, t22, t23, t24, t25


def weighted_sum(vectors, weights):
    """out = sum_k vectors[k] * weights[k] over RGB triples, accumulated in coefficient order."""
    out = [0.0, 0.0, 0.0]
    for v, w in zip(vectors, weights):
        for c in range(3):
            out[c] = out[c] + v[c] * w
    return out


def bake_one(center, scale, quat_xyzw, sh_triples, degree, transform_colmajor16, rotation3x3):
    """One splat.  `sh_triples`: list of RGB triples (3 for degree 1, 8 for degree 2); `rotation3x3`: the normalised rotation of the
    transform (row, col).  Returns (centre, covariance6, sh_triples) as Python floats."""
    e = transform_colmajor16
    x, y, z = center
    w = 1.0 / (e[3] * x + e[7] * y + e[11] * z + e[15])
    c = [(e[0] * x + e[4] * y + e[8] * z + e[12]) * w, (e[1] * x + e[5] * y + e[9] * z + e[13]) * w, (e[2] * x + e[6] * y + e[10] * z + e[14]) * w]
    r = mat3_from_quaternion(quat_xyzw)
    m = [[r[i][j] * scale[j] for j in range(3)] for i in range(3)]
    cov = [[sum(m[i][k] * m[j][k] for k in range(3)) for j in range(3)] for i in range(3)]
    t3 = [[e[0], e[4], e[8]], [e[1], e[5], e[9]], [e[2], e[6], e[10]]]
    tc = [[sum(t3[i][k] * cov[k][j] for k in range(3)) for j in range(3)] for i in range(3)]
    tct = [[sum(tc[i][k] * t3[j][k] for k in range(3)) for j in range(3)] for i in range(3)]
    cov6 = [tct[0][0], tct[0][1], tct[0][2], tct[1][1], tct[1][2], tct[2][2]]
    out_sh = []
    if degree >= 1:
        t11, t12, t13 = band1_rows(rotation3x3)
        out_sh += [weighted_sum(sh_triples[0:3], t) for t in (t11, t12, t13)]
        if degree >= 2:
            out_sh += [weighted_sum(sh_triples[3:8], t) for t in band2_rows(t11, t12, t13)]
    return c, cov6, out_sh
