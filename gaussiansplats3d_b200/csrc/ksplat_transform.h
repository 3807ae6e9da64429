// ksplat_transform.h -- host-only: the doubles k_ksplat_decode<true> needs to bake a static scene transform (plain C++, no CUDA,
// so that the very same code can be compiled and checked on the host by the test suite against the Python restatement).
#pragma once
#include <cmath>
#include <cstring>

namespace gs {

struct KTransform {
    double t[16];        // column-major Matrix4 elements
    double m1[3][3];     // SH band-1 weights  (SplatBuffer.js:632-634): row l = weights of the input coefficients for output l
    double m2[5][5];     // SH band-2 weights  (rotateSphericalHarmonics5 :780-816)
    double sh_lo, sh_hi; // 8-bit SH range of the file
};

// three.js operation order for the rotation a scene transform applies to spherical harmonics (SplatBuffer.js:628-634):
// Matrix4.decompose -> Quaternion.setFromRotationMatrix -> normalize -> makeRotationFromQuaternion; then the band weights.
inline void ksplat_transform_params(const double *e, double sh_lo, double sh_hi, KTransform &K) {
    memcpy(K.t, e, sizeof(K.t));
    K.sh_lo = sh_lo; K.sh_hi = sh_hi;
    double sx = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double sy = std::sqrt(e[4] * e[4] + e[5] * e[5] + e[6] * e[6]);
    const double sz = std::sqrt(e[8] * e[8] + e[9] * e[9] + e[10] * e[10]);
    const double det = e[0] * (e[5] * e[10] - e[9] * e[6]) - e[4] * (e[1] * e[10] - e[9] * e[2]) + e[8] * (e[1] * e[6] - e[5] * e[2]);
    if (det < 0) sx = -sx;
    const double isx = 1.0 / sx, isy = 1.0 / sy, isz = 1.0 / sz;
    const double m11 = e[0] * isx, m21 = e[1] * isx, m31 = e[2] * isx, m12 = e[4] * isy, m22 = e[5] * isy, m32 = e[6] * isy;
    const double m13 = e[8] * isz, m23 = e[9] * isz, m33 = e[10] * isz;
    double x, y, z, w;
    const double tr = m11 + m22 + m33;
    if (tr > 0) { const double k = 0.5 / std::sqrt(tr + 1.0); w = 0.25 / k; x = (m32 - m23) * k; y = (m13 - m31) * k; z = (m21 - m12) * k; }
    else if (m11 > m22 && m11 > m33) { const double k = 2.0 * std::sqrt(1.0 + m11 - m22 - m33); w = (m32 - m23) / k; x = 0.25 * k; y = (m12 + m21) / k; z = (m13 + m31) / k; }
    else if (m22 > m33) { const double k = 2.0 * std::sqrt(1.0 + m22 - m11 - m33); w = (m13 - m31) / k; x = (m12 + m21) / k; y = 0.25 * k; z = (m23 + m32) / k; }
    else { const double k = 2.0 * std::sqrt(1.0 + m33 - m11 - m22); w = (m21 - m12) / k; x = (m13 + m31) / k; y = (m23 + m32) / k; z = 0.25 * k; }
    double ln = std::sqrt(x * x + y * y + z * z + w * w);
    if (ln == 0) { x = y = z = 0; w = 1; } else { ln = 1.0 / ln; x *= ln; y *= ln; z *= ln; w *= ln; }
    const double x2 = x + x, y2 = y + y, z2 = z + z;
    const double xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2, wx = w * x2, wy = w * y2, wz = w * z2;
    const double r[3][3] = {{1 - (yy + zz), xy - wz, xz + wy}, {xy + wz, 1 - (xx + zz), yz - wx}, {xz - wy, yz + wx, 1 - (xx + yy)}};
    const double a[3] = {r[1][1], -r[1][2], r[1][0]}, b[3] = {-r[2][1], r[2][2], -r[2][0]}, c[3] = {r[0][1], -r[0][2], r[0][0]};   // tsh11, tsh12, tsh13
    for (int k = 0; k < 3; ++k) { K.m1[0][k] = a[k]; K.m1[1][k] = b[k]; K.m1[2][k] = c[k]; }
    const double k14 = std::sqrt(1.0 / 4.0), k34 = std::sqrt(3.0 / 4.0), k13 = std::sqrt(1.0 / 3.0), k43 = std::sqrt(4.0 / 3.0), k112 = std::sqrt(1.0 / 12.0);
    auto sym = [&](const double *u, const double *v, double *o) {
        o[0] = k14 * ((u[2] * v[0] + u[0] * v[2]) + (v[2] * u[0] + v[0] * u[2]));
        o[1] = u[1] * v[0] + v[1] * u[0];
        o[2] = k34 * (u[1] * v[1] + v[1] * u[1]);
        o[3] = u[1] * v[2] + v[1] * u[2];
        o[4] = k14 * ((u[2] * v[2] - u[0] * v[0]) + (v[2] * u[2] - v[0] * u[0]));
    };
    sym(c, a, K.m2[0]);
    sym(b, a, K.m2[1]);
    K.m2[2][0] = k13 * (b[2] * b[0] + b[0] * b[2]) + -k112 * ((c[2] * c[0] + c[0] * c[2]) + (a[2] * a[0] + a[0] * a[2]));
    K.m2[2][1] = k43 * b[1] * b[0] + -k13 * (c[1] * c[0] + a[1] * a[0]);
    K.m2[2][2] = b[1] * b[1] + -k14 * (c[1] * c[1] + a[1] * a[1]);
    K.m2[2][3] = k43 * b[1] * b[2] + -k13 * (c[1] * c[2] + a[1] * a[2]);
    K.m2[2][4] = k13 * (b[2] * b[2] - b[0] * b[0]) + -k112 * ((c[2] * c[2] - c[0] * c[0]) + (a[2] * a[2] - a[0] * a[0]));
    sym(b, c, K.m2[3]);
    K.m2[4][0] = k14 * ((c[2] * c[0] + c[0] * c[2]) - (a[2] * a[0] + a[0] * a[2]));
    K.m2[4][1] = c[1] * c[0] - a[1] * a[0];
    K.m2[4][2] = k34 * (c[1] * c[1] - a[1] * a[1]);
    K.m2[4][3] = c[1] * c[2] - a[1] * a[2];
    K.m2[4][4] = k14 * ((c[2] * c[2] - c[0] * c[0]) - (a[2] * a[2] - a[0] * a[0]));
}

} // namespace gs
