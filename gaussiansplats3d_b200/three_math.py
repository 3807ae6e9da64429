"""The slice of three.js (r160) math the hot path depends on, restated in float64 NumPy.

three.js is a peer dependency of the reference (package.json:69-71) and is NOT vendored under /root/reference, so
these follow three's documented conventions (column-major `elements`, right-handed, camera looks down -Z) and
algorithms as published in three@0.160.0 -- PARITY UNPINNED against the real library (no JS engine here).
Call sites in the reference: Viewer.js:338 (PerspectiveCamera fov 50, near 0.1, far 1000), :341-343 (position / up /
lookAt), :1888-1891 (mvp = proj * inv(camera.matrixWorld) * mesh.matrixWorld), :662-673 (focal lengths).
"""
from __future__ import annotations

import numpy as np


def identity() -> np.ndarray:
    return np.eye(4, dtype=np.float64).T.reshape(16).copy()


def to_mat(e) -> np.ndarray:
    """column-major elements[16] -> 4x4 ndarray (row, col)."""
    return np.asarray(e, np.float64).reshape(4, 4).T


def to_elements(m) -> np.ndarray:
    return np.asarray(m, np.float64).T.reshape(16).copy()


def multiply(a, b) -> np.ndarray:
    """Matrix4.multiplyMatrices(a, b) = a * b."""
    return to_elements(to_mat(a) @ to_mat(b))


def invert(e) -> np.ndarray:
    """Matrix4.invert(): cofactor expansion (three evaluates the same closed form; a singular matrix gives zeros)."""
    n11, n21, n31, n41, n12, n22, n32, n42, n13, n23, n33, n43, n14, n24, n34, n44 = (float(v) for v in np.asarray(e, np.float64).reshape(16))
    t11 = n23 * n34 * n42 - n24 * n33 * n42 + n24 * n32 * n43 - n22 * n34 * n43 - n23 * n32 * n44 + n22 * n33 * n44
    t12 = n14 * n33 * n42 - n13 * n34 * n42 - n14 * n32 * n43 + n12 * n34 * n43 + n13 * n32 * n44 - n12 * n33 * n44
    t13 = n13 * n24 * n42 - n14 * n23 * n42 + n14 * n22 * n43 - n12 * n24 * n43 - n13 * n22 * n44 + n12 * n23 * n44
    t14 = n14 * n23 * n32 - n13 * n24 * n32 - n14 * n22 * n33 + n12 * n24 * n33 + n13 * n22 * n34 - n12 * n23 * n34
    det = n11 * t11 + n21 * t12 + n31 * t13 + n41 * t14
    if det == 0:
        return np.zeros(16)
    d = 1.0 / det
    out = np.empty(16)
    out[0] = t11 * d
    out[1] = (n24 * n33 * n41 - n23 * n34 * n41 - n24 * n31 * n43 + n21 * n34 * n43 + n23 * n31 * n44 - n21 * n33 * n44) * d
    out[2] = (n22 * n34 * n41 - n24 * n32 * n41 + n24 * n31 * n42 - n21 * n34 * n42 - n22 * n31 * n44 + n21 * n32 * n44) * d
    out[3] = (n23 * n32 * n41 - n22 * n33 * n41 - n23 * n31 * n42 + n21 * n33 * n42 + n22 * n31 * n43 - n21 * n32 * n43) * d
    out[4] = t12 * d
    out[5] = (n13 * n34 * n41 - n14 * n33 * n41 + n14 * n31 * n43 - n11 * n34 * n43 - n13 * n31 * n44 + n11 * n33 * n44) * d
    out[6] = (n14 * n32 * n41 - n12 * n34 * n41 - n14 * n31 * n42 + n11 * n34 * n42 + n12 * n31 * n44 - n11 * n32 * n44) * d
    out[7] = (n12 * n33 * n41 - n13 * n32 * n41 + n13 * n31 * n42 - n11 * n33 * n42 - n12 * n31 * n43 + n11 * n32 * n43) * d
    out[8] = t13 * d
    out[9] = (n14 * n23 * n41 - n13 * n24 * n41 - n14 * n21 * n43 + n11 * n24 * n43 + n13 * n21 * n44 - n11 * n23 * n44) * d
    out[10] = (n12 * n24 * n41 - n14 * n22 * n41 + n14 * n21 * n42 - n11 * n24 * n42 - n12 * n21 * n44 + n11 * n22 * n44) * d
    out[11] = (n13 * n22 * n41 - n12 * n23 * n41 - n13 * n21 * n42 + n11 * n23 * n42 + n12 * n21 * n43 - n11 * n22 * n43) * d
    out[12] = t14 * d
    out[13] = (n13 * n24 * n31 - n14 * n23 * n31 + n14 * n21 * n33 - n11 * n24 * n33 - n13 * n21 * n34 + n11 * n23 * n34) * d
    out[14] = (n14 * n22 * n31 - n12 * n24 * n31 - n14 * n21 * n32 + n11 * n24 * n32 + n12 * n21 * n34 - n11 * n22 * n34) * d
    out[15] = (n12 * n23 * n31 - n13 * n22 * n31 + n13 * n21 * n32 - n11 * n23 * n32 - n12 * n21 * n33 + n11 * n22 * n33) * d
    return out


def make_perspective(fov_deg: float, aspect: float, near: float, far: float, zoom: float = 1.0) -> np.ndarray:
    """PerspectiveCamera.updateProjectionMatrix + Matrix4.makePerspective (WebGL clip space)."""
    top = near * np.tan(np.deg2rad(0.5 * fov_deg)) / zoom
    height = 2.0 * top
    width = aspect * height
    left = -0.5 * width
    right, bottom = left + width, top - height
    e = np.zeros(16)
    e[0] = 2 * near / (right - left)
    e[5] = 2 * near / (top - bottom)
    e[8] = (right + left) / (right - left)
    e[9] = (top + bottom) / (top - bottom)
    e[10] = -(far + near) / (far - near)
    e[11] = -1.0
    e[14] = -2 * far * near / (far - near)
    return e


def make_orthographic(left, right, top, bottom, near, far, zoom: float = 1.0) -> np.ndarray:
    """OrthographicCamera.updateProjectionMatrix + Matrix4.makeOrthographic."""
    dx, dy = (right - left) / (2 * zoom), (top - bottom) / (2 * zoom)
    cx, cy = (right + left) / 2, (top + bottom) / 2
    left, right, top, bottom = cx - dx, cx + dx, cy + dy, cy - dy
    w, h, p = 1.0 / (right - left), 1.0 / (top - bottom), 1.0 / (far - near)
    e = np.zeros(16)
    e[0], e[5], e[10] = 2 * w, 2 * h, -2 * p
    e[12], e[13], e[14], e[15] = -(right + left) * w, -(top + bottom) * h, -(far + near) * p, 1.0
    return e


def camera_world_matrix(position, target, up) -> np.ndarray:
    """Object3D.lookAt for a camera (Matrix4.lookAt(eye=position, target, up)) composed with the position."""
    eye, tgt, upv = (np.asarray(v, np.float64) for v in (position, target, up))
    z = eye - tgt
    if np.dot(z, z) == 0:
        z = np.array([0.0, 0.0, 1.0])
    z = z / np.linalg.norm(z)
    x = np.cross(upv, z)
    if np.dot(x, x) == 0:  # up parallel to view direction: three nudges z
        z = z + np.array([0.0001, 0.0, 0.0]) if abs(upv[2]) == 1 else z + np.array([0.0, 0.0, 0.0001])
        z = z / np.linalg.norm(z)
        x = np.cross(upv, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return to_elements(m)


def compose(position, quaternion_xyzw, scale) -> np.ndarray:
    """Matrix4.compose."""
    x, y, z, w = (float(v) for v in quaternion_xyzw)
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz, yy, yz, zz, wx, wy, wz = x * x2, x * y2, x * z2, y * y2, y * z2, z * z2, w * x2, w * y2, w * z2
    sx, sy, sz = (float(v) for v in scale)
    e = np.zeros(16)
    e[0], e[1], e[2] = (1 - (yy + zz)) * sx, (xy + wz) * sx, (xz - wy) * sx
    e[4], e[5], e[6] = (xy - wz) * sy, (1 - (xx + zz)) * sy, (yz + wx) * sy
    e[8], e[9], e[10] = (xz + wy) * sz, (yz - wx) * sz, (1 - (xx + yy)) * sz
    e[12], e[13], e[14], e[15] = float(position[0]), float(position[1]), float(position[2]), 1.0
    return e


def rotate_about_axis(v, axis, angle_rad) -> np.ndarray:
    """Vector3.applyAxisAngle (Rodrigues)."""
    v, k = np.asarray(v, np.float64), np.asarray(axis, np.float64)
    k = k / np.linalg.norm(k)
    return v * np.cos(angle_rad) + np.cross(k, v) * np.sin(angle_rad) + k * np.dot(k, v) * (1 - np.cos(angle_rad))


class PerspectiveCamera:
    """Just enough of THREE.PerspectiveCamera: position/up/lookAt -> matrixWorld, projectionMatrix."""

    isOrthographicCamera = False

    def __init__(self, fov=50.0, aspect=1.0, near=0.1, far=1000.0):
        self.fov, self.aspect, self.near, self.far, self.zoom = fov, aspect, near, far, 1.0
        self.position = np.zeros(3)
        self.up = np.array([0.0, 1.0, 0.0])
        self.target = np.array([0.0, 0.0, -1.0])
        self.update()

    def look_at(self, target):
        self.target = np.asarray(target, np.float64)
        self.update()

    def update(self):
        self.projectionMatrix = make_perspective(self.fov, self.aspect, self.near, self.far, self.zoom)
        self.matrixWorld = camera_world_matrix(self.position, self.target, self.up)
        self.matrixWorldInverse = invert(self.matrixWorld)
