"""Host-side quaternion / rotation helpers (numpy, vectorised over a leading axis).

Mirrors dvi_ekf/tools/Quaternion.py for the *inputs* of the hot path (camera
trajectory pre-pass, initial state) and for formatting its outputs.  All
quaternions are xyzw.  The filter arithmetic itself runs on the GPU.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation


def normalise(q):
    """Quaternion.normalise (Quaternion.py:195-206): unit norm, scalar part >= 0."""
    q = np.asarray(q, dtype=float)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return np.where(q[..., 3:4] < 0, -q, q)


def multiply(a, b):
    """Hamilton product a * b of xyzw quaternions (Quaternion.__mul__, Quaternion.py:170-183) WITHOUT the final
    re-normalisation -- callers apply ``normalise`` like the reference's ``do_normalise=True``."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    aw, av, bw, bv = a[..., 3:4], a[..., :3], b[..., 3:4], b[..., :3]
    w = aw * bw - np.sum(av * bv, axis=-1, keepdims=True)
    v = aw * bv + bw * av + np.cross(av, bv)
    return np.concatenate((v, w), axis=-1)


def to_matrix(q):
    """Quaternion.rot (Quaternion.py:101-107) for unit quaternions; shape (...,3,3)."""
    q = np.asarray(q, dtype=float)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = x * x - y * y - z * z + w * w
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = -x * x + y * y - z * z + w * w
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = -x * x - y * y + z * z + w * w
    return R


def from_matrix_markley(M):
    """Quaternion(val=M, do_normalise=True) with scipy-1.10.1's Rotation.from_matrix (Markley's method on
    the raw matrix, no orthogonalisation) -- the version the reference pins (requirements.txt:8)."""
    M = np.asarray(M, dtype=float)
    single = M.ndim == 2
    M = M.reshape(-1, 3, 3)
    d = np.stack([M[:, 0, 0], M[:, 1, 1], M[:, 2, 2], M[:, 0, 0] + M[:, 1, 1] + M[:, 2, 2]], -1)
    choice = np.argmax(d, -1)
    q = np.empty((M.shape[0], 4))
    for n in range(M.shape[0]):
        c, m = choice[n], M[n]
        if c == 3:
            q[n] = [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], 1 + d[n, 3]]
        else:
            i, j, k = c, (c + 1) % 3, (c + 2) % 3
            q[n, i] = 1 - d[n, 3] + 2 * m[i, i]
            q[n, j] = m[j, i] + m[i, j]
            q[n, k] = m[k, i] + m[i, k]
            q[n, 3] = m[k, j] - m[j, k]
    q = normalise(q)
    return q[0] if single else q


def euler_xyz(q, degrees=False):
    """Quaternion.euler_xyz_rad / _deg (Quaternion.py:117-123): extrinsic xyz."""
    return Rotation.from_quat(np.asarray(q, dtype=float)).as_euler("xyz", degrees=degrees)


def euler_zyx_reversed(q):
    """Euler convention of the older revision that produced the reference's golden files (quirk Q11)."""
    return Rotation.from_quat(np.asarray(q, dtype=float)).as_euler("zyx")[..., ::-1]


class Quaternion:
    """Read-only view with the attribute names of dvi_ekf.tools.Quaternion.Quaternion."""

    def __init__(self, xyzw):
        self._q = np.asarray(xyzw, dtype=float).reshape(4)

    x = property(lambda s: float(s._q[0]))
    y = property(lambda s: float(s._q[1]))
    z = property(lambda s: float(s._q[2]))
    w = property(lambda s: float(s._q[3]))
    xyzw = property(lambda s: s._q.copy())
    wxyz = property(lambda s: s._q[[3, 0, 1, 2]].copy())
    v = property(lambda s: s._q[:3].copy())
    rot = property(lambda s: to_matrix(s._q))
    euler_xyz_rad = property(lambda s: euler_xyz(s._q))
    euler_xyz_deg = property(lambda s: euler_xyz(s._q, degrees=True))

    def __repr__(self):
        return f"Quaternion [x={self.x:.3f}, y={self.y:.3f}, z={self.z:.3f}, w={self.w:.3f}]"
