"""Closed-form probe kinematics on the host (numpy, vectorised) for the pre-pass.

Same closed form as the device code (dvi_ekf_b200/csrc/eskf_math.cuh, derived
in DESIGN.md from the DH chain of dvi_ekf/models/Probe.py:147-167):
    p = (L - q4) z6 + q5 e_a(q3) + q6 e_b(q3)
    R = [ -e_a(d) | sa z6 + ca e_b(d) | ca z6 - sa e_b(d) ],  d = q3 - q7
    v = acc = 0,  om = z6 q7',  alp = z6 q7''
"""
from __future__ import annotations

import numpy as np

GT_IMU_DOFS = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 20.0])  # SimpleProbe constraints, Probe.py:385-388


def fwkin(dofs, notch, length, angle):
    """dofs (...,6), notch (...,3) -> p (...,3), R (...,3,3), om (...,3), alp (...,3)."""
    dofs = np.asarray(dofs, dtype=float)
    notch = np.asarray(notch, dtype=float)
    q1, q2, q3, q4, q5, q6 = np.moveaxis(dofs, -1, 0)
    d = q3 - notch[..., 0]
    s1, c1, s2, c2, s3, c3, sd, cd = np.sin(q1), np.cos(q1), np.sin(q2), np.cos(q2), np.sin(q3), np.cos(q3), np.sin(d), np.cos(d)
    sa, ca = np.sin(angle), np.cos(angle)

    def ea(s, c):
        return np.stack([s1 * s2 * s - c1 * c, s1 * c + c1 * s2 * s, c2 * s], -1)

    def eb(s, c):
        return np.stack([s1 * s2 * c + c1 * s, c1 * s2 * c - s1 * s, c2 * c], -1)

    z6 = np.stack([-s1 * c2, -c1 * c2, s2], -1)
    p = (length - q4)[..., None] * z6 + q5[..., None] * ea(s3, c3) + q6[..., None] * eb(s3, c3)
    R = np.stack([-ea(sd, cd), sa * z6 + ca * eb(sd, cd), ca * z6 - sa * eb(sd, cd)], -1)
    return p, R, z6 * notch[..., 1:2], z6 * notch[..., 2:3]
