"""Camera trajectory and the IMU / measurement streams derived from it (host pre-pass).

Mirrors dvi_ekf/models/Camera.py, models/trajectory/{VisualTrajectory,Interpolator}.py and
models/Imu.py for everything the filter consumes; vectorised over time.  (SURVEY section 8f rank 1
moves this pre-pass to the GPU in a later round.)
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import probe as _probe
from . import rotations as rot

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "trajs.npz")


def load_trajectory(name_or_path: str, max_vals: Optional[int] = None, start_frame=None, with_start_index: bool = False):
    """Returns (t, xyz, q_xyzw) from a csv with header ``t,x,y,z,qx,qy,qz,qw`` (tools/files.py:19-28), a
    headerless space-separated txt, or one of the trajectories shipped in data/trajs.npz (by name)."""
    if os.path.exists(name_or_path):
        if name_or_path.endswith(".csv"):
            a = np.loadtxt(name_or_path, delimiter=",", skiprows=1)
        else:
            a = np.loadtxt(name_or_path)
    else:
        key = "traj_" + os.path.splitext(os.path.basename(name_or_path))[0]
        with np.load(_DATA) as z:
            if key not in z:
                raise FileNotFoundError(f"{name_or_path}: no such file and no packaged trajectory {key!r}")
            a = z[key]
    i0 = 0
    if start_frame:
        i0 = int(np.argwhere(a[:, 0] == start_frame).item())
        a = a[i0:]
    if max_vals:
        a = a[:max_vals]
    out = (a[:, 0].copy(), a[:, 1:4].copy(), a[:, 4:8].copy())
    return out + (i0,) if with_start_index else out


def load_notch(name_or_path: str, max_vals: Optional[int] = None, start_index: int = 0):
    """Notch trajectory [n,3] = angle, rate, acceleration per camera frame (config ``notch_traj_name``,
    VisualTrajectory.py:110-118).  The shipped ``notch90.csv`` is COMMA separated; the reference's loader splits on white
    space and appends to ``None`` (``with_notch: true`` cannot run at HEAD) -- read here the way the file
    is written, from a path or from the packaged data/trajs.npz."""
    if os.path.exists(name_or_path):
        with open(name_or_path, "r") as f:
            a = np.array([[float(v) for v in line.strip().split(",")[:3]] for line in f if line.strip()])
    else:
        key = "notch_" + os.path.splitext(os.path.basename(name_or_path))[0]
        with np.load(_DATA) as z:
            if key not in z:
                raise FileNotFoundError(f"{name_or_path}: no such file and no packaged notch trajectory {key!r}")
            a = z[key]
    a = a[start_index:]
    return (a[:max_vals] if max_vals else a).copy()


class Camera:
    """``Camera`` (Camera.py:41-118): t, p (3,n), raw + normalised quaternions, R, and the derived
    v / acc / om / alp obtained with np.gradient (Camera.py:158-170)."""

    def __init__(self, t, xyz, q_xyzw, scale=1.0, euler_mode="xyz", notch=None, _derived=None, interframe_vals=0,
                 _is_rotated=False):
        self.t = np.asarray(t, dtype=float)
        self.p = (np.asarray(xyz, dtype=float) * scale).T.copy() if _derived is None else np.asarray(xyz, dtype=float)
        self.q_raw = np.asarray(q_xyzw, dtype=float)
        self.quats = rot.normalise(self.q_raw)
        self.R = rot.to_matrix(self.quats)
        self.max_vals = len(self.t)
        self.dt = self.t[1] - self.t[0]
        self.min_t, self.max_t = self.t[0], self.t[-1]
        self.euler_mode = euler_mode
        self.interframe_vals = interframe_vals
        self.with_notch = notch is not None
        self.notch3 = np.zeros((self.max_vals, 3)) if notch is None else np.asarray(notch, dtype=float)
        if _derived is None:
            self.v = np.gradient(self.p, self.dt, axis=-1)
            self.acc = np.gradient(self.v, self.dt, axis=-1)
            ang = (rot.euler_xyz(self.quats) if euler_mode == "xyz" else rot.euler_zyx_reversed(self.quats)).T
            self.om = np.gradient(ang, self.dt, axis=-1)
            self.alp = np.gradient(self.om, self.dt, axis=-1)
        else:
            self.v, self.acc, self.om, self.alp = _derived
        # Camera.py:131-134: with a notch trajectory the filter's IMU source and error reference is the ROTATED camera
        self.is_rotated = False
        self.rotated: Optional["Camera"] = None
        if self.with_notch and _derived is None and not _is_rotated:
            self.rotated = self.gen_rotated()

    def gen_rotated(self) -> "Camera":
        """``Camera.gen_rotated`` (Camera.py:172-208): same positions, quaternions ``notch_quat * real_quat`` with
        ``notch_quat = Quaternion([0, 0, ang_notch], euler="xyz")``; the product is re-normalised with w >= 0
        (Quaternion.__mul__, quirk Q10).  The rotated camera re-derives om / alp from ITS Euler angles and carries the
        same notch arrays."""
        h = 0.5 * self.notch3[:, 0]
        nq = np.stack((np.zeros_like(h), np.zeros_like(h), np.sin(h), np.cos(h)), -1)
        r = Camera(self.t, self.p.T, rot.normalise(rot.multiply(nq, self.quats)), scale=1.0, euler_mode=self.euler_mode,
                   notch=self.notch3, _is_rotated=True)
        r.is_rotated = True
        return r

    @property
    def flag_interpolated(self):
        return self.interframe_vals > 0

    @property
    def r_deg(self):
        """VisualTraj._gen_euler_angles (VisualTrajectory.py:151-161): n x 3 Euler xyz in degrees."""
        return rot.euler_xyz(self.quats, degrees=True)

    def get_notch_vec_at(self, i):
        return self.notch3[i].copy()

    def interpolate(self, interframe_vals: int) -> "Camera":
        """Camera.interpolate -> Interpolator (Interpolator.py:25-88): np.linspace time base and np.interp on
        every channel, including the RAW quaternion components (re-normalised afterwards)."""
        n_new = (self.max_vals - 1) * interframe_vals + 1
        t_new = np.linspace(self.t[0], self.t[-1], num=n_new)

        def ip(rows):
            return np.stack([np.interp(t_new, self.t, r) for r in rows])

        return Camera(t_new, ip(self.p), ip(self.q_raw.T).T, euler_mode=self.euler_mode,
                      notch=ip(self.notch3.T).T if self.with_notch else None,
                      _derived=(ip(self.v), ip(self.acc), ip(self.om), ip(self.alp)), interframe_vals=interframe_vals)


def f_imu(p_C, R_WC, v_C, om_C, p_p, R_p, om_p):
    """eqns.f_imu (equations.py:8-13,54-60), vectorised: IMU reference position, rotation, velocity."""
    R_WB = R_WC @ np.swapaxes(R_p, -1, -2)
    Rp = np.einsum("...ij,...j->...i", R_WB, p_p)
    W_om = om_C - np.einsum("...ij,...j->...i", R_WB, om_p)
    return p_C - Rp, R_WB, v_C - np.cross(W_om, Rp)


def f_imu_meas(R_WC, om_C, acc_C, alp_C, p_p, R_p, om_p, alp_p):
    """eqns.f_imu_meas (equations.py:8-41,63-69), vectorised (probe v = acc = 0)."""
    R_WB = R_WC @ np.swapaxes(R_p, -1, -2)
    mv = lambda M, x: np.einsum("...ij,...j->...i", M, x)
    Rp, Rom = mv(R_WB, p_p), mv(R_WB, om_p)
    W_om = om_C - Rom
    W_omxp = np.cross(W_om, Rp)
    W_alp = alp_C - mv(R_WB, alp_p) - np.cross(W_om, Rom)
    W_acc = acc_C - np.cross(W_alp, Rp) - np.cross(W_om, W_omxp)
    R_BW = R_p @ np.swapaxes(R_WC, -1, -2)
    return mv(R_BW, W_om), mv(R_BW, W_acc)


@dataclass
class Streams:
    """Everything Filter.run consumes, as flat arrays (the layout of include/eskf.h)."""

    x0: np.ndarray  # [26]
    u0: np.ndarray  # [6]   first IMU sample (Filter.py:63,78-79)
    dt: np.ndarray  # [T]
    om_acc: np.ndarray  # [T,6]
    t_imu: np.ndarray  # [T]
    n_prop: np.ndarray  # [E] int32
    cam: np.ndarray  # [E,7]
    notch: np.ndarray  # [E]
    cam_ref: np.ndarray  # [E,6]
    imu_ref: np.ndarray  # [E,6]
    imu_ref_rows: np.ndarray  # [T,14] ImuRefTraj rows (ImuRefTraj.py:18-55)
    t_cam: np.ndarray  # [E+1]


def build_streams(cam: Camera, interframe_vals: int, length: float, angle: float, gt_dofs=_probe.GT_IMU_DOFS,
                  ic_dofs=None) -> Streams:
    """Host pre-pass of Simulator.__init__ / Filter.__init__ / Filter.propagate_imu: initial state
    (tools/utils.py:54-75), synthetic IMU samples at every interpolated instant (Imu.py:141-226) and the
    per-epoch membership decided by ``t_interp <= t_frame`` (Camera.py:299-301,320-347; quirk Q14)."""
    meas = cam  # Filter.run is handed the un-rotated camera: its frames are the measurements (Filter.py:144-185)
    if cam.rotated is not None:
        cam = cam.rotated  # IMU source (Imu.py:87-90), initial state (tools/utils.py:63-75), error reference (Filter.py:398)
    ci = cam.interpolate(interframe_vals)
    gt = np.asarray(gt_dofs, dtype=float)
    # ground-truth probe at every interpolated instant (notch joint from the notch trajectory)
    p_p, R_p, om_p, alp_p = _probe.fwkin(np.broadcast_to(gt, (len(ci.t), 6)), ci.notch3, length, angle)
    om, acc = f_imu_meas(ci.R, ci.om.T, ci.acc.T, ci.alp.T, p_p, R_p, om_p, alp_p)
    p_B, R_WB, v_B = f_imu(ci.p.T, ci.R, ci.v.T, ci.om.T, p_p, R_p, om_p)
    eul = rot.Rotation.from_matrix(R_WB).as_euler("xyz", degrees=True)
    qB = rot.from_matrix_markley(R_WB)
    ref_rows = np.hstack((ci.t[:, None], p_B, v_B, eul, qB[:, [3, 0, 1, 2]]))
    # initial state from the UN-interpolated camera frame 0
    p0, R0, om0, _ = _probe.fwkin(gt, cam.notch3[0], length, angle)
    pB0, RWB0, vB0 = f_imu(cam.p[:, 0], cam.R[0], cam.v[:, 0], cam.om[:, 0], p0, R0, om0)
    x0 = np.hstack((pB0, vB0, rot.from_matrix_markley(RWB0), gt if ic_dofs is None else np.asarray(ic_dofs, dtype=float),
                    cam.notch3[0], cam.p[:, 0], cam.quats[0]))
    # epoch membership: index of the last interpolated sample with t <= frame time
    idx = np.searchsorted(ci.t, cam.t, side="right") - 1
    n_prop = np.diff(idx).astype(np.int32)
    first, last = idx[0] + 1, idx[-1]
    sel = slice(first, last + 1)
    t_prev = np.concatenate(([cam.t[0]], ci.t[first:last]))
    # Filter.propagate_imu restarts old_ti at the FRAME time t0 of each epoch (Filter.py:195)
    starts = np.cumsum(np.concatenate(([0], n_prop[:-1])))
    t_prev[starts] = cam.t[:-1]
    dt = ci.t[sel] - t_prev
    ends = idx[1:]
    cam_ref = np.hstack((cam.p.T[1:], cam.r_deg[1:]))
    return Streams(x0=x0, u0=np.hstack((om[0], acc[0])), dt=dt, om_acc=np.hstack((om[sel], acc[sel])), t_imu=ci.t[sel],
                   n_prop=n_prop, cam=np.hstack((meas.p.T[1:], meas.q_raw[1:])), notch=meas.notch3[1:, 0].copy(),
                   cam_ref=cam_ref, imu_ref=ref_rows[ends][:, 4:10], imu_ref_rows=ref_rows[sel], t_cam=cam.t.copy())
