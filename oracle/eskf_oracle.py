"""numpy restatement of dvi-ekf's VI-ESKF hot path (single filter, FP64).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites
the reference file:line it follows (paths relative to the reference repo
root).  The arithmetic is deliberately *generic* (numeric DH chain, geometric
Jacobian, dense 24x24 products, LAPACK inverse) so that it is independent of
the closed forms hard-wired into the CUDA kernels.

Parity: pinned against ``data/trajs/kf_best_mandala0_mono.txt`` and
``imu_ref_mandala0_mono*.txt`` (tests/test_oracle_golden.py) with the
``legacy_golden`` preset; HEAD behaviour differs from that preset by two
documented one-line toggles (``zero_frozen_dofs``, ``euler_mode``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy.spatial.transform import Rotation

# --------------------------------------------------------------------------
# small helpers


def skew(x):
    """dvi_ekf/tools/math.py:29-39."""
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])


def quat_normalise(q):
    """``Quaternion.normalise`` (dvi_ekf/tools/Quaternion.py:195-206):
    divide by the norm, then force the scalar part to be non-negative."""
    q = np.asarray(q, dtype=float)
    d = math.sqrt(q[0] ** 2 + q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    q = q / d
    if q[3] < 0:
        q = -q
    return q


def quat_from_matrix_scipy110(M):
    """``Rotation.from_matrix`` as shipped in scipy 1.10.1 (the version the
    reference pins, requirements.txt:8): Markley's method applied to the RAW
    matrix -- no SVD orthogonalisation -- then a plain normalisation.  The
    reference feeds it the first-order (non-orthonormal) R + R[dt w]x every
    step (dvi_ekf/filter/state.py:67,72 -> tools/Quaternion.py:69), so the
    difference to newer scipy is O(|dt w|^2) per call, not rounding."""
    M = np.asarray(M, dtype=float)
    dec = [M[0, 0], M[1, 1], M[2, 2], M[0, 0] + M[1, 1] + M[2, 2]]
    choice = 0
    for n in range(1, 4):  # first maximum wins, like the cython _argmax4
        if dec[n] > dec[choice]:
            choice = n
    q = np.empty(4)
    if choice != 3:
        i = choice
        j = (i + 1) % 3
        k = (j + 1) % 3
        q[i] = 1 - dec[3] + 2 * M[i, i]
        q[j] = M[j, i] + M[i, j]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
    else:
        q[0] = M[2, 1] - M[1, 2]
        q[1] = M[0, 2] - M[2, 0]
        q[2] = M[1, 0] - M[0, 1]
        q[3] = 1 + dec[3]
    return q / math.sqrt(q @ q)


def quat_from_matrix_svd(M):
    """Newer scipy behaviour (orthogonalise first); only used by tests that
    demonstrate quirk Q1 is load-bearing."""
    return Rotation.from_matrix(np.asarray(M, dtype=float)).as_quat()


def quat_to_matrix(q):
    """``Quaternion.rot`` = ``Rotation.from_quat(xyzw).as_matrix()``
    (tools/Quaternion.py:97-107).  from_quat re-normalises its input."""
    q = np.asarray(q, dtype=float)
    x, y, z, w = q / math.sqrt(q @ q)
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array(
        [
            [x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)],
            [2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)],
            [2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2],
        ]
    )


def quat_mul(a, b):
    """``Quaternion.__mul__`` (tools/Quaternion.py:170-183): Hamilton product,
    result re-normalised with w >= 0 (quirk Q10).  xyzw storage."""
    aw, av = a[3], np.asarray(a[:3], dtype=float)
    bw, bv = b[3], np.asarray(b[:3], dtype=float)
    w = aw * bw - av @ bv
    v = aw * bv + bw * av + np.cross(av, bv)
    return quat_normalise(np.array([v[0], v[1], v[2], w]))


def quat_conj(q):
    """tools/Quaternion.py:133-135 (no normalisation)."""
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_about_axis(angle, axis):
    """``Quaternion.about_axis`` (tools/Quaternion.py:208-224)."""
    q = np.array([0.0, axis[0], axis[1], axis[2]])
    qlen = np.linalg.norm(q)
    eps4 = np.finfo(float).eps * 4.0
    if qlen > eps4:
        q *= math.sin(angle / 2.0) / qlen
    q[0] = math.cos(angle / 2.0)
    return quat_normalise(np.array([q[1], q[2], q[3], q[0]]))


def quat_from_euler_xyz(e):
    """``Quaternion(val=e, euler="xyz")`` (tools/Quaternion.py:72-73);
    extrinsic xyz, no extra normalisation / sign fix (do_normalise=False)."""
    return Rotation.from_euler("xyz", np.asarray(e, dtype=float)).as_quat()


def quat_angle_axis(q):
    """``Quaternion.angle`` / ``.axis`` (tools/Quaternion.py:137-167), quirk
    Q9: angle = asin(|v|) and axis = 0 iff math.isclose(angle, 0)."""
    v = np.asarray(q[:3], dtype=float)
    nv = np.linalg.norm(v)
    ang = math.asin(nv)
    ax = np.zeros(3) if math.isclose(ang, 0) else v / nv
    return ang, ax


def euler_xyz_deg(q):
    """``Quaternion.euler_xyz_deg`` (tools/Quaternion.py:121-123)."""
    return Rotation.from_quat(q).as_euler("xyz", degrees=True)


# --------------------------------------------------------------------------
# probe: 8-joint DH chain (dvi_ekf/models/Probe.py:135-167), numeric


def _rz(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])


def _rx(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1.0, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]])


def _tz(d):
    m = np.eye(4)
    m[2, 3] = d
    return m


@dataclass
class FwKin:
    """p, R, v, om, acc, alp of the camera w.r.t. the IMU, in IMU coordinates
    (the tuple returned by ``Probe.get_sym`` Probe.py:289-306) plus the
    geometric Jacobian columns the error Jacobians need."""

    p: np.ndarray
    R: np.ndarray
    v: np.ndarray
    om: np.ndarray
    acc: np.ndarray
    alp: np.ndarray
    Jv: np.ndarray  # 3x8, d p / d q_k
    Jw: np.ndarray  # 3x8, angular part (d R / d q_k = [Jw_k]x R)


class Probe:
    """Standard-DH serial chain  T = Ry(-180deg) * A_1 ... A_8,
    A_i = Rz(theta_i) Tz(d_i) Tx(0) Rx(alpha_i)  (roboticstoolbox
    ``DHRobot.fkine`` with ``base=SE3.Ry(-180,'deg')``, Probe.py:14,140).

    Joints (Probe.py:147-167): three revolute (IMU orientation), three
    prismatic (IMU translation), revolute notch joint with d = scope length
    and alpha = camera angle, and a final revolute joint fixed to 0."""

    def __init__(self, length: float, angle: float):
        self.length = float(length)
        self.angle = float(angle)
        hp = math.pi / 2
        # (is_revolute, theta offset / fixed theta, d, alpha)
        self.links = [
            (True, hp, 0.0, hp),
            (True, -hp, 0.0, -hp),
            (True, 0.0, 0.0, 0.0),
            (False, 0.0, 0.0, hp),
            (False, hp, 0.0, hp),
            (False, -hp, 0.0, hp),
            (True, 0.0, self.length, self.angle),
            (True, 0.0, 0.0, 0.0),
        ]
        c, s = math.cos(-math.pi), math.sin(-math.pi)
        self.base = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1.0]])

    def fwkin(self, q: Sequence[float], qd: Sequence[float], qdd: Sequence[float]) -> FwKin:
        """fkine + jacob0 + the reference's own Hessian contraction
        (Probe.py:186-259)."""
        q = np.asarray(q, dtype=float)
        qd = np.asarray(qd, dtype=float)
        qdd = np.asarray(qdd, dtype=float)
        n = 8
        T = self.base.copy()
        frames = []
        for i, (rev, th, d, al) in enumerate(self.links):
            frames.append(T.copy())
            if rev:
                A = _rz(q[i] + th) @ _tz(d) @ _rx(al)
            else:
                A = _rz(th) @ _tz(q[i]) @ _rx(al)
            T = T @ A
        pe = T[:3, 3]
        J = np.zeros((6, n))
        for i, (rev, _, _, _) in enumerate(self.links):
            z = frames[i][:3, 2]
            o = frames[i][:3, 3]
            if rev:
                J[:3, i] = np.cross(z, pe - o)
                J[3:, i] = z
            else:
                J[:3, i] = z
        # Probe.hessian_symbolic (Probe.py:219-231)
        H = np.zeros((6, n, n))
        for j in range(n):
            for i in range(j, n):
                H[:3, i, j] = np.cross(J[3:, j], J[:3, i])
                H[3:, i, j] = np.cross(J[3:, j], J[3:, i])
                if i != j:
                    H[:3, j, i] = H[:3, i, j]
        vel = J @ qd
        acc = np.einsum("aij,j,i->a", H, qd, qd) + J @ qdd
        return FwKin(
            p=pe.copy(),
            R=T[:3, :3].copy(),
            v=vel[:3],
            om=vel[3:],
            acc=acc[:3],
            alp=acc[3:],
            Jv=J[:3].copy(),
            Jw=J[3:].copy(),
        )

    def est_fwkin(self, dofs, notch_dofs) -> FwKin:
        """``SymProbe.get_est_fwkin`` (Probe.py:470-480): q = [dofs, notch, 0];
        only the notch joint has a rate / acceleration (Probe.py:401-410,441)."""
        q = [*dofs, notch_dofs[0], 0.0]
        qd = [0.0] * 6 + [notch_dofs[1], 0.0]
        qdd = [0.0] * 6 + [notch_dofs[2], 0.0]
        return self.fwkin(q, qd, qdd)


GT_IMU_DOFS = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 20.0])  # SimpleProbe, Probe.py:385-388


# --------------------------------------------------------------------------
# IMU <- camera kinematics (dvi_ekf/kinematics/equations.py:8-41)


def f_imu(p_C, R_WC, v_C, om_C, fw: FwKin):
    """``eqns.f_imu`` (equations.py:54-60): IMU reference pose / velocity."""
    R_WB = R_WC @ fw.R.T
    W_p = p_C - R_WB @ fw.p
    W_om = om_C - R_WB @ fw.om
    W_omxp = np.cross(W_om, R_WB @ fw.p)
    W_v = v_C - R_WB @ fw.v - W_omxp
    return W_p, R_WB, W_v


def f_imu_meas(R_WC, om_C, acc_C, alp_C, fw: FwKin):
    """``eqns.f_imu_meas`` (equations.py:8-41,63-69): synthetic gyro / accel
    in the IMU frame.  No gravity term anywhere."""
    R_WB = R_WC @ fw.R.T
    W_om = om_C - R_WB @ fw.om
    W_omxp = np.cross(W_om, R_WB @ fw.p)
    W_alp = alp_C - R_WB @ fw.alp - np.cross(W_om, R_WB @ fw.om)
    W_acc = (
        acc_C
        - R_WB @ fw.acc
        - 2 * np.cross(W_om, R_WB @ fw.v)
        - np.cross(W_alp, R_WB @ fw.p)
        - np.cross(W_om, W_omxp)
    )
    R_BW = fw.R @ R_WC.T
    return R_BW @ W_om, R_BW @ W_acc


# --------------------------------------------------------------------------
# configuration (config.yaml + dvi_ekf/tools/config.py), restated without pydantic


@dataclass
class OracleConfig:
    # simulation / camera
    max_vals: Optional[int] = 10  # do_fast_sim => 10 (config.py:246-248)
    interframe_vals: int = 1  # do_fast_sim => 1
    scale: float = 10.0
    frozen_dofs: Sequence[int] = (1, 1, 1, 1, 1, 1)
    # model
    length: float = 50.0
    angle: float = math.radians(30.0)
    # imu (config.py:96-120)
    noise_sample_rate: float = 10.0
    gravity: float = 981.0
    # filter noise (config.yaml:20-23, 76-82)
    meas_pos_std: Sequence[float] = (0.02, 0.002, 0.02)
    meas_theta_std_deg: Sequence[float] = (1e-5, 1e-5, 1e-5)
    meas_notch_std_deg: float = 0.01
    rw_trans: Sequence[float] = (0.25, 0.25, 0.25)
    rw_rot_deg: Sequence[float] = (1.0, 1.0, 10.0)
    rw_notch_acc_deg: float = 0.05
    # cov0 (config.yaml:38-56)
    cov0_imu_pos: Sequence[float] = (0.02, 0.002, 0.002)
    cov0_imu_vel: Sequence[float] = (0.1, 0.1, 0.1)
    cov0_imu_theta_deg: Sequence[float] = (1.0, 1.0, 1.0)
    cov0_dofs_rot_deg: Sequence[float] = (5.0, 5.0, 5.0)
    cov0_dofs_trans: Sequence[float] = (10.0, 10.0, 10.0)
    cov0_notch_deg: Sequence[float] = (0.2, 0.02, 0.02)
    cov0_cam_pos: Sequence[float] = (0.02, 0.002, 0.002)
    cov0_cam_theta_deg: Sequence[float] = (0.2, 0.2, 0.2)
    # ---- quirk toggles (HEAD defaults) ----
    markley: bool = True  # Q1: scipy-1.10.1 from_matrix
    zero_frozen_dofs: bool = True  # Q7: Filter.py:243-245 zeroes the DOF itself
    euler_mode: str = "xyz"  # Q11: "xyz" (HEAD Camera.py:168) | "zyx_legacy"
    fix_q2: bool = False  # v_tr = p_tr bug (Probe.py:458) kept unless True
    fix_q3: bool = False  # column mis-alignment (symbols.py:107) kept unless True
    fix_q4: bool = False  # dqc axis = theta bug (state.py:124) kept unless True

    @staticmethod
    def legacy_golden(**kw) -> "OracleConfig":
        """Preset that reproduces data/trajs/kf_best_mandala0_mono.txt."""
        return OracleConfig(zero_frozen_dofs=False, euler_mode="zyx_legacy", **kw)

    # derived vectors -----------------------------------------------------
    @property
    def stdev_accel(self):  # config.py:115-120 (cm/s^2)
        return np.array([400e-6 * self.gravity * math.sqrt(self.noise_sample_rate)] * 3)

    @property
    def stdev_omega(self):  # config.py:105-113 (rad/s)
        return np.array([np.deg2rad(0.005 * math.sqrt(self.noise_sample_rate))] * 3)

    @property
    def process_noise_rw_std(self):  # config.py:196-198,258
        v = np.hstack((np.deg2rad(self.rw_rot_deg), self.rw_trans, np.deg2rad(self.rw_notch_acc_deg)))
        return v / self.interframe_vals

    @property
    def meas_noise_std(self):  # config.py:75-77,260
        return np.hstack((self.meas_pos_std, np.deg2rad(self.meas_theta_std_deg), np.deg2rad(self.meas_notch_std_deg)))

    @property
    def cov0_matrix(self):  # config.py:154-180
        vec = np.hstack(
            (
                self.cov0_imu_pos,
                self.cov0_imu_vel,
                np.deg2rad(self.cov0_imu_theta_deg),
                np.deg2rad(self.cov0_dofs_rot_deg),
                self.cov0_dofs_trans,
                np.deg2rad(self.cov0_notch_deg),
                self.cov0_cam_pos,
                np.deg2rad(self.cov0_cam_theta_deg),
            )
        )
        return np.square(np.diag(vec))


# --------------------------------------------------------------------------
# camera (dvi_ekf/models/Camera.py, models/trajectory/*.py)


@dataclass
class CameraData:
    """What ``Camera`` exposes (Camera.py:84-118): t, p (3xn), raw and
    normalised quaternions, R list, and derived v / acc / om / alp (3xn)."""

    t: np.ndarray
    p: np.ndarray
    q_raw: np.ndarray  # n x 4 xyzw as read (VisualTraj.at_index uses these)
    quats: np.ndarray  # n x 4 xyzw normalised, w >= 0 (VisualTrajectory.py:144-149)
    R: np.ndarray  # n x 3 x 3
    v: np.ndarray
    acc: np.ndarray
    om: np.ndarray
    alp: np.ndarray
    notch: np.ndarray  # n x 3 (notch, notch_d, notch_dd); zeros without notch
    interframe_vals: int = 0


def _euler_for_gradient(quats, mode):
    if mode == "xyz":  # HEAD: Camera.py:168
        return np.array([Rotation.from_quat(q).as_euler("xyz") for q in quats]).T
    if mode == "zyx_legacy":  # older revision that produced the golden files (Q11)
        return np.array([Rotation.from_quat(q).as_euler("zyx")[::-1] for q in quats]).T
    raise ValueError(mode)


def camera_from_arrays(t, xyz, q_xyzw, cfg: OracleConfig, notch=None) -> CameraData:
    """``Camera.__init__`` for a raw trajectory (Camera.py:58-118,158-170):
    scale positions (VisualTrajectory.py:99-108), normalise quaternions,
    np.gradient for the derived data."""
    t = np.asarray(t, dtype=float)
    p = (np.asarray(xyz, dtype=float) * cfg.scale).T.copy()  # 3 x n
    q_raw = np.asarray(q_xyzw, dtype=float)
    quats = np.array([quat_normalise(q) for q in q_raw])
    R = np.array([quat_to_matrix(q) for q in quats])
    dt = t[1] - t[0]
    v = np.gradient(p, dt, axis=-1)
    acc = np.gradient(v, dt, axis=-1)
    ang = _euler_for_gradient(quats, cfg.euler_mode)
    om = np.gradient(ang, dt, axis=-1)
    alp = np.gradient(om, dt, axis=-1)
    n = len(t)
    nt = np.zeros((n, 3)) if notch is None else np.asarray(notch, dtype=float)
    return CameraData(t, p, q_raw, quats, R, v, acc, om, alp, nt)


def camera_interpolate(cam: CameraData, interframe_vals: int) -> CameraData:
    """``Camera.interpolate`` -> ``Interpolator`` -> ``CameraInterpolated``
    (Camera.py:135-141, Interpolator.py:25-88): np.linspace time base,
    np.interp on positions, RAW-normalised quaternion components (then
    re-normalised, VisualTrajectory.py:136-149) and on v / acc / om / alp."""
    t_old = cam.t
    n_new = (len(t_old) - 1) * interframe_vals + 1
    t_new = np.linspace(t_old[0], t_old[-1], num=n_new)
    ip = lambda y: np.interp(t_new, t_old, y)
    p = np.array([ip(cam.p[i]) for i in range(3)])
    # the interpolator reads traj.qx.. which are the RAW file columns
    q_lin = np.array([ip(cam.q_raw[:, i]) for i in range(4)]).T
    quats = np.array([quat_normalise(q) for q in q_lin])
    R = np.array([quat_to_matrix(q) for q in quats])
    v = np.array([ip(cam.v[i]) for i in range(3)])
    acc = np.array([ip(cam.acc[i]) for i in range(3)])
    om = np.array([ip(cam.om[i]) for i in range(3)])
    alp = np.array([ip(cam.alp[i]) for i in range(3)])
    notch = np.array([ip(cam.notch[:, i]) for i in range(3)]).T
    return CameraData(t_new, p, q_lin, quats, R, v, acc, om, alp, notch, interframe_vals)


def load_notch_csv(path, max_vals=None, start_index=0):
    """``VisualTraj.read_notch_from_file`` (VisualTrajectory.py:110-118) with the Q12 defect repaired the only way
    the shipped file allows: ``data/trajs/notch90.csv`` holds ``notch,notch_d,notch_dd`` per line, COMMA separated
    (HEAD splits on white space and appends to ``None``).  Rows are cut like the camera rows
    (``_filter_values``, VisualTrajectory.py:69)."""
    rows = []
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(v) for v in line.split(",")[:3]])
    a = np.array(rows)[start_index:]
    return a[:max_vals] if max_vals else a


def camera_gen_rotated(cam: CameraData, cfg: OracleConfig) -> CameraData:
    """``Camera.gen_rotated`` (Camera.py:172-208): same positions, every quaternion pre-multiplied by the notch
    rotation ``Quaternion([0, 0, ang_notch], euler="xyz") * real_quat`` (product re-normalised, w >= 0: Q10); the
    rotated trajectory becomes a new ``Camera`` whose derived data are the gradients of ITS Euler angles, and it
    carries the same notch arrays (Camera.py:204-206)."""
    q_rot = np.array([quat_mul(quat_from_euler_xyz([0.0, 0.0, cam.notch[i, 0]]), cam.quats[i]) for i in range(len(cam.t))])
    R = np.array([quat_to_matrix(q) for q in q_rot])
    dt = cam.t[1] - cam.t[0]
    ang = _euler_for_gradient(q_rot, cfg.euler_mode)
    om = np.gradient(ang, dt, axis=-1)
    alp = np.gradient(om, dt, axis=-1)
    return CameraData(cam.t, cam.p, q_rot.copy(), q_rot, R, cam.v, cam.acc, om, alp, cam.notch)


def _timestamp_index(ts, max_t):
    """Camera.py:299-301."""
    return max(i for i, t in enumerate(ts) if t <= max_t)


# --------------------------------------------------------------------------
# filter state


@dataclass
class State:
    """dvi_ekf/filter/state.py:11-29; quaternions xyzw."""

    p: np.ndarray
    v: np.ndarray
    q: np.ndarray
    dofs: np.ndarray
    notch_dofs: np.ndarray
    p_cam: np.ndarray
    q_cam: np.ndarray

    def copy(self):
        return State(*(np.array(a, dtype=float).copy() for a in self.vec_list()))

    def vec_list(self):
        return [self.p, self.v, self.q, self.dofs, self.notch_dofs, self.p_cam, self.q_cam]

    def as_vector(self):
        """26-vector in the engine's layout: p v q(xyzw) dofs notch p_cam q_cam(xyzw)."""
        return np.hstack(self.vec_list())

    @staticmethod
    def from_vector(x):
        x = np.asarray(x, dtype=float)
        return State(x[0:3].copy(), x[3:6].copy(), x[6:10].copy(), x[10:16].copy(), x[16:19].copy(), x[19:22].copy(), x[22:26].copy())


def filter_traj_row(t, s: State):
    """``_get_euler_measurement_array`` (FilterTraj.py:12-32): 31 numbers."""
    wxyz = lambda q: [q[3], q[0], q[1], q[2]]
    return np.array(
        [
            t,
            *s.p,
            *s.v,
            *euler_xyz_deg(s.q),
            *wxyz(s.q),
            *np.rad2deg(s.dofs[:3]),
            *s.dofs[3:],
            *s.p_cam,
            *euler_xyz_deg(s.q_cam),
            *wxyz(s.q_cam),
        ]
    )


class OracleFilter:
    """``Filter`` (dvi_ekf/filter/Filter.py:28-395), single instance."""

    NUM_ERR = 24
    NUM_NOISE = 13
    NUM_MEAS = 7

    def __init__(self, cfg: OracleConfig, x0: State, P0, om0, acc0, probe: Optional[Probe] = None):
        self.cfg = cfg
        self.probe = probe or Probe(cfg.length, cfg.angle)
        self.x = x0.copy()
        self.P = np.array(P0, dtype=float).copy()
        self.dt = 0.0  # Filter.py:40
        self.frozen = [bool(f) for f in cfg.frozen_dofs]
        self.stdev_na = np.array(cfg.stdev_accel)
        self.stdev_nom = np.array(cfg.stdev_omega)
        self.rw_std = np.array(cfg.process_noise_rw_std)
        self.update_noise_matrices()  # with dt = 0 => Q[0:6] = 0 (quirk Q5, Filter.py:68-72)
        self.om_old = np.array(om0, dtype=float).copy()
        self.acc_old = np.array(acc0, dtype=float).copy()
        self.R_WB_old = quat_to_matrix(self.x.q)  # Filter.py:80
        self.H = np.zeros((7, 24))
        self.H[0:6, 18:24] = np.eye(6)
        self.H[6, 15] = 1
        self.Fx = None
        self.Fi = None
        self.status = 0

    # Filter.py:110-117
    def update_noise_matrices(self):
        Q = np.eye(13)
        Q[0:3, 0:3] = self.dt ** 2 * self.stdev_na ** 2 * np.eye(3)
        Q[3:6, 3:6] = self.dt ** 2 * self.stdev_nom ** 2 * np.eye(3)
        Q[6:13, 6:13] = np.diag(np.square(self.rw_std))
        self.Q = Q
        self.R = np.diag(np.square(self.cfg.meas_noise_std))

    def _mat2quat(self, M):
        q = quat_from_matrix_scipy110(M) if self.cfg.markley else quat_from_matrix_svd(M)
        return quat_normalise(q)  # Quaternion(val=M, do_normalise=True)

    # Filter.py:219-230
    def propagate(self, dt, om, acc):
        self.dt = float(dt)
        om = np.asarray(om, dtype=float)
        acc = np.asarray(acc, dtype=float)
        self._predict_nominal(om, acc)
        self._predict_error()
        self.P = self.Fx @ self.P @ self.Fx.T + self.Fi @ self.Q @ self.Fi.T  # Filter.py:349
        self.om_old = om.copy()
        self.acc_old = acc.copy()
        self.R_WB_old = quat_to_matrix(self.x.q)

    # Filter.py:232-247 + equations.py:44-50,72-100 + state.py:62-74
    def _predict_nominal(self, om, acc):
        x, dt = self.x, self.dt
        fw = self.probe.est_fwkin(x.dofs, x.notch_dofs)
        R_WB = quat_to_matrix(x.q)
        R_WC = quat_to_matrix(x.q_cam)
        om_avg = (self.om_old + om) / 2
        R_next = R_WB + R_WB @ skew(dt * om_avg)
        acc_avg = (R_WB @ self.acc_old + R_next @ acc) / 2
        p = x.p + dt * x.v + (dt ** 2 / 2) * acc_avg
        v = x.v + dt * acc_avg
        dofs = x.dofs.copy()
        notch = np.array(
            [
                x.notch_dofs[0] + dt * x.notch_dofs[1],
                x.notch_dofs[1] + dt * x.notch_dofs[2],
                x.notch_dofs[2],
            ]
        )
        p_cam = x.p_cam + dt * x.v + dt * R_WB @ (fw.v + np.cross(om_avg, fw.p))
        om_c = fw.R.T @ (self.om_old + fw.om)
        R_WC_next = R_WC + R_WC @ skew(dt * om_c)
        if self.cfg.zero_frozen_dofs:
            for i, fr in enumerate(self.frozen):
                if fr:
                    dofs[i] = 0.0
        self.x = State(p, v, self._mat2quat(R_next), dofs, notch, p_cam, self._mat2quat(R_WC_next))

    # Filter.py:249-342 + symbols.py:134-200 (closed forms of the casadi AD)
    def _predict_error(self):
        dt = self.dt
        x = self.x  # POST-predict dofs / notch (Filter.py:325-326)
        R_old = self.R_WB_old
        Fx = np.eye(24)
        Fx[0:3, 3:6] = dt * np.eye(3)
        Fx[3:6, 6:9] = -R_old @ skew(self.acc_old) * dt
        Om = quat_normalise(np.array([*(0.5 * dt * self.om_old), 1.0]))  # Filter.py:134
        Fx[6:9, 6:9] = quat_to_matrix(Om).T
        Fx[15, 16] += dt
        Fx[16, 17] += dt

        fw = self.probe.est_fwkin(x.dofs, x.notch_dofs)
        p, Rp = fw.p, fw.R
        om_tr = self.om_old - self.stdev_nom  # noise symbols evaluated at sigma (Q6, Filter.py:330-331)
        # Q2: SymProbe.v_tr = _get_tr(p)  (Probe.py:458)
        v_tr = np.zeros(3) if self.cfg.fix_q2 else p
        w = v_tr + np.cross(om_tr, p)
        dp_dq = fw.Jv[:, 0:7]  # d p / d q1..q7
        if self.cfg.fix_q2:
            dw_dq = skew(om_tr) @ dp_dq
        else:
            dw_dq = (np.eye(3) + skew(om_tr)) @ dp_dq
        a = Rp.T @ (om_tr + fw.om)
        b = Rp.T @ (self.om_old + fw.om)
        # d/dq_k [ R(q)^T (om_tr + om_p(q)) ]; dR/dq_k = [Jw_k]x R, d om_p/dq_k = Jw_k x om_p
        # for the joints before the notch joint (om_p = z6 * qd7)
        da_dq = np.zeros((3, 7))
        for k in range(7):
            jw = fw.Jw[:, k]
            d_om_p = np.cross(jw, fw.om) if k < 6 else np.zeros(3)
            da_dq[:, k] = -Rp.T @ np.cross(jw, om_tr + fw.om) + Rp.T @ d_om_p

        # 6 x 22 block w.r.t. err_x = [err_p_B, err_v_B, err_theta, err_dofs(6), err_notch(1), err_p_C, err_theta_C]
        J = np.zeros((6, 22))
        J[0:3, 3:6] = dt * np.eye(3)
        J[0:3, 6:9] = -dt * R_old @ skew(w)
        J[0:3, 9:16] = dt * R_old @ dw_dq
        J[0:3, 16:19] = np.eye(3)
        J[3:6, 9:16] = dt * da_dq
        J[3:6, 19:22] = np.eye(3) - 0.5 * dt * skew(a + b)
        if self.cfg.fix_q3:
            # what the author presumably intended: err_notchdofs (3 wide) in err_x
            Fx[18:24, :] = 0.0
            Fx[18:24, 0:16] = J[:, 0:16]
            Fx[18:24, 18:24] = J[:, 16:22]
        else:
            Fx[18:24, 0:22] = J  # Filter.py:279-285 (running index 2 short from column 16 on, Q3)
        self.Fx = Fx

        Fi = np.zeros((24, 13))
        Fi[3:15, 0:12] = np.eye(12)
        Fi[17, 12] = 1
        Jn = np.zeros((6, 13))
        Jn[0:3, 3:6] = dt * R_old @ skew(p)
        Jn[3:6, 3:6] = -dt * Rp.T
        Fi[18:24, :] = Jn
        self.Fi = Fi

    # Filter.py:351-395 + state.py:46-60,105-129
    def update(self, cam_pos, cam_q, ang_notch):
        H, P = self.H, self.P
        S = H @ P @ H.T + self.R
        try:
            K = P @ H.T @ np.linalg.inv(S)
        except np.linalg.LinAlgError:
            self.status |= 1
            return None
        if not np.all(np.isfinite(K)):
            # numpy raises only for exactly singular S; the engine flags
            # non-finite gains as well and skips the update the same way.
            self.status |= 1
            return None
        x = self.x
        notch_quat = quat_from_euler_xyz([0.0, 0.0, ang_notch])
        cam_rot_corrected = quat_mul(notch_quat, cam_q)
        res_p = np.asarray(cam_pos, dtype=float) - x.p_cam
        err_q = quat_mul(quat_conj(cam_rot_corrected), x.q_cam)
        ang, ax = quat_angle_axis(err_q)
        res_q = ang * ax
        res_notch = ang_notch - x.notch_dofs[0]
        res = np.hstack((res_p, res_q, res_notch))
        d = K @ res
        theta = d[6:9].copy()
        theta_c = d[21:24].copy()
        dq = quat_about_axis(np.linalg.norm(theta), theta)
        # Q4: axis = theta (IMU), state.py:124
        dqc = quat_about_axis(np.linalg.norm(theta_c), theta_c if self.cfg.fix_q4 else theta)
        ddofs = d[9:15].copy()
        for i, fr in enumerate(self.frozen):
            if fr:
                ddofs[i] = 0.0
        self.x = State(
            x.p + d[0:3],
            x.v + d[3:6],
            quat_normalise(quat_mul(x.q, dq)),
            x.dofs + ddofs,
            x.notch_dofs + d[15:18],
            x.p_cam + d[18:21],
            quat_normalise(quat_mul(x.q_cam, dqc)),
        )
        I = np.eye(24)
        self.P = (I - K @ H) @ P @ (I - K @ H).T + K @ self.R @ K.T
        G = np.eye(24)
        G[6:9, 6:9] = np.eye(3) - skew(0.5 * theta)
        G[21:24, 21:24] = np.eye(3) - skew(0.5 * theta_c)
        self.P = G @ self.P @ G.T
        return K

    # engine-layout accessors ------------------------------------------------
    def get_vectors(self):
        return (
            self.x.as_vector(),
            self.P.copy(),
            np.hstack((self.om_old, self.acc_old)),
            self.R_WB_old.reshape(9).copy(),
        )


# --------------------------------------------------------------------------
# the main.py flow (Simulator.__init__ + Filter.run), restated


@dataclass
class RunResult:
    kf_rows: np.ndarray  # (n_frames, 31)  FilterTraj rows at update instants (+ IC row)
    imu_ref_rows: np.ndarray  # (n_imu, 14)  ImuRefTraj rows
    dt: np.ndarray  # (T,)
    om_acc: np.ndarray  # (T, 6)
    n_prop: np.ndarray  # (E,)
    cam_meas: np.ndarray  # (E, 7) scaled pos + raw quaternion (xyzw)
    notch_meas: np.ndarray  # (E,)
    x0: np.ndarray
    P0: np.ndarray
    u0: np.ndarray
    Rold0: np.ndarray
    x_steps: List[np.ndarray] = field(default_factory=list)  # state after every propagate
    x_final: Optional[np.ndarray] = None
    P_final: Optional[np.ndarray] = None
    update_mse: Optional[np.ndarray] = None  # per epoch
    dof_metric: Optional[float] = None


def imu_ref_row(t, p_C, R_WC, v_C, om_C, fw_gt):
    """``ImuRefTraj.append_value`` (ImuRefTraj.py:40-55)."""
    p, R_WB, v = f_imu(p_C, R_WC, v_C, om_C, fw_gt)
    eul = Rotation.from_matrix(R_WB).as_euler("xyz", degrees=True)
    q = quat_normalise(quat_from_matrix_scipy110(R_WB))
    return np.array([t, *p, *v, *eul, q[3], q[0], q[1], q[2]])


def build_streams(cam: CameraData, cfg: OracleConfig, probe: Probe, rotated: Optional[CameraData] = None):
    """Everything ``Simulator.__init__`` + ``Filter.__init__`` +
    ``Filter.propagate_imu`` derive from the camera before touching the
    filter state: initial state (tools/utils.py:54-75), first IMU sample
    (Imu.py:198-226), per-step dt / IMU samples (Filter.py:187-217) and the
    per-epoch sample counts decided by float comparison (Camera.py:320-347).
    With ``with_notch`` the IMU is synthesised from the ROTATED camera (``Imu.create``, Imu.py:87-90) and the initial
    state comes from it too (tools/utils.py:63-75); the measurements handed to ``Filter.update`` stay those of the
    un-rotated camera (``Filter.run`` is given ``camera``, Filter.py:144-185)."""
    meas = cam
    if rotated is not None:
        cam = rotated
    cam_i = camera_interpolate(cam, cfg.interframe_vals)
    notch0 = cam.notch[0]
    fw0 = probe.fwkin([*GT_IMU_DOFS, notch0[0], 0.0], [0.0] * 6 + [notch0[1], 0.0], [0.0] * 6 + [notch0[2], 0.0])
    p_B0, R_WB0, v_B0 = f_imu(cam.p[:, 0], cam.R[0], cam.v[:, 0], cam.om[:, 0], fw0)
    q0 = quat_normalise(quat_from_matrix_scipy110(R_WB0) if cfg.markley else quat_from_matrix_svd(R_WB0))
    x0 = State(p_B0, v_B0, q0, GT_IMU_DOFS.copy(), notch0.copy(), cam.p[:, 0].copy(), quat_normalise(cam.quats[0]))
    om0, acc0 = f_imu_meas(cam_i.R[0], cam_i.om[:, 0], cam_i.acc[:, 0], cam_i.alp[:, 0], fw0)

    dts, oas, n_prop, ref_rows = [], [], [], []
    old_t = cam.t[0]
    for t in cam.t[1:]:
        old_i = _timestamp_index(cam_i.t, old_t)
        new_i = _timestamp_index(cam_i.t, t)
        old_ti = old_t
        for k in range(old_i + 1, new_i + 1):
            nt = cam_i.notch[k]
            fw = probe.fwkin([*GT_IMU_DOFS, nt[0], 0.0], [0.0] * 6 + [nt[1], 0.0], [0.0] * 6 + [nt[2], 0.0])
            om, acc = f_imu_meas(cam_i.R[k], cam_i.om[:, k], cam_i.acc[:, k], cam_i.alp[:, k], fw)
            ref_rows.append(imu_ref_row(cam_i.t[k], cam_i.p[:, k], cam_i.R[k], cam_i.v[:, k], cam_i.om[:, k], fw))
            dts.append(cam_i.t[k] - old_ti)
            oas.append(np.hstack((om, acc)))
            old_ti = cam_i.t[k]
        n_prop.append(new_i - old_i)
        old_t = t
    cam_meas = np.hstack((meas.p.T[1:], meas.q_raw[1:]))
    notch_meas = meas.notch[1:, 0].copy()
    return x0, np.hstack((om0, acc0)), np.array(dts), np.array(oas), np.array(n_prop), cam_meas, notch_meas, np.array(ref_rows)


def cam_euler_deg(cam: CameraData):
    """``VisualTraj._gen_euler_angles`` (VisualTrajectory.py:151-161)."""
    return np.array([Rotation.from_quat(q).as_euler("xyz", degrees=True) for q in cam.quats])


def run_reference_flow(t, xyz, q_xyzw, cfg: OracleConfig, keep_steps=False, notch=None) -> RunResult:
    """``main.py``: Config -> Simulator -> ``Filter.run`` (Filter.py:144-185).  ``notch`` [n,3] = the rows of the notch
    trajectory file: the ``with_notch: true`` flow (rotated camera as IMU source and as the error reference,
    Filter.py:398)."""
    probe = Probe(cfg.length, cfg.angle)
    cam = camera_from_arrays(t, xyz, q_xyzw, cfg, notch=notch)
    rotated = camera_gen_rotated(cam, cfg) if notch is not None else None
    x0, u0, dts, oas, n_prop, cam_meas, notch_meas, ref_rows = build_streams(cam, cfg, probe, rotated)
    if rotated is not None:
        cam = rotated  # ``cam_reference = camera.rotated if camera.rotated else camera`` (Filter.py:398)
    kf = OracleFilter(cfg, x0, cfg.cov0_matrix, u0[:3], u0[3:], probe)
    res = RunResult(
        kf_rows=None, imu_ref_rows=ref_rows, dt=dts, om_acc=oas, n_prop=n_prop, cam_meas=cam_meas,
        notch_meas=notch_meas, x0=x0.as_vector(), P0=cfg.cov0_matrix.copy(), u0=u0.copy(),
        Rold0=kf.R_WB_old.reshape(9).copy(),
    )
    rows = [filter_traj_row(cam.t[0], kf.x)]
    cam_eul = cam_euler_deg(cam)
    mses = []
    k = 0
    for e, t_e in enumerate(cam.t[1:]):
        for _ in range(n_prop[e]):
            kf.propagate(dts[k], oas[k, :3], oas[k, 3:])
            if keep_steps:
                res.x_steps.append(kf.x.as_vector())
            k += 1
        kf.update(cam_meas[e, :3], cam_meas[e, 3:], notch_meas[e])
        row = filter_traj_row(t_e, kf.x)
        rows.append(row)
        # Filter.calculate_update_mse (Filter.py:397-418)
        i_cam = e + 1
        cam_ref = np.hstack((cam.p[:, i_cam], cam_eul[i_cam]))
        kf_cam = np.hstack((row[20:23], row[23:26]))
        imu_ref_last = ref_rows[k - 1]
        kf_imu = row[4:10]  # vx vy vz rx ry rz
        s_cam = np.sum(np.square(cam_ref - kf_cam))
        s_imu = np.sum(np.square(kf_imu - imu_ref_last[4:10]))
        mses.append((s_cam + s_imu) / 12)
    res.kf_rows = np.array(rows)
    res.x_final = kf.x.as_vector()
    res.P_final = kf.P.copy()
    res.update_mse = np.array(mses)
    r = kf.x.dofs - GT_IMU_DOFS
    res.dof_metric = float(r @ r / 6)  # Filter.calculate_dof_metric (Filter.py:452-455)
    return res
