"""Literal sympy transcription of the reference's *symbolic* camera-error
Jacobian (TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``).

The reference builds rows 18:24 of Fx / Fi by casadi automatic differentiation
of expressions assembled in ``dvi_ekf/kinematics/symbols.py:134-200`` and
``dvi_ekf/filter/Filter.py:270-342`` on top of the symbolic probe kinematics
(``dvi_ekf/models/Probe.py:186-306,431-468``).  casadi is not installable
here, so this module rebuilds the very same expressions with sympy (symbol for
symbol, including the ``v_tr = p_tr`` slip and the scalar ``err_notch`` in
``err_x``), differentiates them with ``sympy.Matrix.jacobian`` and lambdifies
the result.  ``tests/test_oracle_jacobian.py`` compares the closed forms used
by ``oracle/eskf_oracle.py`` against it at random operating points.
"""
from __future__ import annotations

import functools

import numpy as np
import sympy as sp


def _rz(t):
    return sp.Matrix([[sp.cos(t), -sp.sin(t), 0, 0], [sp.sin(t), sp.cos(t), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def _rx(a):
    return sp.Matrix([[1, 0, 0, 0], [0, sp.cos(a), -sp.sin(a), 0], [0, sp.sin(a), sp.cos(a), 0], [0, 0, 0, 1]])


def _ry(a):
    return sp.Matrix([[sp.cos(a), 0, sp.sin(a), 0], [0, 1, 0, 0], [-sp.sin(a), 0, sp.cos(a), 0], [0, 0, 0, 1]])


def _tz(d):
    m = sp.eye(4)
    m[2, 3] = d
    return m


def _skew(v):
    return sp.Matrix([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def probe_symbolic(q, qd7, L, ang):
    """Probe.get_sym for SymProbe (q8 = 0; only the notch joint moves):
    returns p, R, om  (v = J_v qd is evaluated too, to show it is zero)."""
    hp = sp.pi / 2
    links = [
        _rz(q[0] + hp) * _tz(0) * _rx(hp),
        _rz(q[1] - hp) * _tz(0) * _rx(-hp),
        _rz(q[2]) * _tz(0) * _rx(0),
        _rz(0) * _tz(q[3]) * _rx(hp),
        _rz(hp) * _tz(q[4]) * _rx(hp),
        _rz(-hp) * _tz(q[5]) * _rx(hp),
        _rz(q[6]) * _tz(L) * _rx(ang),
        _rz(0) * _tz(0) * _rx(0),
    ]
    T = _ry(-sp.pi)
    frames = []
    for A in links:
        frames.append(T)
        T = T * A
    p = T[:3, 3]
    R = T[:3, :3]
    z6 = frames[6][:3, 2]
    o6 = frames[6][:3, 3]
    v = z6.cross(p - o6) * qd7  # jacob0 column 7 (linear part) times q7_dot
    om = z6 * qd7
    return p, R, v, om


@functools.lru_cache(maxsize=None)
def build(fix_q2=False):
    """Returns f(dt, dofs6, notch3, R_WB(3x3), om(3), n_om(3), L, ang) ->
    (Jx 6x22, Jn 6x13), both evaluated at zero error state."""
    dt, L, ang = sp.symbols("dt L ang", real=True)
    q = sp.symbols("q1:8", real=True)
    qd7 = sp.Symbol("q7_dot", real=True)
    eq = sp.symbols("err_q1:8", real=True)
    R_WB = sp.Matrix(3, 3, sp.symbols("R_WB_0:9", real=True))
    om = sp.Matrix(sp.symbols("om_0:3", real=True))
    n_a = sp.Matrix(sp.symbols("n_a_0:3", real=True))
    n_om = sp.Matrix(sp.symbols("n_om_0:3", real=True))
    n_dofs = sp.Matrix(sp.symbols("n_dofs_0:6", real=True))
    n_notch = sp.Symbol("n_notch_acc", real=True)
    err_p_B = sp.Matrix(sp.symbols("err_p_B_0:3", real=True))
    err_v_B = sp.Matrix(sp.symbols("err_v_B_0:3", real=True))
    err_th = sp.Matrix(sp.symbols("err_theta_0:3", real=True))
    err_p_C = sp.Matrix(sp.symbols("err_p_C_0:3", real=True))
    err_th_C = sp.Matrix(sp.symbols("err_theta_C_0:3", real=True))

    p, Rp, v, om_p = probe_symbolic(q, qd7, L, ang)
    sub = {q[i]: q[i] + eq[i] for i in range(7)}  # SymProbe._get_tr, Probe.py:465-468
    p_tr = p.subs(sub, simultaneous=True)
    R_tr = Rp.subs(sub, simultaneous=True)
    v_tr = v.subs(sub, simultaneous=True) if fix_q2 else p_tr  # Probe.py:458
    om_p_tr = om_p.subs(sub, simultaneous=True)

    # symbols.py:129-131
    R_WB_tr = R_WB * (sp.eye(3) + _skew(err_th))
    om_tr = om - n_om

    # get_err_pc_dot, symbols.py:134-160
    p_CB_dot = R_WB * (v + om.cross(p))
    p_CB_dot_tr = R_WB_tr * (v_tr + om_tr.cross(p_tr))
    err_p_C_dot = err_v_B + p_CB_dot_tr - p_CB_dot

    # get_err_theta_c_dot, symbols.py:186-200
    om_c = Rp.T * (om + om_p)
    om_c_tr = R_tr.T * (om_tr + om_p_tr)

    def quat_matrix(v3, direction):  # symbols.py:168-183 with w = 0
        m = sp.zeros(4, 4)
        m[0, 1:] = -v3.T
        m[1:, 0] = v3
        m[1:, 1:] = -_skew(v3) if direction == "r" else _skew(v3)
        return m

    err_q_C = sp.Matrix([1, *(sp.Rational(1, 2) * err_th_C)])
    M_om = quat_matrix(om_c_tr, "r") - quat_matrix(om_c, "l")
    err_theta_c_dot = (M_om * err_q_C)[1:, 0]

    # Filter._cam_error_jacobian, Filter.py:274-285
    err_p_C_next = err_p_C + dt * err_p_C_dot
    err_theta_C_next = err_th_C + dt * err_theta_c_dot
    err_x = [*err_p_B, *err_v_B, *err_th, *eq[0:6], eq[6], *err_p_C, *err_th_C]  # symbols.py:107
    n = [*n_a, *n_om, *n_dofs, n_notch]
    stacked = sp.Matrix([*err_p_C_next, *err_theta_C_next])
    Jx = stacked.jacobian(err_x)
    Jn = stacked.jacobian(n)
    zero = {s: 0 for s in [*eq, *err_p_B, *err_v_B, *err_th, *err_p_C, *err_th_C]}
    Jx = Jx.subs(zero)
    Jn = Jn.subs(zero)
    args = [dt, *q, qd7, *R_WB, *om, *n_om, L, ang]
    fx = sp.lambdify(args, Jx, modules="numpy", cse=True)
    fn = sp.lambdify(args, Jn, modules="numpy", cse=True)

    def f(dt_, dofs, notch, R_WB_, om_, n_om_, L_, ang_):
        a = [dt_, *dofs, notch[0], notch[1], *np.asarray(R_WB_).reshape(9), *om_, *n_om_, L_, ang_]
        return np.array(fx(*a), dtype=float), np.array(fn(*a), dtype=float)

    return f
