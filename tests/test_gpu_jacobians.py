"""Filter.Fx / Filter.Fi read-out (Filter.py:249-342): the Jacobian record the kernels file for the last IMU step, expanded by
eskf_get_jacobians into the dense matrices the reference holds, against (a) the oracle's Fx / Fi and (b) automatic
differentiation of the literal transcription of symbols.py / Filter._cam_error_jacobian (oracle/symbolic_check.py: what
CasADi evaluates numerically in the reference)."""
import numpy as np
import pytest

from oracle import symbolic_check
from oracle.eskf_oracle import OracleConfig
from tests.helpers import mandala_scenario, model_kwargs, random_filter_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("imu_q", [False, True])
@pytest.mark.parametrize("fpc", [28, 4])
def test_dense_jacobians_match_oracle_and_autodiff(golden, imu_q, fpc):
    from dvi_ekf_b200 import BatchFilter

    rng = np.random.default_rng(21)
    cfg = OracleConfig(interframe_vals=10, frozen_dofs=(0, 0, 0, 0, 0, 0))
    sc = mandala_scenario(golden, n_frames=10, ifv=10, frozen_dofs=(0, 0, 0, 0, 0, 0))
    n = 300
    xs, Ps, us = random_filter_inputs(rng, n, cfg)
    R_old = np.array([np.linalg.qr(rng.normal(size=(3, 3)))[0].reshape(9) for _ in range(n)])  # R_WB_old != rot(q): quirk Q8
    oa = np.hstack((rng.normal(0, 0.05, (n, 3)), rng.normal(0, 0.5, (n, 3))))
    dt = 0.07
    Qd = sc.Qd.copy()
    if imu_q:
        Qd[0:6] = rng.uniform(1e-6, 1e-4, 6)  # Filter.update_noise_matrices with dt != 0 (Filter.py:110-117)
    f = symbolic_check.build(fix_q2=False)
    with BatchFilter(n, **model_kwargs(sc.cfg)) as bf:
        bf.set_tuning(fpc)
        bf.set_noise(Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(xs, Ps, us, R_old)
        bf.keep_jacobians(True)
        bf.propagate(np.array([0.3, dt]), np.stack((oa * 0.5, oa), axis=1))  # two steps: the record of the LAST one is kept
        Fx, Fi = bf.get_jacobians()
    worst = worst_ad = 0.0
    for i in range(n):
        kf = sc.new_oracle(xs[i], Ps[i], us[i])
        kf.R_WB_old = R_old[i].reshape(3, 3).copy()
        kf.Q = np.diag(Qd)
        kf.propagate(0.3, 0.5 * oa[i, :3], 0.5 * oa[i, 3:])
        R_pre, om_pre = kf.R_WB_old.copy(), kf.om_old.copy()
        kf.propagate(dt, oa[i, :3], oa[i, 3:])
        sx, si = max(1.0, np.abs(kf.Fx).max()), max(1.0, np.abs(kf.Fi).max())
        worst = max(worst, np.abs(Fx[i] - kf.Fx).max() / sx, np.abs(Fi[i] - kf.Fi).max() / si)
        # structure: exact ones and zeros where the reference has them
        assert np.array_equal(Fx[i][kf.Fx == 0.0], kf.Fx[kf.Fx == 0.0]) and np.array_equal(Fx[i][kf.Fx == 1.0], kf.Fx[kf.Fx == 1.0])
        if i < 25:  # the CasADi-evaluated blocks (rows 18:24) against automatic differentiation at the same operating point
            Jx, Jn = f(dt, kf.x.dofs, kf.x.notch_dofs, R_pre, om_pre, kf.stdev_nom, cfg.length, cfg.angle)
            s = max(1.0, np.abs(Jx).max())
            worst_ad = max(worst_ad, np.abs(Fx[i][18:24, 0:22] - Jx).max() / s, np.abs(Fi[i][18:24] - Jn).max() / s)
    print(f"Jacobian read-out: vs oracle {worst:.2e}, vs autodiff {worst_ad:.2e}")
    assert worst < 1e-12 and worst_ad < 1e-12


def test_filter_mirror_exposes_fx_fi(golden, tmp_path):
    """``Filter.Fx`` / ``Filter.Fi`` of the mirror class: None before the first propagate (the reference creates them there),
    24 x 24 / 24 x 13 afterwards, and equal to the oracle's after the default run."""
    from dvi_ekf_b200 import Config, Simulator

    sim = Simulator(Config("config.yaml"))
    assert sim.kf.Fx is None and sim.kf.Fi is None
    sim.run_once()
    Fx, Fi = sim.kf.Fx, sim.kf.Fi
    assert Fx.shape == (24, 24) and Fi.shape == (24, 13)
    sc = mandala_scenario(golden, n_frames=10, ifv=1)
    kf = sc.new_oracle()
    k = 0
    for e in range(len(sc.n_prop)):
        for _ in range(sc.n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
    assert np.abs(Fx - kf.Fx).max() < 1e-9 * max(1.0, np.abs(kf.Fx).max()) and np.abs(Fi - kf.Fi).max() < 1e-9
