"""Error statistics of eskf_run against the oracle: Filter.calculate_update_mse (Filter.py:397-418: twelve squared errors
against the camera trajectory and the IMU reference, Euler angles in degrees) per epoch, and Filter.calculate_dof_metric
(Filter.py:452-455).  Row layout of the statistics: include/eskf.h."""
import numpy as np
import pytest

from oracle.eskf_oracle import euler_xyz_deg
from tests.helpers import mandala_scenario, model_kwargs

pytestmark = pytest.mark.gpu


def _oracle_update_mse(sc, n_prop, cam_ref, imu_ref):
    kf = sc.new_oracle()
    k, out = 0, []
    for e in range(len(n_prop)):
        for _ in range(n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        assert kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e]) is not None
        x = kf.x
        s_cam = np.sum(np.square(cam_ref[e] - np.hstack((x.p_cam, euler_xyz_deg(x.q_cam)))))
        s_imu = np.sum(np.square(np.hstack((x.v, euler_xyz_deg(x.q))) - imu_ref[e]))
        out.append((s_cam + s_imu) / 12)
    r = kf.x.dofs - np.array([0, 0, 0, 0, 0, 20.0])
    return np.array(out), float(r @ r / 6)


@pytest.mark.parametrize("variant,fpc", [(3, 28), (3, 8), (1, 28)])
@pytest.mark.parametrize("ragged", [False, True])
def test_update_mse_and_dof_metric_match_oracle(golden, variant, fpc, ragged):
    from dvi_ekf_b200 import BatchFilter
    from dvi_ekf_b200.camera import Camera, build_streams

    frames, ifv = 12, 10
    sc = mandala_scenario(golden, n_frames=frames, ifv=ifv)
    a = golden["traj_mandala0_mono"][:frames]
    s = build_streams(Camera(a[:, 0], a[:, 1:4], a[:, 4:8], scale=10.0), ifv, sc.cfg.length, sc.cfg.angle)
    n_prop = s.n_prop.copy()
    if ragged:  # epochs of 0, 1 and 2 steps: the deferred evaluation of the statistics has to catch up before the next update
        n_prop = np.array([3, 0, 1, 2, 24, 10, 10, 10, 10, 10, 30], dtype=np.int32)
        assert n_prop.sum() == len(s.dt) and len(n_prop) == len(s.n_prop)
    n = 33
    with BatchFilter(n, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_tuning(fpc)
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(s.x0[None], sc.P0[None], s.u0[None], None)
        st, sm = bf.run(s.dt, s.om_acc, n_prop, s.cam, s.notch, cam_ref=s.cam_ref, imu_ref=s.imu_ref)
        x = bf.get_state()[0]
    mse, dof = _oracle_update_mse(sc, n_prop, s.cam_ref, s.imu_ref)
    for i in (0, 7, n - 1):
        assert abs(st[i, 7] - mse[-1]) <= 1e-8 * abs(mse[-1]), (st[i, 7], mse[-1])      # update_mse of the last epoch
        assert abs(st[i, 8] - mse.sum()) <= 1e-8 * abs(mse.sum()), (st[i, 8], mse.sum())  # summed over the epochs
        assert abs(st[i, 6] - dof) <= 1e-9 * max(dof, 1e-12) + 1e-18
        assert st[i, 9] == len(n_prop) and st[i, 10] == 0 and st[i, 11] == 1
    assert np.allclose(sm[7], st[:, 7].sum(), rtol=1e-12) and np.allclose(sm[8], st[:, 8].sum(), rtol=1e-12) and sm[11] == n
    assert np.all(np.isfinite(x))
