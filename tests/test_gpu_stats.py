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


def _streams(golden, frames=12, ifv=10):
    from dvi_ekf_b200.camera import Camera, build_streams

    sc = mandala_scenario(golden, n_frames=frames, ifv=ifv)
    a = golden["traj_mandala0_mono"][:frames]
    return sc, build_streams(Camera(a[:, 0], a[:, 1:4], a[:, 4:8], scale=10.0), ifv, sc.cfg.length, sc.cfg.angle)


@pytest.mark.parametrize("budget", [1 << 30, 0])
def test_reduced_vector_masks_diverged_and_flagged_filters(golden, budget):
    """stats_sum (include/eskf.h) sums the HEALTHY filters and counts the others: one filter with a non-finite state and one
    whose update is skipped (singular S: zero prior and zero measurement noise -- the reference's LinAlgError branch,
    Filter.py:358-361) must neither poison the reduced vector nor be counted as healthy.  Both the pre- / post-pass path and
    the in-kernel path (budget 0)."""
    from dvi_ekf_b200 import BatchFilter

    sc, s = _streams(golden)
    n = 40
    x0 = np.repeat(s.x0[None], n, 0)
    x0[5, 3] = np.nan  # diverged from the start (the velocity enters Filter.calculate_update_mse; a NaN position alone never leaves x[0:3])
    P0 = np.repeat(sc.P0[None], n, 0)
    P0[9] = 0.0
    Qd, Rd = np.repeat(sc.Qd[None], n, 0), np.repeat(sc.Rd[None], n, 0)
    Qd[9] = 0.0
    Rd[9] = 0.0  # S = H P H^T + R = 0: every update of filter 9 is skipped
    with BatchFilter(n, **model_kwargs(sc.cfg)) as bf:
        bf.set_prepass_budget(budget)
        bf.set_noise(Qd, Rd, sc.sig_om[None])
        bf.set_state(x0, P0, s.u0[None], None)
        st, sm = bf.run(s.dt, s.om_acc, s.n_prop, s.cam, s.notch, cam_ref=s.cam_ref, imu_ref=s.imu_ref, seed=3,
                        imu_noise_std=np.full(6, 1e-4), cam_noise_std=np.full(7, 1e-5))
        status = bf.get_state()[4]
    assert status[9] != 0 and st[9, 9] == 0 and not np.isfinite(st[5, :10]).all()
    healthy = np.ones(n, bool)
    healthy[[5, 9]] = False
    assert np.isfinite(sm).all() and sm[11] == n - 2 and sm[10] == np.sum(status != 0) and sm[12] == 1
    assert np.allclose(sm[:10], st[healthy][:, :10].sum(axis=0), rtol=1e-12, atol=0)
    from dvi_ekf_b200.sharding import summarise_stats

    out = summarise_stats(sm)
    assert out["filters"] == n - 2 and out["filters_nonfinite"] == 1 and np.isfinite(out["dof_metric_mean"])


@pytest.mark.parametrize("fpc", [28, 8])
def test_prepass_path_is_bit_identical_to_the_in_kernel_path(golden, fpc):
    """eskf_run with the Monte-Carlo generator / the update-MSE in pre- / post-pass kernels (the default when the buffers fit,
    eskf_set_prepass_budget) against the same run with everything inside the persistent kernel: identical bits in the final
    states, covariances and statistics rows; the reduced vector is deterministic in both."""
    from dvi_ekf_b200 import BatchFilter

    sc, s = _streams(golden, frames=16)
    n = 61
    rng = np.random.default_rng(5)
    x0 = np.repeat(s.x0[None], n, 0)
    x0[1:, 10:16] += rng.normal(0, 0.01, (n - 1, 6))
    out = []
    for budget in (1 << 30, 0, 1 << 30):
        with BatchFilter(n, scope_length=sc.cfg.length, cam_angle_rad=sc.cfg.angle, frozen_dofs=(0,) * 6) as bf:
            bf.set_tuning(fpc)
            bf.set_prepass_budget(budget)
            bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
            bf.set_state(x0, sc.P0[None], s.u0[None], None)
            st, sm = bf.run(s.dt, s.om_acc, s.n_prop, s.cam, s.notch, cam_ref=s.cam_ref, imu_ref=s.imu_ref, seed=77,
                            filter_id0=1000, imu_noise_std=np.array([1e-3] * 3 + [0.5] * 3),
                            cam_noise_std=np.array([0.02, 0.002, 0.02, 1e-4, 1e-4, 1e-4, 1e-3]), noise_free_filter0=False)
            x, P, u, R, status = bf.get_state()
        out.append((x, P, u, R, status, st, sm))
    for a, b in zip(out[0], out[1]):
        assert a.tobytes() == b.tobytes()
    assert out[0][6].tobytes() == out[2][6].tobytes()
    assert np.isfinite(out[0][0]).all() and out[0][5][:, 8].min() > 0


def test_run_rejects_inconsistent_stream_geometry(golden, monkeypatch):
    """eskf_run's argument checks (the kernels index trajectories by the GLOBAL filter id and the sample stream by the running
    sum of n_prop): a shard that does not start at a trajectory boundary, filters beyond the last trajectory and epochs
    that add up to more steps than the stream holds are refused with ESKF_EINVAL instead of reading out of bounds."""
    from dvi_ekf_b200 import BatchFilter
    from dvi_ekf_b200._lib import EskfError

    sc, s = _streams(golden)
    two = lambda a: np.concatenate([a, a])
    with BatchFilter(40, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(s.x0[None], sc.P0[None], s.u0[None], None)
        args = (two(s.dt), two(s.om_acc), two(s.n_prop), two(s.cam), two(s.notch))
        with pytest.raises((EskfError, ValueError), match="exceeds n_traj"):
            bf.run(*args, n_traj=2, filters_per_traj=10)  # 40 filters, streams for 20
        with pytest.raises((EskfError, ValueError), match="multiple of filters_per_traj"):
            bf.run(*args, n_traj=2, filters_per_traj=32, filter_id0=8)
        bad = s.n_prop.copy()
        bad[3] += 5
        with pytest.raises((EskfError, ValueError), match="exceeds n_steps"):
            bf.run(s.dt, s.om_acc, bad, s.cam, s.notch)
        import dvi_ekf_b200.engine as eng

        with monkeypatch.context() as m:  # past the host-side mirror of the checks: the C ABI refuses on its own
            m.setattr(eng, "check_run_geometry", lambda *a, **k: None)
            with pytest.raises(EskfError, match="exceeds n_traj"):
                bf.run(*args, n_traj=2, filters_per_traj=10)
            with pytest.raises(EskfError, match="exceeds n_steps"):
                bf.run(s.dt, s.om_acc, bad, s.cam, s.notch)
        st, sm = bf.run(*args, n_traj=2, filters_per_traj=20)  # the consistent call still runs
        assert sm[11] == 40
