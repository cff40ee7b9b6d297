"""The batch-vectorised oracle (CPU baseline) equals the single-filter oracle."""
import numpy as np

from oracle.batch_oracle import BatchOracle
from tests.helpers import cov_err, mandala_scenario, random_filter_inputs, state_err


def test_batch_oracle_matches_single_oracle_on_trajectory(golden):
    sc = mandala_scenario(golden, n_frames=12, ifv=4)
    kf = sc.new_oracle()
    bo = BatchOracle(sc.cfg, np.repeat(sc.x0[None], 3, 0), sc.P0, sc.u0)
    k = 0
    for e in range(len(sc.n_prop)):
        for _ in range(sc.n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            bo.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
        bo.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
    xr, Pr, ur, Rr = kf.get_vectors()
    for i in range(3):
        assert state_err(bo.x[i], xr) < 1e-10
        assert cov_err(bo.P[i], Pr, sc.Rd) < 1e-10
    assert np.abs(bo.R_old[0].reshape(9) - Rr).max() < 1e-14


def test_batch_oracle_matches_single_oracle_random(golden):
    rng = np.random.default_rng(3)
    sc = mandala_scenario(golden, n_frames=10, ifv=10, frozen_dofs=(0, 1, 0, 0, 1, 0))
    n = 40
    xs, Ps, us = random_filter_inputs(rng, n, sc.cfg)
    oa = np.hstack((rng.normal(0, 0.05, (n, 3)), rng.normal(0, 0.5, (n, 3))))
    bo = BatchOracle(sc.cfg, xs, Ps, us)
    bo.propagate(0.1, oa[:, :3], oa[:, 3:])
    kfs = []
    for i in range(n):
        kf = sc.new_oracle(xs[i], Ps[i], us[i])
        kf.propagate(0.1, oa[i, :3], oa[i, 3:])
        xr, Pr, *_ = kf.get_vectors()
        assert state_err(bo.x[i], xr) < 1e-11 and cov_err(bo.P[i], Pr) < 1e-11
        kfs.append(kf)
    cams = np.hstack((bo.x[:, 19:22] + rng.normal(0, 0.05, (n, 3)), bo.x[:, 22:26] * 1.3 + rng.normal(0, 0.01, (n, 4))))
    notch = rng.normal(0, 0.2, n)
    bo.update(cams[:, :3], cams[:, 3:], notch)
    for i in range(n):
        kfs[i].update(cams[i, :3], cams[i, 3:], notch[i])
        xr, Pr, *_ = kfs[i].get_vectors()
        assert state_err(bo.x[i], xr) < 1e-10 and cov_err(bo.P[i], Pr, sc.Rd) < 1e-10
