"""Closed-form error Jacobians of the oracle vs automatic differentiation of
the literal transcription of symbols.py / Filter._cam_error_jacobian."""
import numpy as np
import pytest

from oracle import symbolic_check
from oracle.eskf_oracle import OracleConfig, OracleFilter, State, quat_normalise


@pytest.mark.parametrize("fix_q2", [False, True])
def test_cam_error_jacobian_matches_autodiff(fix_q2):
    f = symbolic_check.build(fix_q2=fix_q2)
    rng = np.random.default_rng(0)
    cfg = OracleConfig(interframe_vals=10, frozen_dofs=(0,) * 6, fix_q2=fix_q2)
    for _ in range(25):
        dofs = np.hstack((rng.normal(0, 0.5, 3), rng.normal(0, 5, 2), 20 + rng.normal(0, 5)))
        notch = rng.normal(0, 0.3, 3)
        x0 = State(rng.normal(0, 1, 3), rng.normal(0, 1, 3), quat_normalise(rng.normal(0, 1, 4)), dofs, notch,
                   rng.normal(0, 1, 3), quat_normalise(rng.normal(0, 1, 4)))
        kf = OracleFilter(cfg, x0, cfg.cov0_matrix, rng.normal(0, 0.2, 3), rng.normal(0, 1, 3))
        kf.dt = float(rng.uniform(0.01, 1.0))
        kf._predict_error()
        Jx, Jn = f(kf.dt, kf.x.dofs, kf.x.notch_dofs, kf.R_WB_old, kf.om_old, kf.stdev_nom, cfg.length, cfg.angle)
        scale = max(1.0, np.abs(Jx).max())
        assert np.abs(kf.Fx[18:24, 0:22] - Jx).max() <= 1e-13 * scale
        assert np.abs(kf.Fi[18:24] - Jn).max() <= 1e-13 * scale
        # columns 22:24 of rows 18:24 keep their identity values (Filter.py:279-285)
        assert np.array_equal(kf.Fx[18:24, 22:24], np.eye(24)[18:24, 22:24])
