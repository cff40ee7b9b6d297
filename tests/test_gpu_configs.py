"""BASELINE.json configurations 3 and 4 as parity cases (they are not bench lines):
  config 3 -- noise / tuning sweep: every filter has its own Q, R and P0 (per-filter parameter rows);
  config 4 -- all data/trajs trajectories stacked in one launch, 33 IMU samples per camera frame (1 kHz / 30 Hz),
              several filters per trajectory, each CTA following its own trajectory's epoch structure."""
import numpy as np
import pytest

from tests.helpers import Scenario, cov_err, mandala_scenario, model_kwargs, state_err
from oracle.eskf_oracle import OracleConfig

# free-running bounds: ~3x the worst value measured on the B200 (printed with -s as "MEASURED ..."; profiles/r02_parity.md)
TOL_SWEEP, TOL_STACK, TOL_RAGGED = 1e-10, 5e-11, 2e-11  # measured 3.4e-11, 1.3e-11, 3.7e-12

pytestmark = pytest.mark.gpu


def _run_oracle(sc, x0=None, P0=None, Qd=None, Rd=None):
    kf = sc.new_oracle(x0, P0)
    if Qd is not None:
        kf.Q = np.diag(Qd)
    if Rd is not None:
        kf.R = np.diag(Rd)
    k = 0
    for e in range(len(sc.n_prop)):
        for _ in range(sc.n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
    return kf


@pytest.mark.parametrize("variant", [3, 1])
def test_config3_tuning_sweep_per_filter_q_r_p0(golden, variant):
    from dvi_ekf_b200 import BatchFilter

    sc = mandala_scenario(golden, n_frames=8, ifv=10)
    n = 36  # 3 x 3 x 4 grid: random-walk scale, measurement-noise scale, P0 scale (+ ragged CTA)
    qs = np.logspace(-2, 2, 3)
    rs = np.logspace(-3, 3, 3)
    ps = np.logspace(-1, 1, 4)
    grid = [(a, b, c) for a in qs for b in rs for c in ps]
    Qd = np.array([sc.Qd * np.hstack((np.ones(6), np.full(7, a))) for a, b, c in grid])
    Rd = np.array([sc.Rd * b for a, b, c in grid])
    P0 = np.array([sc.P0 * c for a, b, c in grid])
    with BatchFilter(n, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(Qd, Rd, sc.sig_om[None])
        bf.set_state(sc.x0[None], P0, sc.u0[None], None)
        st, sm = bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas)
        xg, Pg, ug, Rg, status = bf.get_state()
    assert np.all(status == 0) and sm[11] == n
    worst_s = worst_P = 0.0
    for i in (0, 5, 13, 22, 35):
        kf = _run_oracle(sc, P0=P0[i], Qd=Qd[i], Rd=Rd[i])
        xr, Pr, _, _ = kf.get_vectors()
        worst_s = max(worst_s, state_err(xg[i], xr))
        worst_P = max(worst_P, cov_err(Pg[i], Pr, Rd[i]))
    print(f"MEASURED config3 sweep (70 free-running steps): state {worst_s:.2e} P {worst_P:.2e}")
    assert worst_s < TOL_SWEEP and worst_P < TOL_SWEEP, (worst_s, worst_P)  # free-running, 70 steps
    assert np.abs(xg[0] - xg[35]).max() > 1e-9  # the tuning really changes the estimate


# (v1 has the shapes 28 and 4: fpt 12 runs on 4; eskf_kernel3 cuts every trajectory into CTAs of its own, ragged last one:
# 13 and 37 filters per trajectory run on the 28-filter shape)
@pytest.mark.parametrize("variant,fpt", [(3, 8), (3, 28), (1, 12), (3, 13), (3, 37)])
def test_config4_stacked_trajectories_33_samples_per_frame(golden, variant, fpt):
    from dvi_ekf_b200 import BatchFilter

    names = ["traj_trans_x", "traj_rot_z", "traj_from_prop", "traj_mandala0_gt"]
    frames, ifv = 7, 33
    scs = [Scenario(golden[nm][:frames], OracleConfig(max_vals=frames, interframe_vals=ifv)) for nm in names]
    T, E = len(scs[0].dt), len(scs[0].n_prop)
    assert all(len(s.dt) == T and len(s.n_prop) == E for s in scs)
    cat = lambda f: np.concatenate([f(s) for s in scs])
    n = fpt * len(scs)
    x0 = np.concatenate([np.repeat(s.x0[None], fpt, 0) for s in scs])
    u0 = np.concatenate([np.repeat(s.u0[None], fpt, 0) for s in scs])
    sc0 = scs[0]
    with BatchFilter(n, variant=variant, **model_kwargs(sc0.cfg)) as bf:
        bf.set_noise(sc0.Qd[None], sc0.Rd[None], sc0.sig_om[None])
        bf.set_state(x0, sc0.P0[None], u0, None)
        st, sm = bf.run(cat(lambda s: s.dt), cat(lambda s: s.om_acc), cat(lambda s: s.n_prop), cat(lambda s: s.cam_meas),
                        cat(lambda s: s.notch_meas), n_traj=len(scs), filters_per_traj=fpt)
        xg, Pg, ug, Rg, status = bf.get_state()
    assert np.all(status == 0) and sm[11] == n and np.allclose(st[:, 9], E)
    for j, s in enumerate(scs):
        kf = _run_oracle(s)
        xr, Pr, ur, Rr = kf.get_vectors()
        for i in (j * fpt, (j + 1) * fpt - 1):
            print(f"MEASURED config4 stacked {names[j]} (198 free-running steps): state {state_err(xg[i], xr):.2e} P {cov_err(Pg[i], Pr, s.Rd):.2e}")
            assert state_err(xg[i], xr) < TOL_STACK, (names[j], state_err(xg[i], xr))  # free-running, 198 steps
            assert cov_err(Pg[i], Pr, s.Rd) < TOL_STACK, (names[j], cov_err(Pg[i], Pr, s.Rd))
            assert np.abs(ug[i] - ur).max() < 1e-12


@pytest.mark.parametrize("variant", [3, 1])
def test_trace_mode_reproduces_the_reference_golden_file_through_eskf_run(golden, variant, tmp_path):
    """one persistent launch with trace_x, the FilterTraj rows written with the reference's text format
    (files.py:68-82) and read back: every number of kf_best_mandala0_mono.txt to its 9 printed decimals."""
    from dvi_ekf_b200 import BatchFilter
    from dvi_ekf_b200.trace import save_filter_traj

    sc = mandala_scenario(golden, n_frames=10, ifv=1, zero_frozen_dofs=False, euler_mode="zyx_legacy")
    n, T = 5, len(sc.dt)
    trace = np.full((n, T, 26), np.nan)
    with BatchFilter(n, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas, trace=trace)
        xg = bf.get_state()[0]
    assert np.all(np.isfinite(trace)) and np.array_equal(trace[:, -1], xg)  # last row = final state
    t_imu = sc.cam.t[1:]  # interframe 1: one IMU step per camera frame
    fn = tmp_path / "kf_best_mandala0_mono.txt"
    save_filter_traj(str(fn), sc.cam.t[0], sc.x0, t_imu, trace[n - 1])
    ours = np.loadtxt(fn)
    ref = golden["kf_best_mandala0_mono"]
    assert ours.shape == ref.shape == (10, 30)
    assert np.abs(ours - ref).max() <= 1.0e-9 + 1e-15  # both sides rounded to 9 decimals


def test_trace_mode_rows_between_updates(golden):
    """interframe 10: rows of propagation steps hold the propagated state, the last row of an epoch the updated one"""
    from dvi_ekf_b200 import BatchFilter

    sc = mandala_scenario(golden, n_frames=5, ifv=10)
    T = len(sc.dt)
    trace = np.zeros((2, T, 26))
    with BatchFilter(2, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas, trace=trace)
    kf = sc.new_oracle()
    k = 0
    worst = 0.0
    for e in range(len(sc.n_prop)):
        for j in range(sc.n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            if j < sc.n_prop[e] - 1:
                worst = max(worst, state_err(trace[1, k], kf.get_vectors()[0]))
            k += 1
        kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
        worst = max(worst, state_err(trace[1, k - 1], kf.get_vectors()[0]))
    assert worst < 1e-9, worst


def test_batched_tuner_population_in_one_launch():
    """the objective of Simulator.optimise for a whole population of (rw_std, meas_std) candidates in
    one launch equals evaluating the candidates one by one, the candidates share their noise realisations, and one
    generation of differential evolution runs on top of it."""
    import os

    from dvi_ekf_b200 import Config, Simulator

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config(os.path.join(root, "config.yaml"))
    cfg.frozen_dofs = [0, 0, 0, 0, 0, 0]  # let the filter estimate the DOFs: the DOF MSE then depends on the tuning
    sim = Simulator(cfg)
    base = np.array([*cfg.process_noise_rw_std, *cfg.meas_noise_std], dtype=float)
    X = np.array([base, base * 3.0, base * 0.3, base * np.linspace(0.5, 2.0, 14)])
    f_all = sim.evaluate_candidates(X, runs_per_candidate=6)
    assert f_all.shape == (4,) and np.all(np.isfinite(f_all)) and np.all(f_all >= 0)
    f_one = np.array([sim.evaluate_candidates(X[j : j + 1], runs_per_candidate=6)[0] for j in range(4)])
    assert np.array_equal(f_all, f_one)  # common random numbers: a candidate's score does not depend on its neighbours
    assert len(set(np.round(f_all, 14))) > 1  # the parameters matter
    ret = sim.optimise(maxiter=1, popsize=1, runs_per_candidate=3, seed=1)
    assert ret.x.shape == (14,) and np.isfinite(ret.fun)
    assert all(lo <= v <= hi for v, (lo, hi) in zip(ret.x, sim.OPTIM_BOUNDS))
    assert ret.fun <= sim.evaluate_candidates(np.array([ret.x]), runs_per_candidate=3)[0] + 1e-12


def test_tuner_objective_matches_oracle(golden, tmp_path):
    """SURVEY 8f rank 3, parity proper: ``Simulator.evaluate_candidates`` (the objective of Simulator.optimise,
    Simulator.py:163-245; Filter.calculate_dof_metric Filter.py:452-455 and the update MSE Filter.py:397-418) for
    2 candidates x 3 runs with all six DOFs estimated, against the ORACLE run filter by filter on the same inputs: the noise
    the kernels drew is read back with eskf_noise_dump (ids 0..r-1: noise_id_modulus = r), the initial-condition perturbation
    is rng([seed, i % r]) and every candidate has its own Q / R."""
    import os

    import yaml

    from dvi_ekf_b200 import Config, Simulator
    from dvi_ekf_b200.engine import NOISE_CAM, NOISE_IMU, noise_samples
    from oracle.eskf_oracle import euler_xyz_deg, quat_about_axis, quat_mul

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    y = yaml.safe_load(open(os.path.join(root, "config.yaml")))
    frames, ifv, r = 14, 5, 3
    y["simulation"].update(do_fast_sim=False, frozen_dofs=[0] * 6)
    y["camera"]["total_frames"] = frames
    y["imu"]["interframe_vals"] = ifv
    fp = tmp_path / "config.yaml"
    fp.write_text(yaml.safe_dump(y))
    cfg = Config(str(fp))
    sim = Simulator(cfg)
    b = cfg.batch
    base = np.array([*cfg.process_noise_rw_std, *cfg.meas_noise_std], dtype=float)
    X = np.array([base, base * np.linspace(0.5, 2.0, 14)])
    f_dof = sim.evaluate_candidates(X, runs_per_candidate=r, metric="dof")
    f_upd = sim.evaluate_candidates(X, runs_per_candidate=r, metric="update")

    sc = mandala_scenario(golden, n_frames=frames, ifv=ifv, frozen_dofs=(0,) * 6)
    s = sim.streams
    T, E = len(sc.dt), len(sc.n_prop)
    assert T == len(s.dt) and E == len(s.n_prop) and np.abs(sc.om_acc - s.om_acc).max() < 1e-9
    imu_std = np.hstack((cfg.imu.stdev_omega, cfg.imu.stdev_accel))
    cam_std = np.array(cfg.meas_noise_std)
    zi = noise_samples(b.seed, 0, r, 0, T, NOISE_IMU)
    zc = noise_samples(b.seed, 0, r, 0, E, NOISE_CAM)
    gt = np.asarray(cfg.gt_imu_dofs, dtype=float)
    ref_dof, ref_upd = np.zeros((2, r)), np.zeros((2, r))
    for c in range(2):
        for j in range(r):
            x0 = sc.x0.copy()
            oa, cam, notch = sc.om_acc.copy(), sc.cam_meas.copy(), sc.notch_meas.copy()
            if j:  # run 0 of every candidate: nominal initial state, noise free
                rng = np.random.default_rng([b.seed, j])
                x0[10:13] += rng.normal(0.0, np.deg2rad(b.dof_ic_std_deg), 3)
                x0[13:16] += rng.normal(0.0, b.dof_ic_std_cm, 3)
                oa = oa + imu_std[None, :] * zi[j][:, :6]
                for e in range(E):
                    cam[e, :3] += cam_std[:3] * zc[j][e, :3]
                    dth = cam_std[3:6] * zc[j][e, 3:6]
                    nq = np.linalg.norm(cam[e, 3:])
                    cam[e, 3:] = quat_mul(cam[e, 3:], quat_about_axis(np.linalg.norm(dth), dth)) * nq
                    notch[e] += cam_std[6] * zc[j][e, 6]
            kf = sc.new_oracle(x0=x0)
            kf.Q = np.diag(np.hstack((np.zeros(6), np.square(X[c, 0:7]))))
            kf.R = np.diag(np.square(X[c, 7:14]))
            k, acc = 0, 0.0
            for e in range(E):
                for _ in range(sc.n_prop[e]):
                    kf.propagate(sc.dt[k], oa[k, :3], oa[k, 3:])
                    k += 1
                assert kf.update(cam[e, :3], cam[e, 3:], notch[e]) is not None
                x = kf.x
                acc += (np.sum(np.square(s.cam_ref[e] - np.hstack((x.p_cam, euler_xyz_deg(x.q_cam)))))
                        + np.sum(np.square(np.hstack((x.v, euler_xyz_deg(x.q))) - s.imu_ref[e]))) / 12
            d = kf.x.dofs - gt
            ref_dof[c, j] = d @ d / 6
            ref_upd[c, j] = acc / E
    rd, ru = ref_dof.mean(axis=1), ref_upd.mean(axis=1)
    print(f"tuner objective vs oracle: dof {np.abs(f_dof / rd - 1).max():.2e}, update {np.abs(f_upd / ru - 1).max():.2e}")
    assert np.abs(f_dof / rd - 1).max() < 1e-10 and np.abs(f_upd / ru - 1).max() < 1e-10  # measured 2.4e-11 / 6.6e-12
    assert abs(rd[0] / rd[1] - 1) > 1e-3 and ref_dof[0, 1] != ref_dof[0, 0]  # candidates and runs really differ


def test_main_py_flow_reproduces_the_reference_artefacts(golden, tmp_path, capsys):
    """The reference's main.py (main.py:3-8: Config -> Simulator -> run_once) followed by Filter.save (Filter.py:457-465),
    through the mirror classes with ``batch.legacy_golden: true`` (the presets of the commit that wrote the artefacts):
    the two files it writes are data/trajs/kf_best_mandala0_mono.txt and imu_ref_mandala0_mono.txt, number for number at the
    9 printed decimals."""
    import os

    import yaml

    from dvi_ekf_b200 import Config, Simulator

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    y = yaml.safe_load(open(os.path.join(root, "config.yaml")))
    y["batch"]["legacy_golden"] = True
    y["simulation"]["traj_path"] = str(tmp_path)  # Filter.save writes next to the trajectories
    fp = tmp_path / "config.yaml"
    fp.write_text(yaml.safe_dump(y))
    config = Config(str(fp))
    sim = Simulator(config)
    sim.run_once()
    sim.kf.save()
    out = capsys.readouterr().out
    assert "MSE" in out and "Singular" not in out
    ours = np.loadtxt(tmp_path / "kf_best_mandala0_mono.txt")
    ref = golden["kf_best_mandala0_mono"]
    assert ours.shape == ref.shape == (10, 30)
    assert np.abs(ours - ref).max() <= 1.0e-9 + 1e-15, np.abs(ours - ref).max()  # both sides rounded to 9 decimals
    ours_imu = np.loadtxt(tmp_path / "imu_ref_mandala0_mono.txt")
    ref_imu = golden["imu_ref_mandala0_mono"]
    assert ours_imu.shape == ref_imu.shape == (9, 14)
    assert np.abs(ours_imu - ref_imu).max() <= 1.0e-9 + 1e-15, np.abs(ours_imu - ref_imu).max()
    # the attributes main.py's consumers read afterwards
    assert len(sim.kf.traj.rows) == 10 and len(sim.kf.imu.ref_rows) == 9
    assert sim.kf.Fx.shape == (24, 24) and sim.kf.Fi.shape == (24, 13) and np.isfinite(sim.kf._P).all()


@pytest.mark.parametrize("variant", [3, 1])
def test_ragged_and_empty_epochs(golden, variant):
    """epochs with 0, 1, 2 and many IMU samples in one launch (the record pipeline, the mid-step waits and the barriers
    around the update must hold for every epoch length), against the oracle driven by the same n_prop."""
    from dvi_ekf_b200 import BatchFilter

    sc = mandala_scenario(golden, n_frames=8, ifv=10)  # 70 samples, 7 camera frames
    n_prop = np.array([3, 0, 1, 17, 2, 0, 47], dtype=np.int32)
    assert n_prop.sum() == len(sc.dt)
    with BatchFilter(9, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        st, sm = bf.run(sc.dt, sc.om_acc, n_prop, sc.cam_meas, sc.notch_meas)
        xg, Pg, ug, Rg, status = bf.get_state()
    kf = sc.new_oracle()
    k = 0
    for e in range(len(n_prop)):
        for _ in range(n_prop[e]):
            kf.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        kf.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
    xr, Pr, ur, Rr = kf.get_vectors()
    assert np.all(status == 0) and np.allclose(st[:, 9], len(n_prop))
    for i in (0, 8):
        print(f"MEASURED ragged epochs (70 free-running steps): state {state_err(xg[i], xr):.2e} P {cov_err(Pg[i], Pr, sc.Rd):.2e}")
        assert state_err(xg[i], xr) < TOL_RAGGED and cov_err(Pg[i], Pr, sc.Rd) < TOL_RAGGED
        assert np.abs(ug[i] - ur).max() < 1e-12 and np.abs(Rg[i] - Rr).max() < 1e-9
