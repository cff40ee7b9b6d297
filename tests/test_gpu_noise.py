"""The Monte-Carlo path of eskf_run (BASELINE config 2: in-kernel Philox noise on the IMU and camera streams,
the workload bench.py times) against the numpy oracle.  The kernels draw the noise with the device's SFU
approximations, so the samples are read back with eskf_noise_dump (include/eskf.h) and the SAME samples are
added to the oracle's input streams: what is compared is the filter arithmetic on identical inputs."""
import numpy as np
import pytest

from oracle.eskf_oracle import quat_about_axis, quat_mul
from tests.helpers import cov_err, mandala_scenario, model_kwargs, state_err

pytestmark = pytest.mark.gpu

SEED = 4321
IMU_STD = np.array([2.8e-4, 2.8e-4, 2.8e-4, 1.24, 1.24, 1.24])  # rad/s, cm/s^2 (config.py:105-120 at 10 samples / frame)
CAM_STD = np.array([0.1, 0.1, 0.1, 0.005, 0.005, 0.005, 0.01])


def _noisy_streams(sc, zi, zc):
    """The kernel's noise model (eskf_kernel3.cuh role3_stage / eskf_kernel.cuh): additive on om, acc and the camera
    position, a small body rotation on the measured quaternion (norm kept), additive on the notch angle."""
    oa = sc.om_acc + IMU_STD[None, :] * zi[:, :6]
    cam = sc.cam_meas.copy()
    notch = sc.notch_meas.copy()
    for e in range(len(cam)):
        cam[e, :3] += CAM_STD[:3] * zc[e, :3]
        dth = CAM_STD[3:6] * zc[e, 3:6]
        dq = quat_about_axis(np.linalg.norm(dth), dth)
        nq = np.linalg.norm(cam[e, 3:])
        cam[e, 3:] = quat_mul(cam[e, 3:], dq) * nq
        notch[e] += CAM_STD[6] * zc[e, 6]
    return oa, cam, notch


@pytest.mark.parametrize("variant,fpc", [(3, 28), (3, 4), (1, 28)])
def test_monte_carlo_run_matches_oracle_on_the_same_noise(golden, variant, fpc):
    from dvi_ekf_b200 import BatchFilter
    from dvi_ekf_b200.engine import NOISE_CAM, NOISE_IMU, noise_samples

    sc = mandala_scenario(golden, n_frames=12, ifv=10)
    T, E = len(sc.dt), len(sc.n_prop)
    n, id0 = 37, 5  # ragged last CTA, global ids 5..41 (sharding: the noise is keyed by the global id)
    with BatchFilter(n, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_tuning(fpc)
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas, seed=SEED, filter_id0=id0, imu_noise_std=IMU_STD,
               cam_noise_std=CAM_STD, noise_free_filter0=True)
        xg, Pg, ug, Rg, st = bf.get_state()
    assert np.all(st == 0)
    zi = noise_samples(SEED, id0, n, 0, T, NOISE_IMU)
    zc = noise_samples(SEED, id0, n, 0, E, NOISE_CAM)
    worst_s = worst_P = 0.0
    for i in (0, 1, 17, n - 1):
        oa, cam, notch = _noisy_streams(sc, zi[i], zc[i])
        kf = sc.new_oracle()
        k = 0
        for e in range(E):
            for _ in range(sc.n_prop[e]):
                kf.propagate(sc.dt[k], oa[k, :3], oa[k, 3:])
                k += 1
            assert kf.update(cam[e, :3], cam[e, 3:], notch[e]) is not None
        xr, Pr, ur, Rr = kf.get_vectors()
        worst_s = max(worst_s, state_err(xg[i], xr), np.abs(ug[i] - ur).max(), np.abs(Rg[i] - Rr).max())
        worst_P = max(worst_P, cov_err(Pg[i], Pr, sc.Rd))
    # free-running over 110 noisy steps: the horizon-dependent tolerance of DESIGN.md section 2 (B)
    print(f"MEASURED Monte-Carlo run vs oracle on the same noise (110 free-running steps): state {worst_s:.2e} P {worst_P:.2e}")
    assert worst_s < 5e-11 and worst_P < 5e-12, (worst_s, worst_P)  # measured 1.4e-11 / 5.5e-13
    # the filters really saw different inputs
    assert np.abs(xg[1] - xg[2]).max() > 1e-6


def test_noise_free_filter0_and_sharding(golden):
    """global filter 0 stays noise free; a shard starting at global id g reproduces filters g.. of the full batch"""
    from dvi_ekf_b200 import BatchFilter

    sc = mandala_scenario(golden, n_frames=6, ifv=10)

    def run(n, id0):
        with BatchFilter(n, **model_kwargs(sc.cfg)) as bf:
            bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
            bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
            bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas, seed=SEED, filter_id0=id0, imu_noise_std=IMU_STD,
                   cam_noise_std=CAM_STD, noise_free_filter0=True)
            return bf.get_state()

    full = run(64, 0)
    shard = run(32, 32)
    assert np.array_equal(full[0][32:], shard[0]) and np.array_equal(full[1][32:], shard[1])
    with BatchFilter(1, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas)
        clean = bf.get_state()
    assert np.array_equal(full[0][0], clean[0][0]) and np.array_equal(full[1][0], clean[1][0])


def test_noise_generator_statistics():
    """Philox4x32-10 + single-precision Box-Muller: mean, variance, kurtosis, independence across draws / filters /
    steps, and a Kolmogorov-Smirnov distance against the normal CDF on 4e6 samples."""
    from math import erf, sqrt

    from dvi_ekf_b200.engine import NOISE_CAM, NOISE_IMU, noise_samples

    z = noise_samples(99, 0, 1000, 0, 500, NOISE_IMU)  # [1000, 500, 8]
    flat = z.reshape(-1)
    n = flat.size
    assert abs(flat.mean()) < 5.0 / sqrt(n)
    assert abs(flat.var() - 1.0) < 5.0 * sqrt(2.0 / n)
    assert abs((flat ** 4).mean() - 3.0) < 5.0 * sqrt(96.0 / n)
    assert np.abs(flat).max() < 6.8  # 32-bit uniforms: |z| <= sqrt(2 ln 2^33)
    c = np.corrcoef(z.reshape(-1, 8).T)
    assert np.abs(c - np.eye(8)).max() < 5.0 / sqrt(n / 8)
    assert abs(np.corrcoef(z[:, :-1, 0].ravel(), z[:, 1:, 0].ravel())[0, 1]) < 5.0 / sqrt(n / 8)  # consecutive steps
    assert abs(np.corrcoef(z[:-1, :, 0].ravel(), z[1:, :, 0].ravel())[0, 1]) < 5.0 / sqrt(n / 8)  # neighbouring filters
    xs = np.sort(flat[::7])
    cdf = 0.5 * (1.0 + np.vectorize(erf)(xs / sqrt(2.0)))
    ks = np.abs(cdf - (np.arange(xs.size) + 0.5) / xs.size).max()
    assert ks < 1.95 / sqrt(xs.size)  # alpha ~ 0.001
    # different streams / seeds are different sequences
    assert not np.array_equal(z[:4, :4], noise_samples(99, 0, 4, 0, 4, NOISE_CAM))
    assert not np.array_equal(z[:4, :4], noise_samples(100, 0, 4, 0, 4, NOISE_IMU))
    assert np.array_equal(z[3:7, 10:14], noise_samples(99, 3, 4, 10, 4, NOISE_IMU))


def test_bench_sized_run_is_deterministic_and_shape_independent(golden):
    """4096 noisy filters over the whole trajectory (the launch bench.py times), three times and with two CTA shapes: every
    state and covariance bit-identical -- a data race in the role pipelines (mbarrier record slots, named barriers, the
    shared exchange records) would show up as run-to-run differences.  Also P stays symmetric and positive on the diagonal."""
    from dvi_ekf_b200 import BatchFilter

    sc = mandala_scenario(golden, n_frames=140, ifv=10)
    n = 4096
    outs = []
    for fpc in (28, 28, 16, 28):
        with BatchFilter(n, **model_kwargs(sc.cfg)) as bf:
            bf.set_tuning(fpc)
            bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
            bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
            bf.run(sc.dt, sc.om_acc, sc.n_prop, sc.cam_meas, sc.notch_meas, seed=SEED, imu_noise_std=IMU_STD, cam_noise_std=CAM_STD)
            outs.append(bf.get_state())
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o, outs[0]))
    x, P, u, Ro, st = outs[0]
    assert np.all(st == 0) and np.all(np.isfinite(x)) and np.all(np.isfinite(P))
    asym = np.abs(P - np.swapaxes(P, 1, 2)).max(axis=(1, 2)) / np.abs(P).max(axis=(1, 2))
    # (the reference's matrix is symmetric only up to its own rounding: over 4096 noisy filters the worst asymmetry is a few
    # 1e-9 in BOTH kernels -- median 2.5e-15, 99th percentile 1e-13 -- and 3e-12 over 256 noisy filters of the oracle)
    assert asym.max() < 1e-7 and np.median(asym) < 1e-13
    assert np.all(np.diagonal(P, axis1=1, axis2=2) > 0)
    assert np.abs(np.linalg.norm(x[:, 6:10], axis=1) - 1).max() < 1e-12 and np.abs(np.linalg.norm(x[:, 22:26], axis=1) - 1).max() < 1e-12
