"""the ``with_notch: true`` flow (notch trajectory file, rotated camera as IMU source / initial state /
error reference, notch angle as the seventh measurement).  The reference cannot run it at HEAD (quirk Q12: the loader
splits a comma-separated file on white space), so there is NO reference artefact for it: parity here is the oracle's
restatement of Camera.gen_rotated / Imu.create / get_ic / Filter.run against the product's pre-pass on the CPU, and
(tests/test_gpu_notch.py) the engine against the oracle on the GPU -- "parity unpinned" by the reference, as DESIGN.md says."""
import numpy as np
import pytest

from oracle.eskf_oracle import OracleConfig, load_notch_csv, quat_to_matrix, run_reference_flow
from tests.helpers import mandala_scenario


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def test_notch_loader_reads_the_comma_separated_file(golden, tmp_path):
    from dvi_ekf_b200.camera import load_notch

    ref = golden["notch_notch90"]
    fp = tmp_path / "notch90.csv"
    with open(fp, "w") as f:  # the layout of data/trajs/notch90.csv: "a,b,c \n"
        for r in ref:
            f.write(f"{r[0]:.9f},{r[1]:.9f},{r[2]:.9f} \n")
    a = load_notch(str(fp))
    assert a.shape == (140, 3) and np.abs(a - ref).max() < 1e-9
    assert np.array_equal(load_notch(str(fp), max_vals=10, start_index=5), a[5:15])
    assert np.array_equal(load_notch_csv(str(fp), max_vals=10, start_index=5), a[5:15])
    assert np.array_equal(load_notch("notch90"), ref)  # packaged copy
    with pytest.raises(FileNotFoundError):
        load_notch("no_such_notch")


def test_rotated_camera_is_rz_of_the_notch_angle(golden):
    sc = mandala_scenario(golden, n_frames=140, ifv=1, with_notch=True)
    assert np.abs(sc.cam.notch[:, 0]).max() > 1.4  # the file sweeps to 0.9 * pi / 2
    for i in (0, 20, 63, 100, 139):
        Rr = quat_to_matrix(sc.rotated.quats[i])
        assert np.abs(Rr - _rz(sc.cam.notch[i, 0]) @ sc.cam.R[i]).max() < 1e-14
        assert sc.rotated.quats[i][3] >= 0 and abs(np.linalg.norm(sc.rotated.quats[i]) - 1) < 1e-15
    assert np.array_equal(sc.rotated.p, sc.cam.p) and np.array_equal(sc.rotated.v, sc.cam.v)
    # measurements stay those of the un-rotated camera
    assert np.array_equal(sc.cam_meas[:, 3:], sc.cam.q_raw[1:]) and np.array_equal(sc.notch_meas, sc.cam.notch[1:, 0])


@pytest.mark.parametrize("frames,ifv", [(140, 1), (60, 10), (140, 10)])
def test_product_prepass_matches_the_oracle_with_notch(golden, frames, ifv):
    from dvi_ekf_b200.camera import Camera, build_streams

    sc = mandala_scenario(golden, n_frames=frames, ifv=ifv, with_notch=True)
    a = golden["traj_mandala0_mono"][:frames]
    cam = Camera(a[:, 0], a[:, 1:4], a[:, 4:8], scale=10.0, notch=golden["notch_notch90"][:frames])
    assert cam.rotated is not None and cam.rotated.is_rotated and cam.rotated.rotated is None
    s = build_streams(cam, ifv, sc.cfg.length, sc.cfg.angle)

    def close(x, y, tol, what):
        err = np.abs(np.asarray(x) - np.asarray(y)).max() / max(np.abs(y).max(), 1e-12)
        assert err < tol, (what, err)

    assert np.array_equal(s.n_prop, sc.n_prop)
    close(s.x0, sc.x0, 1e-13, "x0")
    close(s.u0, sc.u0, 1e-11, "u0")
    close(s.dt, sc.dt, 1e-15, "dt")
    close(s.om_acc[:, :3], sc.om_acc[:, :3], 1e-11, "om")
    close(s.om_acc[:, 3:], sc.om_acc[:, 3:], 1e-11, "acc")
    close(s.cam, sc.cam_meas, 1e-15, "cam")
    close(s.notch, sc.notch_meas, 1e-15, "notch")
    close(s.imu_ref_rows, sc.imu_ref_rows, 1e-11, "imu_ref_rows")
    # the notch rate enters the synthetic IMU (om_p = z6 * notch_d): the streams differ from the notch-free ones
    s0 = build_streams(Camera(a[:, 0], a[:, 1:4], a[:, 4:8], scale=10.0), ifv, sc.cfg.length, sc.cfg.angle)
    assert np.abs(s.om_acc - s0.om_acc).max() > 1e-3


def test_oracle_filter_tracks_the_notch_angle(golden):
    """Sanity of the restated flow: with the notch measured every frame the estimated notch angle follows the file, and the
    estimated camera orientation follows the ROTATED camera (the error reference of Filter.calculate_update_mse)."""
    a = golden["traj_mandala0_mono"][:140]
    cfg = OracleConfig(max_vals=140, interframe_vals=10)
    notch = golden["notch_notch90"][:140]
    res = run_reference_flow(a[:, 0], a[:, 1:4], a[:, 4:8], cfg, notch=notch)
    est = res.x_final
    assert abs(est[16] - notch[-1, 0]) < 2e-3
    assert np.all(np.isfinite(res.update_mse)) and np.all(np.isfinite(res.P_final))
    res0 = run_reference_flow(a[:, 0], a[:, 1:4], a[:, 4:8], cfg)
    assert np.abs(res0.x_final[16:19]).max() < 1e-3  # notch-free flow: the notch states stay at the zero measurement
