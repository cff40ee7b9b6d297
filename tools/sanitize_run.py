"""Small engine run for compute-sanitizer (memcheck / racecheck / synccheck): a ragged batch, ragged epochs, noise on,
trace on, followed by step-wise propagate / update calls.  python tools/sanitize_run.py [variant]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from dvi_ekf_b200 import BatchFilter  # noqa: E402
from tests.helpers import mandala_scenario, model_kwargs  # noqa: E402


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    sc = mandala_scenario(golden, n_frames=6, ifv=10)
    n_prop = np.array([3, 0, 1, 17, 29], dtype=np.int32)
    assert n_prop.sum() == len(sc.dt)
    n = 37
    trace = np.zeros((n, len(sc.dt), 26))
    with BatchFilter(n, variant=variant, **model_kwargs(sc.cfg)) as bf:
        bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
        bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
        st, sm = bf.run(sc.dt, sc.om_acc, n_prop, sc.cam_meas, sc.notch_meas, cam_ref=np.zeros((5, 6)), imu_ref=np.zeros((5, 6)),
                        seed=7, imu_noise_std=[1e-4] * 3 + [1.0] * 3, cam_noise_std=[0.1] * 3 + [0.005] * 3 + [0.01], trace=trace)
        bf.propagate(sc.dt[:4], sc.om_acc[:4])
        K = bf.update(sc.cam_meas[0], sc.notch_meas[0], want_gain=True)
        x = bf.get_state()[0]
    print("ok", float(sm[11]), bool(np.all(np.isfinite(x))), bool(np.all(np.isfinite(K))))


if __name__ == "__main__":
    main()
