"""Race hunting without compute-sanitizer: runs the ragged-epoch case and the bench-sized case on whichever library
ESKF_B200_LIB names and prints one SHA-1 per (case, jitter seed) of everything the launch leaves behind (x, P, u_old,
R_old, status, statistics rows, reduced vector).  With the ESKF_EXP_JITTER build (dvi_ekf_b200/libeskf_b200_jit.so,
`python tools/build_variants.py jit:ESKF_EXP_JITTER`) every synchronisation point of the role pipelines is surrounded by
pseudo-random delays; a plain build ignores the seeds.  tests/test_gpu_jitter.py compares the two.

    ESKF_B200_LIB=dvi_ekf_b200/libeskf_b200_jit.so python tools/jitter_run.py --seeds 20 [--max-ns 4000]
"""
import argparse
import ctypes as C
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def set_jitter(lib, seed, max_ns):
    """True when the loaded library is a jitter build (it exports eskf_debug_set_jitter_<F> per CTA shape)."""
    ok = False
    for f in (4, 8, 16, 28):
        fn = getattr(lib, f"eskf_debug_set_jitter_{f}", None)
        if fn is not None:
            fn.argtypes = [C.c_uint, C.c_uint]
            if fn(int(seed), int(max_ns)) != 0:
                raise RuntimeError("eskf_debug_set_jitter failed")
            ok = True
    return ok


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--max-ns", type=int, default=4000)
    a = ap.parse_args()
    from dvi_ekf_b200 import BatchFilter, _lib
    from tests.helpers import mandala_scenario, model_kwargs

    lib = _lib.load()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    jit = set_jitter(lib, 0, 0)
    print(f"library {_lib.LIB_PATH} jitter_build={int(jit)}")
    imu_std = np.array([2.8e-4] * 3 + [1.24] * 3)
    cam_std = np.array([0.1, 0.1, 0.1, 0.005, 0.005, 0.005, 0.01])
    # case 1: epochs of 0, 1, 2 and many IMU samples, ragged last CTA, two CTA shapes, in-kernel noise AND pre-pass noise
    sc = mandala_scenario(golden, n_frames=8, ifv=10)
    n_prop = np.array([3, 0, 1, 17, 2, 0, 47], dtype=np.int32)
    # case 2: the launch bench.py times (4096 noisy filters, whole trajectory, error statistics)
    sc2 = mandala_scenario(golden, n_frames=140, ifv=10)
    from dvi_ekf_b200.camera import Camera, build_streams

    tr = golden["traj_mandala0_mono"][:140]
    s2 = build_streams(Camera(tr[:, 0], tr[:, 1:4], tr[:, 4:8], scale=10.0), 10, sc2.cfg.length, sc2.cfg.angle)
    for seed in range(a.seeds):
        set_jitter(lib, seed + 1, a.max_ns)
        for fpc, budget in ((28, 1 << 30), (28, 0), (8, 1 << 30), (4, 0)):
            with BatchFilter(61, **model_kwargs(sc.cfg)) as bf:
                bf.set_tuning(fpc)
                bf.set_prepass_budget(budget)
                bf.set_noise(sc.Qd[None], sc.Rd[None], sc.sig_om[None])
                bf.set_state(sc.x0[None], sc.P0[None], sc.u0[None], None)
                st, sm = bf.run(sc.dt, sc.om_acc, n_prop, sc.cam_meas, sc.notch_meas, seed=77, imu_noise_std=imu_std,
                                cam_noise_std=cam_std)
                print(f"ragged fpc={fpc} budget={int(budget > 0)} seed={seed} {digest(*bf.get_state(), st, sm)}", flush=True)
        with BatchFilter(4096, **model_kwargs(sc2.cfg)) as bf:
            bf.set_noise(sc2.Qd[None], sc2.Rd[None], sc2.sig_om[None])
            bf.set_state(s2.x0[None], sc2.P0[None], s2.u0[None], None)
            st, sm = bf.run(s2.dt, s2.om_acc, s2.n_prop, s2.cam, s2.notch, cam_ref=s2.cam_ref, imu_ref=s2.imu_ref, seed=4321,
                            imu_noise_std=imu_std, cam_noise_std=cam_std)
            print(f"bench seed={seed} {digest(*bf.get_state(), st, sm)}", flush=True)


if __name__ == "__main__":
    main()
