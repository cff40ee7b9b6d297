"""Phase timing of the persistent kernel (profiling builds with -DESKF_EXP_TIMING, tools/build_variants.py):
cycles per phase and warp, averaged over the CTAs, per IMU step / per camera update.

    ESKF_B200_LIB=dvi_ekf_b200/libeskf_b200_<name>.so python tools/timing_run.py [--noise] [--stats]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import SEED, Workload, mc_initial_states  # noqa: E402
from dvi_ekf_b200 import BatchFilter, _lib  # noqa: E402

COV = {15: "tail of epoch", 0: "wait first record", 1: "pass 1 (+sync)", 2: "reload", 3: "record wait (+sync)", 4: "pass 2 + noise + release",
       5: "U0 S + inverse", 6: "barrier U0|U1", 7: "gain", 8: "barrier U1|U2", 9: "W pass", 10: "finish (Joseph + reset)"}
SCA = {0: "acquire wait", 1: "work", 2: "scalar barrier wait", 3: "CTA barriers of the update", 4: "pk wait", 5: "U0 work", 6: "U2 work (injection)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=4096)
    ap.add_argument("--noise", action="store_true")
    ap.add_argument("--stats", action="store_true")
    a = ap.parse_args()
    wl = Workload()
    s = wl.s
    dev = torch.device("cuda", 0)
    t = lambda x, dt=torch.float64: torch.tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    d = dict(dt=t(s.dt), oa=t(s.om_acc), npr=t(s.n_prop, torch.int32), cam=t(s.cam), notch=t(s.notch), cam_ref=t(s.cam_ref),
             imu_ref=t(s.imu_ref), x0=t(mc_initial_states(s.x0, a.filters, 0)), P0=t(wl.P0[None]), u0=t(s.u0[None]))
    lib = _lib.load()
    bf = BatchFilter(a.filters, **wl.model)
    bf.set_noise(wl.Qd[None], wl.Rd[None], wl.sig_om[None])
    tab = np.zeros(256 * 12 * 16, dtype=np.int64)
    kw = dict(seed=SEED, imu_noise_std=wl.imu_std, cam_noise_std=wl.cam_std) if a.noise else {}
    for rep in range(3):
        bf.set_state(d["x0"], d["P0"], d["u0"], None)
        bf.sync()
        lib.eskf_debug_timing(tab.ctypes.data_as(C.c_void_p))  # clears
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if a.stats:
            bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], cam_ref=d["cam_ref"], imu_ref=d["imu_ref"], stats_on_device=True, **kw)
        else:
            bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], want_stats=False, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    lib.eskf_debug_timing(tab.ctypes.data_as(C.c_void_p))
    T, E = len(s.dt), len(s.n_prop)
    n_cta = (a.filters + 27) // 28
    tb = tab.reshape(256, 12, 16)[: min(n_cta - 1, 256)].astype(float)  # full CTAs only
    print(f"lib={os.environ.get('ESKF_B200_LIB')} noise={a.noise} stats={a.stats}: {ms:.3f} ms = {ms * 1e-3 * 1.965e9 / T:.0f} cycles @1965 MHz per step incl. updates")
    cov = tb[:, 4:11].mean(axis=(0, 1))
    tot = cov.sum()
    print(f"covariance warps (mean of 7 x {len(tb)} warps): total {tot / 1e6:.2f} Mcycles")
    for k, name in COV.items():
        per = cov[k] / (T if k in (1, 2, 3, 4) else E)
        print(f"   {name:28s} {cov[k] / tot * 100:5.1f} %   {per:9.0f} cycles per {'step' if k in (1, 2, 3, 4) else 'update'}")
    step = sum(cov[k] for k in (1, 2, 3, 4)) / T
    upd = sum(cov[k] for k in (15, 0, 5, 6, 7, 8, 9, 10)) / E
    print(f"   => step {step:.0f} cycles, update (all non-step phases) {upd:.0f} cycles")
    # per-warp spread of the step time
    per_w = tb[:, 4:11, 1:5].sum(axis=2).mean(axis=0) / T
    print("   step cycles by covariance warp (sub-partition 0,1,2,3,0,1,2):", " ".join(f"{v:.0f}" for v in per_w))
    for w, name in enumerate(("IMU", "CAMERA", "STAGER", "JACOB")):
        r = tb[:, w].mean(axis=0)
        print(f"{name}: " + ", ".join(f"{SCA[k]} {r[k] / (T if k in (0, 1, 2, 4) else E):.0f}/{'step' if k in (0, 1, 2, 4) else 'upd'}" for k in SCA))
    bf.close()


if __name__ == "__main__":
    main()
