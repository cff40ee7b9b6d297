"""Small driver for ncu: a few passes of the persistent kernel on the bench workload.

    python tools/profile_run.py [--filters 4096] [--passes 3] [--fpc 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import SEED, Workload, mc_initial_states  # noqa: E402
from dvi_ekf_b200 import BatchFilter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=4096)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--fpc", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0, help="truncate the trajectory to this many camera epochs (0 = all)")
    ap.add_argument("--no-noise", action="store_true", help="noise-free run (the Monte-Carlo generator is not executed)")
    ap.add_argument("--no-stats", action="store_true", help="no error statistics")
    a = ap.parse_args()
    wl = Workload()
    s = wl.s
    E = a.frames or len(s.n_prop)
    T = int(np.sum(s.n_prop[:E]))
    dev = torch.device("cuda", 0)
    t = lambda x, dt=torch.float64: torch.tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    d = dict(dt=t(s.dt[:T]), oa=t(s.om_acc[:T]), npr=t(s.n_prop[:E], torch.int32), cam=t(s.cam[:E]), notch=t(s.notch[:E]),
             cam_ref=t(s.cam_ref[:E]), imu_ref=t(s.imu_ref[:E]), x0=t(mc_initial_states(s.x0, a.filters, 0)),
             P0=t(wl.P0[None]), u0=t(s.u0[None]))
    bf = BatchFilter(a.filters, **wl.model)
    bf.set_tuning(a.fpc)
    bf.set_noise(wl.Qd[None], wl.Rd[None], wl.sig_om[None])
    ms = []
    for _ in range(a.passes):
        bf.set_state(d["x0"], d["P0"], d["u0"], None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kw = {} if a.no_noise else dict(seed=SEED, imu_noise_std=wl.imu_std, cam_noise_std=wl.cam_std)
        if a.no_stats:
            bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], want_stats=False, **kw)
        else:
            bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], cam_ref=d["cam_ref"], imu_ref=d["imu_ref"],
                   stats_on_device=True, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"filters={a.filters} T={T} E={E} ms/pass={ms} -> {a.filters * T / min(ms) * 1e3:.3e} filter-steps/s")


if __name__ == "__main__":
    main()
