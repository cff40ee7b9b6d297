"""A/B timing of the kernel variants on the bench workload (CUDA events, one launch per pass).

    python tools/variant_bench.py [--n 4096,32768] [--variants 1,2] [--fpc 0] [--reps 3]
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
    ap.add_argument("--n", default="4096,32768")
    ap.add_argument("--variants", default="1,2")
    ap.add_argument("--fpc", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--stats", action="store_true", help="also time the launch with the error statistics on (what bench.py times)")
    ap.add_argument("--digest", action="store_true", help="print a SHA-1 of the final (x, P, status) of the last timed launch: builds "
                    "that only re-schedule the same arithmetic must agree bit for bit")
    a = ap.parse_args()
    wl = Workload()
    s = wl.s
    dev = torch.device("cuda")
    t = lambda x, dt=torch.float64: torch.tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    d = dict(dt=t(s.dt), oa=t(s.om_acc), npr=t(s.n_prop, torch.int32), cam=t(s.cam), notch=t(s.notch), cam_ref=t(s.cam_ref),
             imu_ref=t(s.imu_ref))
    for N in [int(v) for v in a.n.split(",")]:
        x0, P0, u0 = t(mc_initial_states(s.x0, N, 0)), t(wl.P0[None]), t(s.u0[None])
        for var in [int(v) for v in a.variants.split(",")]:
            for noise, stats in ((False, False), (True, False)) + (((True, True),) if a.stats else ()):
                bf = BatchFilter(N, variant=var, **wl.model)
                bf.set_tuning(a.fpc)
                bf.set_noise(wl.Qd[None], wl.Rd[None], wl.sig_om[None])
                best = 1e9
                for rep in range(a.reps):
                    bf.set_state(x0, P0, u0, None)
                    bf.sync()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    kw = dict(imu_noise_std=wl.imu_std, cam_noise_std=wl.cam_std, seed=SEED) if noise else {}
                    e0.record()
                    if stats:
                        ret = bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], cam_ref=d["cam_ref"], imu_ref=d["imu_ref"],
                                     stats_on_device=True, **kw)
                    else:
                        bf.run(d["dt"], d["oa"], d["npr"], d["cam"], d["notch"], want_stats=False, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                dg = ""
                if a.digest:
                    import hashlib

                    xs, Ps, _, _, sts = bf.get_state()
                    dg = " digest=" + hashlib.sha1(xs.tobytes() + Ps.tobytes() + sts.tobytes()).hexdigest()[:12]
                    dg += f" nonfinite={int((~np.isfinite(xs)).any(1).sum())}"
                    if stats:
                        dg += " stats=" + hashlib.sha1(ret[0].cpu().numpy().tobytes()).hexdigest()[:12]
                        dg += " sum78=%.10e,%.10e" % tuple(ret[1].cpu().numpy()[7:9])
                print(f"variant={var} N={N} noise={noise} stats={stats}: {best:.3f} ms -> {N * len(s.dt) / best * 1e3:.3e} filter-steps/s{dg}",
                      flush=True)
                bf.close()


if __name__ == "__main__":
    main()
