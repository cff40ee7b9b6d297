"""Packs the reference's numeric artefacts for the hot path into one fixture.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

Inputs (all under /root/reference/data/trajs, read-only):
  * camera trajectories (t x y z qx qy qz qw): mandala0_mono.csv and the
    space-separated *.txt files without header;
  * kf_best_mandala0_mono.txt / imu_ref_mandala0_mono.txt -- the only numeric
    pins of Filter.propagate/update (FilterTraj / ImuRefTraj rows written by
    Filter.save, dvi_ekf/filter/Filter.py:457-465);
  * the legacy imu_ref_mandala0_mono_upd_Kp*_Km1.000.txt files, which pin the
    interframe>1 interpolation / IMU-reference pose path;
  * notch90.csv, the notch trajectory config.yaml names (an INPUT of the
    with_notch flow; the reference holds no output for it).
Output: tests/golden/reference_golden.npz (float64 arrays, compressed).
"""
import os
import sys

import numpy as np

REF = "/root/reference/data/trajs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: fixtures can only be regenerated in the build container")
    out = {}
    out["traj_mandala0_mono"] = np.loadtxt(os.path.join(REF, "mandala0_mono.csv"), delimiter=",", skiprows=1)
    for name in ["mandala0_gt", "trans_x", "trans_y", "trans_z", "rot_x", "rot_y", "rot_z", "from_prop"]:
        out[f"traj_{name}"] = np.loadtxt(os.path.join(REF, f"{name}.txt"))
    out["kf_best_mandala0_mono"] = np.loadtxt(os.path.join(REF, "kf_best_mandala0_mono.txt"))
    out["imu_ref_mandala0_mono"] = np.loadtxt(os.path.join(REF, "imu_ref_mandala0_mono.txt"))
    # notch trajectory of config.yaml (notch_traj_name: notch90): "notch,notch_d,notch_dd" per camera frame, comma separated
    out["notch_notch90"] = np.loadtxt(os.path.join(REF, "notch90.csv"), delimiter=",")
    for kp in ["0.006", "0.01", "2.0", "1.0"]:
        a = np.loadtxt(os.path.join(REF, f"imu_ref_mandala0_mono_upd_Kp{kp}_Km1.000.txt"))
        out[f"imu_ref_legacy_Kp{kp}"] = a
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(f"{k:34s} {v.shape}")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
