"""GPU pre-pass (include/eskf.h eskf_prepass): camera trajectory -> the streams
``BatchFilter.run`` consumes, computed on the device and left there (CUDA tensors), so that many long trajectories
(BASELINE config 4: every trajectory of data/trajs at 1 kHz IMU / 30 Hz camera) never round-trip through numpy.
The numpy pre-pass ``dvi_ekf_b200.camera.build_streams`` is the same arithmetic on the host."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from . import probe as _probe


@dataclass
class DeviceStreams:
    """CUDA tensors in the layout of include/eskf.h (first ``n_steps`` rows of the per-step arrays are valid)."""

    n_steps: int
    x0: "object"
    u0: "object"
    dt: "object"
    om_acc: "object"
    t_imu: "object"
    n_prop: "object"
    cam: "object"
    notch: "object"
    cam_ref: "object"
    imu_ref: "object"
    imu_ref_rows: "object"


def build_streams_gpu(t, xyz, q_xyzw, interframe_vals: int, length: float, angle: float, scale: float = 1.0,
                      gt_dofs=_probe.GT_IMU_DOFS, ic_dofs=None, notch3=None, euler_mode: str = "xyz", device: int = 0) -> DeviceStreams:
    import torch

    lib = _lib.load()
    t = np.ascontiguousarray(t, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    q = np.ascontiguousarray(q_xyzw, dtype=np.float64)
    n = len(t)
    if xyz.shape != (n, 3) or q.shape != (n, 4):
        raise ValueError("xyz must be [n,3] and q_xyzw [n,4]")
    n3 = None if notch3 is None else np.ascontiguousarray(notch3, dtype=np.float64)
    T, E = (n - 1) * int(interframe_vals), n - 1
    dev = torch.device("cuda", device)
    z = lambda *shape, dtype=torch.float64: torch.zeros(shape, dtype=dtype, device=dev)
    out = dict(x0=z(26), u0=z(6), dt=z(T), om_acc=z(T, 6), t_imu=z(T), n_prop=z(E, dtype=torch.int32), cam=z(E, 7), notch=z(E),
               cam_ref=z(E, 6), imu_ref=z(E, 6), imu_ref_rows=z(T, 14))
    gt = np.asarray(gt_dofs, dtype=float)
    ic = gt if ic_dofs is None else np.asarray(ic_dofs, dtype=float)
    pin = _lib.EskfPrepassIn()
    pin.n_frames, pin.interframe_vals = n, int(interframe_vals)
    pin.euler_mode = 0 if euler_mode == "xyz" else 1
    pin.scale = float(scale)
    pin.gt_dofs = (C.c_double * 6)(*gt)
    pin.ic_dofs = (C.c_double * 6)(*ic)
    pin.t, pin.xyz, pin.q_xyzw = t.ctypes.data, xyz.ctypes.data, q.ctypes.data
    pin.notch3 = None if n3 is None else n3.ctypes.data
    pout = _lib.EskfPrepassOut()
    for k, v in out.items():
        setattr(pout, k, v.data_ptr())
    model = _lib.EskfModel(float(length), float(angle), 0, 0)
    nsteps = C.c_int64(0)
    rc = lib.eskf_prepass(int(device), None, C.byref(model), C.byref(pin), C.byref(pout), C.byref(nsteps))
    if rc != 0:
        raise _lib.EskfError(f"eskf_prepass failed (code {rc}): {lib.eskf_prepass_last_error().decode(errors='replace')}")
    return DeviceStreams(n_steps=int(nsteps.value), **out)
