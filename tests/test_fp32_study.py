"""Why there is no FP32 mode (north star: "a stated end-of-trajectory tolerance for an optional FP32 mode"): a numpy study of
what single precision does to THIS filter, on the oracle (test infrastructure), noise free, BASELINE config 2's trajectory.

Three variants of the covariance path -- the nominal state, the Jacobians and the residual stay FP64 in all of them:
  store32   P rounded to float32 after every propagate / update (FP64 arithmetic): what an FP32 register tile would hold
  arith32   every product of the covariance algebra (Fx P Fx^T + Fi Q Fi^T, S, inv(S), K, Joseph form, reset) in float32
  scaled32  arith32 on the unit-scaled covariance D^-1 P D^-1, D = sqrt(diag(P0)) (per-state unit scaling)
The numbers printed here are quoted in DESIGN.md section 8.  The asserts only pin the qualitative outcome: single precision
leaves the 1e-9 contract by five to seven orders of magnitude at the FIRST update (prior >> R: 1 - K cancels at 1e-6 .. 1e-12,
float32 resolves 6e-8), and the error grows along the trajectory because the reference's filter amplifies perturbations
(DESIGN.md section 2)."""
import numpy as np

from oracle.batch_oracle import BatchOracle, _about_axis, _markley, _q2R, _qmul, _qnorm, _skew
from tests.helpers import cov_err, mandala_scenario, state_err


class ReducedPrecisionOracle(BatchOracle):
    def __init__(self, *a, mode="arith32", **k):
        super().__init__(*a, **k)
        self.mode = mode
        self.d = np.sqrt(np.diag(self.P[0])) if mode == "scaled32" else np.ones(24)
        self.D, self.Di = np.diag(self.d), np.diag(1.0 / self.d)

    def _f32(self, a):
        return np.asarray(a, dtype=np.float32)

    def _round_P(self):
        self.P = (self.Di @ self.P @ self.Di).astype(np.float32).astype(np.float64)
        self.P = self.D @ self.P @ self.D

    def propagate(self, dt, om, acc):
        if self.mode == "store32":
            super().propagate(dt, om, acc)
            self._round_P()
            return
        P0 = self.P.copy()
        cap = {}
        orig = np.swapaxes

        def spy(a, i, j):  # the parent forms Fx P Fx^T + Fi Q Fi^T with two swapaxes calls: first Fx, then Fi
            if a.shape[-2:] == (24, 24) and "Fx" not in cap and i == 1 and j == 2 and a.ndim == 3 and np.allclose(a[:, 9, 9], 1.0):
                cap["Fx"] = a
            elif a.shape[-2:] == (24, 13):
                cap["Fi"] = a
            return orig(a, i, j)

        import oracle.batch_oracle as bo

        bo.np.swapaxes = spy
        try:
            super().propagate(dt, om, acc)
        finally:
            bo.np.swapaxes = orig
        Fx, Fi = cap["Fx"], cap["Fi"]
        f = self._f32
        Fs, Gs = f(self.Di @ Fx @ self.D), f(self.Di @ Fi)  # (scaled: Fx_s = D^-1 Fx D, Fi_s = D^-1 Fi)
        Ps = f(self.Di @ P0 @ self.Di)
        Pn = Fs @ Ps @ orig(Fs, 1, 2) + (Gs * f(self.Qd)[:, None, :]) @ orig(Gs, 1, 2)
        self.P = self.D @ Pn.astype(np.float64) @ self.D

    def update(self, cam_pos, cam_q, notch):
        if self.mode == "store32":
            K = super().update(cam_pos, cam_q, notch)
            self._round_P()
            return K
        n, x, H = self.n, self.x, self.H
        f = self._f32
        cam_pos = np.broadcast_to(np.asarray(cam_pos, dtype=float), (n, 3))
        cam_q = np.broadcast_to(np.asarray(cam_q, dtype=float), (n, 4))
        notch = np.broadcast_to(np.asarray(notch, dtype=float), (n,))
        Ps = f(self.Di @ self.P @ self.Di)
        Hs = f(H @ self.D)  # measurement of the scaled error state; its noise stays R
        Rm = np.zeros((n, 7, 7), dtype=np.float32)
        Rm[:, np.arange(7), np.arange(7)] = f(self.Rd)
        S = Hs @ Ps @ Hs.T + Rm
        Ks = Ps @ Hs.T @ np.linalg.inv(S)  # float32 LAPACK
        K = self.D @ Ks.astype(np.float64)
        nq = np.stack([np.zeros(n), np.zeros(n), np.sin(notch / 2), np.cos(notch / 2)], -1)
        err_q = _qmul(_qmul(nq, cam_q) * np.array([-1.0, -1.0, -1.0, 1.0]), x[:, 22:26])
        nv = np.sqrt(np.sum(err_q[:, :3] ** 2, -1))
        ang = np.arcsin(nv)
        fac = np.where(ang == 0.0, 0.0, ang / np.where(nv > 0, nv, 1.0))
        res = np.concatenate([cam_pos - x[:, 19:22], err_q[:, :3] * fac[:, None], (notch - x[:, 16])[:, None]], -1)
        d = np.einsum("nij,nj->ni", K, res)
        th, thc = d[:, 6:9], d[:, 21:24]
        dq = _about_axis(np.sqrt(np.sum(th * th, -1)), th)
        dqc = _about_axis(np.sqrt(np.sum(thc * thc, -1)), th)
        dd = d[:, 9:15].copy()
        dd[:, self.frozen] = 0.0
        self.x = np.concatenate([x[:, 0:3] + d[:, 0:3], x[:, 3:6] + d[:, 3:6], _qmul(x[:, 6:10], dq), x[:, 10:16] + dd,
                                 x[:, 16:19] + d[:, 15:18], x[:, 19:22] + d[:, 18:21], _qmul(x[:, 22:26], dqc)], -1)
        M = np.eye(24, dtype=np.float32) - Ks @ Hs
        Pn = M @ Ps @ np.swapaxes(M, 1, 2) + (Ks * f(self.Rd)[:, None, :]) @ np.swapaxes(Ks, 1, 2)
        G = np.broadcast_to(np.eye(24), (n, 24, 24)).copy()
        G[:, 6:9, 6:9] = np.eye(3) - _skew(0.5 * th)
        G[:, 21:24, 21:24] = np.eye(3) - _skew(0.5 * thc)
        Gs = f(self.Di @ G @ self.D)
        Pn = Gs @ Pn @ np.swapaxes(Gs, 1, 2)
        self.P = self.D @ Pn.astype(np.float64) @ self.D
        return K


def test_single_precision_covariance_leaves_the_contract_at_the_first_update(golden):
    sc = mandala_scenario(golden, n_frames=140, ifv=10)
    ref = BatchOracle(sc.cfg, sc.x0[None], sc.P0, sc.u0)
    alt = {m: ReducedPrecisionOracle(sc.cfg, sc.x0[None], sc.P0, sc.u0, mode=m) for m in ("store32", "arith32", "scaled32")}
    marks = (0, 9, 39, 138)
    out = {m: [] for m in alt}
    k = 0
    for e in range(len(sc.n_prop)):
        for _ in range(sc.n_prop[e]):
            for o in (ref, *alt.values()):
                o.propagate(sc.dt[k], sc.om_acc[k, :3], sc.om_acc[k, 3:])
            k += 1
        for o in (ref, *alt.values()):
            o.update(sc.cam_meas[e, :3], sc.cam_meas[e, 3:], sc.notch_meas[e])
        if e in marks:
            for m, o in alt.items():
                ok = np.isfinite(o.x).all() and np.isfinite(o.P).all()
                out[m].append((state_err(o.x[0], ref.x[0]) if ok else np.inf, cov_err(o.P[0], ref.P[0], sc.Rd) if ok else np.inf))
    print("\nsingle-precision covariance vs the FP64 oracle (state / covariance error) after update 1, 10, 40, 139:")
    for m, rows in out.items():
        print(f"  {m:9s} " + "   ".join(f"{s:.1e} / {p:.1e}" for s, p in rows))
    for m, rows in out.items():
        assert rows[0][0] > 1e-5, (m, rows[0])  # four orders of magnitude outside the 1e-9 contract at the FIRST update
        assert rows[-1][0] > 1e-3, (m, rows[-1])  # and nowhere near a calibration-grade tolerance at the end
