// Per-filter scalar math of the VI-ESKF hot path (FP64), written as
// straight-line host/device functions so that the very same arithmetic can be
// exercised on the CPU by tests/hostcheck (no GPU needed) and by the sm_100a
// kernels in eskf_kernels.cu.
//
// Reference behaviour restated here (paths relative to the reference repo):
//   Quaternion ops ............ dvi_ekf/tools/Quaternion.py:56-224
//   State / ErrorState ........ dvi_ekf/filter/state.py:11-129
//   f_predict ................. dvi_ekf/kinematics/equations.py:44-50,72-100
//   probe forward kinematics .. dvi_ekf/models/Probe.py:147-167,289-306,431-480
//   error Jacobians ........... dvi_ekf/filter/Filter.py:249-342,
//                               dvi_ekf/kinematics/symbols.py:134-200
//   update .................... dvi_ekf/filter/Filter.py:351-395
// The CasADi-generated Jacobians are re-emitted as closed forms; DESIGN.md
// derives them and tests/ checks them against the oracle's generic DH chain
// and against sympy autodiff of the literal symbolic expressions.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ESKF_HD __host__ __device__ __forceinline__
#else
#define ESKF_HD inline
#endif

namespace eskf {

// ---- sizes and layouts ---------------------------------------------------
constexpr int NX = 26;  // nominal state: p v q(xyzw) dofs notch(3) p_cam q_cam(xyzw)
constexpr int NE = 24;  // error state:  dp dv dth ddofs(6) dnotch(3) dpc dthc
constexpr int NM = 7;   // measurement:  cam pos, cam theta, notch
constexpr int NQ = 13;  // process noise: n_a n_om n_dofs(6) n_notch_acc

// measurement rows of H (Filter.py:83-85): H[0:6,18:24] = I, H[6,15] = 1
#define ESKF_HSET(m) ((m) < 6 ? 18 + (m) : 15)

// model flags
constexpr int FLAG_ZERO_FROZEN = 1;  // quirk Q7 (HEAD): Filter.py:243-245 zeroes frozen DOFs

// Per-step Jacobian blocks handed from the scalar role to the covariance role
// ("fx record", doubles):
constexpr int FX_DT = 0;
constexpr int FX_A = 1;    // 3x3  Fx[3:6,6:9]   = -dt R_old [acc_old]x
constexpr int FX_B = 10;   // 3x3  Fx[6:9,6:9]   = rot(Om_old)^T
constexpr int FX_C1 = 19;  // 3x3  Fx[18:21,6:9] = -dt R_old [w]x
constexpr int FX_C2 = 28;  // 3x6  Fx[18:21,9:15]
constexpr int FX_D = 46;   // 3x4  Fx[21:24,{9,10,11,15}]
constexpr int FX_E = 58;   // 3x3  Fx[21:24,19:22]
constexpr int FX_NP = 67;  // 3x3  Fi[18:21,3:6] = dt R_old [p]x     (only used when Q[3:6] != 0)
constexpr int FX_NT = 76;  // 3x3  Fi[21:24,3:6] = -dt R_p^T
constexpr int FX_SIZE = 85;
constexpr int FX_STRIDE = 86;  // doubles per filter per buffer

// Per-filter constant parameters kept next to P ("par record", doubles)
constexpr int PAR_QD = 0;     // 13: diag(Q)
constexpr int PAR_RD = 13;    // 7:  diag(R)
constexpr int PAR_SIGOM = 20; // 3:  gyro noise std used inside the Jacobians (quirk Q6)
constexpr int PAR_SIZE = 23;
constexpr int PAR_STRIDE = 24;

// Update scratch ("upd record", doubles)
constexpr int UP_SINV = 0;   // 49
constexpr int UP_RES = 49;   // 7
constexpr int UP_DELTA = 56; // 24
constexpr int UP_KZ = 80;    // 24x7: K with K[h_m][m] zeroed
constexpr int UP_KD = 248;   // 7:  K[h_m][m]
constexpr int UP_CD = 255;   // 24: diag(I - K H)
constexpr int UP_OK = 279;   // 1:  0 => update skipped (LinAlgError branch, Filter.py:358-361)
constexpr int UP_SIZE = 280;

struct Model {
  double L;       // scope length        (config.yaml model.length)
  double sa, ca;  // sin / cos of the camera angle (model.angle)
  int frozen_mask;  // bit i set => DOF i frozen (simulation.frozen_dofs)
  int flags;
};

// What the scalar role carries in registers for one filter
struct Nominal {
  double p[3], v[3], q[4], dofs[6], notch[3], pc[3], qc[4];
  double om_old[3], acc_old[3];
  double R_old[9];  // rot(q) as of the end of the last propagate (quirk Q8)
};

// Cached probe kinematics at the current (dofs, notch)
struct ProbeKin {
  double p[3];
  double R[9];
  double z6[3];
};

// ---- tiny linear algebra ---------------------------------------------------
ESKF_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
ESKF_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// y = M x (row-major 3x3)
ESKF_HD void mv3(const double* M, const double* x, double* y) {
  y[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
  y[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
  y[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
}
// y = M^T x
ESKF_HD void mtv3(const double* M, const double* x, double* y) {
  y[0] = M[0] * x[0] + M[3] * x[1] + M[6] * x[2];
  y[1] = M[1] * x[0] + M[4] * x[1] + M[7] * x[2];
  y[2] = M[2] * x[0] + M[5] * x[1] + M[8] * x[2];
}
// C = M [w]x   (row-major):  column j of [w]x is  w x e_j ... written out
ESKF_HD void mul_skew(const double* M, const double* w, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = M[3 * i], b = M[3 * i + 1], c = M[3 * i + 2];
    C[3 * i + 0] = b * w[2] - c * w[1];
    C[3 * i + 1] = c * w[0] - a * w[2];
    C[3 * i + 2] = a * w[1] - b * w[0];
  }
}

// ---- quaternions (xyzw) ------------------------------------------------------
// Quaternion.rot -> Rotation.from_quat(q).as_matrix() for an already unit q
// (scipy's extra re-normalisation of a unit quaternion is a <=1 ulp effect and
// is skipped; tolerance budget is 1e-9).
ESKF_HD void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2;
  R[1] = 2.0 * (xy - zw);
  R[2] = 2.0 * (xz + yw);
  R[3] = 2.0 * (xy + zw);
  R[4] = -x2 + y2 - z2 + w2;
  R[5] = 2.0 * (yz - xw);
  R[6] = 2.0 * (xz - yw);
  R[7] = 2.0 * (yz + xw);
  R[8] = -x2 - y2 + z2 + w2;
}

// Quaternion.normalise (Quaternion.py:195-206): divide by the norm, force w >= 0.
ESKF_HD void quat_normalise(double* q) {
  const double d = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  // q / d as q * (1 / d): one division instead of four (<= 1 ulp from the reference's quotient, far
  // inside the 1e-9 budget); the sign flip is exact and folded into the factor
  const double r = ((q[3] < 0.0) ? -1.0 : 1.0) / d;
  q[0] = q[0] * r;
  q[1] = q[1] * r;
  q[2] = q[2] * r;
  q[3] = q[3] * r;
}

// Quaternion(val=M, do_normalise=True): scipy-1.10.1 Rotation.from_matrix
// (Markley's method on the RAW, generally non-orthonormal matrix -- quirk Q1)
// followed by normalisation and the w >= 0 convention.
ESKF_HD void quat_from_matrix(const double* M, double* q) {
  const double d0 = M[0], d1 = M[4], d2 = M[8];
  const double tr = d0 + d1 + d2;
  int choice = 0;
  double best = d0;
  if (d1 > best) { best = d1; choice = 1; }
  if (d2 > best) { best = d2; choice = 2; }
  if (tr > best) { choice = 3; }
  if (choice == 3) {
    q[0] = M[7] - M[5];
    q[1] = M[2] - M[6];
    q[2] = M[3] - M[1];
    q[3] = 1.0 + tr;
  } else if (choice == 0) {  // i=0 j=1 k=2
    q[0] = 1.0 - tr + 2.0 * d0;
    q[1] = M[3] + M[1];
    q[2] = M[6] + M[2];
    q[3] = M[7] - M[5];
  } else if (choice == 1) {  // i=1 j=2 k=0
    q[1] = 1.0 - tr + 2.0 * d1;
    q[2] = M[7] + M[5];
    q[0] = M[1] + M[3];
    q[3] = M[2] - M[6];
  } else {  // i=2 j=0 k=1
    q[2] = 1.0 - tr + 2.0 * d2;
    q[0] = M[2] + M[6];
    q[1] = M[5] + M[7];
    q[3] = M[3] - M[1];
  }
  quat_normalise(q);
}

// Quaternion.__mul__ (Quaternion.py:170-183): Hamilton product, re-normalised, w >= 0.
ESKF_HD void quat_mul(const double* a, const double* b, double* r) {
  const double w = a[3] * b[3] - (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
  double c[3];
  cross3(a, b, c);
  r[0] = a[3] * b[0] + b[3] * a[0] + c[0];
  r[1] = a[3] * b[1] + b[3] * a[1] + c[1];
  r[2] = a[3] * b[2] + b[3] * a[2] + c[2];
  r[3] = w;
  quat_normalise(r);
}

// Quaternion.about_axis (Quaternion.py:208-224)
ESKF_HD void quat_about_axis(double angle, const double* axis, double* q) {
  const double qlen = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  double s, c;
  sincos(0.5 * angle, &s, &c);
  double f = 1.0;
  if (qlen > 4.0 * 2.220446049250313e-16) f = s / qlen;
  q[0] = axis[0] * f;
  q[1] = axis[1] * f;
  q[2] = axis[2] * f;
  q[3] = c;
  quat_normalise(q);
}

// ---- probe kinematics -----------------------------------------------------------
// Closed form of the 8-joint DH chain (Probe.py:147-167) with q8 = 0:
//   triad  e_a(t) = (s1 s2 st - c1 ct,  s1 ct + c1 s2 st,  c2 st)
//          e_b(t) = (s1 s2 ct + c1 st,  c1 s2 ct - s1 st,  c2 ct)
//          z6     = (-s1 c2, -c1 c2, s2)                (notch joint axis)
//   p = (L - q4) z6 + q5 e_a(q3) + q6 e_b(q3)
//   R = [ -e_a(d) | sa z6 + ca e_b(d) | ca z6 - sa e_b(d) ],  d = q3 - q7
//   v = acc = 0, om = z6 q7', alp = z6 q7''
struct ProbeTrig {
  double s1, c1, s2, c2, s3, c3, sd, cd;
  double ea3[3], eb3[3], ead[3], ebd[3];
};

ESKF_HD void probe_eval(const Model& m, const double* dofs, const double* notch, ProbeKin& k, ProbeTrig& t) {
  sincos(dofs[0], &t.s1, &t.c1);
  sincos(dofs[1], &t.s2, &t.c2);
  sincos(dofs[2], &t.s3, &t.c3);
  sincos(dofs[2] - notch[0], &t.sd, &t.cd);
  const double s1s2 = t.s1 * t.s2, c1s2 = t.c1 * t.s2;
  t.ea3[0] = s1s2 * t.s3 - t.c1 * t.c3;
  t.ea3[1] = t.s1 * t.c3 + c1s2 * t.s3;
  t.ea3[2] = t.c2 * t.s3;
  t.eb3[0] = s1s2 * t.c3 + t.c1 * t.s3;
  t.eb3[1] = c1s2 * t.c3 - t.s1 * t.s3;
  t.eb3[2] = t.c2 * t.c3;
  t.ead[0] = s1s2 * t.sd - t.c1 * t.cd;
  t.ead[1] = t.s1 * t.cd + c1s2 * t.sd;
  t.ead[2] = t.c2 * t.sd;
  t.ebd[0] = s1s2 * t.cd + t.c1 * t.sd;
  t.ebd[1] = c1s2 * t.cd - t.s1 * t.sd;
  t.ebd[2] = t.c2 * t.cd;
  k.z6[0] = -t.s1 * t.c2;
  k.z6[1] = -t.c1 * t.c2;
  k.z6[2] = t.s2;
  const double lq = m.L - dofs[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    k.p[i] = lq * k.z6[i] + dofs[4] * t.ea3[i] + dofs[5] * t.eb3[i];
    k.R[3 * i + 0] = -t.ead[i];
    k.R[3 * i + 1] = m.sa * k.z6[i] + m.ca * t.ebd[i];
    k.R[3 * i + 2] = m.ca * k.z6[i] - m.sa * t.ebd[i];
  }
}

// dofs / notch part of f_predict (equations.py:87; Filter.py:243-245).  Returns true when the probe
// kinematics have to be re-evaluated (their inputs changed).
ESKF_HD bool dofs_notch_step(const Model& m, double* dofs, double* notch, double dt) {
  bool changed = false;
  const double n0 = notch[0] + dt * notch[1];
  const double n1 = notch[1] + dt * notch[2];
  changed = (n0 != notch[0]);
  notch[0] = n0;
  notch[1] = n1;
  if (m.flags & FLAG_ZERO_FROZEN) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if ((m.frozen_mask >> i) & 1) {
        changed = changed || (dofs[i] != 0.0);
        dofs[i] = 0.0;
      }
  }
  return changed;
}

// ---- propagate: nominal state + Jacobian blocks -------------------------------------
// One Filter.propagate (Filter.py:219-230) minus the covariance product.
//   s        in: pre-step state and buffers; out: post-step state and buffers
//   pk, t    in: probe kinematics at the pre-step (dofs, notch); out: at the post-step ones
//   R_WB     rot(q) of the pre-step quaternion (== s.R_old unless an update intervened)
//   fx       out: Jacobian blocks (FX_* layout)
ESKF_HD void propagate_scalar(const Model& m, Nominal& s, ProbeKin& pk, ProbeTrig& t, const double* R_WB, double dt,
                              const double* om, const double* acc, const double* sig_om, bool want_noise_jac,
                              double* fx) {
  // ---- f_predict (equations.py:44-50,72-100) with the pre-step state ----
  double om_avg[3], dth[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    om_avg[i] = (s.om_old[i] + om[i]) / 2.0;
    dth[i] = dt * om_avg[i];
  }
  double Rn[9], RS[9];
  mul_skew(R_WB, dth, RS);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rn[i] = R_WB[i] + RS[i];
  double a0[3], a1[3], acc_avg[3];
  mv3(R_WB, s.acc_old, a0);
  mv3(Rn, acc, a1);
  const double hdt2 = dt * dt / 2.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) acc_avg[i] = (a0[i] + a1[i]) / 2.0;

  // camera position: p_C + dt v + dt R_WB (v_p + om_avg x p_p), v_p = 0
  double oxp[3], Roxp[3];
  cross3(om_avg, pk.p, oxp);
  mv3(R_WB, oxp, Roxp);
  // camera rotation: R_WC + R_WC [dt R_p^T (om_old + om_p)]x, om_p = z6 * notch_d
  double omt[3], om_c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) omt[i] = s.om_old[i] + pk.z6[i] * s.notch[1];
  mtv3(pk.R, omt, om_c);
#pragma unroll
  for (int i = 0; i < 3; ++i) om_c[i] = dt * om_c[i];
  double Rc[9], RcS[9];
  quat_to_rot(s.qc, Rc);
  mul_skew(Rc, om_c, RcS);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rc[i] = Rc[i] + RcS[i];

#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double vi = s.v[i];
    s.pc[i] = s.pc[i] + dt * vi + dt * Roxp[i];
    s.p[i] = s.p[i] + dt * vi + hdt2 * acc_avg[i];
    s.v[i] = vi + dt * acc_avg[i];
  }
  const bool probe_changed = dofs_notch_step(m, s.dofs, s.notch, dt);
  // State.from_array (state.py:62-74): Markley quaternions of the first-order matrices
  quat_from_matrix(Rn, s.q);
  quat_from_matrix(Rc, s.qc);

  // ---- error Jacobians (Filter.py:249-342) with the buffered R_old / om_old / acc_old
  //      and the POST-predict dofs / notch ----
  if (probe_changed) probe_eval(m, s.dofs, s.notch, pk, t);  // (pk, t) persist: same inputs, same kinematics
  const double* Ro = s.R_old;
  fx[FX_DT] = dt;
  {  // A = (-R_old [acc_old]x) dt
    double T[9];
    mul_skew(Ro, s.acc_old, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) fx[FX_A + i] = -T[i] * dt;
  }
  {  // B = rot(normalise(quat(w=1, v=dt/2 om_old)))^T   (Filter.py:132-134,255)
    double qo[4] = {0.5 * dt * s.om_old[0], 0.5 * dt * s.om_old[1], 0.5 * dt * s.om_old[2], 1.0};
    quat_normalise(qo);
    double Rb[9];
    quat_to_rot(qo, Rb);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) fx[FX_B + 3 * i + j] = Rb[3 * j + i];
  }
  double wt[3];  // om_tr = om_old - sigma_om (noise symbols evaluated at sigma, quirk Q6)
#pragma unroll
  for (int i = 0; i < 3; ++i) wt[i] = s.om_old[i] - sig_om[i];
  {  // C1 = -dt R_old [w]x,  w = p + om_tr x p   (v_tr = p_tr, quirk Q2)
    double w[3], T[9];
    cross3(wt, pk.p, w);
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = pk.p[i] + w[i];
    mul_skew(Ro, w, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) fx[FX_C1 + i] = -dt * T[i];
  }
  {  // C2 = dt R_old (I + [om_tr]x) dp/dq(1..6)
    double Mw[9], S[9];
    mul_skew(Ro, wt, S);
#pragma unroll
    for (int i = 0; i < 9; ++i) Mw[i] = dt * (Ro[i] + S[i]);
    const double lq = m.L - s.dofs[3];
    const double dz2[3] = {t.s1 * t.s2, t.c1 * t.s2, t.c2};  // d z6 / d q2
    const double k2 = s.dofs[4] * t.s3 + s.dofs[5] * t.c3;
    double col[6][3];
    col[0][0] = pk.p[1];
    col[0][1] = -pk.p[0];
    col[0][2] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      col[1][i] = lq * dz2[i] - k2 * pk.z6[i];
      col[2][i] = s.dofs[4] * t.eb3[i] - s.dofs[5] * t.ea3[i];
      col[3][i] = -pk.z6[i];
      col[4][i] = t.ea3[i];
      col[5][i] = t.eb3[i];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double y[3];
      mv3(Mw, col[k], y);
      fx[FX_C2 + 0 * 6 + k] = y[0];
      fx[FX_C2 + 1 * 6 + k] = y[1];
      fx[FX_C2 + 2 * 6 + k] = y[2];
    }
  }
  {  // D = dt d/dq{1,2,3,7} [ R(q)^T om_tr ]   (R^T om_p is constant: no contribution)
    const double al = dot3(t.ead, wt), be = dot3(t.ebd, wt), ze = dot3(pk.z6, wt);
    const double al1 = t.ead[1] * wt[0] - t.ead[0] * wt[1];
    const double be1 = t.ebd[1] * wt[0] - t.ebd[0] * wt[1];
    const double ze1 = pk.z6[1] * wt[0] - pk.z6[0] * wt[1];
    const double al2 = -t.sd * ze, be2 = -t.cd * ze, ze2 = t.sd * al + t.cd * be;
    // rows: (-alpha, sa zeta + ca beta, ca zeta - sa beta)
    fx[FX_D + 0] = dt * (-al1);
    fx[FX_D + 4] = dt * (m.sa * ze1 + m.ca * be1);
    fx[FX_D + 8] = dt * (m.ca * ze1 - m.sa * be1);
    fx[FX_D + 1] = dt * (-al2);
    fx[FX_D + 5] = dt * (m.sa * ze2 + m.ca * be2);
    fx[FX_D + 9] = dt * (m.ca * ze2 - m.sa * be2);
    // d/dq3: alpha' = beta, beta' = -alpha, zeta' = 0 ; d/dq7 = - d/dq3
    fx[FX_D + 2] = dt * (-be);
    fx[FX_D + 6] = dt * (-m.ca * al);
    fx[FX_D + 10] = dt * (m.sa * al);
    fx[FX_D + 3] = -fx[FX_D + 2];
    fx[FX_D + 7] = -fx[FX_D + 6];
    fx[FX_D + 11] = -fx[FX_D + 10];
  }
  {  // E = I - dt/2 [a + b]x,  a = R_p^T (om_tr + om_p), b = R_p^T (om_old + om_p)
    double ua[3], ub[3], a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double omp = pk.z6[i] * s.notch[1];
      ua[i] = wt[i] + omp;
      ub[i] = s.om_old[i] + omp;
    }
    mtv3(pk.R, ua, a);
    mtv3(pk.R, ub, b);
    const double h = 0.5 * dt;
    const double e0 = h * (a[0] + b[0]), e1 = h * (a[1] + b[1]), e2 = h * (a[2] + b[2]);
    fx[FX_E + 0] = 1.0;
    fx[FX_E + 1] = e2;
    fx[FX_E + 2] = -e1;
    fx[FX_E + 3] = -e2;
    fx[FX_E + 4] = 1.0;
    fx[FX_E + 5] = e0;
    fx[FX_E + 6] = e1;
    fx[FX_E + 7] = -e0;
    fx[FX_E + 8] = 1.0;
  }
  if (want_noise_jac) {  // rows 18:24 of Fi (only matter when Q[3:6] != 0)
    double T[9];
    mul_skew(Ro, pk.p, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) fx[FX_NP + i] = dt * T[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) fx[FX_NT + 3 * i + j] = -dt * pk.R[3 * j + i];
  }

  // ---- buffers (Filter.py:224-227) ----
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.om_old[i] = om[i];
    s.acc_old[i] = acc[i];
  }
  quat_to_rot(s.q, s.R_old);
}

// ---- covariance propagation:  y = Fx x  applied in place to three 24-vectors ----
// Element i of vector v lives at x[v * VS + i * ES].  Column pass of P:
// ES = row stride, VS = 1; row pass: ES = 1, VS = row stride.
template <int ES, int VS>
ESKF_HD void fx_apply3(double* x, const double* fx) {
  const double dt = fx[FX_DT];
  double xv[3][3], xt[3][3], xd[3][6], xn[3][3], xpc[3][3], xtc[3][3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xv[v][i] = x[v * VS + (3 + i) * ES];
      xt[v][i] = x[v * VS + (6 + i) * ES];
      xn[v][i] = x[v * VS + (15 + i) * ES];
      xpc[v][i] = x[v * VS + (18 + i) * ES];
      xtc[v][i] = x[v * VS + (21 + i) * ES];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) xd[v][i] = x[v * VS + (9 + i) * ES];
  }
  // rows 18:21 (camera position error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = dt * xv[v][i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double c = fx[FX_C1 + 3 * i + k];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c * xt[v][k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double c = fx[FX_C2 + 6 * i + k];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c * xd[v][k];
    }
    // mis-aligned identity block (quirk Q3): Fx[18+i, 16+i] = 1
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      y[v] += (i == 0) ? xn[v][1] : (i == 1) ? xn[v][2] : xpc[v][0];
      x[v * VS + (18 + i) * ES] = y[v];
    }
  }
  // rows 21:24 (camera orientation error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = (i == 0) ? 0.0 : xtc[v][i];  // Fx[22,22] = Fx[23,23] = 1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double c = fx[FX_D + 4 * i + k];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c * xd[v][k];
    }
    {
      const double c = fx[FX_D + 4 * i + 3];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c * xn[v][0];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double c = fx[FX_E + 3 * i + k];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c * ((k == 0) ? xpc[v][1] : (k == 1) ? xpc[v][2] : xtc[v][0]);
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) x[v * VS + (21 + i) * ES] = y[v];
  }
  // rows 0:3  p += dt v
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) x[v * VS + i * ES] = x[v * VS + i * ES] + dt * xv[v][i];
  // rows 3:6  v += A theta ; rows 6:9 theta = B theta
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double yv[3], yt[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      yv[v] = xv[v][i];
      yt[v] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double a = fx[FX_A + 3 * i + k], b = fx[FX_B + 3 * i + k];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        yv[v] += a * xt[v][k];
        yt[v] += b * xt[v][k];
      }
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      x[v * VS + (3 + i) * ES] = yv[v];
      x[v * VS + (6 + i) * ES] = yt[v];
    }
  }
  // rows 15:17 notch chain
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    x[v * VS + 15 * ES] = xn[v][0] + dt * xn[v][1];
    x[v * VS + 16 * ES] = xn[v][1] + dt * xn[v][2];
  }
}

// Fi Q Fi^T for rows r0..r0+2 of P (row pass epilogue; Filter.py:349).
// With Q[3:6] == 0 (the reference's normal flow, quirk Q5) this is a diagonal add.
template <int ES, int VS>
ESKF_HD void add_process_noise3(double* x, int r0, const double* fx, const double* qd, bool imu_q) {
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int r = r0 + v;
    double add = 0.0;
    if (r >= 3 && r < 15) add = qd[r - 3];  // Fi[3:15,0:12] = I
    if (r == 17) add = qd[12];               // Fi[17,12] = 1
    if (add != 0.0) x[v * VS + r * ES] += add;
  }
  if (imu_q && (r0 == 6 || r0 == 18 || r0 == 21)) {
    // n_om drives theta (I), p_C (Np) and theta_C (Nt): L Q_om L^T on rows/cols {6:9,18:21,21:24}
    // the diagonal part of the theta rows was added above.
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double Lr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        Lr[k] = (r0 == 6) ? ((k == v) ? 1.0 : 0.0) : (r0 == 18) ? fx[FX_NP + 3 * v + k] : fx[FX_NT + 3 * v + k];
#pragma unroll
      for (int cb = 0; cb < 3; ++cb) {
        const int c0 = (cb == 0) ? 6 : (cb == 1) ? 18 : 21;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (r0 == 6 && cb == 0) continue;  // already added (diagonal of Q_om)
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double Lc = (cb == 0) ? ((k == j) ? 1.0 : 0.0) : (cb == 1) ? fx[FX_NP + 3 * j + k] : fx[FX_NT + 3 * j + k];
            acc += Lr[k] * qd[3 + k] * Lc;
          }
          x[v * VS + (c0 + j) * ES] += acc;
        }
      }
    }
  }
}

// ---- update: scalar part ---------------------------------------------------------------
// 7x7 inverse by LU with partial pivoting (np.linalg.inv -> LAPACK gesv, Filter.py:357).
// Fully unrolled so everything stays in registers.  Returns false if a pivot is exactly
// zero or the result is not finite (the reference's LinAlgError branch).
ESKF_HD bool inv7(const double* S, double* Sinv) {
  double a[7][7], b[7][7];
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      a[i][j] = S[7 * i + j];
      b[i][j] = (i == j) ? 1.0 : 0.0;
    }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    // pivot search: first maximum of |a[i][k]|, i >= k (idamax)
    int piv = k;
    double best = fabs(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      const double v = fabs(a[i][k]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    ok = ok && (best != 0.0);
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      if (piv == i) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          double tmp = a[k][j];
          a[k][j] = a[i][j];
          a[i][j] = tmp;
          tmp = b[k][j];
          b[k][j] = b[i][j];
          b[i][j] = tmp;
        }
      }
    }
    const double rp = 1.0 / a[k][k];
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      const double l = a[i][k] * rp;
#pragma unroll
      for (int j = k + 1; j < 7; ++j) a[i][j] -= l * a[k][j];
#pragma unroll
      for (int j = 0; j < 7; ++j) b[i][j] -= l * b[k][j];
    }
  }
  // back substitution U X = B
#pragma unroll
  for (int i = 6; i >= 0; --i) {
    const double rd = 1.0 / a[i][i];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      double v = b[i][j];
#pragma unroll
      for (int k = i + 1; k < 7; ++k) v -= a[i][k] * b[k][j];
      b[i][j] = v * rd;
    }
  }
  double chk = 0.0;
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      Sinv[7 * i + j] = b[i][j];
      chk += b[i][j] * 0.0;  // NaN / inf detector
    }
  return ok && (chk == 0.0);
}

// Residual of Filter.update (Filter.py:363-375).  cam_q is the RAW file quaternion (xyzw).
ESKF_HD bool update_residual(const Nominal& s, const double* cam_pos, const double* cam_q, double ang_notch,
                             double* res) {
  double sn, cn;
  sincos(0.5 * ang_notch, &sn, &cn);
  const double nq[4] = {0.0, 0.0, sn, cn};  // Rotation.from_euler("xyz", [0, 0, ang]).as_quat()
  double qm[4], qmc[4], e[4];
  quat_mul(nq, cam_q, qm);
  qmc[0] = -qm[0];
  qmc[1] = -qm[1];
  qmc[2] = -qm[2];
  qmc[3] = qm[3];
  quat_mul(qmc, s.qc, e);
  const double nv = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  const double ang = asin(nv);  // quirk Q9: asin|v|, not 2 acos w
  // axis = 0 iff math.isclose(angle, 0) (rel_tol 1e-9, abs_tol 0) <=> angle == 0
  const double f = (ang == 0.0) ? 0.0 : ang / nv;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    res[i] = cam_pos[i] - s.pc[i];
    res[3 + i] = f * e[i];
  }
  res[6] = ang_notch - s.notch[0];
  return nv <= 1.0;  // math.asin raises for |v| > 1
}

// state (+) error state  (state.py:46-60,105-129) incl. the dqc axis slip (quirk Q4)
ESKF_HD void inject_error(const Model& m, Nominal& s, const double* d) {
  const double* th = d + 6;
  const double* thc = d + 21;
  const double nth = sqrt(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
  const double nthc = sqrt(thc[0] * thc[0] + thc[1] * thc[1] + thc[2] * thc[2]);
  double dq[4], dqc[4], qn[4];
  quat_about_axis(nth, th, dq);
  quat_about_axis(nthc, th, dqc);  // axis = theta (IMU), not theta_c: state.py:124
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.p[i] += d[i];
    s.v[i] += d[3 + i];
    s.notch[i] += d[15 + i];
    s.pc[i] += d[18 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (!((m.frozen_mask >> i) & 1)) s.dofs[i] += d[9 + i];  // Filter.py:377-379
  quat_mul(s.q, dq, qn);
#pragma unroll
  for (int i = 0; i < 4; ++i) s.q[i] = qn[i];
  quat_mul(s.qc, dqc, qn);
#pragma unroll
  for (int i = 0; i < 4; ++i) s.qc[i] = qn[i];
}

// ---- update: covariance part ---------------------------------------------------------------
// Gain rows r0..r0+2:  K = (P H^T) inv(S);  delta = K res;  bookkeeping for Joseph.
// P is addressed as P[i * RS + j].
template <int RS>
ESKF_HD void gain_rows3(const double* P, int r0, double* up) {
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int r = r0 + v;
    double ph[7], k[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) ph[j] = P[r * RS + ESKF_HSET(j)];
    double dl = 0.0;
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) acc += ph[j] * up[UP_SINV + 7 * j + mm];
      k[mm] = acc;
      dl += acc * up[UP_RES + mm];
    }
    up[UP_DELTA + r] = dl;
    double cd = 1.0;
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) {
      const bool diag = (ESKF_HSET(mm) == r);
      up[UP_KZ + 7 * r + mm] = diag ? 0.0 : k[mm];
      if (diag) {
        up[UP_KD + mm] = k[mm];
        cd = 1.0 - k[mm];  // (I - K H)[r][r]
      }
    }
    up[UP_CD + r] = cd;
  }
}

// y = (I - K H) x on three 24-vectors, in place (Joseph factor, Filter.py:384).
template <int ES, int VS>
ESKF_HD void joseph_apply3(double* x, const double* up) {
  double xh[3][7];
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) xh[v][mm] = x[v * VS + ESKF_HSET(mm) * ES];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const double cd = up[UP_CD + i];
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = cd * x[v * VS + i * ES];
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) {
      const double k = up[UP_KZ + 7 * i + mm];
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] -= k * xh[v][mm];
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) x[v * VS + i * ES] = y[v];
  }
}

// Row pass of the update for rows r0..r0+2 (ES = 1):
//   P2 = P1 (I-KH)^T + (K R) K^T ;  P3 = G P2 G^T   (Filter.py:384-390)
// The reset matrix G is block diagonal with 3x3 blocks aligned to the row triples, so both
// G P2 (mixes the three rows held here) and (.) G^T (mixes entries 6:9 / 21:24 of a row)
// are local to the caller.
template <int VS>
ESKF_HD void joseph_rows_finish3(double* x, int r0, const double* up, const double* rd) {
  double xh[3][7], kr[3][7];
#pragma unroll
  for (int v = 0; v < 3; ++v)
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) {
      xh[v][mm] = x[v * VS + ESKF_HSET(mm)];
      const int r = r0 + v;
      const double k = (ESKF_HSET(mm) == r) ? up[UP_KD + mm] : up[UP_KZ + 7 * r + mm];
      kr[v][mm] = k * rd[mm];
    }
  // reset blocks: G = I - [delta_theta / 2]x  (rows 6:9 from delta[6:9], rows 21:24 from delta[21:24])
  const bool mix = (r0 == 6) || (r0 == 21);
  double g[3] = {0.0, 0.0, 0.0};
  if (mix) {
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = 0.5 * up[UP_DELTA + r0 + i];
  }
  const double gt[3] = {0.5 * up[UP_DELTA + 6], 0.5 * up[UP_DELTA + 7], 0.5 * up[UP_DELTA + 8]};
  const double gc[3] = {0.5 * up[UP_DELTA + 21], 0.5 * up[UP_DELTA + 22], 0.5 * up[UP_DELTA + 23]};
  double hold[3][6];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const double cd = up[UP_CD + i];
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = cd * x[v * VS + i];
    double ki[7];
#pragma unroll
    for (int mm = 0; mm < 7; ++mm) {
      const double kz = up[UP_KZ + 7 * i + mm];
      ki[mm] = kz;
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] -= kz * xh[v][mm];
    }
    // K R K^T column i: K[i][m] needs the diagonal entries back
    {
      double z[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int mm = 0; mm < 7; ++mm) {
        const double k = (ESKF_HSET(mm) == i) ? up[UP_KD + mm] : ki[mm];
#pragma unroll
        for (int v = 0; v < 3; ++v) z[v] += kr[v][mm] * k;
      }
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += z[v];
    }
    if (mix) {  // G P2: rows (I - [g]x) applied across the three rows held here
      const double y0 = y[0], y1 = y[1], y2 = y[2];
      // (I - [g]x) = [[1, g2, -g1], [-g2, 1, g0], [g1, -g0, 1]]
      y[0] = y0 + g[2] * y1 - g[1] * y2;
      y[1] = -g[2] * y0 + y1 + g[0] * y2;
      y[2] = g[1] * y0 - g[0] * y1 + y2;
    }
    const bool held = (i >= 6 && i < 9) || (i >= 21);
    if (held) {
      const int hi = (i < 9) ? i - 6 : i - 18;
#pragma unroll
      for (int v = 0; v < 3; ++v) hold[v][hi] = y[v];
    } else {
#pragma unroll
      for (int v = 0; v < 3; ++v) x[v * VS + i] = y[v];
    }
  }
  // (.) G^T on columns 6:9 and 21:24
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const double a0 = hold[v][0], a1 = hold[v][1], a2 = hold[v][2];
    x[v * VS + 6] = a0 + gt[2] * a1 - gt[1] * a2;
    x[v * VS + 7] = -gt[2] * a0 + a1 + gt[0] * a2;
    x[v * VS + 8] = gt[1] * a0 - gt[0] * a1 + a2;
    const double b0 = hold[v][3], b1 = hold[v][4], b2 = hold[v][5];
    x[v * VS + 21] = b0 + gc[2] * b1 - gc[1] * b2;
    x[v * VS + 22] = -gc[2] * b0 + b1 + gc[0] * b2;
    x[v * VS + 23] = gc[1] * b0 - gc[0] * b1 + b2;
  }
}

// =====================================================================================
// v2 kernel building blocks: the same arithmetic as propagate_scalar / fx_apply3, cut along
// the lines of the warp-specialised kernel (eskf_kernel2.cuh): IMU nominal state, camera
// nominal state, Jacobian blocks, and the covariance transform on a register-resident tile.
// tests/hostcheck replays them on the CPU against propagate_scalar / fx_apply3.
// =====================================================================================

// "fx2 record": Jacobian blocks laid out in the order the covariance role consumes them,
// every row group starting on an even index so that it is fetched with 16-byte loads.
constexpr int FX2_DT = 0;    // dt, pad
constexpr int FX2_AB = 2;    // 3 x [A(i,0..2) B(i,0..2)]                     Fx[3:6,6:9], Fx[6:9,6:9]
constexpr int FX2_R18 = 20;  // 3 x [C1(i,0..2) C2(i,0..5) pad]               Fx[18:21, 6:15]
constexpr int FX2_R21 = 50;  // 3 x [D(i,0..3) E(i,0..2) pad]                 Fx[21:24, {9,10,11,15}], Fx[21:24,19:22]
constexpr int FX2_NP = 74;   // 3x3 (+pad)  Fi[18:21,3:6]
constexpr int FX2_NT = 84;   // 3x3 (+pad)  Fi[21:24,3:6]
constexpr int FX2_SIZE = 94;
constexpr int FX2_STRIDE = 94;

// Filter._predict_nominal, IMU part (equations.py:72-86; state.py:62-69): p, v, q.
//   R_WB   rot(q) of the pre-step quaternion;  Rn_out  = rot(q+) (the new R_WB_old, Filter.py:227)
ESKF_HD void imu_nominal_step(double* p, double* v, double* q, const double* R_WB, double dt, const double* om_old,
                              const double* acc_old, const double* om, const double* acc, double* R_new) {
  double dth[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) dth[i] = dt * ((om_old[i] + om[i]) / 2.0);
  double Rn[9], RS[9];
  mul_skew(R_WB, dth, RS);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rn[i] = R_WB[i] + RS[i];
  double a0[3], a1[3];
  mv3(R_WB, acc_old, a0);
  mv3(Rn, acc, a1);
  const double hdt2 = dt * dt / 2.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double aa = (a0[i] + a1[i]) / 2.0;
    const double vi = v[i];
    p[i] = p[i] + dt * vi + hdt2 * aa;
    v[i] = vi + dt * aa;
  }
  quat_from_matrix(Rn, q);
  quat_to_rot(q, R_new);
}

// Filter._predict_nominal, camera part (equations.py:88-98; state.py:70-74): p_cam, q_cam.
//   v_pre, R_WB   pre-step IMU velocity and rotation;  pk/notch_d at the pre-step (dofs, notch)
ESKF_HD void cam_nominal_step(double* pc, double* qc, const double* v_pre, const double* R_WB, double dt,
                              const double* om_old, const double* om, const double* pk_p, const double* pk_R,
                              const double* pk_z6, double notch_d) {
  double om_avg[3], oxp[3], Roxp[3], omt[3], om_c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) om_avg[i] = (om_old[i] + om[i]) / 2.0;
  cross3(om_avg, pk_p, oxp);
  mv3(R_WB, oxp, Roxp);
#pragma unroll
  for (int i = 0; i < 3; ++i) omt[i] = om_old[i] + pk_z6[i] * notch_d;
  mtv3(pk_R, omt, om_c);
#pragma unroll
  for (int i = 0; i < 3; ++i) om_c[i] = dt * om_c[i];
  double Rc[9], RcS[9];
  quat_to_rot(qc, Rc);
  mul_skew(Rc, om_c, RcS);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rc[i] = Rc[i] + RcS[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc[i] = pc[i] + dt * v_pre[i] + dt * Roxp[i];
  quat_from_matrix(Rc, qc);
}

// Filter._predict_error (Filter.py:249-342): Jacobian blocks in the fx2 layout, from the buffered
// R_old / om_old / acc_old and the POST-predict (dofs, notch) whose probe kinematics are (pk, t).
ESKF_HD void jacobian_blocks(const Model& m, const double* dofs, double notch_d, const ProbeKin& pk, const ProbeTrig& t,
                             const double* Ro, double dt, const double* om_old, const double* acc_old,
                             const double* sig_om, bool want_noise_jac, double* fx) {
  fx[FX2_DT] = dt;
  {  // A = (-R_old [acc_old]x) dt
    double T[9];
    mul_skew(Ro, acc_old, T);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) fx[FX2_AB + 6 * i + k] = -T[3 * i + k] * dt;
  }
  {  // B = rot(normalise(quat(w=1, v=dt/2 om_old)))^T
    double qo[4] = {0.5 * dt * om_old[0], 0.5 * dt * om_old[1], 0.5 * dt * om_old[2], 1.0};
    quat_normalise(qo);
    double Rb[9];
    quat_to_rot(qo, Rb);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) fx[FX2_AB + 6 * i + 3 + j] = Rb[3 * j + i];
  }
  double wt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) wt[i] = om_old[i] - sig_om[i];
  {  // C1 = -dt R_old [w]x,  w = p + om_tr x p
    double w[3], T[9];
    cross3(wt, pk.p, w);
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = pk.p[i] + w[i];
    mul_skew(Ro, w, T);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) fx[FX2_R18 + 10 * i + k] = -dt * T[3 * i + k];
  }
  {  // C2 = dt R_old (I + [om_tr]x) dp/dq(1..6)
    double Mw[9], S[9];
    mul_skew(Ro, wt, S);
#pragma unroll
    for (int i = 0; i < 9; ++i) Mw[i] = dt * (Ro[i] + S[i]);
    const double lq = m.L - dofs[3];
    const double dz2[3] = {t.s1 * t.s2, t.c1 * t.s2, t.c2};
    const double k2 = dofs[4] * t.s3 + dofs[5] * t.c3;
    double col[6][3];
    col[0][0] = pk.p[1];
    col[0][1] = -pk.p[0];
    col[0][2] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      col[1][i] = lq * dz2[i] - k2 * pk.z6[i];
      col[2][i] = dofs[4] * t.eb3[i] - dofs[5] * t.ea3[i];
      col[3][i] = -pk.z6[i];
      col[4][i] = t.ea3[i];
      col[5][i] = t.eb3[i];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double y[3];
      mv3(Mw, col[k], y);
      fx[FX2_R18 + 0 * 10 + 3 + k] = y[0];
      fx[FX2_R18 + 1 * 10 + 3 + k] = y[1];
      fx[FX2_R18 + 2 * 10 + 3 + k] = y[2];
    }
  }
  {  // D = dt d/dq{1,2,3,7} [ R(q)^T om_tr ]
    const double al = dot3(t.ead, wt), be = dot3(t.ebd, wt), ze = dot3(pk.z6, wt);
    const double al1 = t.ead[1] * wt[0] - t.ead[0] * wt[1];
    const double be1 = t.ebd[1] * wt[0] - t.ebd[0] * wt[1];
    const double ze1 = pk.z6[1] * wt[0] - pk.z6[0] * wt[1];
    const double al2 = -t.sd * ze, be2 = -t.cd * ze, ze2 = t.sd * al + t.cd * be;
    double* D0 = fx + FX2_R21;
    double* D1 = fx + FX2_R21 + 8;
    double* D2 = fx + FX2_R21 + 16;
    D0[0] = dt * (-al1);
    D1[0] = dt * (m.sa * ze1 + m.ca * be1);
    D2[0] = dt * (m.ca * ze1 - m.sa * be1);
    D0[1] = dt * (-al2);
    D1[1] = dt * (m.sa * ze2 + m.ca * be2);
    D2[1] = dt * (m.ca * ze2 - m.sa * be2);
    D0[2] = dt * (-be);
    D1[2] = dt * (-m.ca * al);
    D2[2] = dt * (m.sa * al);
    D0[3] = -D0[2];
    D1[3] = -D1[2];
    D2[3] = -D2[2];
  }
  {  // E = I - dt/2 [a + b]x
    double ua[3], ub[3], a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double omp = pk.z6[i] * notch_d;
      ua[i] = wt[i] + omp;
      ub[i] = om_old[i] + omp;
    }
    mtv3(pk.R, ua, a);
    mtv3(pk.R, ub, b);
    const double h = 0.5 * dt;
    const double e0 = h * (a[0] + b[0]), e1 = h * (a[1] + b[1]), e2 = h * (a[2] + b[2]);
    double* E0 = fx + FX2_R21 + 4;
    double* E1 = fx + FX2_R21 + 8 + 4;
    double* E2 = fx + FX2_R21 + 16 + 4;
    E0[0] = 1.0;
    E0[1] = e2;
    E0[2] = -e1;
    E1[0] = -e2;
    E1[1] = 1.0;
    E1[2] = e0;
    E2[0] = e1;
    E2[1] = -e0;
    E2[2] = 1.0;
  }
  if (want_noise_jac) {
    double T[9];
    mul_skew(Ro, pk.p, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) fx[FX2_NP + i] = dt * T[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) fx[FX2_NT + 3 * i + j] = -dt * pk.R[3 * j + i];
  }
}

// X <- Fx X for three 24-vectors held in registers (X[i][v] = element i of vector v): the sparse
// transition matrix of Filter._predict_error applied as straight-line code.  Both covariance passes
// of a step use this one function (column tile of P, then row tile of Fx P), see eskf_kernel2.cuh.
// Coefficients come from the fx2 record through 16-byte loads.
struct alignas(16) d2 {
  double x, y;
};
// PS = stride between consecutive coefficient pairs of one record (1: plain array; F: the
// [pair][filter] layout of the v2 kernel, conflict free for writer and readers).
template <int PS>
ESKF_HD double fx2_at(const d2* f2, int j) {
  const d2 v = f2[(j >> 1) * PS];
  return (j & 1) ? v.y : v.x;
}
template <int PS>
ESKF_HD void fx_apply_reg(double (&X)[24][3], const d2* fx2) {
  const d2* f2 = fx2;
  const double dt = f2[(FX2_DT / 2) * PS].x;
  double y18[3][3], y21[3][3];
  // rows 18:21 (camera position error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const d2 c0 = f2[((FX2_R18 + 10 * i) / 2 + 0) * PS], c1 = f2[((FX2_R18 + 10 * i) / 2 + 1) * PS], c2 = f2[((FX2_R18 + 10 * i) / 2 + 2) * PS],
             c3 = f2[((FX2_R18 + 10 * i) / 2 + 3) * PS], c4 = f2[((FX2_R18 + 10 * i) / 2 + 4) * PS];
    const double c[9] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y, c4.x};
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double y = dt * X[3 + i][v];
#pragma unroll
      for (int k = 0; k < 9; ++k) y += c[k] * X[6 + k][v];  // C1 on theta (6:9), C2 on dofs (9:15)
      y18[i][v] = y + X[16 + i][v];                          // mis-aligned identity block (quirk Q3)
    }
  }
  // rows 21:24 (camera orientation error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const d2 c0 = f2[((FX2_R21 + 8 * i) / 2 + 0) * PS], c1 = f2[((FX2_R21 + 8 * i) / 2 + 1) * PS], c2 = f2[((FX2_R21 + 8 * i) / 2 + 2) * PS],
             c3 = f2[((FX2_R21 + 8 * i) / 2 + 3) * PS];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double y = (i == 0) ? 0.0 : X[21 + i][v];  // Fx[22,22] = Fx[23,23] = 1
      y += c0.x * X[9][v];
      y += c0.y * X[10][v];
      y += c1.x * X[11][v];
      y += c1.y * X[15][v];
      y += c2.x * X[19][v];
      y += c2.y * X[20][v];
      y += c3.x * X[21][v];
      y21[i][v] = y;
    }
  }
  // rows 0:3  p += dt v
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) X[i][v] = X[i][v] + dt * X[3 + i][v];
  // rows 3:6  v += A theta ; rows 6:9  theta = B theta
  {
    double yt[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const d2 c0 = f2[((FX2_AB + 6 * i) / 2 + 0) * PS], c1 = f2[((FX2_AB + 6 * i) / 2 + 1) * PS], c2 = f2[((FX2_AB + 6 * i) / 2 + 2) * PS];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        double yv = X[3 + i][v];
        yv += c0.x * X[6][v];
        yv += c0.y * X[7][v];
        yv += c1.x * X[8][v];
        X[3 + i][v] = yv;
        double t = 0.0;
        t += c1.y * X[6][v];
        t += c2.x * X[7][v];
        t += c2.y * X[8][v];
        yt[i][v] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int v = 0; v < 3; ++v) X[6 + i][v] = yt[i][v];
  }
  // rows 15:17 notch chain
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    X[15][v] = X[15][v] + dt * X[16][v];
    X[16][v] = X[16][v] + dt * X[17][v];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      X[18 + i][v] = y18[i][v];
      X[21 + i][v] = y21[i][v];
    }
}

// Pass 1 of a covariance step: T(:, tile) = Fx X, every finished row stored straight away as
// out[i * OS + v] (no in-place update: X is dead after this pass, the next pass starts from the
// transposed tile).  Interleaving the 72 stores with the FMAs lets the shared-memory pipe and the FP64
// pipe overlap inside one warp.
template <int PS, int OS>
ESKF_HD void fx_apply_store(const double (&X)[24][3], const d2* f2, double* out) {
  const double dt = f2[(FX2_DT / 2) * PS].x;
  // untouched rows (identity rows of Fx): dofs 9:15 and notch'' 17
#pragma unroll
  for (int i = 9; i < 15; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) out[i * OS + v] = X[i][v];
#pragma unroll
  for (int v = 0; v < 3; ++v) out[17 * OS + v] = X[17][v];
  // rows 18:21 (camera position error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const d2 c0 = f2[((FX2_R18 + 10 * i) / 2 + 0) * PS], c1 = f2[((FX2_R18 + 10 * i) / 2 + 1) * PS],
             c2 = f2[((FX2_R18 + 10 * i) / 2 + 2) * PS], c3 = f2[((FX2_R18 + 10 * i) / 2 + 3) * PS],
             c4 = f2[((FX2_R18 + 10 * i) / 2 + 4) * PS];
    const double c[9] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y, c4.x};
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = dt * X[3 + i][v];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int v = 0; v < 3; ++v) y[v] += c[k] * X[6 + k][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) out[(18 + i) * OS + v] = y[v] + X[16 + i][v];
  }
  // rows 21:24 (camera orientation error)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const d2 c0 = f2[((FX2_R21 + 8 * i) / 2 + 0) * PS], c1 = f2[((FX2_R21 + 8 * i) / 2 + 1) * PS],
             c2 = f2[((FX2_R21 + 8 * i) / 2 + 2) * PS], c3 = f2[((FX2_R21 + 8 * i) / 2 + 3) * PS];
    double y[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      y[v] = (i == 0) ? 0.0 : X[21 + i][v];
      y[v] += c0.x * X[9][v];
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] += c0.y * X[10][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] += c1.x * X[11][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] += c1.y * X[15][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] += c2.x * X[19][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] += c2.y * X[20][v];
#pragma unroll
    for (int v = 0; v < 3; ++v) out[(21 + i) * OS + v] = y[v] + c3.x * X[21][v];
  }
  // rows 0:3  p += dt v
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) out[i * OS + v] = X[i][v] + dt * X[3 + i][v];
  // rows 3:6  v += A theta ; rows 6:9  theta = B theta
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const d2 c0 = f2[((FX2_AB + 6 * i) / 2 + 0) * PS], c1 = f2[((FX2_AB + 6 * i) / 2 + 1) * PS],
             c2 = f2[((FX2_AB + 6 * i) / 2 + 2) * PS];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double yv = X[3 + i][v];
      yv += c0.x * X[6][v];
      yv += c0.y * X[7][v];
      yv += c1.x * X[8][v];
      out[(3 + i) * OS + v] = yv;
      double t = 0.0;
      t += c1.y * X[6][v];
      t += c2.x * X[7][v];
      t += c2.y * X[8][v];
      out[(6 + i) * OS + v] = t;
    }
  }
  // rows 15:17 notch chain
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    out[15 * OS + v] = X[15][v] + dt * X[16][v];
    out[16 * OS + v] = X[16][v] + dt * X[17][v];
  }
}

// Fi Q Fi^T for the row tile of lane group g (X[j][v] = P'[3g+v][j]); Filter.py:349.
template <int PS, typename QD>
ESKF_HD void process_noise_reg(double (&X)[24][3], int g, const d2* f2, const QD& qd, bool imu_q) {
  // diagonal: rows 3..14 get qd[r-3] (Fi[3:15,0:12] = I), row 17 gets qd(12) (Fi[17,12] = 1)
#pragma unroll
  for (int r = 3; r < 15; ++r)
    if (g == r / 3) X[r][r % 3] += qd(r - 3);
  if (g == 5) X[17][2] += qd(12);
  if (imu_q && (g == 2 || g == 6 || g == 7)) {
    // n_om drives theta (I), p_C (Np) and theta_C (Nt): L Q_om L^T on rows/cols {6:9,18:21,21:24};
    // the diagonal of the theta block was added above.
    double Lr[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        Lr[v][k] = (g == 2) ? ((k == v) ? 1.0 : 0.0) : (g == 6) ? fx2_at<PS>(f2, FX2_NP + 3 * v + k) : fx2_at<PS>(f2, FX2_NT + 3 * v + k);
#pragma unroll
    for (int cb = 0; cb < 3; ++cb) {
      const int c0 = (cb == 0) ? 6 : (cb == 1) ? 18 : 21;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double Lc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
          Lc[k] = (cb == 0) ? ((k == j) ? 1.0 : 0.0) : (cb == 1) ? fx2_at<PS>(f2, FX2_NP + 3 * j + k) : fx2_at<PS>(f2, FX2_NT + 3 * j + k);
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) acc += Lr[v][k] * qd(3 + k) * Lc[k];
          if (!(g == 2 && cb == 0)) X[c0 + j][v] += acc;
        }
      }
    }
  }
}

}  // namespace eskf
