// GPU pre-pass: from a camera trajectory to everything eskf_run consumes.
//
// Reference behaviour restated (paths relative to the reference repo), the same arithmetic as the numpy pre-pass
// dvi_ekf_b200/camera.py that it replaces on the critical path:
//   Camera derived data ......... dvi_ekf/models/Camera.py:84-118,158-170 (np.gradient of position and Euler angles)
//   Interpolator ................ dvi_ekf/models/trajectory/Interpolator.py:25-88 (np.linspace + np.interp on every
//                                 channel, raw quaternion components included, re-normalised afterwards)
//   synthetic IMU ............... dvi_ekf/models/Imu.py:141-226, dvi_ekf/kinematics/equations.py:8-41,54-69
//   epoch membership ............ dvi_ekf/models/Camera.py:299-301,320-347 (t_interp <= t_frame, quirk Q14)
//   initial state ............... dvi_ekf/tools/utils.py:54-75
//   rotated camera (with_notch) . dvi_ekf/models/Camera.py:172-208 (IMU source, initial state and error reference; the
//                                 measurements stay those of the un-rotated camera, Filter.py:144-185)
// One thread per camera frame for the derived data, one thread per interpolated instant for the IMU synthesis.
#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/eskf.h"
#include "eskf_math.cuh"

using namespace eskf;

namespace {

struct PP {
  int64_t n;       // camera frames
  int64_t n_new;   // interpolated instants
  int ifv;
  int euler_mode;  // 0: extrinsic xyz (HEAD), 1: zyx reversed (the revision behind the golden files, quirk Q11)
  int with_notch;
  double scale;
  Model model;
  double gt[6], ic[6];
  // frame arrays (device)
  const double *t, *xyz, *q, *notch3;
  double *p, *qn, *ang, *rdeg, *v, *om, *acc, *alp;  // [n,3] / [n,4]
  double* qsrc;  // [n,4] quaternion channel the Interpolator reads: the raw file columns, or the rotated camera's
  // per interpolated instant
  double *tn, *oa_all, *ref_all;  // [n_new], [n_new,6], [n_new,14]
  int64_t* idx;                   // [n] index of the last interpolated instant with t <= frame time
  int64_t* meta;                  // [0] = T, [1] = first
  int32_t* starts;                // [n-1] first step of every epoch
};

__device__ __forceinline__ void euler_of(const double* R, int mode, double* e) {
  if (mode == 0) {  // Rotation.as_euler("xyz")
    e[0] = atan2(R[7], R[8]);
    e[1] = -asin(fmin(1.0, fmax(-1.0, R[6])));
    e[2] = atan2(R[3], R[0]);
  } else {  // Rotation.as_euler("zyx")[::-1]
    e[0] = atan2(-R[5], R[8]);
    e[1] = asin(fmin(1.0, fmax(-1.0, R[2])));
    e[2] = atan2(-R[1], R[0]);
  }
}

// frames: scaled position, normalised quaternion, Euler angles (radians for the gradient, xyz degrees for the reference rows)
__global__ void pp_frames(PP a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double q[4], R[9], e[3];
  for (int j = 0; j < 3; ++j) a.p[3 * i + j] = a.xyz[3 * i + j] * a.scale;
  for (int j = 0; j < 4; ++j) q[j] = a.q[4 * i + j];
  quat_normalise(q);
  if (a.with_notch) {
    // Camera.gen_rotated (Camera.py:172-208): notch_quat * real_quat with notch_quat = Rz(ang_notch); the rotated camera is
    // the IMU source, the initial state and the error reference; its quaternion columns are already normalised
    double sh, ch, nq[4], qr[4];
    sincos(0.5 * a.notch3[3 * i], &sh, &ch);
    nq[0] = 0.0;
    nq[1] = 0.0;
    nq[2] = sh;
    nq[3] = ch;
    quat_mul(nq, q, qr);
    for (int j = 0; j < 4; ++j) q[j] = qr[j];
    for (int j = 0; j < 4; ++j) a.qsrc[4 * i + j] = q[j];
  } else {
    for (int j = 0; j < 4; ++j) a.qsrc[4 * i + j] = a.q[4 * i + j];
  }
  for (int j = 0; j < 4; ++j) a.qn[4 * i + j] = q[j];
  quat_to_rot(q, R);
  euler_of(R, a.euler_mode, e);
  for (int j = 0; j < 3; ++j) a.ang[3 * i + j] = e[j];
  euler_of(R, 0, e);
  const double r2d = 57.295779513082320876798154814105;
  for (int j = 0; j < 3; ++j) a.rdeg[3 * i + j] = e[j] * r2d;
}

// np.gradient(f, dt, axis=time) with uniform spacing: central differences inside, one-sided first order at the ends
__global__ void pp_gradient(const double* f0, double* g0, const double* f1, double* g1, int64_t n, double dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int s = 0; s < 2; ++s) {
    const double* f = s ? f1 : f0;
    double* g = s ? g1 : g0;
    for (int j = 0; j < 3; ++j) {
      double d;
      if (n == 1)
        d = 0.0;
      else if (i == 0)
        d = (f[3 + j] - f[j]) / dt;
      else if (i == n - 1)
        d = (f[3 * (n - 1) + j] - f[3 * (n - 2) + j]) / dt;
      else
        d = (f[3 * (i + 1) + j] - f[3 * (i - 1) + j]) / (2.0 * dt);
      g[3 * i + j] = d;
    }
  }
}

// np.interp(x, xp, fp) for one x and a strided channel: slope * (x - xp[k]) + fp[k] on the segment that holds x
__device__ __forceinline__ double interp1(double x, const double* xp, int64_t k, int64_t n, const double* fp, int stride, int off) {
  if (k >= n - 1) return fp[(n - 1) * stride + off];
  const double f0 = fp[k * stride + off], f1 = fp[(k + 1) * stride + off];
  const double slope = (f1 - f0) / (xp[k + 1] - xp[k]);
  return __dadd_rn(__dmul_rn(slope, x - xp[k]), f0);  // numpy's C loop: a product and a sum, not a fused multiply-add
}

// interpolated instants: camera channels by np.interp, probe kinematics at the ground-truth DOFs, IMU synthesis
__global__ void pp_samples(PP a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_new) return;
  const double t0 = a.t[0], t1 = a.t[a.n - 1];
  const double step = (a.n_new > 1) ? (t1 - t0) / (double)(a.n_new - 1) : 0.0;
  // np.linspace: arange * step + start (two roundings: no FMA contraction, the epoch membership compares these
  // values with the frame stamps), last point set to stop
  const double x = (j == a.n_new - 1) ? t1 : __dadd_rn(__dmul_rn((double)j, step), t0);
  a.tn[j] = x;
  // segment: largest k with xp[k] <= x
  int64_t lo = 0, hi = a.n - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (a.t[mid] <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  const int64_t k = lo;
  double p[3], q[4], v[3], acc[3], om[3], alp[3], nt[3] = {0, 0, 0};
  for (int c = 0; c < 3; ++c) {
    p[c] = interp1(x, a.t, k, a.n, a.p, 3, c);
    v[c] = interp1(x, a.t, k, a.n, a.v, 3, c);
    acc[c] = interp1(x, a.t, k, a.n, a.acc, 3, c);
    om[c] = interp1(x, a.t, k, a.n, a.om, 3, c);
    alp[c] = interp1(x, a.t, k, a.n, a.alp, 3, c);
    if (a.with_notch) nt[c] = interp1(x, a.t, k, a.n, a.notch3, 3, c);
  }
  for (int c = 0; c < 4; ++c) q[c] = interp1(x, a.t, k, a.n, a.qsrc, 4, c);  // RAW components, re-normalised below
  quat_normalise(q);
  double R_WC[9];
  quat_to_rot(q, R_WC);
  // ground-truth probe (SimpleProbe constraints, Probe.py:385-388) with the notch joint from the notch trajectory
  ProbeKin pk;
  ProbeTrig tr;
  probe_eval(a.model, a.gt, nt, pk, tr);
  double om_p[3], alp_p[3];
  for (int c = 0; c < 3; ++c) {
    om_p[c] = pk.z6[c] * nt[1];
    alp_p[c] = pk.z6[c] * nt[2];
  }
  // R_WB = R_WC R_p^T
  double R_WB[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R_WB[3 * r + c] = R_WC[3 * r] * pk.R[3 * c] + R_WC[3 * r + 1] * pk.R[3 * c + 1] + R_WC[3 * r + 2] * pk.R[3 * c + 2];
  double Rp[3], Rom[3], Ralp[3], W_om[3], W_omxp[3], t3[3], W_alp[3], W_acc[3];
  mv3(R_WB, pk.p, Rp);
  mv3(R_WB, om_p, Rom);
  mv3(R_WB, alp_p, Ralp);
  for (int c = 0; c < 3; ++c) W_om[c] = om[c] - Rom[c];
  cross3(W_om, Rp, W_omxp);
  cross3(W_om, Rom, t3);
  for (int c = 0; c < 3; ++c) W_alp[c] = alp[c] - Ralp[c] - t3[c];
  double c1[3], c2[3];
  cross3(W_alp, Rp, c1);
  cross3(W_om, W_omxp, c2);
  for (int c = 0; c < 3; ++c) W_acc[c] = acc[c] - c1[c] - c2[c];
  // f_imu_meas (equations.py:8-41,63-69): om_B = R_BW W_om, acc_B = R_BW W_acc, R_BW = R_WB^T
  double omB[3], accB[3];
  mtv3(R_WB, W_om, omB);
  mtv3(R_WB, W_acc, accB);
  for (int c = 0; c < 3; ++c) {
    a.oa_all[6 * j + c] = omB[c];
    a.oa_all[6 * j + 3 + c] = accB[c];
  }
  // f_imu (equations.py:8-13,54-60): IMU reference pose / velocity; ImuRefTraj row (ImuRefTraj.py:18-55)
  double e[3], qB[4];
  euler_of(R_WB, 0, e);
  quat_from_matrix(R_WB, qB);
  double* r = a.ref_all + 14 * j;
  const double r2d = 57.295779513082320876798154814105;
  r[0] = x;
  for (int c = 0; c < 3; ++c) {
    r[1 + c] = p[c] - Rp[c];
    r[4 + c] = v[c] - W_omxp[c];
    r[7 + c] = e[c] * r2d;
  }
  r[10] = qB[3];
  r[11] = qB[0];
  r[12] = qB[1];
  r[13] = qB[2];
}

// idx[e] = np.searchsorted(t_interp, t_frame[e], side="right") - 1
__global__ void pp_membership(PP a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  const double x = a.t[e];
  int64_t lo = -1, hi = a.n_new - 1;  // largest j with tn[j] <= x (-1: none)
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (a.tn[mid] <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  a.idx[e] = lo;
}

// epochs: samples per epoch, first step of every epoch (sequential scan: E is a few thousand at most)
__global__ void pp_epochs(PP a, int32_t* n_prop) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t acc = 0;
  for (int64_t e = 0; e + 1 < a.n; ++e) {
    const int64_t c = a.idx[e + 1] - a.idx[e];
    n_prop[e] = (int32_t)c;
    a.starts[e] = (int32_t)acc;
    acc += c;
  }
  a.meta[0] = acc;            // T
  a.meta[1] = a.idx[0] + 1;   // first interpolated instant that is a step
}

struct PPOut {
  double *x0, *u0, *dt, *om_acc, *t_imu, *cam, *notch, *cam_ref, *imu_ref, *imu_ref_rows;
};

__global__ void pp_steps(PP a, PPOut o) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t T = a.meta[0], first = a.meta[1];
  if (k >= T) return;
  const int64_t j = first + k;
  // Filter.propagate_imu restarts old_ti at the FRAME time of each epoch (Filter.py:195): find the epoch of step k
  int64_t lo = 0, hi = a.n - 2;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if ((int64_t)a.starts[mid] <= k)
      lo = mid;
    else
      hi = mid - 1;
  }
  // (epochs without samples share their start with the next one: the LAST epoch starting at k owns it)
  const bool is_start = ((int64_t)a.starts[lo] == k);
  const double t_prev = is_start ? a.t[lo] : a.tn[j - 1];
  o.dt[k] = a.tn[j] - t_prev;
  o.t_imu[k] = a.tn[j];
  for (int c = 0; c < 6; ++c) o.om_acc[6 * k + c] = a.oa_all[6 * j + c];
  for (int c = 0; c < 14; ++c) o.imu_ref_rows[14 * k + c] = a.ref_all[14 * j + c];
}

__global__ void pp_epoch_rows(PP a, PPOut o) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e + 1 >= a.n) return;
  const int64_t f = e + 1;  // camera frame of the update
  for (int c = 0; c < 3; ++c) {
    o.cam[7 * e + c] = a.p[3 * f + c];
    o.cam_ref[6 * e + c] = a.p[3 * f + c];
    o.cam_ref[6 * e + 3 + c] = a.rdeg[3 * f + c];
  }
  for (int c = 0; c < 4; ++c) o.cam[7 * e + 3 + c] = a.q[4 * f + c];  // RAW quaternion (VisualTrajectory.py:120-134)
  o.notch[e] = a.with_notch ? a.notch3[3 * f] : 0.0;
  const double* r = a.ref_all + 14 * a.idx[f];
  for (int c = 0; c < 6; ++c) o.imu_ref[6 * e + c] = r[4 + c];
  if (e == 0) {
    // initial state (tools/utils.py:54-75) from frame 0 = interpolated instant 0; first IMU sample (Filter.py:63,78-79)
    const double* r0 = a.ref_all;
    for (int c = 0; c < 3; ++c) {
      o.x0[c] = r0[1 + c];
      o.x0[3 + c] = r0[4 + c];
      o.x0[16 + c] = a.with_notch ? a.notch3[c] : 0.0;
      o.x0[19 + c] = a.p[c];
    }
    o.x0[6] = r0[11];
    o.x0[7] = r0[12];
    o.x0[8] = r0[13];
    o.x0[9] = r0[10];
    for (int c = 0; c < 6; ++c) o.x0[10 + c] = a.ic[c];
    for (int c = 0; c < 4; ++c) o.x0[22 + c] = a.qn[c];
    for (int c = 0; c < 6; ++c) o.u0[c] = a.oa_all[c];
  }
}

thread_local std::string g_pp_err;

}  // namespace

#define PCK(call)                                                    \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      g_pp_err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      for (void* p_ : tmp) cudaFree(p_);                             \
      return ESKF_ECUDA;                                             \
    }                                                                \
  } while (0)

extern "C" {

const char* eskf_prepass_last_error(void) { return g_pp_err.c_str(); }

int eskf_prepass(int device, void* cuda_stream, const eskf_model_t* model, const eskf_prepass_in_t* in, const eskf_prepass_out_t* out,
                 int64_t* n_steps_out) {
  nvtxRangePushA("eskf_prepass");
  struct Pop { ~Pop() { nvtxRangePop(); } } nvtx_pop_;
  std::vector<void*> tmp;
  if (!model || !in || !out || !n_steps_out || in->n_frames < 2 || in->interframe_vals < 1 || !in->t || !in->xyz || !in->q_xyzw ||
      !out->x0 || !out->u0 || !out->dt || !out->om_acc || !out->n_prop || !out->cam || !out->notch || !out->cam_ref || !out->imu_ref) {
    g_pp_err = "eskf_prepass: bad argument";
    return ESKF_EINVAL;
  }
  PCK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t n = in->n_frames, n_new = (n - 1) * in->interframe_vals + 1;
  auto dalloc = [&](size_t bytes, void** p) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) tmp.push_back(*p);
    return e;
  };
  PP a{};
  a.n = n;
  a.n_new = n_new;
  a.ifv = in->interframe_vals;
  a.euler_mode = in->euler_mode;
  a.with_notch = in->notch3 ? 1 : 0;
  a.scale = in->scale;
  a.model.L = model->scope_length;
  a.model.sa = sin(model->cam_angle_rad);
  a.model.ca = cos(model->cam_angle_rad);
  a.model.frozen_mask = model->frozen_mask;
  a.model.flags = model->flags;
  for (int i = 0; i < 6; ++i) {
    a.gt[i] = in->gt_dofs[i];
    a.ic[i] = in->ic_dofs[i];
  }
  double *dt_ = nullptr, *dxyz = nullptr, *dq = nullptr, *dn3 = nullptr, *frame = nullptr, *samp = nullptr;
  PCK(dalloc(n * sizeof(double), (void**)&dt_));
  PCK(dalloc(3 * n * sizeof(double), (void**)&dxyz));
  PCK(dalloc(4 * n * sizeof(double), (void**)&dq));
  PCK(cudaMemcpyAsync(dt_, in->t, n * sizeof(double), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(dxyz, in->xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(dq, in->q_xyzw, 4 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (in->notch3) {
    PCK(dalloc(3 * n * sizeof(double), (void**)&dn3));
    PCK(cudaMemcpyAsync(dn3, in->notch3, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  a.t = dt_;
  a.xyz = dxyz;
  a.q = dq;
  a.notch3 = dn3;
  PCK(dalloc((size_t)n * 29 * sizeof(double), (void**)&frame));  // p 3, qn 4, ang 3, rdeg 3, v 3, om 3, acc 3, alp 3, qsrc 4
  a.p = frame;
  a.qn = frame + 3 * n;
  a.ang = frame + 7 * n;
  a.rdeg = frame + 10 * n;
  a.v = frame + 13 * n;
  a.om = frame + 16 * n;
  a.acc = frame + 19 * n;
  a.alp = frame + 22 * n;
  a.qsrc = frame + 25 * n;
  PCK(dalloc((size_t)n_new * 21 * sizeof(double), (void**)&samp));
  a.tn = samp;
  a.oa_all = samp + n_new;
  a.ref_all = samp + 7 * n_new;
  PCK(dalloc(n * sizeof(int64_t), (void**)&a.idx));
  PCK(dalloc(2 * sizeof(int64_t), (void**)&a.meta));
  PCK(dalloc(n * sizeof(int32_t), (void**)&a.starts));
  double *imu_rows = out->imu_ref_rows, *t_imu = out->t_imu;
  if (!imu_rows) PCK(dalloc((size_t)n_new * 14 * sizeof(double), (void**)&imu_rows));
  if (!t_imu) PCK(dalloc((size_t)n_new * sizeof(double), (void**)&t_imu));

  const double dt0 = in->t[1] - in->t[0];  // Camera.dt (Camera.py:96): the spacing np.gradient is given
  const unsigned bf = (unsigned)((n + 127) / 128), bs = (unsigned)((n_new + 127) / 128);
  pp_frames<<<bf, 128, 0, st>>>(a);
  pp_gradient<<<bf, 128, 0, st>>>(a.p, a.v, a.ang, a.om, n, dt0);
  pp_gradient<<<bf, 128, 0, st>>>(a.v, a.acc, a.om, a.alp, n, dt0);
  pp_samples<<<bs, 128, 0, st>>>(a);
  pp_membership<<<bf, 128, 0, st>>>(a);
  pp_epochs<<<1, 32, 0, st>>>(a, out->n_prop);
  PPOut o{out->x0, out->u0, out->dt, out->om_acc, t_imu, out->cam, out->notch, out->cam_ref, out->imu_ref, imu_rows};
  pp_steps<<<bs, 128, 0, st>>>(a, o);  // n_new - 1 >= T threads
  pp_epoch_rows<<<bf, 128, 0, st>>>(a, o);
  PCK(cudaGetLastError());
  int64_t meta[2] = {0, 0};
  PCK(cudaMemcpyAsync(meta, a.meta, sizeof(meta), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  *n_steps_out = meta[0];
  for (void* p_ : tmp) cudaFree(p_);
  return ESKF_OK;
}

}  // extern "C"
