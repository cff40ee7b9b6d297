// eskf_kernel2: the warp-specialised persistent VI-ESKF kernel for sm_100a.
//
// One CTA owns F filters for a whole launch (a whole trajectory in eskf_run).  Warps are roles:
//   warp 0  IMU     lane = filter: p, v, q of the nominal state (Filter._predict_nominal, Filter.py:232-247,
//                   equations.py:72-86) and R_WB_old (Filter.py:227)
//   warp 1  CAMERA  lane = filter: p_cam, q_cam (equations.py:88-98), the measurement residual
//                   (Filter.py:363-375), status word
//   warp 2  JACOB   lane = filter: dofs, notch chain, probe forward kinematics (Probe.py:470-480) and the
//                   Jacobian blocks of Fx / Fi (Filter._predict_error, Filter.py:249-342)
//   warp 3  STAGER  lane = filter: stages the IMU sample stream (dt, om, acc) into a 4-slot shared-memory
//                   ring two steps ahead and adds the Monte-Carlo noise (Philox4x32-10 + Box-Muller);
//                   stages the camera measurement of the epoch (Imu.eval_expr_single / Filter.propagate_imu,
//                   Imu.py:141-196, Filter.py:187-217; VisualTraj.at_index, VisualTrajectory.py:120-134)
//   warps 4.. COVARIANCE  eight lanes per filter, lane g keeps columns 3g..3g+2 of the 24x24 covariance in
//                   REGISTERS (72 doubles).  One step of Filter._predict_error_covariance (Filter.py:344-349):
//                       X <- Fx X            (column tile of P  ->  column tile of T = Fx P)
//                       transpose through shared memory (72 STS.64 + 36 LDS.128 per lane)
//                       X <- Fx X + Q-terms  (row tile of T  ->  (Fx T^T) = row tile of P' = Fx P Fx^T)
//                   and because P' is symmetric the row tile IS the column tile of the next step: P itself
//                   never goes through shared memory.  The camera update (Filter.update, Filter.py:351-395)
//                   spills the tile to shared memory and uses the cooperative LU / gain / Joseph code of the
//                   v1 kernel (eskf_kernel.cuh, eskf_math.cuh).
// The scalar roles of step k run concurrently with the covariance role of step k-1; all roles meet at one
// CTA barrier per step.  Records exchanged through shared memory are double buffered:
//   RING[4]  samples (slot (k+1)&3 = new sample of step k, slot k&3 = old sample)       STAGER -> IMU, CAMERA, JACOB
//   RO/RW/V[2] R_WB_old, R_WB and v at the start of step k (slot k&1)                   IMU -> CAMERA, JACOB
//   PK[2]    probe kinematics + notch, notch' at the start of step k (slot k&1)         JACOB -> CAMERA
//   FXB[2]   Jacobian blocks of the step (fx2 layout, [pair][filter] so that the writer's STS.128 and the
//            readers' broadcast LDS.128 are both conflict free)                         JACOB -> COVARIANCE
// Register budget by role (setmaxnreg): scalar warps shrink to REG_S, covariance warps grow to REG_C.
#pragma once
#include "eskf_kernel.cuh"

// profiling experiment switches (python -m dvi_ekf_b200.build --exp=...): time one side of the kernel alone
#ifdef ESKF_EXP_NO_SCALAR
#define ESKF2_SCALAR_ON(it) ((it) < 2)  // scalar roles only fill both record slots once per epoch
#else
#define ESKF2_SCALAR_ON(it) true
#endif
#ifdef ESKF_EXP_NO_COV
#define ESKF2_COV_ON false
#else
#define ESKF2_COV_ON true
#endif

namespace eskf {

constexpr int RS2 = 26;          // row stride of the covariance tile buffer: even (16-byte rows) and
constexpr int TB_STRIDE = 632;   // 24*26 + 8; = 8 (mod 16) doubles => conflict-free STS.64 / LDS.128
constexpr int TB_TAIL = 624;     // 8 doubles behind the tile: diag(R) (7)
constexpr int UPD_STRIDE = UP_SIZE + 2;  // update record per filter (+ UP_OK2)
constexpr int NPAIR = FX2_SIZE / 2;      // 47

// element j of the per-filter parameter block kept in the pad columns (24, 25) of the tile rows
__host__ __device__ constexpr int tb_pad(int j) { return (j >> 1) * RS2 + 24 + (j & 1); }

// scalar exchange block, element-major: element j of filter f at SX[j * F + f]
constexpr int SX_RING = 0;    // 4 x 8: om(3) acc(3) dt pad
constexpr int SX_RO = 32;     // 2 x 9
constexpr int SX_RW = 50;     // 2 x 9
constexpr int SX_V = 68;      // 2 x 3
constexpr int SX_PK = 74;     // 2 x 17: p(3) R(9) z6(3) notch notch'
constexpr int SX_MEAS = 108;  // 8: cam pos(3) quat(4) notch
constexpr int SX_SIZE = 116;

template <int F>
struct Lay2 {
  static constexpr int TB = 0;                   // [F][TB_STRIDE]
  static constexpr int SX = F * TB_STRIDE;       // [SX_SIZE][F]
  static constexpr int UN = SX + SX_SIZE * F;    // union { FXB [2][NPAIR][F] d2 ; UPD [F][UPD_STRIDE] }
  static constexpr int UN_SIZE = (2 * NPAIR * 2 > UPD_STRIDE ? 2 * NPAIR * 2 : UPD_STRIDE) * F;
  static constexpr int TOTAL = UN + UN_SIZE;     // doubles
  static_assert((SX % 2) == 0 && (UN % 2) == 0, "16-byte alignment");
};

template <int N>
__device__ __forceinline__ void reg_inc() {
  if constexpr (N > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  if constexpr (N > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}

struct Ctx2 {
  double* smem;
  int64_t f0;      // first local filter of the CTA
  int nf;          // filters of this CTA that exist
  int64_t gid0;    // global id of the CTA's first filter
  int64_t traj;
  const int32_t* n_prop;
  const double* dtp;
};

// ---------------------------------------------------------------------------------------------
// cooperative, coalesced tile store (all threads of the CTA)
template <int F, int NTHR>
__device__ __forceinline__ void store_tiles(const KArgs& a, const Ctx2& c, int tid) {
  const double* sT = c.smem + Lay2<F>::TB;
  for (int idx = tid; idx < c.nf * 576; idx += NTHR) {
    const int f = idx / 576, r = idx - f * 576;
    const int i = r / 24, j = r - i * 24;
    a.P[(c.f0 + f) * 576 + r] = sT[f * TB_STRIDE + i * RS2 + j];
  }
}

// statistics rows (Filter.calculate_dof_metric / update_mse) assembled by warp 0 from the partials the
// scalar roles left in the union region: [f][0] mseA_last [1] mseA_sum [2] mseB_last [3] mseB_sum
// [4] n_upd [5] status [6..11] (dofs - gt)^2
template <int F>
__device__ __forceinline__ void write_stats(const KArgs& a, const Ctx2& c, int lane) {
  if (!(a.stats_out || a.stats_sum)) return;
  const double* un = c.smem + Lay2<F>::UN;
  double row[ESKF_NSTAT];
#pragma unroll
  for (int i = 0; i < ESKF_NSTAT; ++i) row[i] = 0.0;
  if (lane < c.nf) {
    const double* r = un + lane * UPD_STRIDE;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      row[i] = r[6 + i];
      acc += r[6 + i];
    }
    row[6] = acc / 6.0;
    row[7] = (r[0] + r[2]) / 12.0;
    row[8] = (r[1] + r[3]) / 12.0;
    row[9] = r[4];
    row[10] = r[5];
    row[11] = 1.0;
    if (a.stats_out) {
#pragma unroll
      for (int i = 0; i < ESKF_NSTAT; ++i) a.stats_out[(c.f0 + lane) * ESKF_NSTAT + i] = row[i];
    }
  }
  if (a.stats_sum) {
#pragma unroll
    for (int i = 0; i < ESKF_NSTAT; ++i) {
      double v = row[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(a.stats_sum + i, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// role 0: IMU nominal state
template <int F, int NTHR>
__device__ __forceinline__ void role_imu(const KArgs& a, const Ctx2& c, int lane) {
  using L = Lay2<F>;
  const bool act = lane < c.nf;
  double* sx = c.smem + L::SX + lane;
  double p[3], v[3], q[4], Rwb[9], Rold[9];
  double mse_last = 0.0, mse_sum = 0.0;
  if (act) {
    const double* xg = a.x + (c.f0 + lane) * NX;
    const double* rg = a.Ro + (c.f0 + lane) * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = xg[i];
      v[i] = xg[3 + i];
      sx[(SX_V + i) * F] = v[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = xg[6 + i];
    quat_to_rot(q, Rwb);  // R_WB of the first step is rot(q); R_WB_old may be stale (quirk Q8)
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      Rold[i] = rg[i];
      sx[(SX_RO + i) * F] = Rold[i];
      sx[(SX_RW + i) * F] = Rwb[i];
    }
  }
  __syncthreads();  // prologue
  int64_t k = 0;
  for (int64_t e = 0; e < a.E; ++e) {
    const int n = c.n_prop ? c.n_prop[e] : (int)a.T;
    for (int it = 0; it <= n; ++it) {
      if (act && it < n && ESKF2_SCALAR_ON(it)) {
        const int64_t kk = k + it;
        const double* uo = sx + (SX_RING + 8 * (int)(kk & 3)) * F;
        const double* un = sx + (SX_RING + 8 * (int)((kk + 1) & 3)) * F;
        double om_old[3], acc_old[3], om[3], acc[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          om_old[i] = uo[i * F];
          acc_old[i] = uo[(3 + i) * F];
          om[i] = un[i * F];
          acc[i] = un[(3 + i) * F];
        }
        const double dt = un[6 * F];
        imu_nominal_step(p, v, q, Rwb, dt, om_old, acc_old, om, acc, Rold);
        const int s = (int)((kk + 1) & 1);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          Rwb[i] = Rold[i];
          sx[(SX_RO + 9 * s + i) * F] = Rold[i];
          sx[(SX_RW + 9 * s + i) * F] = Rold[i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) sx[(SX_V + 3 * s + i) * F] = v[i];
      }
      __syncthreads();
    }
    k += n;
    if (!a.do_update) continue;
    __syncthreads();  // U0 | U1
    __syncthreads();  // U1 | U2
    if (act) {
      const double* up = c.smem + L::UN + lane * UPD_STRIDE;
      if (up[UP_OK2] != 0.0) {  // state (+) error state, IMU part (state.py:46-53,116-121)
        const double th[3] = {up[UP_DELTA + 6], up[UP_DELTA + 7], up[UP_DELTA + 8]};
        double dq[4], qn[4];
        quat_about_axis(sqrt(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]), th, dq);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          p[i] += up[UP_DELTA + i];
          v[i] += up[UP_DELTA + 3 + i];
        }
        quat_mul(q, dq, qn);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = qn[i];
        quat_to_rot(q, Rwb);  // R_WB of the next step; R_WB_old keeps the pre-update value (quirk Q8)
        const int s = (int)(k & 1);
#pragma unroll
        for (int i = 0; i < 9; ++i) sx[(SX_RW + 9 * s + i) * F] = Rwb[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) sx[(SX_V + 3 * s + i) * F] = v[i];
      }
      if (a.cam_ref && a.imu_ref) {  // IMU half of Filter.calculate_update_mse (Filter.py:408-413)
        const double* ir = a.imu_ref + (c.traj * a.E + e) * 6;
        double ei[3], acc2 = 0.0;
        euler_xyz_deg(q, ei);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double d2_ = v[i] - ir[i], d3 = ei[i] - ir[3 + i];
          acc2 += d2_ * d2_ + d3 * d3;
        }
        mse_last = acc2;
        mse_sum += acc2;
      }
    }
    __syncthreads();  // U2 done
  }
  // ---- write back ----
  if (act) {
    double* xg = a.x + (c.f0 + lane) * NX;
    double* ug = a.u + (c.f0 + lane) * 6;
    double* rg = a.Ro + (c.f0 + lane) * 9;
    const double* uo = sx + (SX_RING + 8 * (int)(k & 3)) * F;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      xg[i] = p[i];
      xg[3 + i] = v[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) xg[6 + i] = q[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) ug[i] = uo[i * F];
#pragma unroll
    for (int i = 0; i < 9; ++i) rg[i] = Rold[i];
    double* r = c.smem + L::UN + lane * UPD_STRIDE;
    r[0] = mse_last;
    r[1] = mse_sum;
  }
  __syncthreads();  // tiles dumped, statistics partials written
  store_tiles<F, NTHR>(a, c, threadIdx.x);
  write_stats<F>(a, c, lane);
}

// ---------------------------------------------------------------------------------------------
// role 1: camera nominal state + measurement residual
template <int F, int NTHR>
__device__ __forceinline__ void role_cam(const KArgs& a, const Ctx2& c, int lane) {
  using L = Lay2<F>;
  const bool act = lane < c.nf;
  double* sx = c.smem + L::SX + lane;
  double pc[3], qc[4];
  double mse_last = 0.0, mse_sum = 0.0, n_upd = 0.0;
  int32_t st = 0;
  if (act) {
    const double* xg = a.x + (c.f0 + lane) * NX;
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = xg[19 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) qc[i] = xg[22 + i];
    st = a.status[c.f0 + lane];
  }
  __syncthreads();  // prologue
  int64_t k = 0;
  for (int64_t e = 0; e < a.E; ++e) {
    const int n = c.n_prop ? c.n_prop[e] : (int)a.T;
    for (int it = 0; it <= n; ++it) {
      if (act && it < n && ESKF2_SCALAR_ON(it)) {
        const int64_t kk = k + it;
        const double* uo = sx + (SX_RING + 8 * (int)(kk & 3)) * F;
        const double* un = sx + (SX_RING + 8 * (int)((kk + 1) & 3)) * F;
        const int s = (int)(kk & 1);
        double om_old[3], om[3], vpre[3], Rwb[9], pkp[3], pkR[9], pkz[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          om_old[i] = uo[i * F];
          om[i] = un[i * F];
          vpre[i] = sx[(SX_V + 3 * s + i) * F];
          pkp[i] = sx[(SX_PK + 17 * s + i) * F];
          pkz[i] = sx[(SX_PK + 17 * s + 12 + i) * F];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          Rwb[i] = sx[(SX_RW + 9 * s + i) * F];
          pkR[i] = sx[(SX_PK + 17 * s + 3 + i) * F];
        }
        const double dt = un[6 * F];
        const double notch_d = sx[(SX_PK + 17 * s + 16) * F];
        cam_nominal_step(pc, qc, vpre, Rwb, dt, om_old, om, pkp, pkR, pkz, notch_d);
      }
      __syncthreads();
    }
    k += n;
    if (!a.do_update) continue;
    // ---- U0: residual (Filter.py:363-375) ----
    if (lane < F) {
      double* up = c.smem + L::UN + lane * UPD_STRIDE;
      bool ok = false;
      if (act) {
        double cam[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) cam[i] = sx[(SX_MEAS + i) * F];
        const double notch_meas = sx[(SX_MEAS + 7) * F];
        const double notch0 = sx[(SX_PK + 17 * (int)(k & 1) + 15) * F];
        Nominal s;  // only pc, qc, notch[0] are read by update_residual
#pragma unroll
        for (int i = 0; i < 3; ++i) s.pc[i] = pc[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) s.qc[i] = qc[i];
        s.notch[0] = notch0;
        ok = update_residual(s, cam, cam + 3, notch_meas, up + UP_RES);
        if (!ok) st |= ESKF_STATUS_ASIN_DOMAIN;
      }
      up[UP_OK] = ok ? 1.0 : 0.0;
    }
    __syncthreads();  // U0 | U1
    __syncthreads();  // U1 | U2
    if (act) {
      const double* up = c.smem + L::UN + lane * UPD_STRIDE;
      if (up[UP_OK2] != 0.0) {  // camera part of the injection, incl. the dqc axis slip (quirk Q4, state.py:124)
        const double th[3] = {up[UP_DELTA + 6], up[UP_DELTA + 7], up[UP_DELTA + 8]};
        const double thc[3] = {up[UP_DELTA + 21], up[UP_DELTA + 22], up[UP_DELTA + 23]};
        double dqc[4], qn[4];
        quat_about_axis(sqrt(thc[0] * thc[0] + thc[1] * thc[1] + thc[2] * thc[2]), th, dqc);
#pragma unroll
        for (int i = 0; i < 3; ++i) pc[i] += up[UP_DELTA + 18 + i];
        quat_mul(qc, dqc, qn);
#pragma unroll
        for (int i = 0; i < 4; ++i) qc[i] = qn[i];
        n_upd += 1.0;
      } else {
        st |= ESKF_STATUS_UPDATE_SKIPPED;
      }
      if (a.cam_ref && a.imu_ref) {  // camera half of Filter.calculate_update_mse (Filter.py:401-406)
        const double* cr = a.cam_ref + (c.traj * a.E + e) * 6;
        double ec[3], acc2 = 0.0;
        euler_xyz_deg(qc, ec);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double d0 = cr[i] - pc[i], d1 = cr[3 + i] - ec[i];
          acc2 += d0 * d0 + d1 * d1;
        }
        mse_last = acc2;
        mse_sum += acc2;
      }
    }
    __syncthreads();  // U2 done
  }
  if (act) {
    double* xg = a.x + (c.f0 + lane) * NX;
#pragma unroll
    for (int i = 0; i < 3; ++i) xg[19 + i] = pc[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) xg[22 + i] = qc[i];
    a.status[c.f0 + lane] = st;
    double* r = c.smem + L::UN + lane * UPD_STRIDE;
    r[2] = mse_last;
    r[3] = mse_sum;
    r[4] = n_upd;
    r[5] = (double)st;
  }
  __syncthreads();
  store_tiles<F, NTHR>(a, c, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// role 2: dofs / notch, probe kinematics, Jacobian blocks
template <int F>
__device__ __forceinline__ void publish_pk(double* sx, int s, const ProbeKin& pk, const double* notch) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sx[(SX_PK + 17 * s + i) * F] = pk.p[i];
    sx[(SX_PK + 17 * s + 12 + i) * F] = pk.z6[i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) sx[(SX_PK + 17 * s + 3 + i) * F] = pk.R[i];
  sx[(SX_PK + 17 * s + 15) * F] = notch[0];
  sx[(SX_PK + 17 * s + 16) * F] = notch[1];
}

template <int F, int NTHR>
__device__ __forceinline__ void role_jac(const KArgs& a, const Ctx2& c, int lane) {
  using L = Lay2<F>;
  const bool act = lane < c.nf;
  double* sx = c.smem + L::SX + lane;
  d2* fxb = reinterpret_cast<d2*>(c.smem + L::UN) + lane;  // pair j2 of slot s at fxb[(s * NPAIR + j2) * F]
  double dofs[6], notch[3], sig_om[3] = {0, 0, 0};
  ProbeKin pk;
  ProbeTrig tr;
  bool imu_q = false;
  if (act) {
    const double* xg = a.x + (c.f0 + lane) * NX;
    const double* pg = a.par + (c.f0 + lane) * PAR_STRIDE;
#pragma unroll
    for (int i = 0; i < 6; ++i) dofs[i] = xg[10 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      notch[i] = xg[16 + i];
      sig_om[i] = pg[PAR_SIGOM + i];
    }
    imu_q = (pg[PAR_QD + 3] != 0.0) || (pg[PAR_QD + 4] != 0.0) || (pg[PAR_QD + 5] != 0.0);
    probe_eval(a.model, dofs, notch, pk, tr);
    publish_pk<F>(sx, 0, pk, notch);
  }
  __syncthreads();  // prologue
  int64_t k = 0;
  for (int64_t e = 0; e < a.E; ++e) {
    const int n = c.n_prop ? c.n_prop[e] : (int)a.T;
    for (int it = 0; it <= n; ++it) {
      if (act && it < n && ESKF2_SCALAR_ON(it)) {
        const int64_t kk = k + it;
        const double* uo = sx + (SX_RING + 8 * (int)(kk & 3)) * F;
        const double dt = sx[(SX_RING + 8 * (int)((kk + 1) & 3) + 6) * F];
        const int s = (int)(kk & 1);
        double om_old[3], acc_old[3], Ro[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          om_old[i] = uo[i * F];
          acc_old[i] = uo[(3 + i) * F];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Ro[i] = sx[(SX_RO + 9 * s + i) * F];
        if (dofs_notch_step(a.model, dofs, notch, dt)) probe_eval(a.model, dofs, notch, pk, tr);
        alignas(16) double fx[FX2_SIZE];
        fx[FX2_DT + 1] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          fx[FX2_R18 + 10 * i + 9] = 0.0;
          fx[FX2_R21 + 8 * i + 7] = 0.0;
        }
        fx[FX2_NP + 9] = 0.0;
        fx[FX2_NT + 9] = 0.0;
        jacobian_blocks(a.model, dofs, notch[1], pk, tr, Ro, dt, om_old, acc_old, sig_om, imu_q, fx);
        d2* dst = fxb + ((it & 1) * NPAIR) * F;
#pragma unroll
        for (int j = 0; j < FX2_NP / 2; ++j) dst[j * F] = d2{fx[2 * j], fx[2 * j + 1]};
        if (imu_q) {
#pragma unroll
          for (int j = FX2_NP / 2; j < NPAIR; ++j) dst[j * F] = d2{fx[2 * j], fx[2 * j + 1]};
        }
        publish_pk<F>(sx, (int)((kk + 1) & 1), pk, notch);
      }
      __syncthreads();
    }
    k += n;
    if (!a.do_update) continue;
    __syncthreads();  // U0 | U1
    __syncthreads();  // U1 | U2
    if (act) {
      const double* up = c.smem + L::UN + lane * UPD_STRIDE;
      if (up[UP_OK2] != 0.0) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
          if (!((a.model.frozen_mask >> i) & 1)) dofs[i] += up[UP_DELTA + 9 + i];  // Filter.py:377-379
#pragma unroll
        for (int i = 0; i < 3; ++i) notch[i] += up[UP_DELTA + 15 + i];
        probe_eval(a.model, dofs, notch, pk, tr);
        publish_pk<F>(sx, (int)(k & 1), pk, notch);
      }
    }
    __syncthreads();  // U2 done
  }
  if (act) {
    double* xg = a.x + (c.f0 + lane) * NX;
#pragma unroll
    for (int i = 0; i < 6; ++i) xg[10 + i] = dofs[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) xg[16 + i] = notch[i];
    double* r = c.smem + L::UN + lane * UPD_STRIDE;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double d = dofs[i] - a.gt_dofs[i];
      r[6 + i] = d * d;
    }
  }
  __syncthreads();
  store_tiles<F, NTHR>(a, c, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// role 3: sample-stream stager + Monte-Carlo noise
template <int F, int NTHR>
__device__ __forceinline__ void role_stage(const KArgs& a, const Ctx2& c, int lane) {
  using L = Lay2<F>;
  const bool act = lane < c.nf;
  double* sx = c.smem + L::SX + lane;
  const int64_t gid = c.gid0 + lane;
  const bool noisy = a.noise_on && !(a.noise_free0 && gid == 0);
  const double* oap =
      a.om_acc ? (a.stream_per_filter ? a.om_acc + (c.f0 + lane) * a.T * 6 : a.om_acc + c.traj * a.T * 6) : nullptr;
  auto stage_sample = [&](int64_t j) {  // sample of step j -> ring slot (j + 1) & 3
    double* dst = sx + (SX_RING + 8 * (int)((j + 1) & 3)) * F;
    double u[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = oap[j * 6 + i];
    if (noisy) {
      double z[6];
      normal6(a.seed, (uint64_t)gid, (uint64_t)j, RNG_KIND_IMU, z);
#pragma unroll
      for (int i = 0; i < 6; ++i) u[i] += a.imu_noise[i] * z[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i * F] = u[i];
    dst[6 * F] = c.dtp[j];
  };
  auto stage_meas = [&](int64_t e) {
    const int64_t mrow = a.meas_per_filter ? (c.f0 + lane) : (c.traj * a.E + e);
    double cam[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) cam[i] = a.cam[mrow * 7 + i];
    double notch = a.notch[mrow];
    if (noisy) {
      double z[8];
      normal6(a.seed, (uint64_t)gid, (uint64_t)e, RNG_KIND_CAM, z);
      normal2(a.seed, (uint64_t)gid, (uint64_t)e, RNG_KIND_CAM2, z + 6);
      double dth[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        cam[i] += a.cam_noise[i] * z[i];
        dth[i] = a.cam_noise[3 + i] * z[3 + i];
      }
      // orientation noise: small body rotation of the measured quaternion (its norm is kept)
      double dq[4], qn[4];
      quat_about_axis(sqrt(dth[0] * dth[0] + dth[1] * dth[1] + dth[2] * dth[2]), dth, dq);
      const double nq = sqrt(cam[3] * cam[3] + cam[4] * cam[4] + cam[5] * cam[5] + cam[6] * cam[6]);
      quat_mul(cam + 3, dq, qn);
#pragma unroll
      for (int i = 0; i < 4; ++i) cam[3 + i] = qn[i] * nq;
      notch += a.cam_noise[6] * z[6];
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) sx[(SX_MEAS + i) * F] = cam[i];
    sx[(SX_MEAS + 7) * F] = notch;
  };
  if (act) {
    const double* ug = a.u + (c.f0 + lane) * 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) sx[(SX_RING + i) * F] = ug[i];  // slot 0: the buffered previous sample
    sx[(SX_RING + 6) * F] = 0.0;
    if (a.T > 0 && oap) stage_sample(0);
  }
  __syncthreads();  // prologue
  int64_t k = 0;
  for (int64_t e = 0; e < a.E; ++e) {
    const int n = c.n_prop ? c.n_prop[e] : (int)a.T;
    for (int it = 0; it <= n; ++it) {
      if (act) {
        if (it == 0 && a.do_update) stage_meas(e);
        if (it < n && k + it + 1 < a.T && ESKF2_SCALAR_ON(it)) stage_sample(k + it + 1);
      }
      __syncthreads();
    }
    k += n;
    if (!a.do_update) continue;
    __syncthreads();  // U0 | U1
    __syncthreads();  // U1 | U2
    __syncthreads();  // U2 done
  }
  __syncthreads();
  store_tiles<F, NTHR>(a, c, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// covariance role
template <int F, int NTHR>
__device__ __forceinline__ void role_cov(const KArgs& a, const Ctx2& c, int ct) {
  using L = Lay2<F>;
  const int cf = ct >> 3;  // filter of this lane (padded filters run on an identity tile, never stored)
  const int cg = ct & 7;   // state group owned
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  double* Tb = c.smem + L::TB + cf * TB_STRIDE;
  double* up = c.smem + L::UN + cf * UPD_STRIDE;
  const d2* fxb = reinterpret_cast<const d2*>(c.smem + L::UN) + cf;
  const double* rd = Tb + TB_TAIL;
  auto qd = [&](int j) { return Tb[tb_pad(PAR_QD + j)]; };
  const bool imu_q = (qd(3) != 0.0) || (qd(4) != 0.0) || (qd(5) != 0.0);

  // X[i][v] = P[3g+v][i]: rows 3g..3g+2 of P, used as its columns 3g..3g+2 (a covariance is symmetric;
  // the engine never relies on more than that, and a launch ends with exactly the rows it would start
  // the next launch from, so eskf_run == the same sequence of eskf_propagate / eskf_update calls bit for bit)
  double X[24][3];
  auto dump_rows = [&]() {  // tile -> rows 3g..3g+2 of the buffer (16-byte stores)
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      d2* row = reinterpret_cast<d2*>(Tb + (3 * cg + v) * RS2);
#pragma unroll
      for (int j = 0; j < 12; ++j) row[j] = d2{X[2 * j][v], X[2 * j + 1][v]};
    }
  };
  auto load_rows = [&]() {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const d2* row = reinterpret_cast<const d2*>(Tb + (3 * cg + v) * RS2);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const d2 t = row[j];
        X[2 * j][v] = t.x;
        X[2 * j + 1][v] = t.y;
      }
    }
  };

  load_rows();
  __syncthreads();  // prologue

  for (int64_t e = 0; e < a.E; ++e) {
    const int n = c.n_prop ? c.n_prop[e] : (int)a.T;
    for (int it = 0; it <= n; ++it) {
      if (it >= 1 && ESKF2_COV_ON) {
        const d2* f2 = fxb + (((it - 1) & 1) * NPAIR) * F;
        // pass 1: T(:, 3g..3g+2) = Fx P(:, 3g..3g+2), rows stored as they are finished
        fx_apply_store<F, RS2>(X, f2, Tb + 3 * cg);
        __syncwarp(gmask);
        load_rows();  // X[k][v] = T(3g+v, k)
        __syncwarp(gmask);
        // pass 2: P'(3g+v, :) = Fx T(3g+v, :)^T
        fx_apply_reg<F>(X, f2);
        process_noise_reg<F>(X, cg, f2, qd, imu_q);
      }
      __syncthreads();
    }
    if (!a.do_update) continue;
    // ---- U0: spill the tile, invert S ----
    dump_rows();
    __syncwarp(gmask);
    const bool inv_ok = inv7_group<RS2>(Tb, rd, cg, up);
    __syncthreads();  // U0 | U1
    const bool upd_c = inv_ok && (up[UP_OK] != 0.0);
    if (upd_c) gain_rows3<RS2>(Tb, 3 * cg, up);
    if (cg == 0) up[UP_OK2] = upd_c ? 1.0 : 0.0;
    __syncthreads();  // U1 | U2
    if (upd_c) {
      if (a.K_out && cf < c.nf) {
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int r = 3 * cg + v;
#pragma unroll
          for (int m = 0; m < 7; ++m)
            a.K_out[((c.f0 + cf) * 24 + r) * 7 + m] = (ESKF_HSET(m) == r) ? up[UP_KD + m] : up[UP_KZ + 7 * r + m];
        }
      }
      joseph_apply3<RS2, 1>(Tb + 3 * cg, up);  // (I-KH) P
      __syncwarp(gmask);
      joseph_rows_finish3<RS2>(Tb + 3 * cg * RS2, 3 * cg, up, rd);  // (.)(I-KH)^T + K R K^T, reset
      load_rows();
    }
    __syncthreads();  // U2 done
  }
  dump_rows();
  __syncthreads();
  store_tiles<F, NTHR>(a, c, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
template <int F, int REG_S, int REG_C>
__global__ void __launch_bounds__(128 + 8 * F, 1) eskf_kernel2(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) double smem[];
  using L = Lay2<F>;
  constexpr int NTHR = 128 + 8 * F;
  const int tid = threadIdx.x;
  Ctx2 c;
  c.smem = smem;
  c.f0 = (int64_t)blockIdx.x * F;
  c.nf = (int)((a.N - c.f0) < F ? (a.N - c.f0) : F);
  c.gid0 = a.filter_id0 + c.f0;
  c.traj = (a.n_traj > 1) ? (c.gid0 / a.filters_per_traj) : 0;
  c.n_prop = a.n_prop ? a.n_prop + c.traj * a.E : nullptr;
  c.dtp = a.dt ? a.dt + c.traj * a.T : nullptr;

  // ---- covariance tiles and parameters (coalesced) ----
  for (int idx = tid; idx < F * 576; idx += NTHR) {
    const int f = idx / 576, r = idx - f * 576;
    const int i = r / 24, j = r - i * 24;
    smem[L::TB + f * TB_STRIDE + i * RS2 + j] = (f < c.nf) ? a.P[(c.f0 + f) * 576 + r] : ((i == j) ? 1.0 : 0.0);
  }
  for (int idx = tid; idx < F * PAR_STRIDE; idx += NTHR) {
    const int f = idx / PAR_STRIDE, r = idx - f * PAR_STRIDE;
    const int64_t row = (f < c.nf) ? (c.f0 + f) : c.f0;
    const double val = (r < PAR_SIZE) ? a.par[row * PAR_STRIDE + r] : 0.0;
    double* Tb = smem + L::TB + f * TB_STRIDE;
    if (r >= PAR_RD && r < PAR_RD + 7)
      Tb[TB_TAIL + (r - PAR_RD)] = val;
    else
      Tb[tb_pad(r)] = val;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= 4) {
    reg_inc<REG_C>();
    role_cov<F, NTHR>(a, c, tid - 128);
  } else {
    reg_dec<REG_S>();
    if (warp == 0)
      role_imu<F, NTHR>(a, c, lane);
    else if (warp == 1)
      role_cam<F, NTHR>(a, c, lane);
    else if (warp == 2)
      role_jac<F, NTHR>(a, c, lane);
    else
      role_stage<F, NTHR>(a, c, lane);
  }
}

template <int F>
cudaError_t launch_eskf_kernel2(const KArgs& a, cudaStream_t stream);

}  // namespace eskf
