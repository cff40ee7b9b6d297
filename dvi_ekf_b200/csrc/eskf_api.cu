// C ABI of the B200 batched VI-ESKF engine (see include/eskf.h) and its small utility kernels.
// The persistent kernel itself lives in eskf_kernel.cuh and is instantiated per CTA shape in
// eskf_launch.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "eskf_kernel3.cuh"

using namespace eskf;

namespace eskf {
#define ESKF_DECL(F) template <> cudaError_t launch_eskf_kernel<F>(const KArgs& a, cudaStream_t stream);
ESKF_DECL(4) ESKF_DECL(8) ESKF_DECL(12) ESKF_DECL(16) ESKF_DECL(20) ESKF_DECL(24) ESKF_DECL(28)
#undef ESKF_DECL
#define ESKF_DECL3(F) template <> cudaError_t launch_eskf_kernel3<F>(const KArgs& a, cudaStream_t stream);
ESKF_DECL3(4) ESKF_DECL3(8) ESKF_DECL3(16) ESKF_DECL3(28)
#undef ESKF_DECL3
}  // namespace eskf

namespace {

// ---- small utility kernels ----
__global__ void bcast_rows_kernel(double* dst, const double* src, int64_t n, int w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * w) return;
  dst[i] = src[i % w];
}
__global__ void rot_from_state_kernel(double* Ro, const double* x, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9];
  quat_to_rot(x + i * NX + 6, R);
  for (int j = 0; j < 9; ++j) Ro[i * 9 + j] = R[j];
}
__global__ void fill_par_kernel(double* par, const double* q, int64_t nq, const double* r, int64_t nr, const double* s,
                                int64_t ns, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double* p = par + i * PAR_STRIDE;
  if (q)
    for (int j = 0; j < 13; ++j) p[PAR_QD + j] = q[(nq == 1 ? 0 : i) * 13 + j];
  if (r)
    for (int j = 0; j < 7; ++j) p[PAR_RD + j] = r[(nr == 1 ? 0 : i) * 7 + j];
  if (s)
    for (int j = 0; j < 3; ++j) p[PAR_SIGOM + j] = s[(ns == 1 ? 0 : i) * 3 + j];
  p[PAR_SIZE] = 0.0;
}

// FP64 FMA throughput probe: ILP independent DFMA chains per thread, nothing else in the loop.
// Used by bench.py to measure the roofline denominator (MEASURED_PEAKS.json has no FP64 figure).
template <int ILP>
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;  // keeps the chain alive without a store in the common case
}

// eskf_noise_dump: the standard normals of (filter, step, kind), exactly as the persistent kernels draw them
__global__ void noise_dump_kernel(double* out, uint64_t seed, int64_t filter_id0, int64_t n_filters, int64_t step0, int64_t n_steps,
                                  uint32_t kind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_filters * n_steps) return;
  const int64_t f = i / n_steps, k = i - f * n_steps;
  double z[8];
  normal8(seed, (uint64_t)(filter_id0 + f), (uint64_t)(step0 + k), kind, z);
  for (int j = 0; j < 8; ++j) out[i * 8 + j] = z[j];
}

thread_local std::string g_err;

}  // namespace

// ---------------------------------------------------------------------------------------------
constexpr int N_STAGE = 9;
struct eskf_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t N = 0;
  Model model{};
  double *x = nullptr, *P = nullptr, *u = nullptr, *Ro = nullptr, *par = nullptr;
  int32_t* status = nullptr;
  // grow-only device staging for host-side arguments / results
  void* stage[N_STAGE] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t stage_sz[N_STAGE] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double* stats_sum_dev = nullptr;
  int64_t launches = 0;
  int fpc = 0;  // filters per CTA (0 = automatic)
  int variant = 0;  // 0 = default (eskf_kernel3), 1 = eskf_kernel (first version, kept for A/B measurements), 3 = eskf_kernel3
  int sm_count = 148;
};

#define CK(call)                                                  \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) {                                      \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return ESKF_ECUDA;                                          \
    }                                                             \
  } while (0)

static int stage_reserve(eskf_t* h, int slot, size_t bytes) {
  if (h->stage_sz[slot] >= bytes) return ESKF_OK;
  if (h->stage[slot]) CK(cudaFree(h->stage[slot]));
  h->stage[slot] = nullptr;
  h->stage_sz[slot] = 0;
  CK(cudaMalloc(&h->stage[slot], bytes));
  h->stage_sz[slot] = bytes;
  return ESKF_OK;
}

// host pointer -> copied into a staging slot on the handle's stream; device pointer -> used as is
static int stage_in(eskf_t* h, int slot, const void* src, size_t bytes, int mem, const void** out) {
  *out = nullptr;
  if (!src || bytes == 0) return ESKF_OK;
  if (mem == ESKF_MEM_DEVICE) {
    *out = src;
    return ESKF_OK;
  }
  int rc = stage_reserve(h, slot, bytes);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->stage[slot], src, bytes, cudaMemcpyHostToDevice, h->stream));
  *out = h->stage[slot];
  return ESKF_OK;
}

static const int kShapes1[] = {28, 24, 20, 16, 12, 8, 4};  // eskf_kernel  (v1)
static const int kShapes3[] = {28, 16, 8, 4};              // eskf_kernel3

static int kernel_of(const eskf_t* h) { return h->variant == 1 ? 1 : 3; }

// Filters per CTA.  Must divide filters_per_traj when several trajectories are stacked (a CTA follows
// ONE trajectory's epoch structure).  Every shape runs one CTA per SM (registers / shared memory) and a
// wave of CTAs takes about the same time whatever its shape (per-step latency bound, profiles/r01_*): the
// automatic choice minimises the number of waves and then prefers the larger shape.
static int pick_fpc(const eskf_t* h, int64_t fpt, bool multi_traj) {
  auto fits = [&](int c) { return !multi_traj || (fpt % c) == 0; };
  const int* shapes = kernel_of(h) == 3 ? kShapes3 : kShapes1;
  const int ns = kernel_of(h) == 3 ? 4 : 7;
  if (h->fpc > 0) {
    for (int i = 0; i < ns; ++i)
      if (shapes[i] == h->fpc && fits(shapes[i])) return shapes[i];
  }
  int best = 0;
  int64_t best_waves = 0;
  for (int i = 0; i < ns; ++i) {  // descending
    const int c = shapes[i];
    if (!fits(c)) continue;
    const int64_t ctas = (h->N + c - 1) / c;
    const int64_t waves = (ctas + h->sm_count - 1) / h->sm_count;
    if (best == 0 || waves < best_waves) {
      best = c;
      best_waves = waves;
    }
  }
  return best;
}

static int launch(eskf_t* h, const KArgs& a, int64_t fpt, bool multi_traj) {
  cudaError_t e;
  const int fpc = pick_fpc(h, fpt, multi_traj);
  if (kernel_of(h) == 3) {
    switch (fpc) {
      case 28: e = launch_eskf_kernel3<28>(a, h->stream); break;
      case 16: e = launch_eskf_kernel3<16>(a, h->stream); break;
      case 8: e = launch_eskf_kernel3<8>(a, h->stream); break;
      case 4: e = launch_eskf_kernel3<4>(a, h->stream); break;
      default:
        g_err = "filters_per_traj must be a multiple of 4 when several trajectories are stacked";
        return ESKF_EINVAL;
    }
  } else {
    switch (fpc) {
      case 28: e = launch_eskf_kernel<28>(a, h->stream); break;
      case 24: e = launch_eskf_kernel<24>(a, h->stream); break;
      case 20: e = launch_eskf_kernel<20>(a, h->stream); break;
      case 16: e = launch_eskf_kernel<16>(a, h->stream); break;
      case 12: e = launch_eskf_kernel<12>(a, h->stream); break;
      case 8: e = launch_eskf_kernel<8>(a, h->stream); break;
      case 4: e = launch_eskf_kernel<4>(a, h->stream); break;
      default:
        g_err = "filters_per_traj must be a multiple of 4 when several trajectories are stacked";
        return ESKF_EINVAL;
    }
  }
  if (e != cudaSuccess) {
    g_err = std::string("eskf_kernel launch: ") + cudaGetErrorString(e);
    return ESKF_ECUDA;
  }
  h->launches += 1;
  return ESKF_OK;
}

static void base_args(const eskf_t* h, KArgs& a) {
  memset(&a, 0, sizeof(a));
  a.x = h->x;
  a.P = h->P;
  a.u = h->u;
  a.Ro = h->Ro;
  a.status = h->status;
  a.par = h->par;
  a.N = h->N;
  a.model = h->model;
  a.n_traj = 1;
  a.filters_per_traj = h->N;
}

extern "C" {

const char* eskf_last_error(void) { return g_err.c_str(); }
const char* eskf_version(void) { return "eskf_b200 0.3 (sm_100a)"; }

int eskf_create(const eskf_model_t* model, int64_t n_filters, int device, void* cuda_stream, eskf_t** out) {
  if (!model || !out || n_filters <= 0) {
    g_err = "eskf_create: bad argument";
    return ESKF_EINVAL;
  }
  CK(cudaSetDevice(device));
  eskf_t* h = new eskf_handle();
  h->device = device;
  h->stream = (cudaStream_t)cuda_stream;
  h->N = n_filters;
  h->model.L = model->scope_length;
  h->model.sa = sin(model->cam_angle_rad);
  h->model.ca = cos(model->cam_angle_rad);
  h->model.frozen_mask = model->frozen_mask;
  h->model.flags = model->flags;
  int smc = 0;
  CK(cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device));
  h->sm_count = smc > 0 ? smc : 148;
  const size_t n = (size_t)n_filters;
  CK(cudaMalloc(&h->x, n * NX * sizeof(double)));
  CK(cudaMalloc(&h->P, n * 576 * sizeof(double)));
  CK(cudaMalloc(&h->u, n * 6 * sizeof(double)));
  CK(cudaMalloc(&h->Ro, n * 9 * sizeof(double)));
  CK(cudaMalloc(&h->par, n * PAR_STRIDE * sizeof(double)));
  CK(cudaMalloc(&h->status, n * sizeof(int32_t)));
  CK(cudaMalloc(&h->stats_sum_dev, ESKF_NSTAT * sizeof(double)));
  CK(cudaMemsetAsync(h->x, 0, n * NX * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->P, 0, n * 576 * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->u, 0, n * 6 * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->Ro, 0, n * 9 * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->par, 0, n * PAR_STRIDE * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->status, 0, n * sizeof(int32_t), h->stream));
  *out = h;
  return ESKF_OK;
}

int eskf_destroy(eskf_t* h) {
  if (!h) return ESKF_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->x);
  cudaFree(h->P);
  cudaFree(h->u);
  cudaFree(h->Ro);
  cudaFree(h->par);
  cudaFree(h->status);
  cudaFree(h->stats_sum_dev);
  for (int i = 0; i < N_STAGE; ++i)
    if (h->stage[i]) cudaFree(h->stage[i]);
  delete h;
  return ESKF_OK;
}

static int set_rows(eskf_t* h, int slot, double* dst, const double* src, int64_t rows, int w, int mem) {
  if (!src) return ESKF_OK;
  if (rows != 1 && rows != h->N) {
    g_err = "leading dimension must be 1 or n_filters";
    return ESKF_EINVAL;
  }
  const size_t bytes = (size_t)rows * w * sizeof(double);
  if (rows == h->N) {
    CK(cudaMemcpyAsync(dst, src, bytes, mem == ESKF_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                       h->stream));
    return ESKF_OK;
  }
  const void* d = nullptr;
  int rc = stage_in(h, slot, src, bytes, mem, &d);
  if (rc) return rc;
  const int64_t tot = h->N * w;
  bcast_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(dst, (const double*)d, h->N, w);
  CK(cudaGetLastError());
  h->launches += 1;
  return ESKF_OK;
}

int eskf_set_state(eskf_t* h, const double* x, int64_t nx, const double* P, int64_t nP, const double* u_old,
                   int64_t nu, const double* R_old, int64_t nR, int mem) {
  if (!h) return ESKF_EINVAL;
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = set_rows(h, 0, h->x, x, nx, NX, mem))) return rc;
  if ((rc = set_rows(h, 1, h->P, P, nP, 576, mem))) return rc;
  if ((rc = set_rows(h, 2, h->u, u_old, nu, 6, mem))) return rc;
  if (R_old) {
    if ((rc = set_rows(h, 3, h->Ro, R_old, nR, 9, mem))) return rc;
  } else if (x) {
    rot_from_state_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(h->Ro, h->x, h->N);
    CK(cudaGetLastError());
    h->launches += 1;
  }
  if (x) CK(cudaMemsetAsync(h->status, 0, (size_t)h->N * sizeof(int32_t), h->stream));
  return ESKF_OK;
}

int eskf_set_noise(eskf_t* h, const double* Qdiag, int64_t nq, const double* Rdiag, int64_t nr,
                   const double* sigma_om, int64_t ns, int mem) {
  if (!h) return ESKF_EINVAL;
  CK(cudaSetDevice(h->device));
  auto bad = [&](const double* p, int64_t n) { return p && n != 1 && n != h->N; };
  if (bad(Qdiag, nq) || bad(Rdiag, nr) || bad(sigma_om, ns)) {
    g_err = "eskf_set_noise: leading dimension must be 1 or n_filters";
    return ESKF_EINVAL;
  }
  const void *dq = nullptr, *dr = nullptr, *ds = nullptr;
  int rc;
  if ((rc = stage_in(h, 0, Qdiag, (size_t)nq * 13 * sizeof(double), mem, &dq))) return rc;
  if ((rc = stage_in(h, 1, Rdiag, (size_t)nr * 7 * sizeof(double), mem, &dr))) return rc;
  if ((rc = stage_in(h, 2, sigma_om, (size_t)ns * 3 * sizeof(double), mem, &ds))) return rc;
  // the parameter table always has N rows; a broadcast simply fills every row
  fill_par_kernel<<<(unsigned)((h->N + 127) / 128), 128, 0, h->stream>>>(h->par, (const double*)dq, nq, (const double*)dr,
                                                                       nr, (const double*)ds, ns, h->N);
  CK(cudaGetLastError());
  h->launches += 1;
  return ESKF_OK;
}

int eskf_propagate(eskf_t* h, const double* dt, const double* om_acc, int64_t T, int per_filter, int mem) {
  if (!h || !dt || !om_acc || T < 0) {
    g_err = "eskf_propagate: bad argument";
    return ESKF_EINVAL;
  }
  if (T == 0) return ESKF_OK;
  CK(cudaSetDevice(h->device));
  const void *ddt = nullptr, *doa = nullptr;
  int rc;
  if ((rc = stage_in(h, 0, dt, (size_t)T * sizeof(double), mem, &ddt))) return rc;
  if ((rc = stage_in(h, 1, om_acc, (size_t)(per_filter ? h->N : 1) * T * 6 * sizeof(double), mem, &doa))) return rc;
  KArgs a;
  base_args(h, a);
  a.T = T;
  a.E = 1;
  a.dt = (const double*)ddt;
  a.om_acc = (const double*)doa;
  a.stream_per_filter = per_filter ? 1 : 0;
  a.do_update = 0;
  return launch(h, a, h->N, false);
}

int eskf_update(eskf_t* h, const double* cam, const double* notch, int per_filter, double* K_out, int mem) {
  if (!h || !cam || !notch) {
    g_err = "eskf_update: bad argument";
    return ESKF_EINVAL;
  }
  CK(cudaSetDevice(h->device));
  const int64_t rows = per_filter ? h->N : 1;
  const void *dc = nullptr, *dn = nullptr;
  int rc;
  if ((rc = stage_in(h, 0, cam, (size_t)rows * 7 * sizeof(double), mem, &dc))) return rc;
  if ((rc = stage_in(h, 1, notch, (size_t)rows * sizeof(double), mem, &dn))) return rc;
  double* dK = nullptr;
  const size_t kb = (size_t)h->N * 168 * sizeof(double);
  if (K_out) {
    if (mem == ESKF_MEM_DEVICE) {
      dK = K_out;
    } else {
      if ((rc = stage_reserve(h, 2, kb))) return rc;
      dK = (double*)h->stage[2];
    }
    CK(cudaMemsetAsync(dK, 0, kb, h->stream));
  }
  KArgs a;
  base_args(h, a);
  a.T = 0;
  a.E = 1;
  a.cam = (const double*)dc;
  a.notch = (const double*)dn;
  a.meas_per_filter = per_filter ? 1 : 0;
  a.do_update = 1;
  a.K_out = dK;
  if ((rc = launch(h, a, h->N, false))) return rc;
  if (K_out && mem == ESKF_MEM_HOST) {
    CK(cudaMemcpyAsync(K_out, dK, kb, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return ESKF_OK;
}

int eskf_run(eskf_t* h, const eskf_streams_t* sp, double* stats_out, double* stats_sum, int mem) {
  if (!h || !sp || !sp->dt || !sp->om_acc || !sp->n_prop || !sp->cam || !sp->notch || sp->n_traj < 1) {
    g_err = "eskf_run: bad argument";
    return ESKF_EINVAL;
  }
  CK(cudaSetDevice(h->device));
  const int smem_kind = sp->mem;
  const int64_t T = sp->n_steps, E = sp->n_epochs, nt = sp->n_traj;
  const int64_t fpt = nt > 1 ? sp->filters_per_traj : h->N;
  if (nt > 1 && (fpt <= 0 || (sp->filter_id0 % fpt) != 0)) {
    g_err = "eskf_run: filters_per_traj must divide filter_id0";
    return ESKF_EINVAL;
  }
  const void *ddt, *doa, *dnp, *dcam, *dno, *dcr, *dir;
  int rc;
  if ((rc = stage_in(h, 0, sp->dt, (size_t)nt * T * sizeof(double), smem_kind, &ddt))) return rc;
  if ((rc = stage_in(h, 1, sp->om_acc, (size_t)nt * T * 6 * sizeof(double), smem_kind, &doa))) return rc;
  if ((rc = stage_in(h, 2, sp->n_prop, (size_t)nt * E * sizeof(int32_t), smem_kind, &dnp))) return rc;
  if ((rc = stage_in(h, 3, sp->cam, (size_t)nt * E * 7 * sizeof(double), smem_kind, &dcam))) return rc;
  if ((rc = stage_in(h, 4, sp->notch, (size_t)nt * E * sizeof(double), smem_kind, &dno))) return rc;
  if ((rc = stage_in(h, 5, sp->cam_ref, (size_t)nt * E * 6 * sizeof(double), smem_kind, &dcr))) return rc;
  if ((rc = stage_in(h, 6, sp->imu_ref, (size_t)nt * E * 6 * sizeof(double), smem_kind, &dir))) return rc;
  double* dstats = nullptr;
  const size_t sb = (size_t)h->N * ESKF_NSTAT * sizeof(double);
  if (stats_out) {
    if (mem == ESKF_MEM_DEVICE) {
      dstats = stats_out;
    } else {
      if ((rc = stage_reserve(h, 7, sb))) return rc;
      dstats = (double*)h->stage[7];
    }
  }
  double* dtrace = nullptr;
  const size_t tb = (size_t)h->N * (size_t)T * NX * sizeof(double);
  if (sp->trace_x) {
    if (smem_kind == ESKF_MEM_DEVICE) {
      dtrace = sp->trace_x;
    } else {
      if ((rc = stage_reserve(h, 8, tb))) return rc;
      dtrace = (double*)h->stage[8];
    }
  }
  double* dsum = nullptr;
  if (stats_sum) {
    dsum = (mem == ESKF_MEM_DEVICE) ? stats_sum : h->stats_sum_dev;
    CK(cudaMemsetAsync(dsum, 0, ESKF_NSTAT * sizeof(double), h->stream));
  }
  KArgs a;
  base_args(h, a);
  a.T = T;
  a.E = E;
  a.n_traj = (int)nt;
  a.filters_per_traj = fpt;
  a.filter_id0 = sp->filter_id0;
  a.dt = (const double*)ddt;
  a.om_acc = (const double*)doa;
  a.n_prop = (const int32_t*)dnp;
  a.cam = (const double*)dcam;
  a.notch = (const double*)dno;
  a.cam_ref = (const double*)dcr;
  a.imu_ref = (const double*)dir;
  for (int i = 0; i < 6; ++i) a.gt_dofs[i] = sp->gt_dofs[i];
  a.do_update = 1;
  a.stats_out = dstats;
  a.stats_sum = dsum;
  a.trace = dtrace;
  a.seed = sp->seed;
  a.noise_free0 = sp->noise_free_filter0;
  a.noise_mod = sp->noise_id_modulus;
  for (int i = 0; i < 6; ++i) {
    a.imu_noise[i] = sp->imu_noise_std[i];
    if (a.imu_noise[i] != 0.0) a.noise_on = 1;
  }
  for (int i = 0; i < 7; ++i) {
    a.cam_noise[i] = sp->cam_noise_std[i];
    if (a.cam_noise[i] != 0.0) a.noise_on = 1;
  }
  if ((rc = launch(h, a, fpt, nt > 1))) return rc;
  if (dtrace && smem_kind == ESKF_MEM_HOST) {
    CK(cudaMemcpyAsync(sp->trace_x, dtrace, tb, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (mem == ESKF_MEM_HOST && (stats_out || stats_sum)) {
    if (stats_out) CK(cudaMemcpyAsync(stats_out, dstats, sb, cudaMemcpyDeviceToHost, h->stream));
    if (stats_sum) CK(cudaMemcpyAsync(stats_sum, dsum, ESKF_NSTAT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return ESKF_OK;
}

int eskf_get_state(eskf_t* h, double* x, double* P, double* u_old, double* R_old, int32_t* status, int mem) {
  if (!h) return ESKF_EINVAL;
  CK(cudaSetDevice(h->device));
  const cudaMemcpyKind kd = (mem == ESKF_MEM_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const size_t n = (size_t)h->N;
  if (x) CK(cudaMemcpyAsync(x, h->x, n * NX * sizeof(double), kd, h->stream));
  if (P) CK(cudaMemcpyAsync(P, h->P, n * 576 * sizeof(double), kd, h->stream));
  if (u_old) CK(cudaMemcpyAsync(u_old, h->u, n * 6 * sizeof(double), kd, h->stream));
  if (R_old) CK(cudaMemcpyAsync(R_old, h->Ro, n * 9 * sizeof(double), kd, h->stream));
  if (status) CK(cudaMemcpyAsync(status, h->status, n * sizeof(int32_t), kd, h->stream));
  if (mem == ESKF_MEM_HOST) CK(cudaStreamSynchronize(h->stream));
  return ESKF_OK;
}

int eskf_sync(eskf_t* h) {
  if (!h) return ESKF_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return ESKF_OK;
}

int eskf_fp64_peak(int device, void* cuda_stream, int repeats, double* tflops_out, double* ms_out) {
  if (!tflops_out || repeats < 1) return ESKF_EINVAL;
  CK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int smc = 0;
  CK(cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device));
  double* d = nullptr;
  CK(cudaMalloc(&d, 64));
  constexpr int ILP = 16;
  const int blocks = smc * 8, threads = 256, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  dfma_peak_kernel<ILP><<<blocks, threads, 0, st>>>(d, iters, 1.0000001, 1e-9);  // warm-up
  double best = 1e30;
  for (int r = 0; r < repeats; ++r) {
    CK(cudaEventRecord(e0, st));
    dfma_peak_kernel<ILP><<<blocks, threads, 0, st>>>(d, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double flops = 2.0 * ILP * (double)iters * blocks * threads;
  *tflops_out = flops / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return ESKF_OK;
}

int eskf_noise_dump(int device, void* cuda_stream, uint64_t seed, int64_t filter_id0, int64_t n_filters, int64_t step0,
                    int64_t n_steps, int kind, double* out, int mem) {
  if (!out || n_filters <= 0 || n_steps <= 0 || (kind != (int)RNG_KIND_IMU && kind != (int)RNG_KIND_CAM)) {
    g_err = "eskf_noise_dump: bad argument";
    return ESKF_EINVAL;
  }
  CK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t tot = n_filters * n_steps;
  const size_t bytes = (size_t)tot * 8 * sizeof(double);
  double* d = out;
  if (mem == ESKF_MEM_HOST) CK(cudaMalloc(&d, bytes));
  noise_dump_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(d, seed, filter_id0, n_filters, step0, n_steps, (uint32_t)kind);
  CK(cudaGetLastError());
  if (mem == ESKF_MEM_HOST) {
    CK(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d));
  }
  return ESKF_OK;
}

int64_t eskf_launch_count(const eskf_t* h) { return h ? h->launches : 0; }

int eskf_set_variant(eskf_t* h, int variant) {
  if (!h || (variant != 0 && variant != 1 && variant != 3)) return ESKF_EINVAL;
  h->variant = variant;
  return ESKF_OK;
}

int eskf_set_tuning(eskf_t* h, int filters_per_cta) {
  if (!h || filters_per_cta < 0) return ESKF_EINVAL;
  h->fpc = filters_per_cta;
  return ESKF_OK;
}

}  // extern "C"
