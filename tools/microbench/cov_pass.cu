// Micro-benchmark of the covariance-role inner passes (eskf_cov3.cuh) in isolation: cycles per pass as a
// function of warps per SM sub-partition.  nvcc -arch=sm_100a -O3 -I../../dvi_ekf_b200/csrc cov_pass.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "eskf_cov3.cuh"
using namespace eskf;

constexpr int F = 28;
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) double sm[];
  d2* fxb = reinterpret_cast<d2*>(sm);                  // [FX3_NPAIR][F]
  double* tb = sm + 2 * FX3_NPAIR * F;                  // [F][632]
  for (int i = threadIdx.x; i < 2 * FX3_NPAIR * F; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
  const int cf = (threadIdx.x >> 3) % F, cg = threadIdx.x & 7;
  double X[24][3];
#pragma unroll
  for (int i = 0; i < 24; ++i)
#pragma unroll
    for (int v = 0; v < 3; ++v) X[i][v] = 1.0 + 0.01 * (i + v + threadIdx.x);
  __syncthreads();
  const d2* f2 = fxb + cf;
  double* Tb = tb + cf * 632;
  long long ph[3] = {0, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    asm volatile("" ::: "memory");  // the coefficients of every step are new: no hoisting out of the loop
    if (MODE == 0) {  // in-place pass only
      fx3_apply_inplace<F>(X, f2);
    } else if (MODE == 1) {  // full step: store pass, transposition, in-place pass
      const long long a0 = clock64();
      fx3_apply_store<F, 26>(X, f2, Tb + 3 * cg);
      const long long a1 = clock64();
      __syncwarp();
      fx3_load_transposed<26>(X, Tb + 3 * cg * 26);
      __syncwarp();
      const long long a2 = clock64();
      fx3_apply_inplace<F>(X, f2);
      const long long a3 = clock64();
      ph[0] += a1 - a0;
      ph[1] += a2 - a1;
      ph[2] += a3 - a2;
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += X[i][0] + X[i][1] + X[i][2];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0;
    cyc[1] = ph[0];
    cyc[2] = ph[1];
    cyc[3] = ph[2];
  }
}

template <int MODE>
void run(int warps) {
  double* d;
  long long* c;
  cudaMalloc(&d, 64);
  cudaMalloc(&c, 64);
  const size_t smem = (2 * FX3_NPAIR * F + F * 632) * 8;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  k<MODE><<<1, 32 * warps, smem>>>(d, c, iters);
  cudaDeviceSynchronize();
  k<MODE><<<1, 32 * warps, smem>>>(d, c, iters);
  long long h[4];
  cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
  printf("mode %d (%s) warps %d: %.0f cycles per iteration [pass1 %.0f | transpose %.0f | pass2 %.0f] (err %s)\n", MODE,
         MODE ? "store + transpose + in-place" : "in-place pass", warps, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters,
         (double)h[3] / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  for (int w : {1, 2, 4, 7, 8}) run<0>(w);
  for (int w : {1, 2, 4, 7, 8}) run<1>(w);
  return 0;
}
