// FP64 tensor-core (DMMA m8n8k4) throughput on the B200 next to the FP64 FMA pipe, and both together
// (nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_peak.cu -o dmma_peak && ./dmma_peak).
// Question (BASELINE.json north_star): would batched small-matrix products on DMMA beat CUDA-core FP64 for the covariance
// algebra?  One DMMA m8n8k4 warp instruction = 8 x 8 x 4 multiply-adds = 512 flop; one DFMA warp instruction = 64 flop.
//   (a) dmma: ILP independent accumulator fragments per warp, full chip
//   (b) dfma: the same with scalar DFMA chains (the denominator of bench.py's roofline, eskf_fp64_peak)
//   (c) mixed: DMMA and DFMA interleaved in the same warps -- if the two were separate pipes their rates would add
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP, bool MMA, bool FMA>
__global__ void kern(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    c0[i] = threadIdx.x * 1e-3 + i;
    c1[i] = 1.0 - i;
    f[i] = 0.5 * i + threadIdx.x;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MMA) dmma(c0[i], c1[i], a, b);
      if (FMA) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + f[i];
  if (s == 1.2345) out[0] = s;
}

template <int ILP, bool MMA, bool FMA>
void run(const char* what, int sms) {
  double* d;
  cudaMalloc(&d, 64);
  const int iters = 1 << 13, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    kern<ILP, MMA, FMA><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double warps = (double)blocks * threads / 32, n = (double)iters * ILP * warps;
  const double tf_mma = MMA ? n * 512 / (best * 1e-3) * 1e-12 : 0, tf_fma = FMA ? n * 64 / (best * 1e-3) * 1e-12 : 0;
  printf("%-28s ILP %2d: %8.3f ms  DMMA %6.2f TFLOP/s  DFMA %6.2f TFLOP/s  sum %6.2f\n", what, ILP, best, tf_mma, tf_fma, tf_mma + tf_fma);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
  run<8, true, false>("dmma m8n8k4", p.multiProcessorCount);
  run<16, true, false>("dmma m8n8k4", p.multiProcessorCount);
  run<8, false, true>("dfma", p.multiProcessorCount);
  run<16, false, true>("dfma", p.multiProcessorCount);
  run<8, true, true>("dmma + dfma interleaved", p.multiProcessorCount);
  run<16, true, true>("dmma + dfma interleaved", p.multiProcessorCount);
  return 0;
}
