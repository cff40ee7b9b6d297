// Micro-benchmarks behind the kernel design (run on the B200: nvcc -arch=sm_100a -O3 fp64_lat.cu && ./a.out):
//   * DFMA dependent-issue latency and per-sub-partition throughput as a function of warps x ILP
//   * LDS.128 load-to-use latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void lds_chain(int* out, long long* cyc, int iters) {
  __shared__ int4 buf[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = make_int4((i * 7 + 1) & 255, 0, 0, 0);
  __syncthreads();
  int idx = threadIdx.x & 255;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) idx = buf[idx].x;
  const long long t1 = clock64();
  if (idx == 12345) out[0] = idx;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps_per_smsp) {
  double* d;
  long long* c;
  cudaMalloc(&d, 64);
  cudaMalloc(&c, 8 * 1024);
  const int iters = 4096;
  const int threads = 32 * 4 * warps_per_smsp;  // warps are dealt round-robin to the 4 sub-partitions
  dfma_chain<ILP><<<1, threads>>>(d, c, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  dfma_chain<ILP><<<1, threads>>>(d, c, iters, 1.0000001, 1e-9);
  long long h;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double per_iter = (double)h / iters;
  printf("ILP %2d warps/SMSP %d: %.2f cycles per iteration, %.2f cycles per warp-DFMA, SMSP issue interval %.2f cycles\n", ILP,
         warps_per_smsp, per_iter, per_iter / ILP, per_iter / ILP / warps_per_smsp);
  cudaFree(d);
  cudaFree(c);
}

int main() {
  for (int w = 1; w <= 4; ++w) {
    run<1>(w);
    run<2>(w);
    run<4>(w);
    run<8>(w);
    run<9>(w);
    run<12>(w);
    run<16>(w);
    run<18>(w);
  }
  int* o;
  long long* c;
  cudaMalloc(&o, 64);
  cudaMalloc(&c, 64);
  lds_chain<<<1, 32>>>(o, c, 4096);
  cudaDeviceSynchronize();
  lds_chain<<<1, 32>>>(o, c, 4096);
  long long h;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("LDS.128 dependent load-to-use: %.1f cycles\n", (double)h / 4096);
  return 0;
}
