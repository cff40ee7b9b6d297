// One translation unit per CTA shape: compiled with -DESKF_F=<filters per CTA> so the
// instantiations build in parallel (see dvi_ekf_b200/build.py).
#include "eskf_kernel.cuh"

#ifndef ESKF_F
#error "compile with -DESKF_F=<filters per CTA>"
#endif

namespace eskf {

template <>
cudaError_t launch_eskf_kernel<ESKF_F>(const KArgs& a, cudaStream_t stream) {
  constexpr int F = ESKF_F;
  static_assert((8 * F) % 32 == 0, "covariance role must fill whole warps");
  const size_t smem = (size_t)F * SM_PER_FILTER * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(eskf_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((a.N + F - 1) / F);
  eskf_kernel<F><<<grid, 32 + 8 * F, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace eskf
