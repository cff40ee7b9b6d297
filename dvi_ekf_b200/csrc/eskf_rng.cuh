// Counter-based Gaussian noise for the Monte-Carlo extension of the engine
// (the reference itself is noise free; its only np.random call is dead code in
// dvi_ekf/models/trajectory/ImuTrajectory.py:133).  Philox4x32-10 keyed by the
// run seed, counter = (step, kind*16 + draw, filter id lo, filter id hi), so a
// filter's noise depends only on its GLOBAL id: results are independent of how
// the batch is sharded over CTAs or GPUs.  Box-Muller in FP64 so that the
// numpy replica in tests/ reproduces the stream to rounding.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef ESKF_HD
#ifdef __CUDACC__
#define ESKF_HD __host__ __device__ __forceinline__
#else
#define ESKF_HD inline
#endif
#endif

namespace eskf {

constexpr uint32_t RNG_KIND_IMU = 1;   // 6 normals per IMU step
constexpr uint32_t RNG_KIND_CAM = 2;   // 6 normals per camera update (position, orientation)
constexpr uint32_t RNG_KIND_CAM2 = 3;  // 2 normals per camera update (notch, spare)

ESKF_HD void philox4x32_10(uint32_t* c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// two standard normals from one Philox block
ESKF_HD void normal_pair(uint64_t seed, uint64_t filter, uint64_t step, uint32_t kind, uint32_t draw, double* z) {
  uint32_t c[4] = {(uint32_t)step, kind * 16u + draw + ((uint32_t)(step >> 32) << 8), (uint32_t)filter,
                   (uint32_t)(filter >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)c[1] << 32) | c[0];
  const uint64_t b = ((uint64_t)c[3] << 32) | c[2];
  const double u1 = ((double)(a >> 11) + 0.5) * 1.1102230246251565e-16;  // (0,1)
  const double u2 = ((double)(b >> 11) + 0.5) * 1.1102230246251565e-16;
  const double r = sqrt(-2.0 * log(u1));
  double s, co;
  sincos(6.283185307179586476925286766559 * u2, &s, &co);
  z[0] = r * co;
  z[1] = r * s;
}

ESKF_HD void normal6(uint64_t seed, uint64_t filter, uint64_t step, uint32_t kind, double* z) {
  normal_pair(seed, filter, step, kind, 0, z);
  normal_pair(seed, filter, step, kind, 1, z + 2);
  normal_pair(seed, filter, step, kind, 2, z + 4);
}
ESKF_HD void normal2(uint64_t seed, uint64_t filter, uint64_t step, uint32_t kind, double* z) {
  normal_pair(seed, filter, step, kind, 0, z);
}

}  // namespace eskf
