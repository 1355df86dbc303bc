// lg_srgb.h — the 8-bit surface encoding of ORACLE.md 8.7 (host + device; no kernels, so tests/host_geom_check.cu
// can hold it against the oracle on the CPU).  Used by surface_bgra8_srgb_kernel (lg_accum.cuh).
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define LG_SRGB_HD __host__ __device__ __forceinline__
#else
#define LG_SRGB_HD inline
#endif

namespace lg {

struct SrgbThresholds {
  float t[256]; // t[0] = -inf (unused), t[1..255] ascending
};
inline SrgbThresholds srgb_thresholds() {
  SrgbThresholds T;
  T.t[0] = -INFINITY;
  for (int k = 1; k < 256; ++k) {
    const double v = ((double)k - 0.5) / 255.0;
    const double lin = v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4);
    T.t[k] = (float)lin;
  }
  return T;
}
LG_SRGB_HD unsigned char srgb_byte(const float *t, float c) {
  // largest k with t[k] <= c (t[0] = -inf; NaN compares false everywhere -> 0)
  int k = 0;
#pragma unroll
  for (int step = 128; step > 0; step >>= 1)
    if (t[k + step] <= c) k += step;
  return (unsigned char)k;
}
LG_SRGB_HD unsigned char unorm_byte(float a) {
  if (!(a > 0.f)) return 0; // NaN -> 0, like the ROP's clamp
  return (unsigned char)(fminf(a, 1.f) * 255.f + 0.5f);
}
} // namespace lg
