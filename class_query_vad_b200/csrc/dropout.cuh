// Stateless (Philox-counter) dropout pass shared by the decoder and encoder training paths; see dropout.cu.
#pragma once
#include "common.cuh"
namespace cqvad {
// out = (res ? res : 0) + keep(seed, site, index) * x / (1 - p);  n % 8 == 0;  in place when out == x
// the scale dropout_apply gives kept elements (p is quantised to 1/65536)
inline float dropout_keep_scale(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t > 65535.0) t = 65535.0;
  const unsigned thr16 = (unsigned)t;
  return 1.f / (1.f - (float)thr16 / 65536.f);
}
template <typename T>
int dropout_apply(const T* x, const T* res, T* out, long n, float p, uint64_t seed, uint32_t site, cudaStream_t st);
}  // namespace cqvad
