// Counter-based RNG for dropout masks: Philox4x32-10 (Salmon et al., SC'11), keyed by the caller's 64-bit seed, counter =
// (element index / 4, site id).  Stateless: the backward regenerates the forward's mask from (seed, site, index) instead of
// storing it, so a dropout site costs no HBM traffic of its own.
#pragma once
#include <stdint.h>

namespace cqvad {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// 4 uniform 32-bit words for elements 4*q .. 4*q+3 of dropout site `site`
__device__ __forceinline__ uint4 dropout_bits(uint64_t seed, uint32_t site, uint64_t q) {
  return philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), site, 0x5EEDu), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// keep decision for one 32-bit word: keep with probability 1 - p  (threshold = p * 2^32)
__device__ __forceinline__ bool dropout_keep(uint32_t word, uint32_t threshold) { return word >= threshold; }
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

}  // namespace cqvad
