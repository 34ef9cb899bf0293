// Stateless (Philox-counter) dropout pass shared by the decoder and encoder training paths; see dropout.cu.
#pragma once
#include "common.cuh"
namespace cqvad {
// out = (res ? res : 0) + keep(seed, site, index) * x / (1 - p);  n % 8 == 0;  in place when out == x
template <typename T>
int dropout_apply(const T* x, const T* res, T* out, long n, float p, uint64_t seed, uint32_t site, cudaStream_t st);
}  // namespace cqvad
