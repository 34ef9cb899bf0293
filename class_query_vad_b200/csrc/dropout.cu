// Dropout of the native training path (reference: nn.Dropout(p) on every residual branch and FFN hidden activation,
// models/detr/dab_transformer.py:499-519,937,991,995,1043-1044,1062,1076-1077).  The keep mask is never stored: one Philox4x32-10
// call (rng.cuh) yields eight 16-bit uniforms for eight consecutive elements of dropout site `site`, so the backward regenerates
// the forward's mask from (seed, site, element index).  out = (res ? res : 0) + keep * x / (1 - p); in place when out == x.
// HBM-bound elementwise pass: 8 elements (one 16-byte bf16 vector) per thread, grid-stride.
#include <algorithm>
#include "common.cuh"
#include "rng.cuh"
#include "dropout.cuh"

namespace cqvad {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ out, long n8,
                                                      float scale, unsigned thr16, uint64_t seed, uint32_t site) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const uint4 r = dropout_bits(seed, site, (uint64_t)i);
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
    float v[8], o[8];
    load8(x + i * 8, v);
    if (res) load8(res + i * 8, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned u = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
      const float d = u >= thr16 ? v[e] * scale : 0.f;
      o[e] = res ? o[e] + d : d;
    }
    store8(out + i * 8, o);
  }
}

}  // namespace

template <typename T>
int dropout_apply(const T* x, const T* res, T* out, long n, float p, uint64_t seed, uint32_t site, cudaStream_t st) {
  if (n == 0) return 0;
  CQ_CHECK_ARG(n % 8 == 0, "dropout: element count must be a multiple of 8");
  CQ_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  const unsigned thr16 = (unsigned)std::min(65535.0, (double)p * 65536.0 + 0.5);   // keep iff u16 >= thr16: P(keep) = 1 - thr16 / 65536
  const float scale = dropout_keep_scale(p);
  const long n8 = n / 8;
  const unsigned grid = (unsigned)std::min<long>(cdiv(n8, 256), 148L * 16);
  dropout_kernel<T><<<grid, 256, 0, st>>>(x, res, out, n8, scale, thr16, seed, site);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int dropout_apply<float>(const float*, const float*, float*, long, float, uint64_t, uint32_t, cudaStream_t);
template int dropout_apply<bf16>(const bf16*, const bf16*, bf16*, long, float, uint64_t, uint32_t, cudaStream_t);

}  // namespace cqvad

using namespace cqvad;

extern "C" int cqvad_dropout(int dtype, const void* x, const void* res, void* out, long n, float p, uint64_t seed, uint32_t site,
                             void* stream) {
  CQ_CHECK_ARG(n >= 0 && (n == 0 || (x && out)), "dropout: null pointer");
  if (dtype == CQVAD_F32) return dropout_apply<float>((const float*)x, (const float*)res, (float*)out, n, p, seed, site, as_stream(stream));
  if (dtype == CQVAD_BF16) return dropout_apply<bf16>((const bf16*)x, (const bf16*)res, (bf16*)out, n, p, seed, site, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "dropout: unknown dtype %d", dtype);
}
