// Dev tool: cycles per tcgen05.mma (kind::f16, bf16 operands in 128-byte-swizzled shared memory, fp32 accumulators in TMEM) by
// instruction shape, single CTA (cta_group::1, M = 128) and CTA pair (cta_group::2, M = 256), back-to-back accumulating issues of
// ONE thread, timed from the first issue to the completion mbarrier.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I class_query_vad_b200/csrc tools/umma_bench.cu -o tools/build/umma_bench -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
typedef __nv_bfloat16 bf16;
#include "tc_common.cuh"
using namespace cqvad::tc;

struct Args { int n1, n2, rounds; long long* out; };

template <bool PAIR>
__device__ void body(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 65536, bar = base + 65536 + 131072, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 1) { if constexpr (PAIR) tmem_alloc_pair(slot, 512); else tmem_alloc(slot, 512); }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 0 && lane == 0 && rank == 0) {
    const uint32_t id1 = make_idesc_bf16(PAIR ? 256 : 128, a.n1), id2 = make_idesc_bf16(PAIR ? 256 : 128, a.n2);
    long long t0 = clock64();
    for (int r = 0; r < a.rounds; ++r) {
      for (int kb = 0; kb < 4; ++kb) {                      // 16 UMMAs of shape 1 (A: 4 k-blocks of [128 x 64], B the same)
        const uint64_t ad = make_smem_desc_sw128(sA + kb * 16384), bd = make_smem_desc_sw128(sB + kb * 32768);
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) umma_bf16_pair(tmem + 256, ad + 2 * k, bd + 2 * k, id1, 1u); else umma_bf16(tmem + 256, ad + 2 * k, bd + 2 * k, id1, 1u);
        }
      }
      if (a.n2 > 0) {
        const uint64_t ad = make_smem_desc_sw128(sA), bd = make_smem_desc_sw128(sB);
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) umma_bf16_pair(tmem, ad + 2 * k, bd + 2 * k, id2, 1u); else umma_bf16(tmem, ad + 2 * k, bd + 2 * k, id2, 1u);
        }
      }
    }
    long long t1 = clock64();
    if constexpr (PAIR) umma_commit_pair(bar); else umma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    a.out[0] = t1 - t0; a.out[1] = t2 - t0;
  } else if (PAIR && warp == 0 && lane == 0) {
    mbar_wait(bar, 0);     // the multicast commit
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) { __syncwarp(); tc_fence_after(); if constexpr (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}
__global__ void __launch_bounds__(128, 1) k_single(const Args a) { body<false>(a); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_pair(const Args a) { body<true>(a); }

int main() {
  const int smem = 65536 + 131072 + 1024 + 64;
  cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out; cudaMalloc(&out, 16);
  const int rounds = 64;
  const int shapes[][2] = {{64, 0}, {128, 0}, {256, 0}, {64, 256}, {128, 256}, {32, 0}, {16, 0}};
  for (int pair = 0; pair < 2; ++pair)
    for (auto& s : shapes) {
      Args a{s[0], s[1], rounds, out};
      for (int rep = 0; rep < 2; ++rep) {
        if (pair) k_pair<<<2, 128, smem>>>(a); else k_single<<<1, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const int n_mma = rounds * (16 + (s[1] ? 4 : 0));
      printf("%s M=%d: %2d x N=%3d%s : issue %7.1f cyc/round, complete %7.1f cyc/round (%.1f per UMMA; math time %d)\n", pair ? "pair  " : "single",
             pair ? 256 : 128, 16, s[0], s[1] ? " + 4 x N=256" : "            ", (double)h[0] / rounds, (double)h[1] / rounds, (double)h[1] / n_mma,
             (16 * s[0] + 4 * s[1]) / 8);
    }
  return 0;
}
