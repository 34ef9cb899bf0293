// CUDA-core GEMM (fp32 accumulate) with the shared epilogue: the fp32 parity path ("rel 1e-3, TF32 off") and the
// fallback for shapes the tcgen05 kernel does not take (K % 64 != 0, tiny N).  Also does the 3x3 conv as an
// implicit GEMM over the y-padded NHWC layout (models/detr/dab_transformer.py:81,90).
#include "common.cuh"
#include <stdlib.h>

namespace cqvad {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// grid (ceil(N/BN), ceil(M/BM)); 256 threads; thread (ty,tx) owns rows ty*4.., cols tx*4..
template <typename T, bool CONV>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long lda, const T* __restrict__ W,
                                                        T* __restrict__ C, long ldc, long M, int N, int K,
                                                        Epilogue epi, int cw) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long m0 = (long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  // loader mapping: 64 rows x 16 k = 256 threads x 4 consecutive k
  const int lr = tid / 4, lk = (tid % 4) * 4;
  const long a_row = m0 + lr;
  const int w_row = n0 + lr;
  const int a_x = CONV ? (int)(a_row % cw) : 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[4], rw[4];
  auto fetch = [&](int k0) {
    const int k = k0 + lk;
#pragma unroll
    for (int i = 0; i < 4; ++i) { ra[i] = 0.f; rw[i] = 0.f; }
    if (CONV) {
      const int tap = k >> 8, c = k & 255;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const long src = a_row + (long)dy * cw + dx;
      const int xx = a_x + dx;
      if (a_row < M && src >= 0 && src < M && xx >= 0 && xx < cw) load4<T>(A + src * lda + c, ra);
    } else {
      if (a_row < M && k < K) load4<T>(A + a_row * lda + k, ra);
    }
    if (w_row < N && k < K) load4<T>(W + (long)w_row * K + k, rw);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[buf][lk + i][lr] = ra[i]; Ws[buf][lk + i][lr] = rw[i]; }
  };

  fetch(0);
  stash(0);
  __syncthreads();
  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      const float4 bv = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * TN]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

  const T* res = reinterpret_cast<const T*>(epi.res);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long r = m0 + ty * TM + i;
    if (r >= M) continue;
    const bool zero_row = epi.zero_period > 0 && (int)(r % epi.zero_period) >= epi.zero_valid;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = n0 + tx * TN + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (epi.bias) v += epi.bias[c];
      if (epi.dual_gelu) {
        C[r * ldc + c] = from_f<T>(gelu_erf(v));
        reinterpret_cast<T*>(epi.c2)[r * ldc + c] = from_f<T>(gelu_grad(v));
        continue;
      }
      if (epi.act == CQVAD_ACT_RELU) v = fmaxf(v, 0.f);
      else if (epi.act == CQVAD_ACT_GELU) v = gelu_erf(v);
      if (epi.mul_mode) {
        const float a = to_f(reinterpret_cast<const T*>(epi.mul_aux)[r * ldc + c]);
        v *= epi.mul_mode == 1 ? (a > 0.f ? epi.mul_scale : 0.f) : (epi.mul_mode == 3 ? a : gelu_grad(a));
      }
      if (epi.res32) v += epi.res32[r * epi.ldr + c];
      else if (res) v += to_f(res[r * epi.ldr + c]);
      if (zero_row) v = 0.f;
      C[r * ldc + c] = from_f<T>(v);
      if (epi.c2) reinterpret_cast<T*>(epi.c2)[r * ldc + c] = from_f<T>(epi.c2_act == CQVAD_ACT_GELU ? gelu_erf(v) : (epi.c2_act == CQVAD_ACT_RELU ? fmaxf(v, 0.f) : v));
      if (epi.c32) epi.c32[r * ldc + c] = v;
    }
  }
}

}  // namespace

template <typename T>
int gemm_simt(const T* A, long lda, const T* W, T* C, long ldc, long M, int N, int K, const Epilogue& epi,
              const ConvGeom* conv, cudaStream_t st) {
  CQ_CHECK_SHAPE(K % 4 == 0 && lda % 4 == 0, "gemm_simt: K (%d) and lda (%ld) must be multiples of 4", K, lda);
  if (M == 0 || N == 0) return 0;
  dim3 grid((unsigned)cdiv(N, BN), (unsigned)cdiv(M, BM));
  CQ_CHECK_SHAPE(grid.y <= 65535u * 32u, "gemm_simt: M too large");
  Epilogue e = epi;
  e.ln_g = nullptr;  // LayerNorm (if any) is applied by a second kernel below
  if (epi.ln_g) e.c32 = nullptr;   // the fp32 copy must hold the normalised values: only the tensor-core epilogue fuses that
  if (conv) {
    CQ_CHECK_SHAPE(K == 9 * kC, "conv: K must be 9*256");
    gemm_simt_kernel<T, true><<<grid, 256, 0, st>>>(A, lda, W, C, ldc, M, N, K, e, conv->w);
  } else {
    gemm_simt_kernel<T, false><<<grid, 256, 0, st>>>(A, lda, W, C, ldc, M, N, K, e, 0);
  }
  CQ_LAUNCH_CHECK();
  if (epi.ln_g) {
    CQ_CHECK_SHAPE(N == kC && ldc == kC, "fused LayerNorm needs N == ldc == 256");
    CQ_TRY(layernorm_rows<T>(C, nullptr, epi.ln_g, epi.ln_b, epi.ln_eps, C, false, M, st));
  }
  return 0;
}

template int gemm_simt<float>(const float*, long, const float*, float*, long, long, int, int, const Epilogue&,
                              const ConvGeom*, cudaStream_t);
template int gemm_simt<bf16>(const bf16*, long, const bf16*, bf16*, long, long, int, int, const Epilogue&,
                             const ConvGeom*, cudaStream_t);

static bool g_force_simt = false;
void set_force_simt(bool v) { g_force_simt = v; }
bool force_simt() { return g_force_simt; }

template <>
int gemm<float>(const float* A, long lda, const float* W, float* C, long ldc, long M, int N, int K, const Epilogue& epi,
                const ConvGeom* conv, cudaStream_t st) {
  return gemm_simt<float>(A, lda, W, C, ldc, M, N, K, epi, conv, st);
}

template <>
int gemm<bf16>(const bf16* A, long lda, const bf16* W, bf16* C, long ldc, long M, int N, int K, const Epilogue& epi,
               const ConvGeom* conv, cudaStream_t st) {
  static const long small_m = [] { const char* e = getenv("CQVAD_SMALL_M"); return e ? atol(e) : 0L; }();
  if (!g_force_simt && !(M <= small_m && !conv)) {
    int r = gemm_tc(A, lda, W, C, ldc, M, N, K, epi, conv, st);
    if (r <= 0) return r;  // 0 ok, <0 error; 1 = shape not supported by the tensor-core kernel
  }
  return gemm_simt<bf16>(A, lda, W, C, ldc, M, N, K, epi, conv, st);
}

}  // namespace cqvad
