// Shared helpers for libcqvad (sm_100a).  Internal header; the public contract is include/cqvad.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/cqvad.h"

typedef __nv_bfloat16 bf16;

namespace cqvad {

// ---- error plumbing -------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
void reset_launch_count();

#define CQ_CHECK_ARG(cond, ...)                                                  \
  do {                                                                           \
    if (!(cond)) return cqvad::set_error(CQVAD_E_INVALID_ARG, __VA_ARGS__);      \
  } while (0)
#define CQ_CHECK_SHAPE(cond, ...)                                                \
  do {                                                                           \
    if (!(cond)) return cqvad::set_error(CQVAD_E_UNSUPPORTED_SHAPE, __VA_ARGS__);\
  } while (0)
#define CQ_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return cqvad::set_error(CQVAD_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)
#define CQ_LAUNCH_CHECK()                                                                               \
  do {                                                                                                  \
    cqvad::count_launch();                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ != cudaSuccess)                                                                             \
      return cqvad::set_error(CQVAD_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)
#define CQ_TRY(expr)            \
  do {                          \
    int r__ = (expr);           \
    if (r__ != 0) return r__;   \
  } while (0)

// ---- dtype helpers --------------------------------------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float> { static constexpr int id = CQVAD_F32; };
template <> struct DT<bf16> { static constexpr int id = CQVAD_BF16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats (rows are 256 wide: one warp = one row, lane owns channels 8*lane..8*lane+7)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm of one 256-wide row held as 8 values per lane (two-pass in registers, like F.layer_norm's fp32 math).
__device__ __forceinline__ void warp_layernorm256(float (&v)[8], const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + eps);
  float g[8], b[8];
  load8(gamma + lane * 8, g);
  load8(beta + lane * 8, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}

constexpr int kC = 256;  // d_model (fixed by every shipped config: configuration/*.yaml D_MODEL 256)
constexpr int kH = 8;    // heads
constexpr int kL = 4;    // feature levels

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// Auxiliary (non-blocking) streams and timing-free events of the multi-stream schedules, created lazily PER DEVICE (a process may
// drive several GPUs): slot = purpose.  One host thread per device is assumed (the events are reused from call to call).
cudaStream_t aux_stream(int slot);   // slot in [0, 8)
cudaEvent_t aux_event(int slot);     // slot in [0, 96)
inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// ---- internal entry points shared between translation units -----------------------------------------------------
struct Epilogue {
  const float* bias = nullptr;   // [N]
  int act = CQVAD_ACT_NONE;
  const void* res = nullptr;     // [M,N] same dtype as C, added after the activation
  long ldr = 0;
  const float* ln_g = nullptr;   // LayerNorm over N (requires N == 256) applied last
  const float* ln_b = nullptr;
  float ln_eps = 1e-5f;
  // rows of C that must be written as zeros (y-padded NHWC separator rows): row % period >= valid  (period 0 = off)
  int zero_period = 0, zero_valid = 0;
  // fp32 side channels of the class-token stream (bf16 path): residual read in fp32 instead of `res`, and an fp32 copy
  // of C (same leading dimension), so that the values feeding cls_norm_/cls_norm2 skip two un-averaged bf16 roundings
  const float* res32 = nullptr;
  float* c32 = nullptr;
  // training path: (a) second output c2 = act2(v) next to C = v (pre-activation kept for the backward, same ldc);
  // (b) backward through an activation fused into the data-gradient GEMM: v *= act'(aux[row, col]) with aux of C's
  //     shape / ldc -- mul_mode 1: ReLU mask (aux = post-activation, > 0), 2: exact erf-GELU derivative (aux = pre-activation),
  //     3: aux holds the derivative itself (stored by the forward gelu pass)
  void* c2 = nullptr;
  int c2_act = CQVAD_ACT_NONE;
  const void* mul_aux = nullptr;
  int mul_mode = 0;
  // mul_mode 1 only: kept elements are additionally scaled (the 1 / (1 - p) of a dropout that followed the ReLU: the dropped
  // activation doubles as the mask, so the backward needs no dropout pass of its own)
  float mul_scale = 1.f;
  // training forward of a GELU layer in ONE epilogue: C = gelu(v), c2 = gelu'(v) with v = acc + bias (act / c2_act ignored);
  // the pre-activation itself is never written (the backward needs only gelu' -- mul_mode 3 -- and the activation)
  bool dual_gelu = false;
  // MSDA "prepare" folded into the two query projections (tcgen05 path only; fp32 output c32 ONLY through TMA stores -- C is not
  // written; N % 32 == 0, ldc counts c32 elements):
  //   rowop 1: softmax over every aligned group of 32 columns (one head's L*P = 32 attention logits)
  //   rowop 2: sampling locations for L = 4 levels x P = 8 points, column n = ((m*4 + l)*8 + p)*3 + i:
  //            c32[r, n] = ro_ref[r*12 + l*3 + i] + (acc + bias) / ro_norm[l*3 + i]   (IEEE division)
  int rowop = 0;
  const float* ro_ref = nullptr;    // [M, 12]
  const float* ro_norm = nullptr;   // [12]
};
struct ConvGeom {  // implicit-GEMM 3x3 conv on the y-padded NHWC layout [n_img, h+1, w, 256]
  int h = 0, w = 0;
};

// C[M,N] = epi(A[M,K] . W[N,K]^T).  conv != nullptr: A is the padded activation [M,256], K = 9*256.
template <typename T>
int gemm(const T* A, long lda, const T* W, T* C, long ldc, long M, int N, int K, const Epilogue& epi,
         const ConvGeom* conv, cudaStream_t st);
template <typename T>
int gemm_simt(const T* A, long lda, const T* W, T* C, long ldc, long M, int N, int K, const Epilogue& epi,
              const ConvGeom* conv, cudaStream_t st);
// tcgen05 path (bf16 only); returns 1 if the shape is not supported (caller falls back to gemm_simt)
int gemm_tc(const bf16* A, long lda, const bf16* W, bf16* C, long ldc, long M, int N, int K, const Epilogue& epi,
            const ConvGeom* conv, cudaStream_t st);
// fused MLP  Y = LN?( res? + act(X.W1^T + b1).W2^T + b2 )  (tcgen05; returns 1 if unsupported)
int mlp_tc(const bf16* X, const bf16* W1, const float* b1, const bf16* W2, const float* b2, int act, const bf16* res,
           const float* ln_g, const float* ln_b, float ln_eps, bf16* Y, long M, int C, int F, int zero_period,
           int zero_valid, cudaStream_t st, bf16* YT = nullptr, long ldyt = 0, int yt_rows = 0, int yt_pitch = 0,
           const float* res32 = nullptr, float* Y32 = nullptr);

template <typename T>
int layernorm_rows(const T* x, const T* res, const float* g, const float* b, float eps, void* out, bool out_f32,
                   long rows, cudaStream_t st);

template <typename T>
int mha_core(int mode, const T* q, const T* k, const T* v, const uint8_t* kpm, T* o, int L, int S, int Nb, int H, int E,
             int Ev, long q_ls, long q_bs, long k_ls, long k_bs, long k_qs, long v_ls, long v_bs, long v_qs, long o_ls,
             long o_bs, cudaStream_t st);

}  // namespace cqvad
