// Optimizer step of the reference training loop on flat fp32 buffers (train.py:83,158-167):
//   torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)   -> global L2 norm over ALL gradients, coef = min(1, max_norm / (norm + 1e-6))
//   torch.optim.AdamW(lr, betas, eps, weight_decay).step()         -> decoupled weight decay, bias-corrected moments
// Two launches over one flat buffer (the reference launches the multi-tensor foreach kernels over ~700 tensors):
//   adamw_sumsq_kernel : grid = 4 x SMs, float4 loads, one fp64 partial per block (fixed order => deterministic norm)
//   adamw_step_kernel  : every block re-reduces the <= 1024 partials (L2-resident, 8 KB), derives the clip coefficient and the
//                        gradient pre-scale (1 / world size of the all-reduce), then streams p, g, m, v once: 16 B read + 12 B written
//                        per parameter = the HBM floor of the update; optionally refreshes a bf16 copy of the parameter in the same pass
//                        (the tensor-core weight operand) and zero-fills the gradient for the next step.
#include <algorithm>
#include <math.h>
#include "common.cuh"

namespace cqvad {
namespace {

constexpr int kOptThreads = 256;
constexpr int kMaxPartials = 1024;

__global__ void __launch_bounds__(kOptThreads) adamw_sumsq_kernel(const float* __restrict__ g, long n, double* __restrict__ partials) {
  __shared__ double red[kOptThreads / 32];
  double acc = 0.0;
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc += (double)g[i] * (double)g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
  float bc1, bc2_sqrt;     // 1 - beta1^t, sqrt(1 - beta2^t)
  int n_partials, zero_grad;
};

__global__ void __launch_bounds__(kOptThreads) adamw_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                                 float* __restrict__ v, bf16* __restrict__ p16, long n, AdamArgs a,
                                                                 const double* __restrict__ partials, float* __restrict__ norm_out) {
  __shared__ double red[kOptThreads / 32];
  __shared__ float s_coef;
  {
    double acc = 0.0;
    for (int i = threadIdx.x; i < a.n_partials; i += blockDim.x) acc += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < kOptThreads / 32; ++i) s += red[i];
      const float norm = (float)sqrt(s) * a.grad_scale;                      // norm of the averaged gradient
      float coef = 1.f;
      if (a.max_norm > 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.f);  // clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max = 1)
      s_coef = coef * a.grad_scale;
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
    __syncthreads();
  }
  const float gs = s_coef;
  const float step_size = a.lr / a.bc1;
  const float decay = 1.f - a.lr * a.weight_decay;
  const long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gs;
    pp *= decay;                                        // p.mul_(1 - lr * weight_decay)
    mm = a.beta1 * mm + (1.f - a.beta1) * gg;           // exp_avg.lerp_(grad, 1 - beta1)
    vv = a.beta2 * vv + (1.f - a.beta2) * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
    pp -= step_size * (mm / denom);                     // p.addcdiv_(exp_avg, denom, value = -step_size)
  };
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (a.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 r; r.x = *reinterpret_cast<unsigned*>(&lo); r.y = *reinterpret_cast<unsigned*>(&hi);
      reinterpret_cast<uint2*>(p16)[i] = r;
    }
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      upd(p[i], g[i], m[i], v[i]);
      if (a.zero_grad) g[i] = 0.f;
      if (p16) p16[i] = __float2bfloat16_rn(p[i]);
    }
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

extern "C" size_t cqvad_adamw_workspace_bytes(void) { return kMaxPartials * sizeof(double) + 16; }

extern "C" int cqvad_adamw_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, long n,
                                     float lr, float beta1, float beta2, float eps, float weight_decay, long step, float max_norm,
                                     float grad_scale, int zero_grad, float* grad_norm_out, void* workspace, size_t ws_bytes,
                                     void* stream) {
  CQ_CHECK_ARG(n >= 0 && step >= 1, "adamw_clip_step: n >= 0 and step >= 1 (1-based count of this update)");
  if (n == 0) return 0;
  CQ_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && workspace, "adamw_clip_step: null pointer");
  CQ_CHECK_ARG(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0 &&
               ((uintptr_t)params_bf16 % 8) == 0, "adamw_clip_step: buffers must be 16-byte aligned");
  if (ws_bytes < cqvad_adamw_workspace_bytes()) return set_error(CQVAD_E_WORKSPACE, "adamw_clip_step: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  const int blocks = (int)std::min<long>(std::min<long>(kMaxPartials, 4L * num_sms()), std::max<long>(1, (n / 4 + kOptThreads - 1) / kOptThreads));
  adamw_sumsq_kernel<<<blocks, kOptThreads, 0, st>>>(grads, n, partials);
  CQ_LAUNCH_CHECK();
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
  a.grad_scale = grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.n_partials = blocks; a.zero_grad = zero_grad ? 1 : 0;
  const int sblocks = (int)std::min<long>(8L * num_sms(), std::max<long>(1, (n / 4 + kOptThreads - 1) / kOptThreads));
  adamw_step_kernel<<<sblocks, kOptThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (bf16*)params_bf16, n, a, partials, grad_norm_out);
  CQ_LAUNCH_CHECK();
  return 0;
}
