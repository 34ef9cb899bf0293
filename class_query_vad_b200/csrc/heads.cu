// DETR heads of the reference model for the TRAINING step (models/model.py:191-236), forward and backward, fp32 (the reference
// runs them with autocast disabled):
//   pred_logits_b = class_embed_b(hs)                                                   (:192)
//   pred_boxes    = sigmoid(bbox_embed(hs)[..., :4] + inverse_sigmoid(reference))       (:195-199; bbox_embed = MLP 256-256-256-4, ReLU)
//   pred_logits   = Dropout(0.5)(cls_hs).mean(-1)                                       (:103,219-221)
// Rows r = (layer, clip, query) in the order of the decoder outputs hs / refs [Lr, BT, nq, *]; cls_hs [Lr, BT, nq, K, 256].
// Small-row work (R = Lr * BT * nq <= a few thousand rows of 256): one block per row for the MLP (warp per output group,
// coalesced weight reads, shuffle reductions), one warp per class token for the channel mean.  The dropout mask is regenerated
// from a Philox counter in the backward (rng.cuh), nothing is stored for it.
#include <algorithm>
#include "common.cuh"
#include "rng.cuh"

namespace cqvad {
namespace {

constexpr uint32_t kSiteClsHs = 0x100;      // dropout site id of models/model.py:103

__device__ __forceinline__ float inv_sigmoid(float x) {        // utils/misc.py:530-534
  x = fminf(fmaxf(x, 0.f), 1.f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
}
__device__ __forceinline__ float inv_sigmoid_grad(float x) {   // d/dx log(clamp(x, eps) / clamp(1 - x, eps)), x clamped to [0, 1]
  if (x < 0.f || x > 1.f) return 0.f;
  float g = 0.f;
  if (x >= 1e-5f) g += 1.f / x;
  if (1.f - x >= 1e-5f) g += 1.f / (1.f - x);
  return g;
}

// y[o] = act(b[o] + sum_i W[o][i] x[i]) for o in [0, nout): warp w handles outputs w, w + 8, ...; x in shared memory
__device__ __forceinline__ void row_linear(const float* __restrict__ W, const float* __restrict__ b, const float* x, float* y, int nout,
                                           bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < nout; o += 8) {
    const float* w = W + (long)o * kC;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < kC / 32; ++i) a = fmaf(w[lane + 32 * i], x[lane + 32 * i], a);
    a = warp_sum(a);
    if (lane == 0) { a += b[o]; y[o] = relu ? fmaxf(a, 0.f) : a; }
  }
}

struct HeadW { const float *w1, *b1, *w2, *b2, *w3, *b3, *wb, *bb; };
struct HeadG { float *w1, *b1, *w2, *b2, *w3, *b3, *wb, *bb; };

__global__ void __launch_bounds__(256) heads_fwd_kernel(HeadW w, const float* __restrict__ hs, const float* __restrict__ refs,
                                                        float* __restrict__ h1s, float* __restrict__ h2s, float* __restrict__ boxes,
                                                        float* __restrict__ logits_b) {
  __shared__ float x[kC], h1[kC], h2[kC], t[8];
  const long r = blockIdx.x;
  x[threadIdx.x] = hs[r * kC + threadIdx.x];
  __syncthreads();
  row_linear(w.w1, w.b1, x, h1, kC, true);
  row_linear(w.wb, w.bb, x, t + 4, 3, false);
  __syncthreads();
  row_linear(w.w2, w.b2, h1, h2, kC, true);
  __syncthreads();
  row_linear(w.w3, w.b3, h2, t, 4, false);
  __syncthreads();
  h1s[r * kC + threadIdx.x] = h1[threadIdx.x];
  h2s[r * kC + threadIdx.x] = h2[threadIdx.x];
  if (threadIdx.x < 4) boxes[r * 4 + threadIdx.x] = 1.f / (1.f + expf(-(t[threadIdx.x] + inv_sigmoid(refs[r * 4 + threadIdx.x]))));
  if (threadIdx.x >= 4 && threadIdx.x < 7) logits_b[r * 3 + threadIdx.x - 4] = t[threadIdx.x];
}

// pred_logits[row] = mean_c( keep(row, c) * cls_hs[row, c] / (1 - p) ); one warp per class token
__global__ void __launch_bounds__(256) cls_mean_fwd_kernel(const float* __restrict__ cls_hs, float* __restrict__ logits, long rows, float p,
                                                           uint32_t thr, uint64_t seed) {
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  load8(cls_hs + row * kC + lane * 8, v);
  float a = 0.f;
  if (p > 0.f) {
    const uint64_t q = (uint64_t)row * (kC / 4) + lane * 2;
    const uint4 r0 = dropout_bits(seed, kSiteClsHs, q), r1 = dropout_bits(seed, kSiteClsHs, q + 1);
    const uint32_t bits[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) a += dropout_keep(bits[e], thr) ? v[e] : 0.f;
    a *= 1.f / (1.f - p);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) a += v[e];
  }
  a = warp_sum(a);
  if (lane == 0) logits[row] = a * (1.f / kC);
}

__global__ void __launch_bounds__(256) cls_mean_bwd_kernel(const float* __restrict__ g_logits, float* __restrict__ g_cls, long rows, float p,
                                                           uint32_t thr, uint64_t seed) {
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float g = g_logits[row] * (1.f / kC) * (p > 0.f ? 1.f / (1.f - p) : 1.f);
  float v[8];
  if (p > 0.f) {
    const uint64_t q = (uint64_t)row * (kC / 4) + lane * 2;
    const uint4 r0 = dropout_bits(seed, kSiteClsHs, q), r1 = dropout_bits(seed, kSiteClsHs, q + 1);
    const uint32_t bits[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = dropout_keep(bits[e], thr) ? g : 0.f;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = g;
  }
  store8(g_cls + row * kC + lane * 8, v);
}

// y[c] = sum_o W[o][c] * d[o] (o < nout) -- thread c, coalesced over c
__device__ __forceinline__ float col_dot(const float* __restrict__ W, const float* d, int nout) {
  float a = 0.f;
  for (int o = 0; o < nout; ++o) a = fmaf(W[(long)o * kC + threadIdx.x], d[o], a);
  return a;
}

__global__ void __launch_bounds__(256) heads_bwd_kernel(HeadW w, HeadG gw, const float* __restrict__ refs, const float* __restrict__ h1s,
                                                        const float* __restrict__ h2s, const float* __restrict__ boxes,
                                                        const float* __restrict__ g_boxes, const float* __restrict__ g_logits_b,
                                                        float* __restrict__ dh1s, float* __restrict__ dh2s, float* __restrict__ dts,
                                                        float* __restrict__ g_hs, float* __restrict__ g_refs) {
  __shared__ float dt[8], dh2[kC], dh1[kC];
  const long r = blockIdx.x;
  const int c = threadIdx.x;
  if (c < 4) {
    const float b = boxes[r * 4 + c];
    const float d = (g_boxes ? g_boxes[r * 4 + c] : 0.f) * b * (1.f - b);
    dt[c] = d;
    if (g_refs) g_refs[r * 4 + c] = d * inv_sigmoid_grad(refs[r * 4 + c]);
  } else if (c < 7) {
    dt[c] = g_logits_b ? g_logits_b[r * 3 + c - 4] : 0.f;
  } else if (c == 7) {
    dt[7] = 0.f;
  }
  __syncthreads();
  if (c < 8) dts[r * 8 + c] = dt[c];
  dh2[c] = h2s[r * kC + c] > 0.f ? col_dot(w.w3, dt, 4) : 0.f;
  __syncthreads();
  dh1[c] = h1s[r * kC + c] > 0.f ? col_dot(w.w2, dh2, kC) : 0.f;
  __syncthreads();
  g_hs[r * kC + c] = col_dot(w.w1, dh1, kC) + col_dot(w.wb, dt + 4, 3);
  dh1s[r * kC + c] = dh1[c];
  dh2s[r * kC + c] = dh2[c];
}

// dW[o][i] += sum_r A[r*lda + o] * B[r*ldb + i];  db[o] += sum_r A[r*lda + o].  32 x 32 output tile per block, rows split over
// gridDim.z (fp32 atomics into the accumulating parameter gradients).
__global__ void __launch_bounds__(256) outer_acc_kernel(const float* __restrict__ A, int lda, int na, const float* __restrict__ Bm, int ldb,
                                                        int nb, float* __restrict__ dW, float* __restrict__ db, long R) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const long per = (R + gridDim.z - 1) / gridDim.z, r0 = (long)blockIdx.z * per, r1 = min(R, r0 + per);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bsum = 0.f;
  for (long rr = r0; rr < r1; rr += 32) {
    for (int k = ty; k < 32; k += 8) {
      const long r = rr + k;
      sa[k][tx] = (r < r1 && o0 + tx < na) ? A[r * lda + o0 + tx] : 0.f;
      sb[k][tx] = (r < r1 && i0 + tx < nb) ? Bm[r * ldb + i0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = sb[k][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(sa[k][ty + 8 * j], b, acc[j]);
    }
    if (db && blockIdx.x == 0 && ty == 0)
      for (int k = 0; k < 32; ++k) bsum += sa[k][tx];
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = o0 + ty + 8 * j, i = i0 + tx;
    if (o < na && i < nb) atomicAdd(dW + (long)o * nb + i, acc[j]);
  }
  if (db && blockIdx.x == 0 && ty == 0 && o0 + tx < na) atomicAdd(db + o0 + tx, bsum);
}

int outer_acc(const float* A, int lda, int na, const float* B, int ldb, int nb, float* dW, float* db, long R, cudaStream_t st) {
  if (!dW) return 0;
  const dim3 grid((unsigned)cdiv(nb, 32), (unsigned)cdiv(na, 32), (unsigned)std::max<long>(1, std::min<long>(16, R / 64)));
  outer_acc_kernel<<<grid, 256, 0, st>>>(A, lda, na, B, ldb, nb, dW, db, R);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

// workspace: h1, h2, dh1, dh2 [R,256] + boxes [R,4] + dt [R,8]
extern "C" size_t cqvad_heads_train_workspace_bytes(long R) { return (size_t)R * (4 * kC + 4 + 8) * sizeof(float) + 256; }

extern "C" int cqvad_heads_train_forward(const float* const* weights, const float* hs, const float* cls_hs, const float* refs, long R,
                                         int K, float p_drop, uint64_t seed, float* pred_logits, float* pred_boxes, float* pred_logits_b,
                                         void* workspace, size_t ws_bytes, void* stream) {
  CQ_CHECK_ARG(R >= 0 && K >= 1 && p_drop >= 0.f && p_drop < 1.f, "heads_train_forward: bad arguments");
  if (R == 0) return 0;
  CQ_CHECK_ARG(weights && hs && cls_hs && refs && pred_logits && pred_boxes && pred_logits_b && workspace, "heads_train_forward: null pointer");
  for (int i = 0; i < 8; ++i) CQ_CHECK_ARG(weights[i] != nullptr, "heads_train_forward: weights[%d] is NULL", i);
  if (ws_bytes < cqvad_heads_train_workspace_bytes(R)) return set_error(CQVAD_E_WORKSPACE, "heads_train_forward: workspace too small");
  cudaStream_t st = as_stream(stream);
  HeadW w{weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], weights[6], weights[7]};
  float* h1 = (float*)workspace; float* h2 = h1 + R * kC; float* bx = h2 + 3 * R * kC;
  heads_fwd_kernel<<<(unsigned)R, 256, 0, st>>>(w, hs, refs, h1, h2, bx, pred_logits_b);
  CQ_LAUNCH_CHECK();
  CQ_CUDA(cudaMemcpyAsync(pred_boxes, bx, (size_t)R * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const long rows = R * K;
  cls_mean_fwd_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(cls_hs, pred_logits, rows, p_drop, dropout_threshold(p_drop), seed);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int cqvad_heads_train_backward(const float* const* weights, const float* hs, const float* refs, const float* grad_logits,
                                          const float* grad_boxes, const float* grad_logits_b, long R, int K, float p_drop, uint64_t seed,
                                          float* grad_hs, float* grad_cls_hs, float* grad_refs, float* const* grad_weights,
                                          void* workspace, size_t ws_bytes, void* stream) {
  CQ_CHECK_ARG(R >= 0 && K >= 1 && p_drop >= 0.f && p_drop < 1.f, "heads_train_backward: bad arguments");
  if (R == 0) return 0;
  CQ_CHECK_ARG(weights && hs && refs && grad_hs && workspace, "heads_train_backward: null pointer");
  if (ws_bytes < cqvad_heads_train_workspace_bytes(R)) return set_error(CQVAD_E_WORKSPACE, "heads_train_backward: workspace too small");
  cudaStream_t st = as_stream(stream);
  HeadW w{weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], weights[6], weights[7]};
  HeadG g{};
  if (grad_weights) g = HeadG{grad_weights[0], grad_weights[1], grad_weights[2], grad_weights[3], grad_weights[4], grad_weights[5],
                              grad_weights[6], grad_weights[7]};
  float* h1 = (float*)workspace; float* h2 = h1 + R * kC; float* dh1 = h2 + R * kC; float* dh2 = dh1 + R * kC;
  float* bx = dh2 + R * kC; float* dt = bx + R * 4;
  heads_bwd_kernel<<<(unsigned)R, 256, 0, st>>>(w, g, refs, h1, h2, bx, grad_boxes, grad_logits_b, dh1, dh2, dt, grad_hs, grad_refs);
  CQ_LAUNCH_CHECK();
  if (grad_weights) {
    CQ_TRY(outer_acc(dh1, kC, kC, hs, kC, kC, g.w1, g.b1, R, st));
    CQ_TRY(outer_acc(dh2, kC, kC, h1, kC, kC, g.w2, g.b2, R, st));
    CQ_TRY(outer_acc(dt, 8, 4, h2, kC, kC, g.w3, g.b3, R, st));
    CQ_TRY(outer_acc(dt + 4, 8, 3, hs, kC, kC, g.wb, g.bb, R, st));
  }
  if (grad_cls_hs && grad_logits) {
    const long rows = R * K;
    cls_mean_bwd_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(grad_logits, grad_cls_hs, rows, p_drop, dropout_threshold(p_drop), seed);
    CQ_LAUNCH_CHECK();
  }
  return 0;
}
