// HBM-bound kernels of the decoder: LayerNorms, level-weighted memory mix, sine embeddings, small-N linears,
// reference-point refinement, heads, position encoding.  One warp per 256-channel row, 8 channels per lane,
// 16-byte vector accesses; no shared memory is needed (no reuse), reductions by warp shuffle.
#include "common.cuh"
#include "kernels_mem.cuh"

namespace cqvad {

namespace {
constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
inline unsigned row_grid(long rows) { return (unsigned)cdiv(rows, kWarpsPerBlock); }

// ---- LayerNorm(x (+res)) ---------------------------------------------------------------------------------------
template <typename T, typename O>
__global__ void __launch_bounds__(kThreads) layernorm_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                             const float* __restrict__ g, const float* __restrict__ b,
                                                             float eps, O* __restrict__ out, long rows) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  load8(x + row * kC + lane * 8, v);
  if (res) {
    float r[8];
    load8(res + row * kC + lane * 8, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  }
  warp_layernorm256(v, g, b, eps, lane);
  store8(out + row * kC + lane * 8, v);
}

// ---- LayerNorm + permuted store: internal rows (n,b[,k]) -> output [layer][b][n][k][256] ------------------------
template <typename T, typename O>
__global__ void __launch_bounds__(kThreads) layernorm_permute_kernel(const T* __restrict__ x, const float* __restrict__ g,
                                                                     const float* __restrict__ b, float eps,
                                                                     O* __restrict__ out, long rows, int nq, int BT,
                                                                     int K, float* __restrict__ row_mean_out) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  load8(x + row * kC + lane * 8, v);
  warp_layernorm256(v, g, b, eps, lane);
  const long i = row / K;
  const int k = (int)(row % K);
  const int n = (int)(i / BT), bb = (int)(i % BT);
  const long orow = ((long)bb * nq + n) * K + k;
  if (out) store8(out + orow * kC + lane * 8, v);
  if (row_mean_out) {  // class logit = channel mean of cls_norm2 output (models/model.py:219-221)
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    s = warp_sum(s);
    if (lane == 0) row_mean_out[orow] = s * (1.0f / 256.0f);
  }
}

// ---- q_memory = norm_( sum_l lvl_w[i,l] * memory[l,s,b,:] )  (dab_transformer.py:943-946) ------------------------
// also emits the fp32 level weights are read from lvlw [N,4] (softmaxed).  Row = (i, s), i = n*BT + b.
template <typename T>
__global__ void __launch_bounds__(kThreads) lvlmix_ln_kernel(const T* __restrict__ mem, const float* __restrict__ lvlw,
                                                             const float* __restrict__ g, const float* __restrict__ b,
                                                             T* __restrict__ qm, long rows, int S, int Sq, int BT) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long i = row / S;
  const int s = (int)(row % S);
  const int bb = (int)(i % BT);
  const float4 w4 = *reinterpret_cast<const float4*>(lvlw + i * 4);
  const float w[4] = {w4.x, w4.y, w4.z, w4.w};
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int l = 0; l < kL; ++l) {
    float m[8];
    load8(mem + (((long)l * S + s) * BT + bb) * kC + lane * 8, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(w[l], m[j], v[j]);
  }
  warp_layernorm256(v, g, b, 1e-5f, lane);
  store8(qm + (i * Sq + s) * kC + lane * 8, v);   // per-instance row pitch Sq >= S (multiple of 8: TMA alignment)
}

// ---- cls_feature0 = conv_norm(actor[i] + q_memory[i,s])  -> y-padded NHWC (dab_transformer.py:1049-1054) --------
template <typename T>
__global__ void __launch_bounds__(kThreads) add_ln_pad_kernel(const T* __restrict__ actor, const T* __restrict__ qm,
                                                              const float* __restrict__ g, const float* __restrict__ b,
                                                              T* __restrict__ xpad, long rows, int S, int Sq, int Sp) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long i = row / S;
  const int s = (int)(row % S);
  float v[8], a[8];
  load8(qm + (i * Sq + s) * kC + lane * 8, v);
  load8(actor + i * kC + lane * 8, a);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += a[j];
  warp_layernorm256(v, g, b, 1e-5f, lane);
  store8(xpad + (i * Sp + s) * kC + lane * 8, v);
}

// dense [n_img,S,256] <-> y-padded [n_img,Sp,256] copies (module-level ConvBlock API)
template <typename T>
__global__ void __launch_bounds__(kThreads) pad_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, long rows,
                                                            int S, int Sp, int to_padded) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long i = row / S;
  const int s = (int)(row % S);
  const long d = row * kC, p = (i * Sp + s) * kC;
  const uint4* sp = reinterpret_cast<const uint4*>(src + (to_padded ? d : p));
  uint4* dp = reinterpret_cast<uint4*>(dst + (to_padded ? p : d));
  constexpr int n16 = kC * sizeof(T) / 16;
  for (int j = lane; j < n16; j += 32) dp[j] = sp[j];
}

// ---- fp32 -> T conversion ------------------------------------------------------------------------------------
template <typename T>
__global__ void convert_kernel(const float* __restrict__ in, T* __restrict__ out, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(in + i * 8, v);
  store8(out + i * 8, v);
}

// ---- sine embedding of reference points (dab_transformer.py:50-76) ---------------------------------------------
// one block of 128 threads per row; thread j handles dims j of each of the 4 coordinates.  Output order (y,x,w,h).
template <typename O>
__global__ void __launch_bounds__(128) sine_embed_kernel(const float* __restrict__ ref, O* __restrict__ out, long rows) {
  const long row = blockIdx.x;
  const int j = threadIdx.x;
  const float4 r = *reinterpret_cast<const float4*>(ref + row * 4);
  // dim_t = 10000 ** (2*(j//2)/128)
  const float dim_t = powf(10000.0f, (float)(2 * (j / 2)) / 128.0f);
  const float two_pi = 6.283185307179586f;
  const float c[4] = {r.y, r.x, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float p = (c[q] * two_pi) / dim_t;
    const float v = (j & 1) ? cosf(p) : sinf(p);
    out[row * 512 + q * 128 + j] = from_f<O>(v);
  }
}

// ---- query_sine_embed = sine[:, :256] * pos_transformation, modulated by ref_anchor_head (dab_transformer.py:757-763)
// hidden = relu(ref_anchor_head.layers.0(output)) [N,256] (T); w1 [2,256], b1[2] fp32: the 256->2 linear is done here.
template <typename T>
__global__ void __launch_bounds__(kThreads) qse_kernel(const float* __restrict__ ref, const T* __restrict__ scale,
                                                       const T* __restrict__ hidden, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, T* __restrict__ qse, long rows) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float hdn[8], wa[8], wb[8];
  load8(hidden + row * kC + lane * 8, hdn);
  load8(w1 + lane * 8, wa);
  load8(w1 + kC + lane * 8, wb);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0 = fmaf(hdn[j], wa[j], a0); a1 = fmaf(hdn[j], wb[j], a1); }
  a0 = warp_sum(a0) + b1[0];
  a1 = warp_sum(a1) + b1[1];
  a0 = 1.0f / (1.0f + expf(-a0));
  a1 = 1.0f / (1.0f + expf(-a1));
  const float4 r = *reinterpret_cast<const float4*>(ref + row * 4);
  // channels [0,128): y part, scaled by refHW[...,1] / ref_h ; [128,256): x part, scaled by refHW[...,0] / ref_w
  const float mod = (lane < 16) ? (a1 / r.w) : (a0 / r.z);
  const float coord = (lane < 16) ? r.y : r.x;
  const float two_pi = 6.283185307179586f;
  float sc[8];
  if (scale) load8(scale + row * kC + lane * 8, sc);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = (lane * 8 + j) & 127;
    const float dim_t = powf(10000.0f, (float)(2 * (d / 2)) / 128.0f);
    const float p = (coord * two_pi) / dim_t;
    float e = (d & 1) ? cosf(p) : sinf(p);
    if (scale) e *= sc[j];
    v[j] = e * mod;
  }
  store8(qse + row * kC + lane * 8, v);
}

// ---- small-N linear: out[row, n] = x[row,:] . w[n,:] + b[n], n <= 8, fp32 out; optional softmax over n ----------
template <typename T>
__global__ void __launch_bounds__(kThreads) linear_smalln_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ b, float* __restrict__ out,
                                                                 long rows, int n, int softmax) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float xv[8];
  load8(x + row * kC + lane * 8, xv);
  float o[8];
  float mx = -INFINITY;
  for (int c = 0; c < n; ++c) {
    float wv[8];
    load8(w + c * kC + lane * 8, wv);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fmaf(xv[j], wv[j], a);
    a = warp_sum(a) + b[c];
    o[c] = a;
    mx = fmaxf(mx, a);
  }
  if (softmax) {
    float s = 0.f;
    for (int c = 0; c < n; ++c) { o[c] = expf(o[c] - mx); s += o[c]; }
    for (int c = 0; c < n; ++c) o[c] /= s;
  }
  if (lane == 0)
    for (int c = 0; c < n; ++c) out[row * n + c] = o[c];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float inv_sigmoid(float x) {  // utils/misc.py:530-534
  x = fminf(fmaxf(x, 0.f), 1.f);
  const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
  return logf(x1 / x2);
}

// ---- box refinement: r_new = sigmoid(bbox_embed(output)[..., :4] + inverse_sigmoid(r))  (dab_transformer.py:817-823)
// hidden = relu(bbox_embed.layers.1(relu(bbox_embed.layers.0(x)))) [rows,256] (T); w2 [4,256], b2 [4].
// Writes r_new (internal order (n,b)) and optionally the permuted copy [b][n][4] (references / pred_boxes).
template <typename T>
__global__ void __launch_bounds__(kThreads) box_refine_kernel(const T* __restrict__ hidden, const float* __restrict__ w2,
                                                              const float* __restrict__ b2, const float* __restrict__ ref,
                                                              float* __restrict__ ref_new, float* __restrict__ out_perm,
                                                              long rows, int nq, int BT, int ref_is_perm) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float xv[8];
  load8(hidden + row * kC + lane * 8, xv);
  float o[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float wv[8];
    load8(w2 + c * kC + lane * 8, wv);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fmaf(xv[j], wv[j], a);
    o[c] = warp_sum(a) + b2[c];
  }
  if (lane == 0) {
    const int n = (int)(row / BT), bb = (int)(row % BT);
    const long prow = (long)bb * nq + n;
    const float* r = ref + (ref_is_perm ? prow : row) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = sigmoidf_(o[c] + inv_sigmoid(r[c]));
    if (ref_new) *reinterpret_cast<float4*>(ref_new + row * 4) = make_float4(o[0], o[1], o[2], o[3]);
    if (out_perm) *reinterpret_cast<float4*>(out_perm + prow * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void sigmoid4_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out_perm,
                                long rows, int nq, int BT) {
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float4 v = *reinterpret_cast<const float4*>(in + row * 4);
  v.x = sigmoidf_(v.x); v.y = sigmoidf_(v.y); v.z = sigmoidf_(v.z); v.w = sigmoidf_(v.w);
  *reinterpret_cast<float4*>(out + row * 4) = v;
  if (out_perm) {
    const int n = (int)(row / BT), bb = (int)(row % BT);
    *reinterpret_cast<float4*>(out_perm + ((long)bb * nq + n) * 4) = v;
  }
}

// broadcast rows: out[i*K + k, :] = src[k, :]
template <typename T>
__global__ void __launch_bounds__(kThreads) broadcast_rows_kernel(const T* __restrict__ src, T* __restrict__ out,
                                                                  long rows, int K) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  load8(src + (row % K) * kC + lane * 8, v);
  store8(out + row * kC + lane * 8, v);
}

// ---- PositionEmbeddingSine_3D (models/position_encoding.py:32-73) ----------------------------------------------
// one thread per (b, t, y, x); loops over the channels.  cumsums are recomputed per thread (extents are tiny).
__global__ void posenc3d_kernel(const uint8_t* __restrict__ mask, float* __restrict__ pos, int B, int T, int H, int W,
                                int npf) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * T * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H), t = (int)((idx / ((long)W * H)) % T);
  const int bb = (int)(idx / ((long)W * H * T));
  const uint8_t* m = mask + (long)bb * T * H * W;
  float ct = 0, cy = 0, cx = 0, nt = 0, ny = 0, nx = 0;
  for (int i = 0; i < T; ++i) { const float v = m[((long)i * H + y) * W + x] ? 0.f : 1.f; nt += v; if (i <= t) ct += v; }
  for (int i = 0; i < H; ++i) { const float v = m[((long)t * H + i) * W + x] ? 0.f : 1.f; ny += v; if (i <= y) cy += v; }
  for (int i = 0; i < W; ++i) { const float v = m[((long)t * H + y) * W + i] ? 0.f : 1.f; nx += v; if (i <= x) cx += v; }
  const float two_pi = 6.283185307179586f, eps = 1e-6f;
  const float et = ct / (nt + eps) * two_pi, ey = cy / (ny + eps) * two_pi, ex = cx / (nx + eps) * two_pi;
  const int n_t = npf / 8 * 2, n_s = npf / 8 * 3;
  const long plane = (long)T * H * W;
  float* o = pos + (long)bb * npf * plane + ((long)t * H + y) * W + x;
  for (int c = 0; c < npf; ++c) {
    float e, nn; int j;
    if (c < n_t) { e = et; nn = (float)n_t; j = c; }
    else if (c < n_t + n_s) { e = ey; nn = (float)n_s; j = c - n_t; }
    else { e = ex; nn = (float)n_s; j = c - n_t - n_s; }
    // temperature ** (2 * (j / 2) / n) with TRUE division (position_encoding.py:55,60)
    const float dim_t = powf(10000.0f, 2.0f * ((float)j / 2.0f) / nn);
    const float p = e / dim_t;
    o[(long)c * plane] = (j & 1) ? cosf(p) : sinf(p);
  }
}

// heads: pred_logits_b = class_embed_b(hs) (models/model.py:192); x = norm(output) rows in internal (n,b) order,
// written permuted to [b][n][3]
template <typename T>
__global__ void __launch_bounds__(kThreads) logits_b_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ out,
                                                            long rows, int nq, int BT) {
  const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float xv[8];
  load8(x + row * kC + lane * 8, xv);
  const int n = (int)(row / BT), bb = (int)(row % BT);
  const long prow = (long)bb * nq + n;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float wv[8];
    load8(w + c * kC + lane * 8, wv);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fmaf(xv[j], wv[j], a);
    a = warp_sum(a) + b[c];
    if (lane == 0) out[prow * 3 + c] = a;
  }
}

}  // namespace

// ---- host wrappers ---------------------------------------------------------------------------------------------
template <typename T>
int layernorm_rows(const T* x, const T* res, const float* g, const float* b, float eps, void* out, bool out_f32,
                   long rows, cudaStream_t st) {
  if (rows == 0) return 0;
  if (out_f32)
    layernorm_kernel<T, float><<<row_grid(rows), kThreads, 0, st>>>(x, res, g, b, eps, (float*)out, rows);
  else
    layernorm_kernel<T, T><<<row_grid(rows), kThreads, 0, st>>>(x, res, g, b, eps, (T*)out, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int layernorm_rows<float>(const float*, const float*, const float*, const float*, float, void*, bool, long, cudaStream_t);
template int layernorm_rows<bf16>(const bf16*, const bf16*, const float*, const float*, float, void*, bool, long, cudaStream_t);

template <typename T>
int layernorm_permute(const T* x, const float* g, const float* b, float eps, void* out, bool out_f32, long rows, int nq,
                      int BT, int K, float* row_mean_out, cudaStream_t st) {
  if (rows == 0) return 0;
  if (out_f32)
    layernorm_permute_kernel<T, float><<<row_grid(rows), kThreads, 0, st>>>(x, g, b, eps, (float*)out, rows, nq, BT, K, row_mean_out);
  else
    layernorm_permute_kernel<T, T><<<row_grid(rows), kThreads, 0, st>>>(x, g, b, eps, (T*)out, rows, nq, BT, K, row_mean_out);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int layernorm_permute<float>(const float*, const float*, const float*, float, void*, bool, long, int, int, int, float*, cudaStream_t);
// fp32 input (side channel of the bf16 path), output fp32 or bf16
int layernorm_permute_f32in(const float* x, const float* g, const float* b, float eps, void* out, bool out_f32, long rows,
                            int nq, int BT, int K, float* row_mean_out, cudaStream_t st) {
  if (rows == 0) return 0;
  if (out_f32)
    layernorm_permute_kernel<float, float><<<row_grid(rows), kThreads, 0, st>>>(x, g, b, eps, (float*)out, rows, nq, BT, K, row_mean_out);
  else
    layernorm_permute_kernel<float, bf16><<<row_grid(rows), kThreads, 0, st>>>(x, g, b, eps, (bf16*)out, rows, nq, BT, K, row_mean_out);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int layernorm_permute<bf16>(const bf16*, const float*, const float*, float, void*, bool, long, int, int, int, float*, cudaStream_t);

template <typename T>
int lvlmix_ln(const T* mem, const float* lvlw, const float* g, const float* b, T* qm, long N, int S, int Sq, int BT, cudaStream_t st) {
  const long rows = N * S;
  lvlmix_ln_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(mem, lvlw, g, b, qm, rows, S, Sq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int lvlmix_ln<float>(const float*, const float*, const float*, const float*, float*, long, int, int, int, cudaStream_t);
template int lvlmix_ln<bf16>(const bf16*, const float*, const float*, const float*, bf16*, long, int, int, int, cudaStream_t);

template <typename T>
int add_ln_pad(const T* actor, const T* qm, const float* g, const float* b, T* xpad, long N, int S, int Sq, int Sp, cudaStream_t st) {
  const long rows = N * S;
  add_ln_pad_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(actor, qm, g, b, xpad, rows, S, Sq, Sp);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int add_ln_pad<float>(const float*, const float*, const float*, const float*, float*, long, int, int, int, cudaStream_t);
template int add_ln_pad<bf16>(const bf16*, const bf16*, const float*, const float*, bf16*, long, int, int, int, cudaStream_t);

template <typename T>
int pad_copy(const T* src, T* dst, long n_img, int S, int Sp, bool to_padded, cudaStream_t st) {
  const long rows = n_img * S;
  pad_copy_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(src, dst, rows, S, Sp, to_padded ? 1 : 0);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int pad_copy<float>(const float*, float*, long, int, int, bool, cudaStream_t);
template int pad_copy<bf16>(const bf16*, bf16*, long, int, int, bool, cudaStream_t);

template <typename T>
int convert_f32(const float* in, T* out, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "convert: element count must be a multiple of 8");
  const long n8 = n / 8;
  convert_kernel<T><<<(unsigned)cdiv(n8, 256), 256, 0, st>>>(in, out, n8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int convert_f32<float>(const float*, float*, long, cudaStream_t);
template int convert_f32<bf16>(const float*, bf16*, long, cudaStream_t);

template <typename O>
int sine_embed(const float* ref, O* out, long rows, cudaStream_t st) {
  if (rows == 0) return 0;
  sine_embed_kernel<O><<<(unsigned)rows, 128, 0, st>>>(ref, out, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int sine_embed<float>(const float*, float*, long, cudaStream_t);
template int sine_embed<bf16>(const float*, bf16*, long, cudaStream_t);

template <typename T>
int qse_modulate(const float* ref, const T* scale, const T* hidden, const float* w1, const float* b1, T* qse, long rows,
                 cudaStream_t st) {
  qse_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(ref, scale, hidden, w1, b1, qse, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int qse_modulate<float>(const float*, const float*, const float*, const float*, const float*, float*, long, cudaStream_t);
template int qse_modulate<bf16>(const float*, const bf16*, const bf16*, const float*, const float*, bf16*, long, cudaStream_t);

template <typename T>
int linear_smalln(const T* x, const float* w, const float* b, float* out, long rows, int n, bool softmax, cudaStream_t st) {
  CQ_CHECK_SHAPE(n >= 1 && n <= 8, "linear_smalln: n must be in [1,8]");
  linear_smalln_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(x, w, b, out, rows, n, softmax ? 1 : 0);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int linear_smalln<float>(const float*, const float*, const float*, float*, long, int, bool, cudaStream_t);
template int linear_smalln<bf16>(const bf16*, const float*, const float*, float*, long, int, bool, cudaStream_t);

template <typename T>
int box_refine(const T* hidden, const float* w2, const float* b2, const float* ref, float* ref_new, float* out_perm,
               long rows, int nq, int BT, bool ref_is_perm, cudaStream_t st) {
  box_refine_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(hidden, w2, b2, ref, ref_new, out_perm, rows, nq, BT, ref_is_perm ? 1 : 0);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int box_refine<float>(const float*, const float*, const float*, const float*, float*, float*, long, int, int, bool, cudaStream_t);
template int box_refine<bf16>(const bf16*, const float*, const float*, const float*, float*, float*, long, int, int, bool, cudaStream_t);

int sigmoid4(const float* in, float* out, float* out_perm, long rows, int nq, int BT, cudaStream_t st) {
  sigmoid4_kernel<<<(unsigned)cdiv(rows, 128), 128, 0, st>>>(in, out, out_perm, rows, nq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int broadcast_rows(const T* src, T* out, long rows, int K, cudaStream_t st) {
  broadcast_rows_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(src, out, rows, K);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int broadcast_rows<float>(const float*, float*, long, int, cudaStream_t);
template int broadcast_rows<bf16>(const bf16*, bf16*, long, int, cudaStream_t);

int posenc3d(const uint8_t* mask, float* pos, int B, int T, int H, int W, int npf, cudaStream_t st) {
  const long total = (long)B * T * H * W;
  posenc3d_kernel<<<(unsigned)cdiv(total, 128), 128, 0, st>>>(mask, pos, B, T, H, W, npf);
  CQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int logits_b(const T* x, const float* w, const float* b, float* out, long rows, int nq, int BT, cudaStream_t st) {
  logits_b_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(x, w, b, out, rows, nq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int logits_b<float>(const float*, const float*, const float*, float*, long, int, int, cudaStream_t);
template int logits_b<bf16>(const bf16*, const float*, const float*, float*, long, int, int, cudaStream_t);

}  // namespace cqvad
