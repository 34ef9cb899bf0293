// Deformable encoder layer around the MSDA-3D op:
//   DeformableTransformerEncoderLayer.forward (models/detr/dab_transformer.py:513-523) with
//   MSDeformAttn3D.forward (ops/modules/ms_deform_attn.py:167-203) inlined:
//     q      = src + pos                                                  (:515 with_pos_embed)
//     value  = value_proj(src), padded tokens zeroed                      (:181-183)
//     off    = sampling_offsets(q), logit = attention_weights(q)          (:185-186)   fp32 outputs of the tcgen05 GEMMs
//     attn   = softmax_{L*P}(logit); loc = ref + off / (T_l, W_l, H_l)    (:187-192)   one warp per (token, head)
//     samp   = MSDA-3D(value, loc, attn)                                  (:198)       msda.cu
//     x1     = LN1(src + output_proj(samp))                               (:200, :516-517)  LayerNorm in the GEMM epilogue
//     out    = LN2(x1 + linear2(relu(linear1(x1))))                       (:507-510)   fused tcgen05 MLP (hidden stays on the SM)
// Dropout is the identity (eval / the native path's documented divergence).  Heads = 8, d_model = 256 (all shipped configs).
#include "common.cuh"
#include "bwd.cuh"
#include "dropout.cuh"

#include <stdlib.h>

namespace cqvad {
int msda_fwd_fused(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi, const float* off, const float* logit,
                   const float* ref, void* out, int N, int Len, int M, int D, int L, int Lq, int P, cudaStream_t st);
namespace {

enum { E_OFF_W = 0, E_OFF_B, E_ATT_W, E_ATT_B, E_VAL_W, E_VAL_B, E_OUT_W, E_OUT_B, E_N1_W, E_N1_B, E_L1_W, E_L1_B, E_L2_W, E_L2_B,
       E_N2_W, E_N2_B, E_COUNT };
constexpr int kM = 8;   // heads

template <typename T>
__global__ void __launch_bounds__(256) add_rows_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float x[8], y[8];
  load8(a + i * 8, x);
  load8(b + i * 8, y);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] += y[j];
  store8(o + i * 8, x);
}

// value.masked_fill(padding_mask[..., None], 0): one warp per row of 256
template <typename T>
__global__ void __launch_bounds__(256) mask_rows_kernel(T* __restrict__ x, const uint8_t* __restrict__ mask, long rows) {
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows || !mask[row]) return;
  const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  store8(x + row * kC + lane * 8, z);
}

// one warp per (token, head): softmax over the L*P logits and sampling locations of the L*P points.
// off [rows, M*L*P*3], logit [rows, M*L*P] (fp32), ref [rows, L, 3] -> loc [rows, M, L, P, 3], attn [rows, M, L, P]
__global__ void __launch_bounds__(256) msda_prepare_kernel(const float* __restrict__ off, const float* __restrict__ logit,
                                                           const float* __restrict__ ref, const int64_t* __restrict__ shapes,
                                                           float* __restrict__ loc, float* __restrict__ attn, long rows, int L,
                                                           int P) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= rows * kM) return;
  const long row = wid / kM;
  const int LP = L * P;
  const float* lg = logit + wid * LP;
  float mx = -INFINITY;
  for (int j = lane; j < LP; j += 32) mx = fmaxf(mx, lg[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < LP; j += 32) sum += expf(lg[j] - mx);
  sum = warp_sum(sum);
  for (int j = lane; j < LP; j += 32) {
    attn[wid * LP + j] = expf(lg[j] - mx) / sum;
    const int l = j / P;
    // offset normaliser stacked as (T_l, W_l, H_l) against (x, y, t) offsets: the reference's own order, ms_deform_attn.py:190
    const float nx = (float)shapes[l * 3 + 0], ny = (float)shapes[l * 3 + 2], nt = (float)shapes[l * 3 + 1];
    const float* o = off + (wid * LP + j) * 3;
    const float* r = ref + (row * L + l) * 3;
    float* d = loc + (wid * LP + j) * 3;
    d[0] = r[0] + __fdiv_rn(o[0], nx);
    d[1] = r[1] + __fdiv_rn(o[1], ny);
    d[2] = r[2] + __fdiv_rn(o[2], nt);
  }
}

// the 12 offset normalisers (T_l, W_l, H_l) of Epilogue::rowop 2, index l*3 + i
__global__ void msda_norm12_kernel(const int64_t* __restrict__ shapes, float* __restrict__ norm) {
  const int c = threadIdx.x;
  if (c >= 12) return;
  const int i = c % 3, l = c / 3;
  norm[c] = (float)shapes[l * 3 + (i == 0 ? 0 : i == 1 ? 2 : 1)];
}

// loc / attn [rows, M, L, P, (3)] from q in two GEMMs with the prepare arithmetic in their epilogues; false = shape not taken by the
// tcgen05 path (the caller runs the unfused sequence)
static bool msda_projections_prepared(const bf16* q, const bf16* w_off, const float* b_off, const bf16* w_att, const float* b_att,
                                      const float* refp, const int64_t* shapes, float* loc, float* attn, float* norm12, long rows,
                                      cudaStream_t st) {
  const int LP3 = kM * 4 * 8 * 3, LP1 = kM * 4 * 8;
  msda_norm12_kernel<<<1, 32, 0, st>>>(shapes, norm12);
  Epilogue e; e.bias = b_off; e.c32 = loc; e.rowop = 2; e.ro_ref = refp; e.ro_norm = norm12;
  if (gemm_tc(q, kC, w_off, nullptr, LP3, rows, LP3, kC, e, nullptr, st) != 0) return false;
  Epilogue a; a.bias = b_att; a.c32 = attn; a.rowop = 1;
  return gemm_tc(q, kC, w_att, nullptr, LP1, rows, LP1, kC, a, nullptr, st) == 0;
}

struct EncWs {
  char* base; size_t off = 0, cap;
  EncWs(void* p, size_t c) : base((char*)p), cap(c) {}
  void* take(size_t bytes) {
    const size_t a = (off + 255) & ~(size_t)255;
    off = a + bytes;
    return (base && off <= cap) ? base + a : nullptr;
  }
};

template <typename T>
size_t enc_ws_bytes(long rows, int L, int P, int F) {
  EncWs w(nullptr, 0);
  const int LP3 = kM * L * P * 3, LP1 = kM * L * P;
  w.take((size_t)rows * kC * sizeof(T));                        // q
  w.take((size_t)rows * kC * sizeof(T));                        // value
  w.take((size_t)rows * LP3 * sizeof(T));                       // offsets (dtype)
  w.take((size_t)rows * LP1 * sizeof(T));                       // logits (dtype)
  if (sizeof(T) == 2) { w.take((size_t)rows * LP3 * 4); w.take((size_t)rows * LP1 * 4); }   // fp32 copies of both
  w.take((size_t)rows * LP3 * 4);                               // loc
  w.take((size_t)rows * LP1 * 4);                               // attn
  w.take((size_t)rows * kC * sizeof(T));                        // sampled
  w.take((size_t)rows * kC * sizeof(T));                        // x1
  w.take((size_t)rows * F * sizeof(T));                         // FFN hidden (only used off the fused-MLP path)
  w.take((size_t)LP3 * 8);                                      // offset normalisers of the fused location epilogue (12 used)
  return w.off + 256;
}

template <typename T>
int enc_layer_t(const void* const* W, const T* src, const T* pos, const float* refp, const int64_t* shapes, const int64_t* lsi,
                const uint8_t* mask, T* out, T* attn_out, void* ws, size_t ws_bytes, int B, long Len, int L, int P, int F,
                cudaStream_t st) {
  const long rows = (long)B * Len;
  if (rows == 0) return 0;
  const int LP3 = kM * L * P * 3, LP1 = kM * L * P;
  if (ws_bytes < enc_ws_bytes<T>(rows, L, P, F)) return set_error(CQVAD_E_WORKSPACE, "deform_encoder_layer: workspace too small");
  EncWs w(ws, ws_bytes);
  T* q = (T*)w.take((size_t)rows * kC * sizeof(T));
  T* value = (T*)w.take((size_t)rows * kC * sizeof(T));
  T* offT = (T*)w.take((size_t)rows * LP3 * sizeof(T));
  T* lgT = (T*)w.take((size_t)rows * LP1 * sizeof(T));
  float *off32 = (float*)offT, *lg32 = (float*)lgT;
  if (sizeof(T) == 2) { off32 = (float*)w.take((size_t)rows * LP3 * 4); lg32 = (float*)w.take((size_t)rows * LP1 * 4); }
  float* loc = (float*)w.take((size_t)rows * LP3 * 4);
  float* attn = (float*)w.take((size_t)rows * LP1 * 4);
  T* samp = (T*)w.take((size_t)rows * kC * sizeof(T));
  T* x1 = (T*)w.take((size_t)rows * kC * sizeof(T));
  T* hid = (T*)w.take((size_t)rows * F * sizeof(T));
  float* col_norm = (float*)w.take((size_t)LP3 * 8);   // 12 floats used
  auto Wm = [&](int i) { return (const T*)W[i]; };
  auto Wf = [&](int i) { return (const float*)W[i]; };

  const long n8 = rows * kC / 8;
  add_rows_kernel<T><<<(unsigned)cdiv(n8, 256), 256, 0, st>>>(src, pos, q, n8);
  CQ_LAUNCH_CHECK();
  { Epilogue e; e.bias = Wf(E_VAL_B); CQ_TRY(gemm<T>(src, kC, Wm(E_VAL_W), value, kC, rows, kC, kC, e, nullptr, st)); }
  if (mask) {
    mask_rows_kernel<T><<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(value, mask, rows);
    CQ_LAUNCH_CHECK();
  }
  // bf16 path: softmax and sampling-location arithmetic (msda_prepare) run in the epilogues of the two query projections, straight
  // from the fp32 accumulators -- the raw offsets / logits are never written (1.4 GB of traffic per layer at 4 clips)
  static const bool no_rowop = getenv("CQVAD_ENC_NO_ROWOP") != nullptr;
  // CQVAD_ENC_FUSED=1 (opt-in, measured slower): softmax + location arithmetic inside the SAMPLING kernel instead (msda_fwd_fused)
  static const bool fused_sampling = getenv("CQVAD_ENC_FUSED") != nullptr;
  bool prepared = false, sampled = false;
  if (sizeof(T) == 2 && !no_rowop && !fused_sampling && L == 4 && P == 8) {
    prepared = msda_projections_prepared((const bf16*)q, (const bf16*)Wm(E_OFF_W), Wf(E_OFF_B), (const bf16*)Wm(E_ATT_W), Wf(E_ATT_B), refp,
                                         shapes, loc, attn, col_norm, rows, st);
  }
  if (!prepared) {
    // offsets / logits: fp32 straight from the accumulators (bf16: fp32 side output of the epilogue), so that the sampling
    // locations carry no bf16 rounding of their own
    { Epilogue e; e.bias = Wf(E_OFF_B); if (sizeof(T) == 2) e.c32 = off32; CQ_TRY(gemm<T>(q, kC, Wm(E_OFF_W), offT, LP3, rows, LP3, kC, e, nullptr, st)); }
    { Epilogue e; e.bias = Wf(E_ATT_B); if (sizeof(T) == 2) e.c32 = lg32; CQ_TRY(gemm<T>(q, kC, Wm(E_ATT_W), lgT, LP1, rows, LP1, kC, e, nullptr, st)); }
    if (fused_sampling) {
      const int fr = msda_fwd_fused(DT<T>::id, value, shapes, lsi, off32, lg32, refp, samp, B, (int)Len, kM, kC / kM, L, (int)Len, P, st);
      if (fr < 0) return fr;
      sampled = fr == 0;
    }
    if (!sampled) {
      msda_prepare_kernel<<<(unsigned)cdiv(rows * kM * 32, 256), 256, 0, st>>>(off32, lg32, refp, shapes, loc, attn, rows, L, P);
      CQ_LAUNCH_CHECK();
    }
  }
  if (!sampled) CQ_TRY(cqvad_msda3d_forward(DT<T>::id, value, shapes, lsi, loc, attn, samp, B, (int)Len, kM, kC / kM, L, (int)Len, P, (void*)st));
  if (attn_out) {   // the module's own output (tests): output_proj(samp)
    Epilogue e; e.bias = Wf(E_OUT_B);
    CQ_TRY(gemm<T>(samp, kC, Wm(E_OUT_W), attn_out, kC, rows, kC, kC, e, nullptr, st));
  }
  { Epilogue e; e.bias = Wf(E_OUT_B); e.res = src; e.ldr = kC; e.ln_g = Wf(E_N1_W); e.ln_b = Wf(E_N1_B); e.ln_eps = 1e-5f;
    CQ_TRY(gemm<T>(samp, kC, Wm(E_OUT_W), x1, kC, rows, kC, kC, e, nullptr, st)); }
  return cqvad_mlp(DT<T>::id, x1, Wm(E_L1_W), Wf(E_L1_B), Wm(E_L2_W), Wf(E_L2_B), CQVAD_ACT_RELU, x1, Wf(E_N2_W), Wf(E_N2_B), 1e-5f,
                   out, hid, rows, F, (void*)st);
}

// ---- training: forward that keeps what the backward needs, and the backward ------------------------------------------------
// backward of msda_prepare: dlogit = attn * (dattn - sum_j attn_j dattn_j)  (softmax), doff = dloc / (T_l, W_l, H_l); written in
// the GEMM operand type T (they feed the weight- and data-gradient GEMMs of the two query projections)
template <typename T>
__global__ void __launch_bounds__(256) msda_prepare_bwd_kernel(const float* __restrict__ attn, const float* __restrict__ dattn,
                                                               const float* __restrict__ dloc, const int64_t* __restrict__ shapes,
                                                               T* __restrict__ doff, T* __restrict__ dlogit, long rows, int L,
                                                               int P) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= rows * kM) return;
  const int LP = L * P;
  float dot = 0.f;
  for (int j = lane; j < LP; j += 32) dot = fmaf(attn[wid * LP + j], dattn[wid * LP + j], dot);
  dot = warp_sum(dot);
  for (int j = lane; j < LP; j += 32) {
    dlogit[wid * LP + j] = from_f<T>(attn[wid * LP + j] * (dattn[wid * LP + j] - dot));
    const int l = j / P;
    const float nx = (float)shapes[l * 3 + 0], ny = (float)shapes[l * 3 + 2], nt = (float)shapes[l * 3 + 1];
    const float* g = dloc + (wid * LP + j) * 3;
    T* d = doff + (wid * LP + j) * 3;
    d[0] = from_f<T>(g[0] / nx); d[1] = from_f<T>(g[1] / ny); d[2] = from_f<T>(g[2] / nt);
  }
}

// one layout for both passes: [saved by the forward | scratch of the backward]
template <typename T>
struct EncTrainWs {
  T *q, *value, *offT, *lgT, *samp, *z1, *x1, *h, *z2;
  float *off32, *lg32, *loc, *attn;
  T *dz2, *dh, *dx1, *dz1, *dsamp, *doffT, *dlgT, *dq, *dvalT, *wt;
  float *dval32, *dloc, *dattn, *col_norm;
  size_t bytes;
  EncTrainWs(void* base, size_t cap, long rows, int L, int P, int F) {
    EncWs w(base, cap);
    const size_t LP3 = (size_t)kM * L * P * 3, LP1 = (size_t)kM * L * P, R = (size_t)rows, C = kC;
    auto tk = [&](size_t n) { return (T*)w.take(n * sizeof(T)); };
    auto tf = [&](size_t n) { return (float*)w.take(n * 4); };
    q = tk(R * C); value = tk(R * C); offT = tk(R * LP3); lgT = tk(R * LP1);
    off32 = (float*)offT; lg32 = (float*)lgT;
    if (sizeof(T) == 2) { off32 = tf(R * LP3); lg32 = tf(R * LP1); }
    loc = tf(R * LP3); attn = tf(R * LP1);
    samp = tk(R * C); z1 = tk(R * C); x1 = tk(R * C); h = tk(R * F); z2 = tk(R * C);
    dz2 = tk(R * C); dh = tk(R * F); dx1 = tk(R * C); dz1 = tk(R * C); dsamp = tk(R * C);
    doffT = tk(R * LP3); dlgT = tk(R * LP1); dq = tk(R * C); dvalT = tk(R * C);
    dval32 = tf(R * C); dloc = tf(R * LP3); dattn = tf(R * LP1);
    wt = tk(2 * C * C + LP3 * C + LP1 * C + 2 * (size_t)F * C);      // transposed weights for the data gradients
    col_norm = tf(2 * LP3);                                          // offset normalisers of the fused location epilogue (12 used)
    bytes = w.off + 256;
  }
};

template <typename T>
int enc_train_fwd_t(const void* const* W, const T* src, const T* pos, const float* refp, const int64_t* shapes, const int64_t* lsi,
                    const uint8_t* mask, T* out, void* ws, size_t ws_bytes, int B, long Len, int L, int P, int F, float pdrop,
                    uint64_t seed, cudaStream_t st) {
  const long rows = (long)B * Len;
  if (rows == 0) return 0;
  const int LP3 = kM * L * P * 3, LP1 = kM * L * P;
  EncTrainWs<T> w(ws, ws_bytes, rows, L, P, F);
  if (ws_bytes < w.bytes) return set_error(CQVAD_E_WORKSPACE, "deform_encoder_layer (training): workspace too small");
  auto Wm = [&](int i) { return (const T*)W[i]; };
  auto Wf = [&](int i) { return (const float*)W[i]; };
  const bool drop = pdrop > 0.f;   // dropout1 / dropout2 / dropout3 of dab_transformer.py:499-519 = sites 1 / 2 / 3
  const long n8 = rows * kC / 8;
  add_rows_kernel<T><<<(unsigned)cdiv(n8, 256), 256, 0, st>>>(src, pos, w.q, n8);
  CQ_LAUNCH_CHECK();
  { Epilogue e; e.bias = Wf(E_VAL_B); CQ_TRY(gemm<T>(src, kC, Wm(E_VAL_W), w.value, kC, rows, kC, kC, e, nullptr, st)); }
  if (mask) { mask_rows_kernel<T><<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(w.value, mask, rows); CQ_LAUNCH_CHECK(); }
  static const bool no_rowop = getenv("CQVAD_ENC_NO_ROWOP") != nullptr;
  bool prepared = false;
  if (sizeof(T) == 2 && !no_rowop && L == 4 && P == 8)     // softmax / location arithmetic in the projections' epilogues (see enc_layer_t)
    prepared = msda_projections_prepared((const bf16*)w.q, (const bf16*)Wm(E_OFF_W), Wf(E_OFF_B), (const bf16*)Wm(E_ATT_W), Wf(E_ATT_B), refp,
                                         shapes, w.loc, w.attn, w.col_norm, rows, st);
  if (!prepared) {
    { Epilogue e; e.bias = Wf(E_OFF_B); if (sizeof(T) == 2) e.c32 = w.off32; CQ_TRY(gemm<T>(w.q, kC, Wm(E_OFF_W), w.offT, LP3, rows, LP3, kC, e, nullptr, st)); }
    { Epilogue e; e.bias = Wf(E_ATT_B); if (sizeof(T) == 2) e.c32 = w.lg32; CQ_TRY(gemm<T>(w.q, kC, Wm(E_ATT_W), w.lgT, LP1, rows, LP1, kC, e, nullptr, st)); }
    msda_prepare_kernel<<<(unsigned)cdiv(rows * kM * 32, 256), 256, 0, st>>>(w.off32, w.lg32, refp, shapes, w.loc, w.attn, rows, L, P);
    CQ_LAUNCH_CHECK();
  }
  CQ_TRY(cqvad_msda3d_forward(DT<T>::id, w.value, shapes, lsi, w.loc, w.attn, w.samp, B, (int)Len, kM, kC / kM, L, (int)Len, P, (void*)st));
  { Epilogue e; e.bias = Wf(E_OUT_B); if (!drop) { e.res = src; e.ldr = kC; }
    CQ_TRY(gemm<T>(w.samp, kC, Wm(E_OUT_W), w.z1, kC, rows, kC, kC, e, nullptr, st)); }
  if (drop) CQ_TRY(dropout_apply<T>(w.z1, src, w.z1, rows * kC, pdrop, seed, 1, st));            // src + dropout1(src2)
  CQ_TRY(layernorm_rows<T>(w.z1, nullptr, Wf(E_N1_W), Wf(E_N1_B), 1e-5f, w.x1, false, rows, st));
  { Epilogue e; e.bias = Wf(E_L1_B); e.act = CQVAD_ACT_RELU; CQ_TRY(gemm<T>(w.x1, kC, Wm(E_L1_W), w.h, F, rows, F, kC, e, nullptr, st)); }
  if (drop) CQ_TRY(dropout_apply<T>(w.h, nullptr, w.h, rows * F, pdrop, seed, 2, st));            // dropout2(activation(linear1))
  { Epilogue e; e.bias = Wf(E_L2_B); if (!drop) { e.res = w.x1; e.ldr = kC; }
    CQ_TRY(gemm<T>(w.h, F, Wm(E_L2_W), w.z2, kC, rows, kC, F, e, nullptr, st)); }
  if (drop) CQ_TRY(dropout_apply<T>(w.z2, w.x1, w.z2, rows * kC, pdrop, seed, 3, st));           // src + dropout3(src2)
  return layernorm_rows<T>(w.z2, nullptr, Wf(E_N2_W), Wf(E_N2_B), 1e-5f, out, false, rows, st);
}

template <typename T>
int enc_train_bwd_t(const void* const* W, const T* src, const int64_t* shapes, const int64_t* lsi, const uint8_t* mask,
                    const T* gout, T* gsrc, T* gpos, float* const* G, void* ws, size_t ws_bytes, int B, long Len, int L, int P, int F,
                    float pdrop, uint64_t seed, cudaStream_t st) {
  const long rows = (long)B * Len;
  if (rows == 0) return 0;
  const int LP3 = kM * L * P * 3, LP1 = kM * L * P;
  EncTrainWs<T> w(ws, ws_bytes, rows, L, P, F);
  if (ws_bytes < w.bytes) return set_error(CQVAD_E_WORKSPACE, "deform_encoder_layer (training): workspace too small");
  auto Wm = [&](int i) { return (const T*)W[i]; };
  auto Wf = [&](int i) { return (const float*)W[i]; };
  // transposed weights [in][out] for dX = dY . W
  T* wt_val = w.wt; T* wt_out = wt_val + kC * kC; T* wt_off = wt_out + kC * kC; T* wt_att = wt_off + (size_t)LP3 * kC;
  T* wt_l1 = wt_att + (size_t)LP1 * kC; T* wt_l2 = wt_l1 + (size_t)F * kC;
  CQ_TRY(transpose_w<T>(Wm(E_VAL_W), wt_val, kC, kC, st));
  CQ_TRY(transpose_w<T>(Wm(E_OUT_W), wt_out, kC, kC, st));
  CQ_TRY(transpose_w<T>(Wm(E_OFF_W), wt_off, LP3, kC, st));
  CQ_TRY(transpose_w<T>(Wm(E_ATT_W), wt_att, LP1, kC, st));
  CQ_TRY(transpose_w<T>(Wm(E_L1_W), wt_l1, F, kC, st));
  CQ_TRY(transpose_w<T>(Wm(E_L2_W), wt_l2, kC, F, st));
  // out = LN2(z2)
  CQ_TRY(ln_bwd<T>(w.z2, nullptr, Wf(E_N2_W), 1e-5f, gout, false, 0, 0, 0, w.dz2, 0.f, nullptr, 0.f, G[E_N2_W], G[E_N2_B], rows, st));
  // z2 = x1 + dropout3(linear2(h)), h = dropout2(relu(linear1(x1)))
  const bool drop = pdrop > 0.f;
  const T* dbr2 = w.dz2;                                 // gradient entering the linear2 branch (masked copy under dropout3)
  if (drop) { CQ_TRY(dropout_apply<T>(w.dz2, nullptr, w.dz1, rows * kC, pdrop, seed, 3, st)); dbr2 = w.dz1; }
  CQ_TRY(wgrad<T>(dbr2, kC, w.h, F, G[E_L2_W], F, G[E_L2_B], rows, kC, F, nullptr, st));
  // dropout2: h holds the DROPPED activation, so (h > 0) is ReLU mask and keep mask at once; only the 1 / (1 - p) is left, folded
  // into the same epilogue (no dropout pass over the [rows x F] gradient)
  { Epilogue e; e.mul_aux = w.h; e.mul_mode = 1; if (drop) e.mul_scale = dropout_keep_scale(pdrop);
    CQ_TRY(gemm<T>(dbr2, kC, wt_l2, w.dh, F, rows, F, kC, e, nullptr, st)); }
  CQ_TRY(wgrad<T>(w.dh, F, w.x1, kC, G[E_L1_W], kC, G[E_L1_B], rows, F, kC, nullptr, st));
  { Epilogue e; e.res = w.dz2; e.ldr = kC; CQ_TRY(gemm<T>(w.dh, F, wt_l1, w.dx1, kC, rows, kC, F, e, nullptr, st)); }
  // x1 = LN1(z1), z1 = src + dropout1(output_proj(samp))
  CQ_TRY(ln_bwd<T>(w.z1, nullptr, Wf(E_N1_W), 1e-5f, w.dx1, false, 0, 0, 0, w.dz1, 0.f, nullptr, 0.f, G[E_N1_W], G[E_N1_B], rows, st));
  const T* dbr1 = w.dz1;                                 // dx1 is free again: masked copy of dz1 for the output_proj branch
  if (drop) { CQ_TRY(dropout_apply<T>(w.dz1, nullptr, w.dx1, rows * kC, pdrop, seed, 1, st)); dbr1 = w.dx1; }
  CQ_TRY(wgrad<T>(dbr1, kC, w.samp, kC, G[E_OUT_W], kC, G[E_OUT_B], rows, kC, kC, nullptr, st));
  { Epilogue e; CQ_TRY(gemm<T>(dbr1, kC, wt_out, w.dsamp, kC, rows, kC, kC, e, nullptr, st)); }
  // sampling
  CQ_CUDA(cudaMemsetAsync(w.dval32, 0, (size_t)rows * kC * 4, st));
  CQ_TRY(cqvad_msda3d_backward(DT<T>::id, w.value, shapes, lsi, w.loc, w.attn, w.dsamp, w.dval32, w.dloc, w.dattn, B, (int)Len, kM,
                               kC / kM, L, (int)Len, P, (void*)st));
  msda_prepare_bwd_kernel<T><<<(unsigned)cdiv(rows * kM * 32, 256), 256, 0, st>>>(w.attn, w.dattn, w.dloc, shapes, w.doffT, w.dlgT,
                                                                                 rows, L, P);
  CQ_LAUNCH_CHECK();
  // the two query projections: dq = doff . Woff + dlogit . Watt  (written straight into grad_pos: q = src + pos)
  CQ_TRY(wgrad<T>(w.doffT, LP3, w.q, kC, G[E_OFF_W], kC, G[E_OFF_B], rows, LP3, kC, nullptr, st));
  CQ_TRY(wgrad<T>(w.dlgT, LP1, w.q, kC, G[E_ATT_W], kC, G[E_ATT_B], rows, LP1, kC, nullptr, st));
  { Epilogue e; CQ_TRY(gemm<T>(w.doffT, LP3, wt_off, w.dq, kC, rows, kC, LP3, e, nullptr, st)); }
  { Epilogue e; e.res = w.dq; e.ldr = kC; CQ_TRY(gemm<T>(w.dlgT, LP1, wt_att, gpos, kC, rows, kC, LP1, e, nullptr, st)); }
  // value projection (padded tokens: value was overwritten with zeros, so no gradient flows through them)
  CQ_TRY(f32_to_t<T>(w.dval32, w.dvalT, 0.f, rows * kC, st));
  if (mask) { mask_rows_kernel<T><<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(w.dvalT, mask, rows); CQ_LAUNCH_CHECK(); }
  CQ_TRY(wgrad<T>(w.dvalT, kC, src, kC, G[E_VAL_W], kC, G[E_VAL_B], rows, kC, kC, nullptr, st));
  // grad_src = dz1 (residual) + dvalue . Wv + dq
  { Epilogue e; e.res = w.dz1; e.ldr = kC; CQ_TRY(gemm<T>(w.dvalT, kC, wt_val, gsrc, kC, rows, kC, kC, e, nullptr, st)); }
  return axpby<T>(gsrc, gpos, 1.f, rows * kC, st);
}

// ---- encoder -> decoder data format ------------------------------------------------------------
// Transformer.forward (models/detr/dab_transformer.py:349-393) between the encoder and the decoder: un-flatten per level,
// make_interpolated_features (:239-294: F.grid_sample, align_corners = False, zeros padding) onto the (num_frames, H, W) grid of
// level -2, key-frame slice (`eff`, :378-382), rearrange to the decoder's "L (H W) (B T) C".  The reference materialises every
// level at every frame as [B, C, T, H, W] (+ two permuted copies) and then keeps ONE frame; here one warp produces one row of the
// decoder memory straight from the token-major encoder output: channels are the contiguous axis, so the <= 8 corner reads are
// 512-byte coalesced segments and only consumed frames are computed.
__device__ __forceinline__ float lin_m11(int i, int n) {   // torch.linspace(-1, 1, n)[i] (symmetric evaluation, RangeFactories.cu)
  if (n == 1) return -1.f;
  const float step = 2.f / (float)(n - 1);
  return i < n / 2 ? -1.f + step * (float)i : 1.f - step * (float)(n - 1 - i);
}
__device__ __forceinline__ float unnorm(float c, int size) { return ((c + 1.f) * (float)size - 1.f) * 0.5f; }

template <typename T>
__global__ void __launch_bounds__(256) interp_to_decoder_kernel(const T* __restrict__ tok, const T* __restrict__ postok,
                                                                const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                                                                T* __restrict__ mem, T* __restrict__ pos0, int L, int B, long Len,
                                                                int Tt, int H, int W, int nf, int eff) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int Tp = eff ? 1 : nf;
  const long BT = (long)B * Tp, S = (long)H * W;
  if (wid >= (long)L * S * BT) return;
  const long bt = wid % BT;
  const long s = (wid / BT) % S;
  const int l = (int)(wid / (BT * S));
  const int b = (int)(bt / Tp), fr = eff ? nf / 2 : (int)(bt % Tp);
  const int i = (int)(s / W), j = (int)(s % W);
  const int Tl = (int)shapes[l * 3], Hl = (int)shapes[l * 3 + 1], Wl = (int)shapes[l * 3 + 2];
  const T* base = tok + ((long)b * Len + lsi[l]) * kC + lane * 8;
  float x, y, lt = 0.f;
  int t0 = fr, nt = 1;
  if (Tt == nf) {   // per-frame 2-D sampling; the reference stacks the grid as (meshy, meshx): row coordinate -> x, column -> y
    x = unnorm(lin_m11(i, H), Wl); y = unnorm(lin_m11(j, W), Hl);
  } else {          // trilinear, grid (x, y, t) = (dw[j], dh[i], dt[fr])
    x = unnorm(lin_m11(j, W), Wl); y = unnorm(lin_m11(i, H), Hl);
    const float t = unnorm(lin_m11(fr, nf), Tl);
    t0 = (int)floorf(t); lt = t - (float)t0; nt = 2;
  }
  const int x0 = (int)floorf(x), y0 = (int)floorf(y);
  const float lx = x - (float)x0, ly = y - (float)y0;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int kt = 0; kt < nt; ++kt) {
    const int tz = t0 + kt;
    const float wt = nt == 1 ? 1.f : (kt ? lt : 1.f - lt);
    if (tz < 0 || tz >= Tl) continue;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
      const int yz = y0 + ky;
      if (yz < 0 || yz >= Hl) continue;
#pragma unroll
      for (int kx = 0; kx < 2; ++kx) {
        const int xz = x0 + kx;
        if (xz < 0 || xz >= Wl) continue;
        const float wgt = wt * (ky ? ly : 1.f - ly) * (kx ? lx : 1.f - lx);
        float v[8];
        load8(base + ((long)(tz * Hl + yz) * Wl + xz) * kC, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
      }
    }
  }
  store8(mem + wid * kC + lane * 8, acc);
  if (pos0 && l == L - 2) {   // pos of level -2, repeated in time (:285), no resampling
    float v[8];
    load8(postok + ((long)b * Len + lsi[l] + ((long)(fr % Tt) * H + i) * W + j) * kC + lane * 8, v);
    store8(pos0 + (s * BT + bt) * kC + lane * 8, v);
  }
}

// ---- pyramid level -> encoder tokens (Transformer.forward, dab_transformer.py:310-327) --------------------------------------
// src.flatten(2).transpose(1, 2) (+ level_embed[lvl] for the position embedding): channel-first [B, 256, N] -> token-major
// rows [B, level_start + n, 256] of the concatenated sequence, through a 32 x 33 shared-memory tile (coalesced on both sides).
template <typename T>
__global__ void __launch_bounds__(256) level_to_tokens_kernel(const T* __restrict__ x, const float* __restrict__ add, T* __restrict__ out,
                                                              long N, long Len, long level_start, int C = kC) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long n0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const T* xb = x + (long)b * C * N;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {                          // rows = channels, columns = positions (contiguous in x)
    const long n = n0 + tx;
    tile[r][tx] = n < N ? to_f(xb[(long)(c0 + r) * N + n]) : 0.f;
  }
  __syncthreads();
  T* ob = out + ((long)b * Len + level_start) * C;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {                          // rows = positions, columns = channels (contiguous in out)
    const long n = n0 + r;
    if (n < N) ob[n * C + c0 + tx] = from_f<T>(tile[tx][r] + (add ? add[c0 + tx] : 0.f));
  }
}

// Backward of interp_to_decoder_kernel: the same warp-per-memory-row geometry, each corner weight scatters the row's gradient into
// the fp32 token gradient (red.global.add.v4.f32; with `eff` only 4 * H*W * B rows exist, against B*Len token rows to zero-fill).
__device__ __forceinline__ void red_add_v4e(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__global__ void __launch_bounds__(256) interp_to_decoder_bwd_kernel(const float* __restrict__ gmem, const int64_t* __restrict__ shapes,
                                                                    const int64_t* __restrict__ lsi, float* __restrict__ gtok, int L,
                                                                    int B, long Len, int Tt, int H, int W, int nf, int eff) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int Tp = eff ? 1 : nf;
  const long BT = (long)B * Tp, S = (long)H * W;
  if (wid >= (long)L * S * BT) return;
  const long bt = wid % BT;
  const long s = (wid / BT) % S;
  const int l = (int)(wid / (BT * S));
  const int b = (int)(bt / Tp), fr = eff ? nf / 2 : (int)(bt % Tp);
  const int i = (int)(s / W), j = (int)(s % W);
  const int Tl = (int)shapes[l * 3], Hl = (int)shapes[l * 3 + 1], Wl = (int)shapes[l * 3 + 2];
  float* base = gtok + ((long)b * Len + lsi[l]) * kC + lane * 8;
  float x, y, lt = 0.f;
  int t0 = fr, nt = 1;
  if (Tt == nf) {
    x = unnorm(lin_m11(i, H), Wl); y = unnorm(lin_m11(j, W), Hl);
  } else {
    x = unnorm(lin_m11(j, W), Wl); y = unnorm(lin_m11(i, H), Hl);
    const float t = unnorm(lin_m11(fr, nf), Tl);
    t0 = (int)floorf(t); lt = t - (float)t0; nt = 2;
  }
  const int x0 = (int)floorf(x), y0 = (int)floorf(y);
  const float lx = x - (float)x0, ly = y - (float)y0;
  const float4 g0 = *reinterpret_cast<const float4*>(gmem + wid * kC + lane * 8);
  const float4 g1 = *reinterpret_cast<const float4*>(gmem + wid * kC + lane * 8 + 4);
  for (int kt = 0; kt < nt; ++kt) {
    const int tz = t0 + kt;
    const float wt = nt == 1 ? 1.f : (kt ? lt : 1.f - lt);
    if (tz < 0 || tz >= Tl) continue;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
      const int yz = y0 + ky;
      if (yz < 0 || yz >= Hl) continue;
#pragma unroll
      for (int kx = 0; kx < 2; ++kx) {
        const int xz = x0 + kx;
        if (xz < 0 || xz >= Wl) continue;
        const float wgt = wt * (ky ? ly : 1.f - ly) * (kx ? lx : 1.f - lx);
        float* dst = base + ((long)(tz * Hl + yz) * Wl + xz) * kC;
        red_add_v4e(dst, wgt * g0.x, wgt * g0.y, wgt * g0.z, wgt * g0.w);
        red_add_v4e(dst + 4, wgt * g1.x, wgt * g1.y, wgt * g1.z, wgt * g1.w);
      }
    }
  }
}

// Backward of level_to_tokens_kernel: token-major gradient rows -> channel-first gradient of the level (the same 32 x 33 tile the
// other way round) and, for the position embedding, d level_embed[lvl][c] = sum over clips and positions (block partial sums,
// one fp32 atomic per (block, channel)).
template <typename T>
__global__ void __launch_bounds__(256) level_to_tokens_bwd_kernel(const T* __restrict__ gtok, T* __restrict__ gx, float* __restrict__ gadd,
                                                                  long N, long Len, long level_start, int C = kC) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long n0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* gb = gtok + ((long)b * Len + level_start) * C;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {                          // rows = positions, columns = channels
    const long n = n0 + r;
    tile[r][tx] = n < N ? to_f(gb[n * C + c0 + tx]) : 0.f;
  }
  __syncthreads();
  if (gx) {
    T* xb = gx + (long)b * C * N;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {                        // rows = channels, columns = positions
      const long n = n0 + tx;
      if (n < N) xb[(long)(c0 + r) * N + n] = from_f<T>(tile[tx][r]);
    }
  }
  if (gadd && ty == 0) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) a += tile[r][tx];
    atomicAdd(gadd + c0 + tx, a);
  }
}

// ---- input projection of a backbone level (CSN configurations) ----------------------------------
// models/model.py:64-71,162-164: input_proj[l] = Conv3d(C_in, 256, kernel_size = 1) -> GroupNorm(32, 256).  The 1x1x1 conv is a
// GEMM over the level's positions; GroupNorm(32 groups of 8 channels) normalises over (8 channels x T*H*W) per clip: in the
// token-major layout a lane's 8 channels ARE one group, so the statistics are per-lane sums over rows (no cross-lane traffic
// until the block reduction) and the normalised rows land directly in the encoder's token sequence (conv + norm + flatten).
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ y, float* __restrict__ stats, long N, int rows_per_block) {
  __shared__ float red[8][32][2];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float s = 0.f, q = 0.f;
  for (long r = r0 + warp; r < r1; r += 8) {
    float v[8];
    load8(y + ((long)b * N + r) * kC + lane * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += v[e]; q = fmaf(v[e], v[e], q); }
  }
  red[warp][lane][0] = s; red[warp][lane][1] = q;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w][g][k];
    atomicAdd(stats + ((long)b * 32 + g) * 2 + k, a);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ y, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                       T* __restrict__ tokens, long N, long Len, long level_start, int B) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long)B * N) return;
  const int b = (int)(wid / N);
  const long n = wid % N;
  const float cnt = (float)N * 8.f;
  const float mean = stats[((long)b * 32 + lane) * 2] / cnt;
  const float var = fmaxf(stats[((long)b * 32 + lane) * 2 + 1] / cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float v[8], g[8], bt[8];
  load8(y + wid * kC + lane * 8, v);
  load8(gamma + lane * 8, g);
  load8(beta + lane * 8, bt);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd * g[e] + bt[e];
  store8(tokens + ((long)b * Len + level_start + n) * kC + lane * 8, v);
}

template <typename T>
int input_proj_t(const T* x, const T* W, const float* bias, const float* gamma, const float* beta, float eps, T* tokens, void* ws,
                 size_t ws_bytes, int B, int Cin, long N, long Len, long level_start, cudaStream_t st) {
  const long rows = (long)B * N;
  EncWs w(ws, ws_bytes);
  T* xt = (T*)w.take((size_t)rows * Cin * sizeof(T));
  T* y = (T*)w.take((size_t)rows * kC * sizeof(T));
  float* stats = (float*)w.take((size_t)B * 64 * 4);
  if (!xt || !y || !stats) return set_error(CQVAD_E_WORKSPACE, "input_proj: workspace too small");
  const dim3 tg((unsigned)cdiv(N, 32), (unsigned)(Cin / 32), (unsigned)B);
  level_to_tokens_kernel<T><<<tg, 256, 0, st>>>(x, nullptr, xt, N, N, 0, Cin);      // [B, Cin, N] -> [B*N, Cin]
  CQ_LAUNCH_CHECK();
  { Epilogue e; e.bias = bias; CQ_TRY(gemm<T>(xt, Cin, W, y, kC, rows, kC, Cin, e, nullptr, st)); }
  CQ_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * 64 * 4, st));
  const int rpb = 256;
  gn_stats_kernel<T><<<dim3((unsigned)cdiv(N, rpb), (unsigned)B), 256, 0, st>>>(y, stats, N, rpb);
  CQ_LAUNCH_CHECK();
  gn_apply_kernel<T><<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(y, stats, gamma, beta, eps, tokens, N, Len, level_start, B);
  CQ_LAUNCH_CHECK();
  return 0;
}

// extra pyramid level of the non-ViT configurations (models/model.py:72-76,166-170): Conv3d(C_in, 256, kernel_size = 3,
// stride = (1, 2, 2), padding = 1) -> GroupNorm(32, 256).  The level is tiny (7x7 -> 4x4 per frame): an im2col gather (27 taps,
// zero padding) into a [rows, 27 * C_in] matrix, then the same tcgen05 GEMM + group-norm kernels as the 1x1x1 levels.
// col[(b, t, yo, xo), tap * C_in + ci] = x[b, ci, t + kt - 1, 2 yo + ky - 1, 2 xo + kx - 1]
template <typename T>
__global__ void __launch_bounds__(256) im2col_3x3s2_kernel(const T* __restrict__ x, T* __restrict__ col, int Cin, int Tn, int H, int W,
                                                           int Ho, int Wo, long rows) {
  const long row = blockIdx.x;
  const int tap = blockIdx.y;
  if (row >= rows) return;
  const int kt = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
  const int xo = (int)(row % Wo), yo = (int)((row / Wo) % Ho), t = (int)((row / ((long)Wo * Ho)) % Tn);
  const long b = row / ((long)Wo * Ho * Tn);
  const int ti = t + kt - 1, yi = 2 * yo + ky - 1, xi = 2 * xo + kx - 1;
  const bool ok = ti >= 0 && ti < Tn && yi >= 0 && yi < H && xi >= 0 && xi < W;
  const long plane = (long)Tn * H * W;
  const T* src = x + b * Cin * plane + ((long)ti * H + yi) * W + xi;
  T* dst = col + (row * 27 + tap) * Cin;
  for (int ci = threadIdx.x; ci < Cin; ci += 256) dst[ci] = ok ? src[(long)ci * plane] : from_f<T>(0.f);
}

template <typename T>
int input_proj3_t(const T* x, const T* Wr, const float* bias, const float* gamma, const float* beta, float eps, T* tokens, void* ws,
                  size_t ws_bytes, int B, int Cin, int Tn, int H, int W, long Len, long level_start, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long N = (long)Tn * Ho * Wo, rows = (long)B * N;
  EncWs w(ws, ws_bytes);
  T* col = (T*)w.take((size_t)rows * 27 * Cin * sizeof(T));
  T* y = (T*)w.take((size_t)rows * kC * sizeof(T));
  float* stats = (float*)w.take((size_t)B * 64 * 4);
  if (!col || !y || !stats) return set_error(CQVAD_E_WORKSPACE, "input_proj (3x3x3): workspace too small");
  im2col_3x3s2_kernel<T><<<dim3((unsigned)rows, 27), 256, 0, st>>>(x, col, Cin, Tn, H, W, Ho, Wo, rows);
  CQ_LAUNCH_CHECK();
  { Epilogue e; e.bias = bias; CQ_TRY(gemm<T>(col, 27L * Cin, Wr, y, kC, rows, kC, 27 * Cin, e, nullptr, st)); }
  CQ_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * 64 * 4, st));
  gn_stats_kernel<T><<<dim3((unsigned)cdiv(N, 256), (unsigned)B), 256, 0, st>>>(y, stats, N, 256);
  CQ_LAUNCH_CHECK();
  gn_apply_kernel<T><<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(y, stats, gamma, beta, eps, tokens, N, Len, level_start, B);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

extern "C" int cqvad_deform_encoder_layer_num_weights(void) { return E_COUNT; }

extern "C" size_t cqvad_deform_encoder_layer_workspace_bytes(int dtype, int B, long Len, int L, int P, int F) {
  if (B < 0 || Len < 0 || L < 1 || P < 1 || F < 1) return 0;
  return dtype == CQVAD_F32 ? enc_ws_bytes<float>((long)B * Len, L, P, F) : enc_ws_bytes<bf16>((long)B * Len, L, P, F);
}

extern "C" int cqvad_deform_encoder_layer_forward(int dtype, const void* const* weights, const void* src, const void* pos,
                                                  const float* reference_points, const int64_t* shapes,
                                                  const int64_t* level_start, const uint8_t* padding_mask, void* out,
                                                  void* attn_out, void* workspace, size_t workspace_bytes, int B, long Len, int L,
                                                  int P, int F, void* stream) {
  CQ_CHECK_ARG(B >= 0 && Len >= 0 && L >= 1 && P >= 1 && F >= 1, "deform_encoder_layer: bad dimensions");
  if ((long)B * Len == 0) return 0;
  CQ_CHECK_ARG(weights && src && pos && reference_points && shapes && level_start && out && workspace,
               "deform_encoder_layer: null pointer");
  for (int i = 0; i < E_COUNT; ++i) CQ_CHECK_ARG(weights[i] != nullptr, "deform_encoder_layer: weight %d is null", i);
  CQ_CHECK_SHAPE(F % 8 == 0 && (kM * L * P) % 8 == 0, "deform_encoder_layer: F and heads*levels*points must be multiples of 8");
  CQ_CHECK_SHAPE(Len < (1L << 31) / kC, "deform_encoder_layer: Len*256 must fit int32 (as the MSDA op requires)");
  if (dtype == CQVAD_F32)
    return enc_layer_t<float>(weights, (const float*)src, (const float*)pos, reference_points, shapes, level_start, padding_mask,
                              (float*)out, (float*)attn_out, workspace, workspace_bytes, B, Len, L, P, F, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return enc_layer_t<bf16>(weights, (const bf16*)src, (const bf16*)pos, reference_points, shapes, level_start, padding_mask,
                             (bf16*)out, (bf16*)attn_out, workspace, workspace_bytes, B, Len, L, P, F, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "deform_encoder_layer: unknown dtype %d", dtype);
}

extern "C" size_t cqvad_deform_encoder_layer_train_workspace_bytes(int dtype, int B, long Len, int L, int P, int F) {
  if (B < 0 || Len < 0 || L < 1 || P < 1 || F < 1) return 0;
  return dtype == CQVAD_F32 ? EncTrainWs<float>(nullptr, 0, (long)B * Len, L, P, F).bytes
                            : EncTrainWs<bf16>(nullptr, 0, (long)B * Len, L, P, F).bytes;
}

extern "C" int cqvad_deform_encoder_layer_train_forward(int dtype, const void* const* weights, const void* src, const void* pos,
                                                        const float* reference_points, const int64_t* shapes,
                                                        const int64_t* level_start, const uint8_t* padding_mask, void* out,
                                                        void* workspace, size_t workspace_bytes, int B, long Len, int L, int P,
                                                        int F, float dropout_p, uint64_t seed, void* stream) {
  CQ_CHECK_ARG(B >= 0 && Len >= 0 && L >= 1 && P >= 1 && F >= 1 && dropout_p >= 0.f && dropout_p < 1.f, "deform_encoder_layer: bad dimensions");
  if ((long)B * Len == 0) return 0;
  CQ_CHECK_ARG(weights && src && pos && reference_points && shapes && level_start && out && workspace,
               "deform_encoder_layer: null pointer");
  for (int i = 0; i < E_COUNT; ++i) CQ_CHECK_ARG(weights[i] != nullptr, "deform_encoder_layer: weight %d is null", i);
  CQ_CHECK_SHAPE(F % 8 == 0 && (kM * L * P) % 8 == 0, "deform_encoder_layer: F and heads*levels*points must be multiples of 8");
  if (dtype == CQVAD_F32)
    return enc_train_fwd_t<float>(weights, (const float*)src, (const float*)pos, reference_points, shapes, level_start, padding_mask,
                                  (float*)out, workspace, workspace_bytes, B, Len, L, P, F, dropout_p, seed, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return enc_train_fwd_t<bf16>(weights, (const bf16*)src, (const bf16*)pos, reference_points, shapes, level_start, padding_mask,
                                 (bf16*)out, workspace, workspace_bytes, B, Len, L, P, F, dropout_p, seed, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "deform_encoder_layer: unknown dtype %d", dtype);
}

extern "C" int cqvad_deform_encoder_layer_backward(int dtype, const void* const* weights, const void* src, const int64_t* shapes,
                                                   const int64_t* level_start, const uint8_t* padding_mask, const void* grad_out,
                                                   void* grad_src, void* grad_pos, float* const* grad_weights, void* workspace,
                                                   size_t workspace_bytes, int B, long Len, int L, int P, int F, float dropout_p,
                                                   uint64_t seed, void* stream) {
  CQ_CHECK_ARG(B >= 0 && Len >= 0 && L >= 1 && P >= 1 && F >= 1, "deform_encoder_layer: bad dimensions");
  if ((long)B * Len == 0) return 0;
  CQ_CHECK_ARG(weights && src && shapes && level_start && grad_out && grad_src && grad_pos && grad_weights && workspace,
               "deform_encoder_layer_backward: null pointer");
  for (int i = 0; i < E_COUNT; ++i)
    CQ_CHECK_ARG(weights[i] != nullptr && grad_weights[i] != nullptr, "deform_encoder_layer_backward: weight / gradient %d is null", i);
  if (dtype == CQVAD_F32)
    return enc_train_bwd_t<float>(weights, (const float*)src, shapes, level_start, padding_mask, (const float*)grad_out,
                                  (float*)grad_src, (float*)grad_pos, grad_weights, workspace, workspace_bytes, B, Len, L, P, F,
                                  dropout_p, seed, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return enc_train_bwd_t<bf16>(weights, (const bf16*)src, shapes, level_start, padding_mask, (const bf16*)grad_out,
                                 (bf16*)grad_src, (bf16*)grad_pos, grad_weights, workspace, workspace_bytes, B, Len, L, P, F,
                                 dropout_p, seed, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "deform_encoder_layer_backward: unknown dtype %d", dtype);
}

extern "C" int cqvad_encoder_to_decoder_memory(int dtype, const void* tokens, const void* pos_tokens, const int64_t* shapes,
                                               const int64_t* level_start, int L, int B, long Len, int Tt, int H, int W,
                                               int num_frames, int eff, void* memory, void* pos0, void* stream) {
  CQ_CHECK_ARG(L >= 2 && B >= 0 && Len >= 0 && Tt >= 1 && H >= 1 && W >= 1 && num_frames >= 1, "encoder_to_decoder_memory: bad dimensions");
  const long rows = (long)L * H * W * B * (eff ? 1 : num_frames);
  if (rows == 0) return 0;
  CQ_CHECK_ARG(tokens && shapes && level_start && memory && (pos0 == nullptr || pos_tokens != nullptr),
               "encoder_to_decoder_memory: null pointer");
  const unsigned grid = (unsigned)cdiv(rows * 32, 256);
  if (dtype == CQVAD_F32)
    interp_to_decoder_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)tokens, (const float*)pos_tokens, shapes,
                                                                         level_start, (float*)memory, (float*)pos0, L, B, Len, Tt, H,
                                                                         W, num_frames, eff ? 1 : 0);
  else if (dtype == CQVAD_BF16)
    interp_to_decoder_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)tokens, (const bf16*)pos_tokens, shapes,
                                                                        level_start, (bf16*)memory, (bf16*)pos0, L, B, Len, Tt, H, W,
                                                                        num_frames, eff ? 1 : 0);
  else
    return set_error(CQVAD_E_INVALID_ARG, "encoder_to_decoder_memory: unknown dtype %d", dtype);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int cqvad_level_to_tokens(int dtype, const void* x, const float* add, void* tokens, int B, long N, long Len,
                                     long level_start, void* stream) {
  CQ_CHECK_ARG(B >= 0 && N >= 0 && Len >= N && level_start >= 0 && level_start + N <= Len, "level_to_tokens: bad dimensions");
  if ((long)B * N == 0) return 0;
  CQ_CHECK_ARG(x && tokens, "level_to_tokens: null pointer");
  CQ_CHECK_SHAPE(B <= 65535, "level_to_tokens: batch %d > 65535", B);
  const dim3 grid((unsigned)cdiv(N, 32), kC / 32, (unsigned)B);
  if (dtype == CQVAD_F32)
    level_to_tokens_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, add, (float*)tokens, N, Len, level_start);
  else if (dtype == CQVAD_BF16)
    level_to_tokens_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)x, add, (bf16*)tokens, N, Len, level_start);
  else
    return set_error(CQVAD_E_INVALID_ARG, "level_to_tokens: unknown dtype %d", dtype);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int cqvad_encoder_to_decoder_memory_backward(int dtype, const float* grad_memory, const int64_t* shapes,
                                                        const int64_t* level_start, int L, int B, long Len, int Tt, int H, int W,
                                                        int num_frames, int eff, void* grad_tokens, float* workspace, void* stream) {
  CQ_CHECK_ARG(L >= 2 && B >= 0 && Len >= 0 && Tt >= 1 && H >= 1 && W >= 1 && num_frames >= 1,
               "encoder_to_decoder_memory_backward: bad dimensions");
  const long rows = (long)L * H * W * B * (eff ? 1 : num_frames);
  const long n = (long)B * Len * kC;
  if (n == 0) return 0;
  CQ_CHECK_ARG(grad_memory && shapes && level_start && grad_tokens, "encoder_to_decoder_memory_backward: null pointer");
  CQ_CHECK_ARG(dtype == CQVAD_F32 || dtype == CQVAD_BF16, "encoder_to_decoder_memory_backward: unknown dtype");
  CQ_CHECK_ARG(dtype == CQVAD_F32 || workspace != nullptr, "encoder_to_decoder_memory_backward: bf16 needs the fp32 workspace [B*Len*256]");
  cudaStream_t st = as_stream(stream);
  float* acc = dtype == CQVAD_F32 ? (float*)grad_tokens : workspace;
  CQ_CUDA(cudaMemsetAsync(acc, 0, (size_t)n * sizeof(float), st));
  if (rows > 0) {
    interp_to_decoder_bwd_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, st>>>(grad_memory, shapes, level_start, acc, L, B, Len, Tt, H,
                                                                                 W, num_frames, eff ? 1 : 0);
    CQ_LAUNCH_CHECK();
  }
  if (dtype == CQVAD_BF16) CQ_TRY(f32_to_t<bf16>(acc, (bf16*)grad_tokens, 0.f, n, st));
  return 0;
}

extern "C" int cqvad_level_to_tokens_backward(int dtype, const void* grad_tokens, void* grad_x, float* grad_add, int B, long N, long Len,
                                              long level_start, void* stream) {
  CQ_CHECK_ARG(B >= 0 && N >= 0 && Len >= N && level_start >= 0 && level_start + N <= Len, "level_to_tokens_backward: bad dimensions");
  if ((long)B * N == 0) return 0;
  CQ_CHECK_ARG(grad_tokens && (grad_x || grad_add), "level_to_tokens_backward: null pointer");
  CQ_CHECK_SHAPE(B <= 65535, "level_to_tokens_backward: batch %d > 65535", B);
  const dim3 grid((unsigned)cdiv(N, 32), kC / 32, (unsigned)B);
  if (dtype == CQVAD_F32)
    level_to_tokens_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)grad_tokens, (float*)grad_x, grad_add, N, Len,
                                                                           level_start);
  else if (dtype == CQVAD_BF16)
    level_to_tokens_bwd_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)grad_tokens, (bf16*)grad_x, grad_add, N, Len,
                                                                          level_start);
  else
    return set_error(CQVAD_E_INVALID_ARG, "level_to_tokens_backward: unknown dtype %d", dtype);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t cqvad_input_proj_workspace_bytes(int dtype, int B, int Cin, long N) {
  const size_t es = dtype == CQVAD_F32 ? 4 : 2;
  return (size_t)B * N * Cin * es + (size_t)B * N * kC * es + (size_t)B * 64 * 4 + 4 * 256 + 256;
}

extern "C" int cqvad_input_proj_1x1_gn(int dtype, const void* x, const void* weight, const float* bias, const float* gn_weight,
                                       const float* gn_bias, float eps, void* tokens, void* workspace, size_t workspace_bytes,
                                       int B, int Cin, long N, long Len, long level_start, void* stream) {
  CQ_CHECK_ARG(B >= 0 && Cin >= 32 && N >= 0 && Len >= N && level_start >= 0 && level_start + N <= Len, "input_proj: bad dimensions");
  if ((long)B * N == 0) return 0;
  CQ_CHECK_ARG(x && weight && gn_weight && gn_bias && tokens && workspace, "input_proj: null pointer");
  CQ_CHECK_SHAPE(Cin % 32 == 0 && B <= 65535, "input_proj: C_in must be a multiple of 32 (got %d), batch <= 65535", Cin);
  if (dtype == CQVAD_F32)
    return input_proj_t<float>((const float*)x, (const float*)weight, bias, gn_weight, gn_bias, eps, (float*)tokens, workspace,
                               workspace_bytes, B, Cin, N, Len, level_start, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return input_proj_t<bf16>((const bf16*)x, (const bf16*)weight, bias, gn_weight, gn_bias, eps, (bf16*)tokens, workspace,
                              workspace_bytes, B, Cin, N, Len, level_start, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "input_proj: unknown dtype %d", dtype);
}

extern "C" int cqvad_msda3d_prepare(const float* offsets, const float* logits, const float* reference_points, const int64_t* shapes,
                                    float* loc, float* attn, long rows, int L, int P, void* stream) {
  CQ_CHECK_ARG(rows >= 0 && L >= 1 && P >= 1, "msda3d_prepare: bad dimensions");
  if (rows == 0) return 0;
  CQ_CHECK_ARG(offsets && logits && reference_points && shapes && loc && attn, "msda3d_prepare: null pointer");
  msda_prepare_kernel<<<(unsigned)cdiv(rows * kM * 32, 256), 256, 0, as_stream(stream)>>>(offsets, logits, reference_points, shapes, loc,
                                                                                         attn, rows, L, P);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t cqvad_input_proj_3x3s2_workspace_bytes(int dtype, int B, int Cin, int T, int H, int W) {
  const size_t es = dtype == CQVAD_F32 ? 4 : 2;
  const size_t rows = (size_t)B * T * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1);
  return rows * 27 * Cin * es + rows * kC * es + (size_t)B * 64 * 4 + 4 * 256 + 256;
}

extern "C" int cqvad_input_proj_3x3s2_gn(int dtype, const void* x, const void* weight_taps, const float* bias, const float* gn_weight,
                                         const float* gn_bias, float eps, void* tokens, void* workspace, size_t workspace_bytes,
                                         int B, int Cin, int T, int H, int W, long Len, long level_start, void* stream) {
  CQ_CHECK_ARG(B >= 0 && Cin >= 1 && T >= 1 && H >= 1 && W >= 1 && level_start >= 0, "input_proj (3x3x3): bad dimensions");
  const long N = (long)T * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1);
  CQ_CHECK_ARG(level_start + N <= Len, "input_proj (3x3x3): level does not fit the token sequence");
  if ((long)B * N == 0) return 0;
  CQ_CHECK_ARG(x && weight_taps && gn_weight && gn_bias && tokens && workspace, "input_proj (3x3x3): null pointer");
  CQ_CHECK_SHAPE(Cin % 8 == 0 && (long)B * N <= 2147483647L, "input_proj (3x3x3): C_in must be a multiple of 8 (got %d)", Cin);
  if (dtype == CQVAD_F32)
    return input_proj3_t<float>((const float*)x, (const float*)weight_taps, bias, gn_weight, gn_bias, eps, (float*)tokens, workspace,
                                workspace_bytes, B, Cin, T, H, W, Len, level_start, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return input_proj3_t<bf16>((const bf16*)x, (const bf16*)weight_taps, bias, gn_weight, gn_bias, eps, (bf16*)tokens, workspace,
                               workspace_bytes, B, Cin, T, H, W, Len, level_start, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "input_proj (3x3x3): unknown dtype %d", dtype);
}
