// Decoder driver: TransformerDecoder.forward (models/detr/dab_transformer.py:722-852) + DETR heads
// (models/model.py:191-221) as one natively-enqueued kernel sequence, plus the C-ABI wrappers of the building blocks.
//
// Not a translation of the reference graph (about 130 aten ops per layer on (nq,BT,...)-major tensors with permute /
// expand / cat copies): activations live as row-major [rows,256] matrices in instance-major order
// (instance i = n*BT + b; pixels (i,s); class tokens (i,k)), the ConvBlock chain runs on a y-padded NHWC layout
// [N, h+1, w, 256] whose zero separator rows give the 3x3 conv its vertical halo for free, per-head
// [content | position] concatenations are never materialised, and bias / activation / residual / LayerNorm are
// GEMM epilogues.
#include <string.h>
#include <string>
#include <vector>
#include <mutex>
#include "common.cuh"
#include <stdlib.h>
#include "kernels_mem.cuh"
#include "attention.cuh"
#include "prof.cuh"
#include "wtable.cuh"

namespace cqvad {

int cls_xattn_tc(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vt, long ldvt,
                 const float* bv, bf16* out, long N, int K, int S, int Sq, int Sp_rows, int BT, cudaStream_t st);
int cls_sattn_tc(const bf16* x, const bf16* xt, long ldxt, bf16* out, long N, int K, int K8, cudaStream_t st);

// ---- error / launch-count state (thread-local) -------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local long g_launches = 0;
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
namespace {
constexpr int kAuxDev = 32, kAuxStreams = 8, kAuxEvents = 96;
cudaStream_t g_aux_streams[kAuxDev][kAuxStreams];
cudaEvent_t g_aux_events[kAuxDev][kAuxEvents];
std::mutex g_aux_mu;
}  // namespace
cudaStream_t aux_stream(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kAuxDev || slot < 0 || slot >= kAuxStreams) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  if (!g_aux_streams[dev][slot] && cudaStreamCreateWithFlags(&g_aux_streams[dev][slot], cudaStreamNonBlocking) != cudaSuccess)
    g_aux_streams[dev][slot] = nullptr;
  return g_aux_streams[dev][slot];
}
cudaEvent_t aux_event(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kAuxDev || slot < 0 || slot >= kAuxEvents) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  if (!g_aux_events[dev][slot] && cudaEventCreateWithFlags(&g_aux_events[dev][slot], cudaEventDisableTiming) != cudaSuccess)
    g_aux_events[dev][slot] = nullptr;
  return g_aux_events[dev][slot];
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }
long launch_count_now() { return g_launches; }

template <typename T>
struct Decoder {
  const cqvad_decoder_desc& d;
  const void* const* w;
  cudaStream_t st;
  int BT, nq, h, wd, S, Sp, Sq, K, F, Lr;   // Sq: per-instance row pitch of q_memory (S rounded up to 8)
  long N, NS, NSq, Rp, NK;

  // buffers
  T *memc, *pos0c, *e512, *qpos, *tmpN, *pscale, *qse, *saq, *sak, *sav, *sao, *out, *actor, *acls, *qm, *kv, *kp, *qc,
      *qs, *cao, *XA, *XB, *Xn, *Hc, *Qc[2], *cq1, *cq2, *saoc, *kx, *vx, *vt, *qt, *cqp, *caoc, *cls0, *Hf, *hsn, *bb1, *bb2;
  long ldvt, ldqt; int K8;
  float *r_cur, *r_next, *lvlw, *cls0_32, *clsout_32;   // *_32: fp32 side copies of the class-token stream (bf16 path)
  // Two-stream schedule of the inference forward: the localisation chain of layer l+1 (small-row latency-bound kernels, the
  // k/v projection of q_memory, the level mix) depends only on the localisation chain of layer l, so it runs ahead on the
  // caller's stream while the class branch of layer l (ConvBlocks, class attention, class FFN: the large kernels) runs on a side
  // stream.  What the class branch reads from the localisation chain is kept per layer (actor, query sine embedding) or
  // double-buffered with back-pressure (q_memory, 48 MB at B = 32).  CQVAD_INFER_STREAMS=1 restores the single-stream order.
  static constexpr int kMaxLayers = 16;
  T *actor_l[kMaxLayers], *qse_l[kMaxLayers], *qm_l[2], *tmpC;
  cudaStream_t st0 = nullptr, st1 = nullptr;
  bool two_streams = false;

  Decoder(const cqvad_decoder_desc& dd, const void* const* ww, cudaStream_t s) : d(dd), w(ww), st(s) {
    BT = d.BT; nq = d.nq; h = d.h; wd = d.w; S = h * wd; Sp = (h + 1) * wd; K = d.K; F = d.F; Lr = d.layers;
    Sq = (S + 7) & ~7;
    N = (long)nq * BT; NS = N * S; NSq = N * Sq; Rp = N * Sp; NK = N * K;
  }
  size_t plan(Arena& a) {
    auto t = [&](long n) { return (T*)a.take((size_t)n * sizeof(T)); };
    auto f = [&](long n) { return (float*)a.take((size_t)n * sizeof(float)); };
    const long Fm = F > kC ? F : kC;
    memc = t(4L * S * BT * kC); pos0c = t((long)S * BT * kC);
    e512 = t(N * 512); qpos = t(N * kC); tmpN = t(N * Fm); pscale = t(N * kC); qse = t(N * kC);
    saq = t(N * kC); sak = t(N * kC); sav = t(N * kC); sao = t(N * kC); out = t(N * kC); actor = t(N * kC);
    acls = t(N * kC); qm = t(NSq * kC); kv = t(NSq * 2 * kC); kp = t((long)S * BT * kC); qc = t(N * kC); qs = t(N * kC);
    cao = t(N * kC); XA = t(Rp * kC); XB = t(Rp * kC); Xn = t(Rp * kC); Hc = t(Rp * 4 * kC);
    Qc[0] = t(NK * kC); Qc[1] = t(NK * kC); cq1 = t((long)K * kC); cq2 = t((long)K * kC); saoc = t(NK * kC);
    kx = t(Rp * kC); vx = t(NSq * kC); ldvt = NSq + 64; vt = t(kC * ldvt);
    K8 = (K + 7) & ~7; ldqt = N * K8 + 64; qt = t(kC * ldqt); cqp = t(N * kC); caoc = t(NK * kC); cls0 = t(NK * kC); Hf = t(NK * F);
    hsn = t(N * kC); bb1 = t(N * kC); bb2 = t(N * kC);
    r_cur = f(N * 4); r_next = f(N * 4); lvlw = f(N * 4);
    actor_l[0] = actor; qse_l[0] = qse; qm_l[0] = qm;
    for (int l = 1; l < Lr && l < kMaxLayers; ++l) { actor_l[l] = t(N * kC); qse_l[l] = t(N * kC); }
    qm_l[1] = t(NSq * kC); tmpC = t(N * Fm);
    cls0_32 = clsout_32 = nullptr;
    // fp32 side copies of the class-token stream: measured on B200 to leave the bf16 error unchanged (it is dominated by
    // GEMM operand rounding, tools/diag_bf16.py) while costing 0.33 ms/step -> off unless CQVAD_DEC_FP32_CLS_STREAM
    if (DT<T>::id == CQVAD_BF16 && (d.flags & 2)) { cls0_32 = f(NK * kC); clsout_32 = f(NK * kC); }
    return a.off;
  }

  const T* Wm(int i) const { return (const T*)w[i]; }
  const float* Wf(int i) const { return (const float*)w[i]; }
  int loc(int l, int s) const { return l * LOC_COUNT + s; }
  int cls(int l, int s) const { return Lr * LOC_COUNT + l * CLS_COUNT + s; }
  int glob(int s) const { return Lr * (LOC_COUNT + CLS_COUNT) + s; }

  // C[M,N] = act(A . W^T + b) (+res) (-> LN)
  int lin(const T* A, long M, int Kd, int widx, T* C, int Nn, int act = CQVAD_ACT_NONE, const T* res = nullptr,
          int ln_idx = -1, float eps = 1e-5f, long ldc = -1, long lda = -1) {
    Epilogue e;
    e.bias = Wf(widx + 1); e.act = act; e.res = res; e.ldr = ldc < 0 ? Nn : ldc;
    if (ln_idx >= 0) { e.ln_g = Wf(ln_idx); e.ln_b = Wf(ln_idx + 1); e.ln_eps = eps; }
    return gemm<T>(A, lda < 0 ? Kd : lda, Wm(widx), C, ldc < 0 ? Nn : ldc, M, Nn, Kd, e, nullptr, st);
  }
  // Y = LN?( res + W2.act(W1.X + b1) + b2 ), hidden in `hid` ([M,Fh]) unless the fused tensor-core kernel takes it
  int mlp(const T* X, long M, int Fh, int w1, int w2, int act, const T* res, int ln_idx, float eps, T* Y, T* hid,
          int zero_period = 0, int zero_valid = 0, bool emit_qt = false, const float* res32 = nullptr, float* y32 = nullptr);
  bool y32_valid = false;
  bool qt_valid = false;   // qt holds the transposed class tokens of the previous layer
  int sattn(int l, const T* Qprev, T* Qin);

  // class cross-attention (dab_transformer.py:1067-1071): fills caoc [N*K,256] from Qin, X3 (padded), qm, qse
  int xattn(int l, const T* Qin, const T* X3);
  int xattn_generic(int l, const T* Qin, const T* X3);

  int run(const float* tgt, const float* memory, const float* pos, const uint8_t* mask, const float* ref_u, void* hs,
          void* cls_hs, float* refs, float* pred_logits, float* pred_boxes, float* pred_logits_b);
};

template <>
int Decoder<float>::mlp(const float* X, long M, int Fh, int w1, int w2, int act, const float* res, int ln_idx, float eps,
                        float* Y, float* hid, int zero_period, int zero_valid, bool, const float*, float*) {
  CQ_TRY(lin(X, M, kC, w1, hid, Fh, act));
  Epilogue e;
  e.bias = Wf(w2 + 1); e.res = res; e.ldr = kC; e.zero_period = zero_period; e.zero_valid = zero_valid;
  if (ln_idx >= 0) { e.ln_g = Wf(ln_idx); e.ln_b = Wf(ln_idx + 1); e.ln_eps = eps; }
  return gemm<float>(hid, Fh, Wm(w2), Y, kC, M, kC, Fh, e, nullptr, st);
}
template <>
int Decoder<bf16>::mlp(const bf16* X, long M, int Fh, int w1, int w2, int act, const bf16* res, int ln_idx, float eps,
                       bf16* Y, bf16* hid, int zero_period, int zero_valid, bool emit_qt, const float* res32, float* y32) {
  if (emit_qt) qt_valid = false;
  if (y32) y32_valid = false;
  // the fused kernel walks the hidden dimension serially per 128-row tile: with few row tiles (small-row FFNs, M = nq*BT)
  // two plain GEMMs spread the F dimension over more SMs
  // CQVAD_INFER_UNFUSED_MLP=1 (experiment, measured SLOWER: 2 829 vs 2 898 clips/s inference at B = 32): the ConvBlock MLP as two
  // CTA-pair GEMMs (GELU in the lean TMA-store epilogue) instead of the fused kernel that keeps the hidden activation on the SM
  static const bool unfused_big = getenv("CQVAD_INFER_UNFUSED_MLP") != nullptr;
  const bool skip_fused = unfused_big && act == CQVAD_ACT_GELU && !emit_qt && !y32 && M >= 148 * 128;
  if (!force_simt() && M > 2048 && !skip_fused) {
    int r = mlp_tc(X, Wm(w1), Wf(w1 + 1), Wm(w2), Wf(w2 + 1), act, res, ln_idx >= 0 ? Wf(ln_idx) : nullptr,
                   ln_idx >= 0 ? Wf(ln_idx + 1) : nullptr, eps, Y, M, kC, Fh, zero_period, zero_valid, st,
                   emit_qt ? qt : nullptr, ldqt, K, K8, res32, y32);
    if (r == 0 && emit_qt) qt_valid = true;
    if (r == 0 && y32) y32_valid = true;
    if (r <= 0) return r;
  }
  CQ_TRY(lin(X, M, kC, w1, hid, Fh, act));
  Epilogue e;
  e.bias = Wf(w2 + 1); e.res = res; e.ldr = kC; e.zero_period = zero_period; e.zero_valid = zero_valid; e.res32 = res32;
  if (ln_idx >= 0) { e.ln_g = Wf(ln_idx); e.ln_b = Wf(ln_idx + 1); e.ln_eps = eps; }
  return gemm<bf16>(hid, Fh, Wm(w2), Y, kC, M, kC, Fh, e, nullptr, st);   // y32 not produced here: y32_valid stays false
}

// class-query self-attention :1063-1065 for layers >= 1 (q = k = v = previous layer's class tokens)
template <typename T>
static int sattn_generic(Decoder<T>& d, int l, const T* Qprev, T* Qin) {
  StdStrides ss{};
  ss.q_ls = ss.k_ls = ss.v_ls = ss.o_ls = kC;
  ss.q_bs = ss.k_bs = ss.v_bs = ss.o_bs = (long)d.K * kC;
  CQ_TRY(mha_std<T>(Qprev, nullptr, Qprev, nullptr, Qprev, nullptr, d.saoc, d.K, d.K, (int)d.N, kH, 32, 32, ss, d.st));
  return d.lin(d.saoc, d.NK, kC, d.cls(l, C_SA_O), Qin, kC, CQVAD_ACT_NONE, Qprev, d.cls(l, C_NORM1));
}
template <>
int Decoder<float>::sattn(int l, const float* Qprev, float* Qin) { return sattn_generic<float>(*this, l, Qprev, Qin); }
template <>
int Decoder<bf16>::sattn(int l, const bf16* Qprev, bf16* Qin) {
  if (force_simt() || !qt_valid || K > 128) return sattn_generic<bf16>(*this, l, Qprev, Qin);
  int r = cls_sattn_tc(Qprev, qt, ldqt, saoc, N, K, K8, st);
  if (r == 1) return sattn_generic<bf16>(*this, l, Qprev, Qin);
  if (r != 0) return r;
  return lin(saoc, NK, kC, cls(l, C_SA_O), Qin, kC, CQVAD_ACT_NONE, Qprev, cls(l, C_NORM1));
}

// class cross-attention :1067-1071.  512-wide q/k split contiguously into 8 heads of 64 (attention.py:336,339):
// heads 0-3 = (class query . k_proj(conv feature)), heads 4-7 = (actor sine pos . spatial pos).
template <typename T>
int Decoder<T>::xattn_generic(int l, const T* Qin, const T* X3) {
  { ProfScope ps(P_BIG_PROJ, st);
    CQ_TRY(lin(X3, Rp, kC, cls(l, C_KPROJ), kx, kC));
    CQ_TRY(lin(qm, NSq, kC, cls(l, C_VPROJ), vx, kC));
    CQ_TRY(lin(qse, N, kC, cls(l, C_QPS), cqp, kC)); }
  ProfScope ps(P_CLS_XATTN, st);
  StdStrides ss{};
  ss.q_ls = kC; ss.q_bs = (long)K * kC;             // Qin rows (i,k)
  ss.q2_ls = 0; ss.q2_bs = kC;                      // cqp row i, same for every class
  ss.k_ls = kC; ss.k_bs = (long)Sp * kC;            // kx on the padded layout
  ss.k2_ls = (long)BT * kC; ss.k2_bs = kC; ss.k2_bmod = BT;   // pos0[s, b], b = i % BT
  ss.v_ls = kC; ss.v_bs = (long)Sq * kC;
  ss.o_ls = kC; ss.o_bs = (long)K * kC;
  return mha_std<T>(Qin, cqp, kx, pos0c, vx, nullptr, caoc, K, S, (int)N, kH, 64, 32, ss, st);
}
template <>
int Decoder<float>::xattn(int l, const float* Qin, const float* X3) { return xattn_generic(l, Qin, X3); }
template <>
int Decoder<bf16>::xattn(int l, const bf16* Qin, const bf16* X3) {
  if (force_simt() || K > 128 || ((S + 15) & ~15) > 256) return xattn_generic(l, Qin, X3);
  { ProfScope ps(P_BIG_PROJ, st);
    CQ_TRY(lin(X3, Rp, kC, cls(l, C_KPROJ), kx, kC));
    // V^T[256, N*S] = W_v . q_memory^T : the v_proj 1x1 conv with swapped operands (bias added after the softmax-weighted
    // sum, exact because the weights sum to one)
    Epilogue e;
    CQ_TRY(gemm<bf16>(Wm(cls(l, C_VPROJ)), kC, qm, vt, ldvt, kC, (int)NSq, kC, e, nullptr, st));
    CQ_TRY(lin(qse, N, kC, cls(l, C_QPS), cqp, kC)); }
  ProfScope ps(P_CLS_XATTN, st);
  int r = cls_xattn_tc(Qin, cqp, kx, pos0c, vt, ldvt, Wf(cls(l, C_VPROJ) + 1), caoc, N, K, S, Sq, Sp, BT, st);
  if (r == 1) return set_error(CQVAD_E_UNSUPPORTED_SHAPE, "class cross-attention: shape rejected by the tensor-core kernel");
  return r;
}

template <typename T>
int Decoder<T>::run(const float* tgt, const float* memory, const float* pos, const uint8_t* mask, const float* ref_u,
                    void* hs, void* cls_hs, float* refs, float* pred_logits, float* pred_boxes, float* pred_logits_b) {
  const bool of32 = d.out_f32 != 0;
  const size_t osz = of32 ? sizeof(float) : sizeof(T);
  // ---- streams / events of the two-stream schedule (process-wide, created once) ----
  cudaStream_t side = aux_stream(0);
  cudaEvent_t ev_loc[kMaxLayers], ev_cls[kMaxLayers];
  bool ev_ok = true;
  for (int i = 0; i < kMaxLayers; ++i) {
    ev_loc[i] = aux_event(i); ev_cls[i] = aux_event(kMaxLayers + i);
    ev_ok = ev_ok && ev_loc[i] && ev_cls[i];
  }
  static const bool one_stream = [] { const char* e = getenv("CQVAD_INFER_STREAMS"); return e && atoi(e) == 1; }();
  st0 = st;
  two_streams = !one_stream && side != nullptr && ev_ok && Lr <= kMaxLayers;
  st1 = two_streams ? side : st0;
  // inputs -> compute dtype.  Only pos[0] is ever used (dab_transformer.py:958, :810).
  ProfScope* ps = new ProfScope(P_INPUT, st);
  CQ_TRY(convert_f32<T>(memory, memc, 4L * S * BT * kC, st));
  CQ_TRY(convert_f32<T>(pos, pos0c, (long)S * BT * kC, st));
  CQ_TRY(convert_f32<T>(tgt, out, N * kC, st));
  CQ_CUDA(cudaMemsetAsync(XA, 0, (size_t)Rp * kC * sizeof(T), st));  // zero separator rows (never written afterwards)
  CQ_CUDA(cudaMemsetAsync(XB, 0, (size_t)Rp * kC * sizeof(T), st));
  CQ_CUDA(cudaMemsetAsync(qt, 0, (size_t)kC * ldqt * sizeof(T), st));   // pad columns of the transposed class tokens stay finite
  qt_valid = false;
  if (Sq != S) CQ_CUDA(cudaMemsetAsync(qm, 0, (size_t)NSq * kC * sizeof(T), st));   // pad rows stay zero (finite)
  CQ_TRY(sigmoid4(ref_u, r_cur, refs, N, nq, BT, st));               // :735; refs[0]
  delete ps;
#define PROF(c) { delete ps; ps = new ProfScope(c, st); }
#define WORK(c, fl, by) { if (prof_enabled()) prof_work(c, fl, by); }
  ps = nullptr;
  const double eb = sizeof(T), NSd = (double)N * S;   // algorithmic work: valid rows only, operands once, result once

  if (Sq != S && two_streams) CQ_CUDA(cudaMemsetAsync(qm_l[1], 0, (size_t)NSq * kC * sizeof(T), st));
  for (int l = 0; l < Lr; ++l) {
    const bool first = (l == 0);
    st = st0;                       // ---- localisation chain: caller's stream ----
    if (two_streams) {
      actor = actor_l[l]; qse = qse_l[l]; qm = qm_l[l & 1];
      if (l >= 2) CQ_CUDA(cudaStreamWaitEvent(st0, ev_cls[l - 2], 0));   // q_memory buffer l & 1 is free again
    }
    // ---- prologue :742-763 ----
    PROF(P_SMALL);
    CQ_TRY(sine_embed<T>(r_cur, e512, N, st));
    CQ_TRY(lin(e512, N, 512, glob(G_RPH0), tmpN, kC, CQVAD_ACT_RELU));
    CQ_TRY(lin(tmpN, N, kC, glob(G_RPH1), qpos, kC));
    if (!first) {
      CQ_TRY(lin(out, N, kC, glob(G_QS0), tmpN, kC, CQVAD_ACT_RELU));
      CQ_TRY(lin(tmpN, N, kC, glob(G_QS1), pscale, kC));
    }
    CQ_TRY(lin(out, N, kC, glob(G_RAH0), tmpN, kC, CQVAD_ACT_RELU));
    CQ_TRY(qse_modulate<T>(r_cur, first ? nullptr : pscale, tmpN, Wf(glob(G_RAH1)), Wf(glob(G_RAH1) + 1), qse, N, st));

    // ---- localisation layer: self-attention over the nq actors of a frame :921-938 ----
    CQ_TRY(lin(out, N, kC, loc(l, SA_QC), saq, kC));
    CQ_TRY(lin(qpos, N, kC, loc(l, SA_QP), saq, kC, CQVAD_ACT_NONE, saq));
    CQ_TRY(lin(out, N, kC, loc(l, SA_KC), sak, kC));
    CQ_TRY(lin(qpos, N, kC, loc(l, SA_KP), sak, kC, CQVAD_ACT_NONE, sak));
    CQ_TRY(lin(out, N, kC, loc(l, SA_V), sav, kC));
    {
      StdStrides ss{};
      ss.q_ls = ss.k_ls = ss.v_ls = ss.o_ls = (long)BT * kC;
      ss.q_bs = ss.k_bs = ss.v_bs = ss.o_bs = kC;
      CQ_TRY(mha_std<T>(saq, nullptr, sak, nullptr, sav, nullptr, sao, nq, nq, BT, kH, 32, 32, ss, st));
    }
    CQ_TRY(lin(sao, N, kC, loc(l, SA_O), out, kC, CQVAD_ACT_NONE, out, loc(l, NORM1)));
    // ---- level-weighted query-specific memory :943-946 ----
    CQ_TRY(linear_smalln<T>(out, Wf(loc(l, LVLW)), Wf(loc(l, LVLW) + 1), lvlw, N, 4, true, st));
    PROF(P_LVLMIX);
    CQ_TRY(lvlmix_ln<T>(memc, lvlw, Wf(loc(l, NORMU)), Wf(loc(l, NORMU) + 1), qm, N, S, Sq, BT, st));
    // ---- cross-attention with per-actor keys :951-988 ----
    PROF(P_BIG_PROJ);
    WORK(P_BIG_PROJ, 2.0 * NSd * 512 * kC + 2.0 * S * BT * kC * kC, eb * (NSd * (kC + 512) + (double)S * BT * 2 * kC + 3.0 * kC * kC));
    if (w[loc(l, CA_KV)] != nullptr) {
      CQ_TRY(lin(qm, NSq, kC, loc(l, CA_KV), kv, 2 * kC));          // [k | v] in one pass over q_memory
    } else {
      CQ_TRY(lin(qm, NSq, kC, loc(l, CA_KC), kv, kC, CQVAD_ACT_NONE, nullptr, -1, 1e-5f, 2 * kC));
      CQ_TRY(lin(qm, NSq, kC, loc(l, CA_V), kv + kC, kC, CQVAD_ACT_NONE, nullptr, -1, 1e-5f, 2 * kC));
    }
    CQ_TRY(lin(pos0c, (long)S * BT, kC, loc(l, CA_KP), kp, kC));
    PROF(P_SMALL);
    CQ_TRY(lin(out, N, kC, loc(l, CA_QC), qc, kC));
    if (first) {
      CQ_CHECK_ARG(w[loc(l, CA_QP)] != nullptr, "layers.0.ca_qpos_proj is required");
      CQ_TRY(lin(qpos, N, kC, loc(l, CA_QP), qc, kC, CQVAD_ACT_NONE, qc));
    }
    CQ_TRY(lin(qse, N, kC, loc(l, CA_QS), qs, kC));
    PROF(P_LOC_QSK);
    CQ_TRY(dec_qsk_attn<T>(qc, qs, kv, kv + kC, 2 * kC, kp, mask, cao, N, S, Sq, BT, first, st));
    PROF(P_SMALL);
    CQ_TRY(lin(cao, N, kC, loc(l, CA_O), actor, kC, CQVAD_ACT_NONE, out, loc(l, NORM2)));   // tgt_temp :992-993
    if (two_streams) { delete ps; ps = nullptr; CQ_CUDA(cudaEventRecord(ev_loc[l], st0)); }   // actor, q_memory, qse of layer l are final
    PROF(P_SMALL);
    CQ_TRY(mlp(actor, N, F, loc(l, LIN1), loc(l, LIN2), CQVAD_ACT_RELU, actor, loc(l, NORM3), 1e-5f, out, tmpN));

    // ---- outputs of this layer :826-827 and heads (models/model.py:192-221) ----
    PROF(P_OUT_LN);
    CQ_TRY(layernorm_rows<T>(out, nullptr, Wf(glob(G_NORM)), Wf(glob(G_NORM) + 1), 1e-5f, hsn, false, N, st));
    CQ_TRY(layernorm_permute<T>(out, Wf(glob(G_NORM)), Wf(glob(G_NORM) + 1), 1e-5f, (char*)hs + (size_t)l * N * kC * osz,
                                of32, N, nq, BT, 1, nullptr, st));
    if (pred_logits_b)
      CQ_TRY(logits_b<T>(hsn, Wf(glob(G_CEB)), Wf(glob(G_CEB) + 1), pred_logits_b + (size_t)l * N * 3, N, nq, BT, st));
    if (pred_boxes) {
      CQ_TRY(lin(hsn, N, kC, glob(G_BB0), bb1, kC, CQVAD_ACT_RELU));
      CQ_TRY(lin(bb1, N, kC, glob(G_BB1), bb2, kC, CQVAD_ACT_RELU));
      CQ_TRY(box_refine<T>(bb2, Wf(glob(G_BB2)), Wf(glob(G_BB2) + 1), r_cur, nullptr, pred_boxes + (size_t)l * N * 4, N, nq,
                           BT, false, st));
    }
    // ---- iterative box refinement :813-823 ----
    PROF(P_SMALL);
    CQ_TRY(lin(out, N, kC, glob(G_BB0), bb1, kC, CQVAD_ACT_RELU));
    CQ_TRY(lin(bb1, N, kC, glob(G_BB1), bb2, kC, CQVAD_ACT_RELU));
    CQ_TRY(box_refine<T>(bb2, Wf(glob(G_BB2)), Wf(glob(G_BB2) + 1), r_cur, r_next,
                         (l != Lr - 1) ? refs + (size_t)(l + 1) * N * 4 : nullptr, N, nq, BT, false, st));
    float* t = r_cur; r_cur = r_next; r_next = t;
    delete ps; ps = nullptr;

    st = st1;                       // ---- class branch: side stream, one layer behind the localisation chain at most two ----
    if (two_streams) CQ_CUDA(cudaStreamWaitEvent(st1, ev_loc[l], 0));
    // ---- class-query layer :1040-1079 ----
    CQ_TRY(mlp(actor, N, F, cls(l, C_L1), cls(l, C_L2), CQVAD_ACT_RELU, actor, cls(l, C_NORM), 1e-5f, acls, tmpC));
    PROF(P_ADDLN);
    CQ_TRY(add_ln_pad<T>(acls, qm, Wf(cls(l, C_CONVNORM)), Wf(cls(l, C_CONVNORM) + 1), XA, N, S, Sq, Sp, st));
    T* xin = XA; T* xout = XB;
    for (int blk = 0; blk < 3; ++blk) {   // the same ConvBlock three times (:1017-1018, :1055-1056)
      Epilogue e;
      e.bias = Wf(cls(l, C_CONV1) + 1);
      e.ln_g = Wf(cls(l, C_CBNORM)); e.ln_b = Wf(cls(l, C_CBNORM) + 1); e.ln_eps = 1e-6f;
      ConvGeom cg; cg.h = h; cg.w = wd;
      PROF(P_CONV);
      WORK(P_CONV, 2.0 * NSd * kC * 9 * kC, eb * (2.0 * NSd * kC + 9.0 * kC * kC));
      CQ_TRY(gemm<T>(xin, kC, Wm(cls(l, C_CONV1)), Xn, kC, Rp, kC, 9 * kC, e, &cg, st));
      PROF(P_CONV_MLP);
      WORK(P_CONV_MLP, 2.0 * NSd * kC * 4 * kC * 2, eb * (3.0 * NSd * kC + 8.0 * kC * kC));
      CQ_TRY(mlp(Xn, Rp, 4 * kC, cls(l, C_CONV2), cls(l, C_CONV3), CQVAD_ACT_GELU, xin, -1, 0.f, xout, Hc, Sp, S));
      T* t = xin; xin = xout; xout = t;
    }
    const T* X3 = xin;
    PROF(P_CLS_SATTN);
    // class-query self-attention :1059-1065
    T* Qin = Qc[l & 1];
    const T* Qprev = Qc[(l + 1) & 1];
    if (first) {   // identical for every actor instance: computed once on K rows, then broadcast
      StdStrides ss{};
      ss.q_ls = ss.k_ls = ss.v_ls = ss.o_ls = kC;
      CQ_TRY(mha_std<T>(Wm(glob(G_CQ)), nullptr, Wm(glob(G_CQ)), nullptr, Wm(glob(G_CQ)), nullptr, cq1, K, K, 1, kH, 32, 32, ss, st));
      CQ_TRY(lin(cq1, K, kC, cls(l, C_SA_O), cq2, kC, CQVAD_ACT_NONE, Wm(glob(G_CQ)), cls(l, C_NORM1)));
      CQ_TRY(broadcast_rows<T>(cq2, Qin, NK, K, st));
    } else {
      CQ_TRY(sattn(l, Qprev, Qin));
    }
    delete ps; ps = nullptr;
    CQ_TRY(xattn(l, Qin, X3));
    PROF(P_CLS_OPROJ);
    {
      Epilogue e;
      e.bias = Wf(cls(l, C_CA_O) + 1); e.c32 = cls0_32;
      CQ_TRY(gemm<T>(caoc, kC, Wm(cls(l, C_CA_O)), cls0, kC, NK, kC, kC, e, nullptr, st));
    }
    PROF(P_CLS_FFN);
    WORK(P_CLS_FFN, 2.0 * (double)NK * kC * F * 2, eb * (2.0 * (double)NK * kC + 2.0 * (double)F * kC));
    T* cls_out = Qc[l & 1];   // Qin is dead after the attention; reuse its buffer for the layer output / next query
    CQ_TRY(mlp(cls0, NK, F, cls(l, C_L1_), cls(l, C_L2_), CQVAD_ACT_RELU, cls0, cls(l, C_NORM_), 1e-5f, cls_out, Hf, 0, 0,
               /*emit_qt=*/l + 1 < Lr, cls0_32, clsout_32));

    PROF(P_OUT_LN);
    {
      void* dst = cls_hs ? (void*)((char*)cls_hs + (size_t)l * NK * kC * osz) : nullptr;
      float* lg = pred_logits ? pred_logits + (size_t)l * NK : nullptr;
      if ((dst || lg) && clsout_32 && y32_valid)   // bf16 path: cls_norm2 reads the fp32 copy of the class tokens
        CQ_TRY(layernorm_permute_f32in(clsout_32, Wf(glob(G_CLSNORM2)), Wf(glob(G_CLSNORM2) + 1), 1e-5f, dst, of32, NK, nq,
                                       BT, K, lg, st));
      else if (dst || lg)
        CQ_TRY(layernorm_permute<T>(cls_out, Wf(glob(G_CLSNORM2)), Wf(glob(G_CLSNORM2) + 1), 1e-5f, dst, of32, NK, nq, BT,
                                    K, lg, st));
    }
    delete ps; ps = nullptr;
    if (two_streams) CQ_CUDA(cudaEventRecord(ev_cls[l], st1));
  }
  st = st0;
  if (two_streams) CQ_CUDA(cudaStreamWaitEvent(st0, ev_cls[Lr - 1], 0));   // join: the caller's stream owns every output again
  delete ps;
#undef PROF
#undef WORK
  return 0;
}

template <typename T>
static int run_decoder(const cqvad_decoder_desc* d, const void* const* weights, const float* tgt, const float* memory,
                       const float* pos, const uint8_t* mask, const float* ref_u, void* hs, void* cls_hs, float* refs,
                       float* pl, float* pb, float* plb, void* ws, size_t ws_bytes, cudaStream_t st) {
  Decoder<T> dec(*d, weights, st);
  const size_t skew = (1024 - (((uintptr_t)ws) & 1023)) & 1023;   // tiles / TMA bases want 1024-byte alignment
  if (ws_bytes < skew) return set_error(CQVAD_E_WORKSPACE, "decoder workspace too small");
  Arena a((char*)ws + skew, ws_bytes - skew);
  dec.plan(a);
  if (a.overflow)
    return set_error(CQVAD_E_WORKSPACE, "decoder workspace too small (%zu needed, %zu given)", a.off + 1024, ws_bytes);
  return dec.run(tgt, memory, pos, mask, ref_u, hs, cls_hs, refs, pl, pb, plb);
}

static int check_desc(const cqvad_decoder_desc* d) {
  CQ_CHECK_ARG(d != nullptr, "decoder: null descriptor");
  CQ_CHECK_ARG(d->dtype == CQVAD_F32 || d->dtype == CQVAD_BF16, "decoder: unknown dtype %d", d->dtype);
  CQ_CHECK_ARG(d->BT >= 1 && d->nq >= 1 && d->h >= 1 && d->w >= 1 && d->K >= 1 && d->layers >= 1, "decoder: bad extents");
  CQ_CHECK_SHAPE(d->F >= 8 && d->F % 8 == 0, "decoder: dim_feedforward must be a multiple of 8");
  CQ_CHECK_SHAPE(d->w <= 128, "decoder: feature-map width %d > 128 not supported", d->w);
  return 0;
}

}  // namespace cqvad

using namespace cqvad;

extern "C" int cqvad_version(void) { return CQVAD_VERSION; }
extern "C" const char* cqvad_last_error(void) { return g_err; }
extern "C" long cqvad_last_launch_count(void) { return g_launches; }
extern "C" void cqvad_debug_force_simt(int v) { set_force_simt(v != 0); }

extern "C" int cqvad_decoder_num_weights(int layers) { return layers * (LOC_COUNT + CLS_COUNT) + GLOB_COUNT; }
extern "C" const char* cqvad_decoder_weight_name(int idx, int layers) {
  if (!weight_slot(idx, layers, &g_name, nullptr)) return nullptr;
  return g_name.c_str();
}
extern "C" int cqvad_decoder_weight_kind(int idx, int layers) {
  int k = -1;
  weight_slot(idx, layers, nullptr, &k);
  return k;
}

extern "C" size_t cqvad_decoder_workspace_bytes(const cqvad_decoder_desc* d) {
  if (check_desc(d) != 0) return 0;
  Arena a(nullptr, 0);
  if (d->dtype == CQVAD_F32) { Decoder<float> dec(*d, nullptr, nullptr); dec.plan(a); }
  else { Decoder<bf16> dec(*d, nullptr, nullptr); dec.plan(a); }
  return a.off + 1024;
}

extern "C" int cqvad_decoder_forward(const cqvad_decoder_desc* d, const void* const* weights, const float* tgt,
                                     const float* memory, const float* pos, const uint8_t* mask,
                                     const float* refpoints_unsigmoid, void* hs, void* cls_hs, float* refs,
                                     float* pred_logits, float* pred_boxes, float* pred_logits_b, void* workspace,
                                     size_t ws_bytes, void* stream) {
  CQ_TRY(check_desc(d));
  CQ_CHECK_ARG(weights && tgt && memory && pos && refpoints_unsigmoid && hs && refs && workspace, "decoder: null pointer");
  CQ_CHECK_ARG(cls_hs || (d->flags & CQVAD_DEC_SKIP_CLS_HS), "decoder: cls_hs is NULL without CQVAD_DEC_SKIP_CLS_HS");
  const int nw = cqvad_decoder_num_weights(d->layers);
  for (int i = 0; i < nw; ++i) {
    if (weights[i] == nullptr) {
      std::string nm; weight_slot(i, d->layers, &nm, nullptr);
      const bool optional = (nm.find("ca_qpos_proj") != std::string::npos && nm.rfind("layers.0.", 0) != 0) ||
                            nm.find(".__") != std::string::npos;   // synthesised (fused) entries are optional
      CQ_CHECK_ARG(optional, "decoder: weight '%s' is NULL", nm.c_str());
    }
  }
  reset_launch_count();
  if (d->dtype == CQVAD_F32)
    return run_decoder<float>(d, weights, tgt, memory, pos, mask, refpoints_unsigmoid, hs, cls_hs, refs, pred_logits,
                              pred_boxes, pred_logits_b, workspace, ws_bytes, as_stream(stream));
  return run_decoder<bf16>(d, weights, tgt, memory, pos, mask, refpoints_unsigmoid, hs, cls_hs, refs, pred_logits,
                           pred_boxes, pred_logits_b, workspace, ws_bytes, as_stream(stream));
}

// ---- building blocks ---------------------------------------------------------------------------------------------
extern "C" int cqvad_layernorm(int dtype, const void* x, const void* res, const float* gamma, const float* beta, float eps,
                               void* out, int out_f32, long rows, int C, void* stream) {
  CQ_CHECK_ARG(x && gamma && beta && out && rows >= 0, "layernorm: bad argument");
  CQ_CHECK_SHAPE(C == kC, "layernorm: C must be 256 (got %d)", C);
  if (dtype == CQVAD_F32) return layernorm_rows<float>((const float*)x, (const float*)res, gamma, beta, eps, out, out_f32 != 0, rows, as_stream(stream));
  if (dtype == CQVAD_BF16) return layernorm_rows<bf16>((const bf16*)x, (const bf16*)res, gamma, beta, eps, out, out_f32 != 0, rows, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "layernorm: unknown dtype %d", dtype);
}

extern "C" int cqvad_linear(int dtype, const void* A, const void* W, const float* bias, const void* res, void* C, long M,
                            int N, int K, int act, void* stream) {
  CQ_CHECK_ARG(A && W && C && M >= 0 && N >= 1 && K >= 1, "linear: bad argument");
  CQ_CHECK_ARG(act >= CQVAD_ACT_NONE && act <= CQVAD_ACT_GELU, "linear: unknown activation %d", act);
  Epilogue e;
  e.bias = bias; e.act = act; e.res = res; e.ldr = N;
  if (dtype == CQVAD_F32) return gemm<float>((const float*)A, K, (const float*)W, (float*)C, N, M, N, K, e, nullptr, as_stream(stream));
  if (dtype == CQVAD_BF16) return gemm<bf16>((const bf16*)A, K, (const bf16*)W, (bf16*)C, N, M, N, K, e, nullptr, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "linear: unknown dtype %d", dtype);
}

extern "C" int cqvad_linear_gelu_train(int dtype, const void* A, const void* W, const float* bias, void* act_out,
                                       void* dact_out, long M, int N, int K, void* stream) {
  CQ_CHECK_ARG(A && W && act_out && dact_out && M >= 0 && N >= 1 && K >= 1, "linear_gelu_train: bad argument");
  Epilogue e;
  e.bias = bias; e.dual_gelu = true; e.c2 = dact_out;
  if (dtype == CQVAD_F32) return gemm<float>((const float*)A, K, (const float*)W, (float*)act_out, N, M, N, K, e, nullptr, as_stream(stream));
  if (dtype == CQVAD_BF16) return gemm<bf16>((const bf16*)A, K, (const bf16*)W, (bf16*)act_out, N, M, N, K, e, nullptr, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "linear_gelu_train: unknown dtype %d", dtype);
}

extern "C" int cqvad_linear_dgrad_act(int dtype, const void* dY, const void* Wt, const void* aux, int mode, void* dX, long M,
                                      int N, int K, void* stream) {
  CQ_CHECK_ARG(dY && Wt && aux && dX && M >= 0 && N >= 1 && K >= 1, "linear_dgrad_act: bad argument");
  CQ_CHECK_ARG(mode == 1 || mode == 3, "linear_dgrad_act: mode must be 1 (ReLU mask) or 3 (stored derivative)");
  Epilogue e;
  e.mul_aux = aux; e.mul_mode = mode;
  if (dtype == CQVAD_F32) return gemm<float>((const float*)dY, K, (const float*)Wt, (float*)dX, N, M, N, K, e, nullptr, as_stream(stream));
  if (dtype == CQVAD_BF16) return gemm<bf16>((const bf16*)dY, K, (const bf16*)Wt, (bf16*)dX, N, M, N, K, e, nullptr, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "linear_dgrad_act: unknown dtype %d", dtype);
}

template <typename T>
static int mlp_t(const T* X, const T* W1, const float* b1, const T* W2, const float* b2, int act, const T* res,
                 const float* g, const float* b, float eps, T* Y, T* hid, long M, int F, cudaStream_t st) {
  if (DT<T>::id == CQVAD_BF16 && !force_simt()) {
    int r = mlp_tc((const bf16*)X, (const bf16*)W1, b1, (const bf16*)W2, b2, act, (const bf16*)res, g, b, eps, (bf16*)Y, M,
                   kC, F, 0, 0, st);
    if (r <= 0) return r;
  }
  CQ_CHECK_ARG(hid != nullptr, "mlp: hidden scratch buffer required on this path");
  Epilogue e1; e1.bias = b1; e1.act = act;
  CQ_TRY(gemm<T>(X, kC, W1, hid, F, M, F, kC, e1, nullptr, st));
  Epilogue e2; e2.bias = b2; e2.res = res; e2.ldr = kC; e2.ln_g = g; e2.ln_b = b; e2.ln_eps = eps;
  return gemm<T>(hid, F, W2, Y, kC, M, kC, F, e2, nullptr, st);
}

extern "C" int cqvad_mlp(int dtype, const void* X, const void* W1, const float* b1, const void* W2, const float* b2, int act,
                         const void* res, const float* ln_g, const float* ln_b, float ln_eps, void* Y, void* hidden, long M,
                         int F, void* stream) {
  CQ_CHECK_ARG(X && W1 && b1 && W2 && b2 && Y && M >= 0 && F >= 8, "mlp: bad argument");
  CQ_CHECK_ARG(act == CQVAD_ACT_RELU || act == CQVAD_ACT_GELU, "mlp: activation must be ReLU or GELU");
  CQ_CHECK_ARG((ln_g == nullptr) == (ln_b == nullptr), "mlp: ln_g and ln_b must both be given or both be NULL");
  if (M == 0) return 0;
  if (dtype == CQVAD_F32)
    return mlp_t<float>((const float*)X, (const float*)W1, b1, (const float*)W2, b2, act, (const float*)res, ln_g, ln_b, ln_eps, (float*)Y, (float*)hidden, M, F, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return mlp_t<bf16>((const bf16*)X, (const bf16*)W1, b1, (const bf16*)W2, b2, act, (const bf16*)res, ln_g, ln_b, ln_eps, (bf16*)Y, (bf16*)hidden, M, F, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "mlp: unknown dtype %d", dtype);
}

extern "C" size_t cqvad_convblock_workspace_bytes(int dtype, long n_img, int h, int w) {
  const size_t es = dtype == CQVAD_F32 ? 4 : 2;
  const size_t rp = (size_t)n_img * (h + 1) * w;
  return (rp * kC * 3 + rp * 4 * kC) * es + 8 * 1024;
}

template <typename T>
static int convblock_t(const T* x, T* y, const T* w1, const float* b1, const float* g, const float* b, const T* w2,
                       const float* b2, const T* w3, const float* b3, long n_img, int h, int w, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  const int S = h * w, Sp = (h + 1) * w;
  const long Rp = n_img * Sp;
  Arena a(ws, ws_bytes);
  T* xp = (T*)a.take((size_t)Rp * kC * sizeof(T));
  T* xn = (T*)a.take((size_t)Rp * kC * sizeof(T));
  T* yp = (T*)a.take((size_t)Rp * kC * sizeof(T));
  T* hid = (T*)a.take((size_t)Rp * 4 * kC * sizeof(T));
  if (a.overflow) return set_error(CQVAD_E_WORKSPACE, "convblock: workspace too small (%zu needed)", a.off);
  CQ_CUDA(cudaMemsetAsync(xp, 0, (size_t)Rp * kC * sizeof(T), st));
  CQ_TRY(pad_copy<T>(x, xp, n_img, S, Sp, true, st));
  Epilogue e;
  e.bias = b1; e.ln_g = g; e.ln_b = b; e.ln_eps = 1e-6f;
  ConvGeom cg; cg.h = h; cg.w = w;
  CQ_TRY(gemm<T>(xp, kC, w1, xn, kC, Rp, kC, 9 * kC, e, &cg, st));
  bool done = false;
  if (DT<T>::id == CQVAD_BF16 && !force_simt()) {
    int r = mlp_tc((const bf16*)xn, (const bf16*)w2, b2, (const bf16*)w3, b3, CQVAD_ACT_GELU, (const bf16*)xp, nullptr,
                   nullptr, 0.f, (bf16*)yp, Rp, kC, 4 * kC, Sp, S, st);
    if (r < 0) return r;
    done = (r == 0);
  }
  if (!done) {
    Epilogue e1; e1.bias = b2; e1.act = CQVAD_ACT_GELU;
    CQ_TRY(gemm<T>(xn, kC, w2, hid, 4 * kC, Rp, 4 * kC, kC, e1, nullptr, st));
    Epilogue e2; e2.bias = b3; e2.res = xp; e2.ldr = kC; e2.zero_period = Sp; e2.zero_valid = S;
    CQ_TRY(gemm<T>(hid, 4 * kC, w3, yp, kC, Rp, kC, 4 * kC, e2, nullptr, st));
  }
  return pad_copy<T>(yp, y, n_img, S, Sp, false, st);
}

extern "C" int cqvad_convblock_forward(int dtype, const void* x, void* y, const void* w1, const float* b1,
                                       const float* ln_g, const float* ln_b, const void* w2, const float* b2,
                                       const void* w3, const float* b3, long n_img, int h, int w, void* workspace,
                                       size_t ws_bytes, void* stream) {
  CQ_CHECK_ARG(x && y && w1 && b1 && ln_g && ln_b && w2 && b2 && w3 && b3 && workspace, "convblock: null pointer");
  CQ_CHECK_SHAPE(n_img >= 0 && h >= 1 && w >= 1 && w <= 128, "convblock: bad extents");
  if (dtype == CQVAD_F32)
    return convblock_t<float>((const float*)x, (float*)y, (const float*)w1, b1, ln_g, ln_b, (const float*)w2, b2, (const float*)w3, b3, n_img, h, w, workspace, ws_bytes, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return convblock_t<bf16>((const bf16*)x, (bf16*)y, (const bf16*)w1, b1, ln_g, ln_b, (const bf16*)w2, b2, (const bf16*)w3, b3, n_img, h, w, workspace, ws_bytes, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "convblock: unknown dtype %d", dtype);
}

extern "C" int cqvad_mha_core(int dtype, int mode, const void* q, const void* k, const void* v,
                              const uint8_t* key_padding_mask, void* o, int L, int S, int Nb, int H, int E, int Ev,
                              void* stream) {
  CQ_CHECK_ARG(q && k && v && o, "mha_core: null pointer");
  CQ_CHECK_ARG(mode == 0 || mode == 1, "mha_core: mode must be 0 or 1");
  const long q_ls = (long)Nb * E, q_bs = E, o_ls = (long)Nb * Ev, o_bs = Ev;
  long k_ls, k_bs, k_qs, v_ls, v_bs, v_qs;
  k_ls = (long)Nb * E; k_bs = E; k_qs = (long)S * Nb * E;
  v_ls = (long)Nb * Ev; v_bs = Ev; v_qs = (long)S * Nb * Ev;
  if (dtype == CQVAD_F32)
    return mha_core<float>(mode, (const float*)q, (const float*)k, (const float*)v, key_padding_mask, (float*)o, L, S, Nb, H, E, Ev, q_ls, q_bs, k_ls, k_bs, k_qs, v_ls, v_bs, v_qs, o_ls, o_bs, as_stream(stream));
  if (dtype == CQVAD_BF16)
    return mha_core<bf16>(mode, (const bf16*)q, (const bf16*)k, (const bf16*)v, key_padding_mask, (bf16*)o, L, S, Nb, H, E, Ev, q_ls, q_bs, k_ls, k_bs, k_qs, v_ls, v_bs, v_qs, o_ls, o_bs, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "mha_core: unknown dtype %d", dtype);
}

extern "C" int cqvad_posenc3d(const uint8_t* mask, float* pos, int B, int T, int H, int W, int num_pos_feats, void* stream) {
  CQ_CHECK_ARG(mask && pos && B >= 0 && T >= 1 && H >= 1 && W >= 1, "posenc3d: bad argument");
  CQ_CHECK_SHAPE(num_pos_feats % 8 == 0, "posenc3d: num_pos_feats must be a multiple of 8");
  if (B == 0) return 0;
  return posenc3d(mask, pos, B, T, H, W, num_pos_feats, as_stream(stream));
}

extern "C" int cqvad_sine_embed(const float* ref, float* out, long rows, void* stream) {
  CQ_CHECK_ARG(ref && out && rows >= 0, "sine_embed: bad argument");
  return sine_embed<float>(ref, out, rows, as_stream(stream));
}
