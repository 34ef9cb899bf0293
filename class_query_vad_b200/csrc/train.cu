// Training path of the decoder: TransformerDecoder.forward (models/detr/dab_transformer.py:722-852) with every
// intermediate the backward needs kept in the caller's workspace, and its backward pass (what torch autograd derives
// from the reference forward: SURVEY.md section 8c "Gradient oracle", section 8d "Config 2").
//
// Structure: the forward is written ONCE as a sequence of ops on `Ten` (activation + gradient buffer).  Each op enqueues
// its forward kernels (mode FWD) or records a backward closure on a tape (mode REPLAY); both modes take identical
// allocations from the workspace arena, so cqvad_decoder_backward re-derives every pointer by replaying the op sequence
// over the workspace the forward filled -- no host-side state survives between the two C-ABI calls.  Mode PLAN only sizes.
//
// Autograd semantics reproduced (dab_transformer.py:810,823): reference points are detached between layers (only layer 0
// back-propagates into refpoints_unsigmoid), the actor feature is detached on entry to the class branch, while q_memory,
// query_sine_embed and the class-query chain are not.  Dropout: desc.dropout_p > 0 applies nn.Dropout at the nine residual-branch /
// FFN-hidden sites of a layer pair (Philox masks, dropout.cu); the dropout on attention probabilities (attention.py:402) is not
// applied; dropout_p = 0 is the parity / eval semantics.
#include <stdlib.h>
#include <deque>
#include <functional>
#include <vector>
#include "common.cuh"
#include "kernels_mem.cuh"
#include "attention.cuh"
#include "bwd.cuh"
#include "dropout.cuh"
#include "prof.cuh"
#include "wtable.cuh"

namespace cqvad {

static thread_local void* const* g_layer_events = nullptr;
static thread_local int g_layer_events_n = 0;
static void* const* layer_events() { return g_layer_events; }
static int layer_events_count() { return g_layer_events_n; }

int cls_xattn_tc(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vt, long ldvt,
                 const float* bv, bf16* out, long N, int K, int S, int Sq, int Sp_rows, int BT, cudaStream_t st);
int cls_sattn_tc(const bf16* x, const bf16* xt, long ldxt, bf16* out, long N, int K, int K8, cudaStream_t st);
int cls_xattn_bwd_tc(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vx, const bf16* dO, bf16* dQin,
                     float beta_q, bf16* dcqp, float beta_q2, bf16* dkx, bf16* dvx, long N, int K, int S, int Sq, int Sp_rows, int BT,
                     cudaStream_t st);
int cls_sattn_bwd_tc(const bf16* x, const bf16* dO, bf16* dx, float beta, long N, int K, cudaStream_t st);
int transpose_tokens(const bf16* x, bf16* xt, long ldxt, long N, int K, int K8, cudaStream_t st);

namespace {

// tensor-core attention paths exist for bf16 only; the float instantiation never calls them
inline int tc_xattn_fwd(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* W, const bf16* qm, bf16* vt,
                        long ldvt, const float* bv, bf16* out, long N, long NSq, int K, int S, int Sq, int Sp, int BT, cudaStream_t st) {
  Epilogue e;   // V^T[256, N*Sq] = W_v . q_memory^T (bias added after the softmax-weighted sum: the weights sum to one)
  CQ_TRY(gemm<bf16>(W, kC, qm, vt, ldvt, kC, (int)NSq, kC, e, nullptr, st));
  return cls_xattn_tc(Qin, cqp, kx, pos0, vt, ldvt, bv, out, N, K, S, Sq, Sp, BT, st);
}
inline int tc_xattn_fwd(const float*, const float*, const float*, const float*, const float*, const float*, float*, long,
                        const float*, float*, long, long, int, int, int, int, int, cudaStream_t) { return 1; }
inline int tc_xattn_bwd(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vx, const bf16* dO, bf16* dQin,
                        float bq, bf16* dcqp, float bq2, bf16* dkx, bf16* dvx, long N, int K, int S, int Sq, int Sp, int BT,
                        cudaStream_t st) {
  return cls_xattn_bwd_tc(Qin, cqp, kx, pos0, vx, dO, dQin, bq, dcqp, bq2, dkx, dvx, N, K, S, Sq, Sp, BT, st);
}
inline int tc_xattn_bwd(const float*, const float*, const float*, const float*, const float*, const float*, float*, float, float*,
                        float, float*, float*, long, int, int, int, int, int, cudaStream_t) { return 1; }
inline int tc_sattn_fwd(const bf16* x, bf16* xt, long ldxt, bf16* out, long N, int K, int K8, cudaStream_t st) {
  CQ_TRY(transpose_tokens(x, xt, ldxt, N, K, K8, st));
  return cls_sattn_tc(x, xt, ldxt, out, N, K, K8, st);
}
inline int tc_sattn_fwd(const float*, float*, long, float*, long, int, int, cudaStream_t) { return 1; }
inline int tc_sattn_bwd(const bf16* x, const bf16* dO, bf16* dx, float beta, long N, int K, cudaStream_t st) {
  return cls_sattn_bwd_tc(x, dO, dx, beta, N, K, st);
}
inline int tc_sattn_bwd(const float*, const float*, float*, float, long, int, cudaStream_t) { return 1; }

enum Mode { PLAN = 0, FWD = 1, REPLAY = 2 };

template <typename T>
struct Ten {
  T* p = nullptr;   // activation
  T* g = nullptr;   // gradient (nullptr: no gradient flows into this tensor)
  long rows = 0;
  int cols = 0;
  bool hg = false;  // a gradient flows into this tensor (g is valid outside PLAN mode)
  bool gi = false;  // gradient buffer holds a value (first writer overwrites, later writers accumulate)
  // this tensor is act(.) of something: the gradient arriving here must be multiplied by act'(gref).  The single consumer's
  // data-gradient GEMM does it in its epilogue (gmasked = true); otherwise the producer's backward runs an elementwise pass.
  int gact = 0;             // 0 none, 1 ReLU (gref = this tensor), 2 GELU (gref = pre-activation), 3 stored derivative (gref = act')
  const T* gref = nullptr;
  bool gmasked = false;
  // a dropout followed this ReLU output in place: its backward is the ReLU mask (dropped units are zero) times gscale, folded into
  // the consumer's data-gradient epilogue (gfolded set there); otherwise the dropout's own backward pass runs
  float gscale = 1.f;
  bool gfolded = false;
  long n() const { return rows * cols; }
};

struct TrainIO {
  // forward
  const float *tgt, *memory, *pos, *ref_u; const uint8_t* mask;
  void *hs, *cls_hs; float* refs;
  // backward
  const void *g_hs, *g_cls; const float* g_refs;
  float* const* gw; float *g_memory, *g_tgt, *g_ref_u;
};

template <typename T>
struct Trainer {
  const cqvad_decoder_desc& d;
  const void* const* w;
  TrainIO io;
  cudaStream_t st;
  Arena& a;
  Mode mode;
  int BT, nq, h, wd, S, Sp, Sq, K, F, Lr;
  long N, NS, NSq, Rp, NK;
  std::deque<Ten<T>> tens;
  struct TapeEntry { std::function<int()> fn; int branch; int layer; bool class_sync = false; };
  struct Tape {
    std::vector<TapeEntry> v; int* cur; int* layer;
    void push_back(std::function<int()> f) { v.push_back(TapeEntry{std::move(f), *cur, *layer, false}); }
  } tape;
  int cur_layer = -1;   // layer whose ops are being recorded (-1: prologue / global ops)
  std::vector<T*> wt;   // transposed / flipped weight copies for the data-gradient GEMMs, by weight index
  std::vector<char> wt_alloc, wt_built;

  Trainer(const cqvad_decoder_desc& dd, const void* const* ww, const TrainIO& i, cudaStream_t s, Arena& ar, Mode m)
      : d(dd), w(ww), io(i), st(s), a(ar), mode(m) {
    BT = d.BT; nq = d.nq; h = d.h; wd = d.w; S = h * wd; Sp = (h + 1) * wd; K = d.K; F = d.F; Lr = d.layers;
    Sq = (S + 7) & ~7;
    N = (long)nq * BT; NS = N * S; NSq = N * Sq; Rp = N * Sp; NK = N * K;
    tape.cur = &cur_branch; tape.layer = &cur_layer;
    static const bool one = [] { const char* e = getenv("CQVAD_TRAIN_STREAMS"); return e && atoi(e) == 1; }();
    two_streams = !one && mode != PLAN && side_stream() != nullptr;
    streams[0] = st; streams[1] = two_streams ? side_stream() : st;
    const size_t nw = (size_t)Lr * (LOC_COUNT + CLS_COUNT) + GLOB_COUNT;
    wt.assign(nw, nullptr); wt_alloc.assign(nw, 0); wt_built.assign(nw, 0);
  }
  // CQVAD_TRAIN_SYNC=1: synchronise after every op and report the first failing one (debug aid)
  int dbg(const char* what, int idx = -1) {
    static const bool on = getenv("CQVAD_TRAIN_SYNC") != nullptr;
    if (!on || mode == PLAN) return 0;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) fprintf(stderr, "train[%s] after %s #%d (arena %zu): %s\n", mode == FWD ? "fwd" : "bwd", what, idx, a.off, cudaGetErrorString(e));
    if (e != cudaSuccess) return set_error(CQVAD_E_CUDA, "train[%s] after %s #%d (arena %zu): %s", mode == FWD ? "fwd" : "bwd", what, idx, a.off, cudaGetErrorString(e));
    return 0;
  }
  // Two-stream schedule.  Branch 0 = localisation chain (prologue, loc layer, hs output, box refinement: mostly small-row
  // latency-bound kernels), branch 1 = class branch (the large GEMM / conv / attention kernels).  Forward: the class branch
  // of layer l waits for the loc layer l only, so the loc chain runs ahead; backward: the class chain runs ahead and the loc
  // layer l waits for the class branch l.  Events fork from / join into the caller's stream; CQVAD_TRAIN_STREAMS=1 disables.
  cudaStream_t streams[2];
  int cur_branch = 0;
  bool two_streams = false;
  float* wg_scr[2] = {nullptr, nullptr};
  static cudaStream_t side_stream() { return aux_stream(1); }          // per device (common.cuh)
  static cudaEvent_t sync_event(int i) { return aux_event(40 + i); }
  // Third stream: the large weight-gradient GEMMs only feed the (atomically accumulated) parameter gradients, so they leave
  // the dgrad chain's stream as soon as their dY is final and fill the SMs that chain's kernel tails leave idle; joined at
  // the end of the backward.  CQVAD_TRAIN_WG_STREAM=0 disables.
  static cudaStream_t wg_stream() { return aux_stream(2); }
  bool wg_used = false;
  cudaStream_t wgrad_stream(long rows) {
    static const bool off = [] { const char* e = getenv("CQVAD_TRAIN_WG_STREAM"); return e && atoi(e) == 0; }();
    cudaEvent_t ev = aux_event(42);
    static const long min_rows = [] { const char* e = getenv("CQVAD_TRAIN_WG_MIN_ROWS"); return e ? atol(e) : 1L; }();   // small-row wgrads too: 124 launches of ~8 us leave the loc chain
    if (off || !two_streams || rows < min_rows || wg_stream() == nullptr || ev == nullptr || sizeof(T) != 2) return st;   // the profiler times each kernel on the stream it runs on
    if (cudaEventRecord(ev, st) != cudaSuccess || cudaStreamWaitEvent(wg_stream(), ev, 0) != cudaSuccess) return st;
    wg_used = true;
    return wg_stream();
  }
  // error path: whatever was already enqueued on the side / weight-gradient streams is ordered before anything the caller
  // enqueues next on its own stream (e.g. the release of the workspace); errors here are ignored, the first one is reported
  void join_after_error() {
    if (mode == PLAN) return;
    if (two_streams && sync_event(1) && cudaEventRecord(sync_event(1), streams[1]) == cudaSuccess)
      cudaStreamWaitEvent(st, sync_event(1), 0);
    if (wg_used && aux_event(43) && cudaEventRecord(aux_event(43), wg_stream()) == cudaSuccess)
      cudaStreamWaitEvent(st, aux_event(43), 0);
    wg_used = false;
  }
  int wgrad_join() {
    if (!wg_used) return 0;
    cudaEvent_t ev = aux_event(43);
    CQ_CHECK_ARG(ev != nullptr, "training: could not create the weight-gradient join event");
    CQ_CUDA(cudaEventRecord(ev, wg_stream()));
    CQ_CUDA(cudaStreamWaitEvent(streams[0], ev, 0));
    wg_used = false;
    return 0;
  }
  // Gradient-bucket signalling (cqvad_decoder_backward_layer_events): once every backward op of layer l is enqueued, the caller's
  // event l is recorded behind all three streams, so a communication stream can all-reduce that layer's parameter gradients
  // while the backward of the layers below still runs.
  int signal_layer(int l) {
    void* const* evs = layer_events();
    if (!evs || l >= layer_events_count() || !evs[l]) return 0;
    cudaStream_t sig = aux_stream(3);
    cudaEvent_t tmp[3] = {aux_event(44), aux_event(45), aux_event(46)};
    CQ_CHECK_ARG(sig && tmp[0] && tmp[1] && tmp[2], "training: could not create the gradient-bucket signalling stream / events");
    cudaStream_t srcs[3] = {streams[0], streams[1], wg_used ? wg_stream() : streams[0]};
    for (int i = 0; i < 3; ++i) {
      if (i > 0 && srcs[i] == srcs[0]) continue;
      CQ_CUDA(cudaEventRecord(tmp[i], srcs[i]));
      CQ_CUDA(cudaStreamWaitEvent(sig, tmp[i], 0));
    }
    CQ_CUDA(cudaEventRecord((cudaEvent_t)evs[l], sig));
    return 0;
  }
  int link(int from, int to) {   // stream `to` waits for everything enqueued on stream `from` so far
    if (!two_streams) return 0;
    cudaEvent_t e = sync_event(from);
    CQ_CUDA(cudaEventRecord(e, streams[from]));
    CQ_CUDA(cudaStreamWaitEvent(streams[to], e, 0));
    return 0;
  }
  int set_branch(int b) {
    if (b == cur_branch) return 0;
    if (fwd() && b == 1) CQ_TRY(link(0, 1));   // forward: the class branch consumes what the loc chain has produced
    cur_branch = b;
    st = streams[two_streams ? b : 0];
    return 0;
  }
  bool fwd() const { return mode == FWD; }
  bool rec() const { return mode == REPLAY; }
  const T* Wm(int i) const { return (const T*)w[i]; }
  const float* Wf(int i) const { return (const float*)w[i]; }
  float* G(int i) const { return io.gw ? io.gw[i] : nullptr; }
  static bool big(long rows) { return rows >= 8192; }   // profiler classes: large-M GEMMs vs the small-row (launch-bound) ones
  int loc(int l, int s) const { return l * LOC_COUNT + s; }
  int cls(int l, int s) const { return Lr * LOC_COUNT + l * CLS_COUNT + s; }
  int glob(int s) const { return Lr * (LOC_COUNT + CLS_COUNT) + s; }

  T* take(long n) { return (T*)a.take((size_t)n * sizeof(T)); }
  float* takef(long n) { return (float*)a.take((size_t)n * sizeof(float)); }
  Ten<T>* mk(long rows, int cols, bool grad = true) {
    tens.emplace_back();
    Ten<T>* t = &tens.back();
    t->rows = rows; t->cols = cols;
    t->p = take(rows * cols);
    t->g = grad ? take(rows * cols) : nullptr;
    t->hg = grad;
    return t;
  }
  Ten<T>* detach(Ten<T>* x) {
    tens.emplace_back();
    Ten<T>* t = &tens.back();
    *t = *x; t->g = nullptr; t->hg = false; t->gi = false;
    return t;
  }
  static float beta(Ten<T>* t) { const float b = t->gi ? 1.f : 0.f; t->gi = true; return b; }

  // transposed weight for dX = dY . W  (built at backward time: the weights may have changed since the forward)
  struct WtJob { int widx, out, in; bool conv; };
  std::vector<WtJob> wt_jobs;     // every transposed / flipped copy this backward needs, built in one launch before the tape runs
  const T* WT(int widx, long elems, int out = 0, int in = 0, bool conv = false) {
    if (!wt_alloc[widx]) {
      wt[widx] = take(elems); wt_alloc[widx] = 1;
      if (rec()) wt_jobs.push_back(WtJob{widx, out, in, conv});
    }
    return wt[widx];
  }
  int build_all_wt() {            // REPLAY mode, before the tape: the weights may have changed since the forward
    std::vector<TransposeJob> jobs;
    for (const WtJob& j : wt_jobs) {
      if (!w[j.widx] || wt_built[j.widx]) continue;
      wt_built[j.widx] = 1;
      if (j.conv) CQ_TRY(conv_w_flip<T>(Wm(j.widx), wt[j.widx], st));
      else jobs.push_back(TransposeJob{w[j.widx], wt[j.widx], j.out, j.in});
    }
    ProfScope ps(P_T_ACT_BWD, st);
    return transpose_w_batch<T>(jobs.data(), (int)jobs.size(), st);
  }
  int build_wt(int widx, int out, int in, bool conv) {
    if (wt_built[widx]) return 0;   // shared weights (the ConvBlock, the global MLP heads) are transposed once per backward
    wt_built[widx] = 1;
    if (conv) return conv_w_flip<T>(Wm(widx), wt[widx], st);
    return transpose_w<T>(Wm(widx), wt[widx], out, in, st);
  }

  // ---- ops -------------------------------------------------------------------------------------------------------
  // Y = act(X.W^T + b) (+ res)   [rows, Nout];  zero_period/valid: rows written as zeros (y-padded layout)
  // Fusing gelu / act' into the GEMM epilogue (one thread per output row) measured SLOWER on B200 than a plain epilogue plus
  // a full-occupancy elementwise pass (245 us vs 65 + 65 us on the [94 080 x 1024] ConvBlock hidden): off by default.
  const bool fuse_act = getenv("CQVAD_TRAIN_FUSE_ACT") != nullptr;
  // Backward through an activation: the data-gradient GEMM of the single consumer multiplies by act' in its (lean, TMA-store)
  // epilogue -- a ReLU mask read from the activation itself, or gelu'(H) that the forward gelu pass stored next to gelu(H) --
  // so the 578 MB read-modify-write pass over the [94 080 x 1024] ConvBlock hidden gradient disappears.
  const bool fuse_act_bwd = getenv("CQVAD_TRAIN_NO_FUSE_ACT_BWD") == nullptr;
  // ... and the forward of a GELU layer writes gelu(H) and gelu'(H) from the GEMM epilogue itself (Epilogue::dual_gelu): the
  // pre-activation H never reaches HBM (768 -> 384 MB of traffic per ConvBlock hidden layer, one launch less)
  const bool dual_gelu = fuse_act_bwd && !fuse_act && getenv("CQVAD_TRAIN_NO_DUAL_GELU") == nullptr;
  Ten<T>* last_c2 = nullptr;   // second output of the last lin(..., c2_act)
  Ten<T>* lin(Ten<T>* X, int widx, int Nout, int act = CQVAD_ACT_NONE, Ten<T>* res = nullptr, int zp = 0, int zv = 0,
              int* rc = nullptr, int c2_act = CQVAD_ACT_NONE) {
    Ten<T>* Y = mk(X->rows, Nout);
    const int Kd = X->cols;
    if (act == CQVAD_ACT_RELU) { Y->gact = 1; Y->gref = Y->p; }
    Ten<T>* A2 = nullptr;
    if (c2_act != CQVAD_ACT_NONE) {   // A2 = act(Y) written by the same epilogue; shares Y's gradient buffer
      tens.emplace_back();
      A2 = &tens.back();
      A2->rows = Y->rows; A2->cols = Y->cols; A2->p = take(Y->n()); A2->g = Y->g; A2->hg = true;
      A2->gact = c2_act == CQVAD_ACT_GELU ? 2 : 1; A2->gref = Y->p;
      if (c2_act == CQVAD_ACT_GELU && !fuse_act && fuse_act_bwd) { A2->gact = 3; A2->gref = take(Y->n()); }   // gelu'(Y) stored by gelu_fwd
      last_c2 = A2;
    }
    const T* Wt = X->hg ? WT(widx, (long)Nout * Kd, Nout, Kd, false) : nullptr;
    const bool dual = A2 && A2->gact == 3 && dual_gelu && !res && zp == 0 && act == CQVAD_ACT_NONE;
    if (fwd()) {
      Epilogue e;
      e.bias = Wf(widx + 1); e.act = act; e.res = res ? res->p : nullptr; e.ldr = Nout; e.zero_period = zp; e.zero_valid = zv;
      if (A2 && fuse_act) { e.c2 = A2->p; e.c2_act = c2_act; }
      if (dual) { e.dual_gelu = true; e.c2 = const_cast<T*>(A2->gref); }
      int r;
      {
        const double fl = 2.0 * X->rows * Nout * Kd, by = sizeof(T) * ((double)X->rows * (Kd + (dual ? 2.0 : 1.0) * Nout + (res ? Nout : 0)) + (double)Nout * Kd);
        ProfScope ps(dual ? P_T_FWD_GEMM_GELU : big(X->rows) ? P_T_FWD_GEMM : P_T_FWD_GEMM_SMALL, st, fl, by);
        r = gemm<T>(X->p, Kd, Wm(widx), dual ? A2->p : Y->p, Nout, X->rows, Nout, Kd, e, nullptr, st);
      }
      if (r == 0 && A2 && !fuse_act && !dual) { ProfScope ps(P_T_FWD_OTHER, st); r = gelu_fwd<T>(Y->p, A2->p, A2->gact == 3 ? const_cast<T*>(A2->gref) : nullptr, Y->n(), st); }
      if (r == 0) r = dbg("lin", widx);
      if (r != 0 && rc && *rc == 0) *rc = r;
    }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (A2) {   // the gradient arrived through the activated copy
          if (!A2->gi) return 0;
          Y->gi = true;
          if (!A2->gmasked) {
            if (dual) return set_error(CQVAD_E_INVALID_ARG, "backward: a dual-GELU activation needs a single GEMM consumer");
            ProfScope ps(P_T_ACT_BWD, st);
            CQ_TRY(act_bwd<T>(Y->g, Y->p, c2_act, Y->n(), st));
          }
        }
        if (!Y->gi) return 0;
        {
          ProfScope ps(P_T_ACT_BWD, st);
          if (act == CQVAD_ACT_RELU && !Y->gmasked) CQ_TRY(act_bwd<T>(Y->g, Y->p, CQVAD_ACT_RELU, Y->n(), st));
          if (res && res->hg) CQ_TRY(axpby<T>(res->g, Y->g, beta(res), Y->n(), st));
        }
        if (X->hg) {
          CQ_TRY(build_wt(widx, Nout, Kd, false));
          Epilogue e;
          const float b = beta(X);
          if (b != 0.f) { e.res = X->g; e.ldr = Kd; }
          const bool actmul = X->gact && (fuse_act || fuse_act_bwd) && b == 0.f;
          if (actmul) {   // dX = (dY . W) * act'(.) in the epilogue
            e.mul_aux = X->gref; e.mul_mode = X->gact; X->gmasked = true;
            if (X->gact == 1 && X->gscale != 1.f) { e.mul_scale = X->gscale; X->gfolded = true; }
          }
          const double fl = 2.0 * X->rows * Nout * Kd, by = sizeof(T) * ((double)X->rows * (Nout + Kd * (1.0 + (actmul ? 1 : 0) + (b != 0.f ? 1 : 0))) + (double)Nout * Kd);
          ProfScope ps(!big(X->rows) ? P_T_DGRAD_SMALL : (actmul && X->gact == 3) ? P_T_DGRAD_ACT : P_T_DGRAD, st, fl, by);
          CQ_TRY(gemm<T>(Y->g, Nout, Wt, X->g, Kd, X->rows, Kd, Nout, e, nullptr, st));
        }
        if (G(widx) || G(widx + 1)) {
          cudaStream_t ws = wgrad_stream(X->rows);
          const double fl = 2.0 * X->rows * Nout * Kd, by = sizeof(T) * (double)X->rows * (Nout + Kd) + 4.0 * 2.0 * Nout * Kd;
          ProfScope ps(big(X->rows) ? P_T_WGRAD : P_T_WGRAD_SMALL, ws, fl, by);
          CQ_TRY(wgrad<T>(Y->g, Nout, X->p, Kd, G(widx), Kd, G(widx + 1), X->rows, Nout, Kd, nullptr, ws));
        }
        return 0;
      });
    }
    return Y;
  }
  // nn.Dropout(p) on X, in place (activation in the forward, its gradient in the backward; the Philox mask is regenerated from
  // (seed, site), dropout.cu).  A ReLU output keeps working as its own derivative mask: dropped units are zero.
  Ten<T>* drop(Ten<T>* X, int l, int k, int* rc) {
    const float p = d.dropout_p;
    if (!(p > 0.f)) return X;
    const uint64_t seed = ((uint64_t)d.seed_hi << 32) | d.seed_lo;
    const uint32_t site = 0x1000u + 16u * (uint32_t)l + (uint32_t)k;
    if (fwd()) {
      ProfScope ps(P_T_FWD_OTHER, st);
      int r = dropout_apply<T>(X->p, nullptr, X->p, X->n(), p, seed, site, st);
      if (r != 0 && rc && *rc == 0) *rc = r;
    }
    if (rec()) {
      if (X->gact == 1 && X->gref == X->p) X->gscale = dropout_keep_scale(p);
      tape.push_back([=]() -> int {
        if (!X->gi || X->gfolded) return 0;     // folded: the consumer's ReLU-mask epilogue already applied mask and scale
        ProfScope ps(P_T_ACT_BWD, st);
        return dropout_apply<T>(X->g, nullptr, X->g, X->n(), p, seed, site, st);
      });
    }
    return X;
  }
  // Y = LN(X (+ res))
  Ten<T>* ln(Ten<T>* X, Ten<T>* res, int lnidx, float eps, int* rc) {
    Ten<T>* Y = mk(X->rows, kC);
    if (fwd()) {
      int r;
      { ProfScope ps(P_T_FWD_OTHER, st); r = layernorm_rows<T>(X->p, res ? res->p : nullptr, Wf(lnidx), Wf(lnidx + 1), eps, Y->p, false, X->rows, st); }
      if (r == 0) r = dbg("ln", lnidx);
      if (r != 0 && *rc == 0) *rc = r;
    }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (!Y->gi) return 0;
        const float bx = X->hg ? beta(X) : 0.f;
        const bool rg = res && res->hg;
        const float br = rg ? beta(res) : 0.f;
        ProfScope ps(P_T_LN_BWD, st);
        if (X->hg)
          return ln_bwd<T>(X->p, res ? res->p : nullptr, Wf(lnidx), eps, Y->g, false, 0, 0, 0, X->g, bx, rg ? res->g : nullptr,
                           br, G(lnidx), G(lnidx + 1), X->rows, st);
        // X carries no gradient (cannot happen in this graph)
        return set_error(CQVAD_E_INVALID_ARG, "ln backward: input without gradient");
      });
    }
    return Y;
  }
  // A = gelu(Hpre); the two tensors share one gradient buffer (dHpre = dA * gelu'(Hpre) in place)
  Ten<T>* gelu(Ten<T>* Hpre, int* rc) {
    tens.emplace_back();
    Ten<T>* A = &tens.back();
    A->rows = Hpre->rows; A->cols = Hpre->cols;
    A->p = take(A->n());
    A->g = Hpre->g; A->hg = Hpre->hg;
    if (fwd()) { ProfScope ps(P_T_FWD_OTHER, st); int r = gelu_fwd<T>(Hpre->p, A->p, nullptr, A->n(), st); if (r != 0) *rc = r; }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (!A->gi) return 0;
        Hpre->gi = true;
        ProfScope ps(P_T_ACT_BWD, st);
        return act_bwd<T>(A->g, Hpre->p, CQVAD_ACT_GELU, A->n(), st);
      });
    }
    return A;
  }
  // algorithmic work of one 3x3 conv over the N*S valid positions: 2*N*S*256*2304 FLOPs; X + Z + W bytes
  double conv_flops() const { return 2.0 * (double)NS * kC * 9 * kC; }
  double conv_bytes() const { return sizeof(T) * (2.0 * (double)Rp * kC + 9.0 * kC * kC); }
  // Z = conv3x3(X) + b on the y-padded layout
  Ten<T>* conv(Ten<T>* X, int widx, int* rc) {
    Ten<T>* Z = mk(X->rows, kC);
    const T* Wd = WT(widx, 256L * 9 * 256, 0, 0, true);
    ConvGeom cg; cg.h = h; cg.w = wd;
    if (fwd()) {
      Epilogue e;
      e.bias = Wf(widx + 1);
      int r;
      { ProfScope ps(P_T_CONV_FWD, st, conv_flops(), conv_bytes()); r = gemm<T>(X->p, kC, Wm(widx), Z->p, kC, X->rows, kC, 9 * kC, e, &cg, st); }
      if (r == 0) r = dbg("conv", widx);
      if (r != 0 && *rc == 0) *rc = r;
    }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (!Z->gi) return 0;
        if (X->hg) {
          CQ_TRY(build_wt(widx, 0, 0, true));
          ProfScope ps(P_T_CONV_DGRAD, st, conv_flops(), conv_bytes());
          Epilogue e;
          e.zero_period = Sp; e.zero_valid = S;
          const float b = beta(X);
          if (b != 0.f) { e.res = X->g; e.ldr = kC; }
          CQ_TRY(gemm<T>(Z->g, kC, Wd, X->g, kC, X->rows, kC, 9 * kC, e, &cg, st));
        }
        cudaStream_t ws = wgrad_stream(X->rows);
        ProfScope ps(P_T_CONV_WGRAD, ws, conv_flops(), conv_bytes());
        return wgrad<T>(Z->g, kC, X->p, kC, G(widx), 9 * kC, G(widx + 1), X->rows, kC, kC, &cg, ws);
      });
    }
    return Z;
  }
  // projection-free MHA, mode A.  q2/k2: second source for heads >= H/2 (class cross-attention); k2 gets no gradient.
  Ten<T>* mha(Ten<T>* q, Ten<T>* q2, Ten<T>* k, const T* k2, Ten<T>* v, long orows, int L, int Skeys, int Nb, int hd, int vd,
              StdStrides ss, int* rc) {
    Ten<T>* O = mk(orows, kC);
    if (fwd()) {
      int r;
      { ProfScope ps(P_T_FWD_OTHER, st); r = mha_std<T>(q->p, q2 ? q2->p : nullptr, k->p, k2, v->p, nullptr, O->p, L, Skeys, Nb, kH, hd, vd, ss, st); }
      if (r == 0) r = dbg("mha", L);
      if (r != 0 && *rc == 0) *rc = r;
    }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (!O->gi) return 0;
        // aliasing (q == k == v for the self-attentions): the kernel writes dv, then dk, then dq of one (batch, head) from
        // the same block, so later writers simply accumulate
        const float bv = beta(v), bk = beta(k), bq = beta(q);
        const float bq2 = q2 ? beta(q2) : 0.f;
        ProfScope ps(P_T_ATTN_BWD, st);
        return mha_std_bwd<T>(q->p, q2 ? q2->p : nullptr, k->p, k2, v->p, nullptr, O->g, q->g, bq, q2 ? q2->g : nullptr, bq2, k->g,
                              bk, v->g, bv, L, Skeys, Nb, kH, hd, vd, ss, st);
      });
    }
    return O;
  }

  int run();
};

template <typename T>
int Trainer<T>::run() {
  int rc = 0;
  const bool of32 = d.out_f32 != 0;
  const size_t osz = of32 ? sizeof(float) : sizeof(T);
  // ---- inputs ------------------------------------------------------------------------------------------------------
  Ten<T>* memc = mk(4L * S * BT, kC, false);
  Ten<T>* pos0 = mk((long)S * BT, kC, false);
  Ten<T>* out = mk(N, kC);
  Ten<T>* out_in = out;
  float* dmem32 = io.g_memory;                       // fp32 gradient of memory accumulates directly in the caller's buffer
  const int K8 = (K + 7) & ~7;
  const long ldvt = NSq + 64, ldxt = N * K8 + 64;
  T* vt = take(kC * ldvt);          // V^T of the forward cross-attention kernel (transient, shared by the layers)
  T* xt = take(kC * ldxt);          // transposed class tokens of the forward self-attention kernel
  const bool use_tc = DT<T>::id == CQVAD_BF16 && !force_simt() && K <= 128 && ((S + 15) & ~15) <= 256;
  if (fwd() && use_tc) CQ_CUDA(cudaMemsetAsync(xt, 0, (size_t)kC * ldxt * sizeof(T), st));
  wg_scr[0] = takef((long)(wgrad_scratch_bytes() / sizeof(float)));
  wg_scr[1] = takef((long)(wgrad_scratch_bytes() / sizeof(float)));
  float* dkp32 = takef((long)S * BT * kC);           // fp32 staging of d(ca_kpos_proj(pos)) (summed over the nq actors)
  std::vector<float*> rl(Lr + 1), drl(Lr + 1);
  for (int l = 0; l <= Lr; ++l) { rl[l] = takef(N * 4); drl[l] = takef(N * 4); }
  // class queries as an activation (its gradient is folded into the weight gradient at the end)
  tens.emplace_back();
  Ten<T>* cq = &tens.back();
  cq->rows = K; cq->cols = kC; cq->p = w ? const_cast<T*>(Wm(glob(G_CQ))) : nullptr; cq->g = take((long)K * kC); cq->hg = true;
  if (fwd()) {
    CQ_TRY(convert_f32<T>(io.memory, memc->p, memc->n(), st));
    CQ_TRY(convert_f32<T>(io.pos, pos0->p, pos0->n(), st));
    CQ_TRY(convert_f32<T>(io.tgt, out->p, out->n(), st));
    CQ_TRY(sigmoid4(io.ref_u, rl[0], io.refs, N, nq, BT, st));
  }
  if (rec()) {
    tape.push_back([=]() -> int {   // runs LAST in the backward
      if (io.g_ref_u)
        CQ_TRY(sigmoid4_bwd(rl[0], drl[0], io.g_refs, io.g_ref_u, N, nq, BT, st));
      if (io.g_tgt) {
        if (out_in->gi) CQ_TRY(acc_to_f32<T>(io.g_tgt, out_in->g, out_in->n(), st));
      }
      if (G(glob(G_CQ)) && cq->gi) CQ_TRY(acc_to_f32<T>(G(glob(G_CQ)), cq->g, cq->n(), st));
      return 0;
    });
  }

  Ten<T>* Qprev = nullptr;
  for (int l = 0; l < Lr; ++l) {
    const bool first = (l == 0);
    cur_layer = l;
    float* r_cur = rl[l];
    float* dr_cur = first ? drl[0] : nullptr;   // reference points are detached between layers (:823)
    // ---- prologue :742-763 ----
    Ten<T>* e512 = mk(N, 512, first);
    if (fwd()) CQ_TRY(sine_embed<T>(r_cur, e512->p, N, st));
    if (rec() && first) {
      tape.push_back([=]() -> int {
        if (!e512->gi) return 0;
        return sine_embed_bwd<T>(r_cur, e512->g, dr_cur, N, st);
      });
    }
    Ten<T>* qpos = lin(lin(e512, glob(G_RPH0), kC, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), glob(G_RPH1), kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* pscale = nullptr;
    if (!first) pscale = lin(lin(out, glob(G_QS0), kC, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), glob(G_QS1), kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* rah = lin(out, glob(G_RAH0), kC, CQVAD_ACT_RELU, nullptr, 0, 0, &rc);
    Ten<T>* qse = mk(N, kC);
    if (fwd())
      CQ_TRY(qse_modulate<T>(r_cur, pscale ? pscale->p : nullptr, rah->p, Wf(glob(G_RAH1)), Wf(glob(G_RAH1) + 1), qse->p, N, st));
    if (rec()) {
      const int wi = glob(G_RAH1);
      tape.push_back([=]() -> int {
        if (!qse->gi) return 0;
        const float bs = pscale ? beta(pscale) : 0.f, bh = beta(rah);
        return qse_bwd<T>(r_cur, pscale ? pscale->p : nullptr, rah->p, Wf(wi), Wf(wi + 1), qse->g, pscale ? pscale->g : nullptr,
                          bs, rah->g, bh, G(wi), G(wi + 1), dr_cur, N, st);
      });
    }
    // ---- localisation layer: self-attention over the nq actors of a frame :921-938 ----
    Ten<T>* saq = lin(qpos, loc(l, SA_QP), kC, 0, lin(out, loc(l, SA_QC), kC, 0, nullptr, 0, 0, &rc), 0, 0, &rc);
    Ten<T>* sak = lin(qpos, loc(l, SA_KP), kC, 0, lin(out, loc(l, SA_KC), kC, 0, nullptr, 0, 0, &rc), 0, 0, &rc);
    Ten<T>* sav = lin(out, loc(l, SA_V), kC, 0, nullptr, 0, 0, &rc);
    StdStrides s1{};
    s1.q_ls = s1.k_ls = s1.v_ls = s1.o_ls = (long)BT * kC;
    s1.q_bs = s1.k_bs = s1.v_bs = s1.o_bs = kC;
    Ten<T>* sao = mha(saq, nullptr, sak, nullptr, sav, N, nq, nq, BT, 32, 32, s1, &rc);
    Ten<T>* out1 = ln(drop(lin(sao, loc(l, SA_O), kC, 0, nullptr, 0, 0, &rc), l, 0, &rc), out, loc(l, NORM1), 1e-5f, &rc);   // dropout1 :937
    // ---- level-weighted query-specific memory :943-946 ----
    float* lvlw = takef(N * 4);
    float* dlvlw = takef(N * 4);
    Ten<T>* qm = mk(NSq, kC);
    if (fwd()) {
      CQ_TRY(linear_smalln<T>(out1->p, Wf(loc(l, LVLW)), Wf(loc(l, LVLW) + 1), lvlw, N, 4, true, st));
      if (Sq != S) CQ_CUDA(cudaMemsetAsync(qm->p, 0, (size_t)NSq * kC * sizeof(T), st));
      CQ_TRY(lvlmix_ln<T>(memc->p, lvlw, Wf(loc(l, NORMU)), Wf(loc(l, NORMU) + 1), qm->p, N, S, Sq, BT, st));
    }
    if (rec()) {
      const int wn = loc(l, NORMU), wl = loc(l, LVLW);
      tape.push_back([=]() -> int {
        if (!qm->gi) return 0;
        CQ_CUDA(cudaMemsetAsync(dlvlw, 0, (size_t)N * 4 * sizeof(float), st));
        CQ_TRY(lvlmix_ln_bwd<T>(memc->p, lvlw, Wf(wn), qm->g, dmem32, dlvlw, G(wn), G(wn + 1), N, nq, S, Sq, BT, st));
        return lvlw_bwd<T>(out1->p, Wf(wl), lvlw, dlvlw, out1->g, beta(out1), G(wl), G(wl + 1), N, st);
      });
    }
    // ---- cross-attention with per-actor keys :951-988 ----
    CQ_CHECK_ARG(mode == PLAN || w[loc(l, CA_KV)] != nullptr, "training path needs the stacked layers.%d.__ca_kv weights", l);
    Ten<T>* kv = lin(qm, loc(l, CA_KV), 2 * kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* kp = lin(pos0, loc(l, CA_KP), kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* qc = lin(out1, loc(l, CA_QC), kC, 0, nullptr, 0, 0, &rc);
    if (first) qc = lin(qpos, loc(l, CA_QP), kC, 0, qc, 0, 0, &rc);
    Ten<T>* qs = lin(qse, loc(l, CA_QS), kC, 0, nullptr, 0, 0, &rc);
    // Backward schedule: everything recorded AFTER this op in the loc layer (cross-attention core, its output projection, the FFN,
    // the two LayerNorms) consumes only gradients produced on the loc stream, so it may run while the class branch of this layer is
    // still in flight; this op's backward is the first to touch a buffer the class branch writes (qse->g, then qm->g).
    if (rec() && !tape.v.empty()) tape.v.back().class_sync = true;
    Ten<T>* cao = mk(N, kC);
    if (fwd()) {
      ProfScope ps(P_T_FWD_OTHER, st);
      CQ_TRY(dec_qsk_attn<T>(qc->p, qs->p, kv->p, kv->p + kC, 2 * kC, kp->p, io.mask, cao->p, N, S, Sq, BT, first, st));
    }
    if (rec()) {
      tape.push_back([=]() -> int {
        if (!cao->gi) return 0;
        if (Sq != S) CQ_CUDA(cudaMemsetAsync(kv->g, 0, (size_t)kv->n() * sizeof(T), st));
        CQ_CUDA(cudaMemsetAsync(dkp32, 0, (size_t)S * BT * kC * sizeof(float), st));
        const float bqc = beta(qc), bqs = beta(qs);
        ProfScope ps(P_T_ATTN_BWD, st);
        CQ_TRY(dec_qsk_bwd<T>(qc->p, qs->p, kv->p, kv->p + kC, 2 * kC, kp->p, io.mask, cao->g, qc->g, bqc, qs->g, bqs, kv->g,
                              kv->g + kC, dkp32, N, S, Sq, BT, first, st));
        kv->gi = true;
        CQ_TRY(f32_to_t<T>(dkp32, kp->g, beta(kp), kp->n(), st));
        return 0;
      });
    }
    Ten<T>* actor = ln(drop(lin(cao, loc(l, CA_O), kC, 0, nullptr, 0, 0, &rc), l, 1, &rc), out1, loc(l, NORM2), 1e-5f, &rc);   // tgt_temp :991-993 (dropout2)
    Ten<T>* ffn = drop(lin(drop(lin(actor, loc(l, LIN1), F, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), l, 2, &rc), loc(l, LIN2), kC, 0, nullptr, 0, 0, &rc), l, 3, &rc);   // dropout, dropout3 :995-996
    Ten<T>* out2 = ln(ffn, actor, loc(l, NORM3), 1e-5f, &rc);

    if (rc != 0) return rc;
    CQ_TRY(dbg("loc layer", l));
    // ---- class-query layer :1040-1079 (actor feature detached, :810) ----
    CQ_TRY(set_branch(1));
    Ten<T>* actor_d = detach(actor);
    Ten<T>* cffn = drop(lin(drop(lin(actor_d, cls(l, C_L1), F, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), l, 4, &rc), cls(l, C_L2), kC, 0, nullptr, 0, 0, &rc), l, 5, &rc);   // dropout1, dropout2 :1043-1044
    Ten<T>* acls = ln(cffn, actor_d, cls(l, C_NORM), 1e-5f, &rc);
    Ten<T>* X = mk(Rp, kC);
    if (fwd()) {
      CQ_CUDA(cudaMemsetAsync(X->p, 0, (size_t)Rp * kC * sizeof(T), st));
      CQ_TRY(add_ln_pad<T>(acls->p, qm->p, Wf(cls(l, C_CONVNORM)), Wf(cls(l, C_CONVNORM) + 1), X->p, N, S, Sq, Sp, st));
    }
    if (rec()) {
      const int wi = cls(l, C_CONVNORM);
      Ten<T>* XA = X;
      tape.push_back([=]() -> int {
        if (!XA->gi) return 0;
        const float bq = beta(qm), ba = beta(acls);
        return add_ln_pad_bwd<T>(acls->p, qm->p, Wf(wi), XA->g, qm->g, bq, acls->g, ba, G(wi), G(wi + 1), N, S, Sq, Sp, st);
      });
    }
    for (int blk = 0; blk < 3; ++blk) {   // the same ConvBlock three times (:1017-1018, :1055-1056)
      Ten<T>* Z = conv(X, cls(l, C_CONV1), &rc);
      Ten<T>* Xn = ln(Z, nullptr, cls(l, C_CBNORM), 1e-6f, &rc);
      lin(Xn, cls(l, C_CONV2), 4 * kC, 0, nullptr, 0, 0, &rc, CQVAD_ACT_GELU);   // pre-activation kept + gelu in one epilogue
      Ten<T>* A = last_c2;
      X = lin(A, cls(l, C_CONV3), kC, 0, X, Sp, S, &rc);
    }
    Ten<T>* X3 = X;
    // class-query self-attention :1059-1065
    Ten<T>* Qin;
    StdStrides s2{};
    s2.q_ls = s2.k_ls = s2.v_ls = s2.o_ls = kC;
    if (first) {   // identical for every actor instance: computed once on K rows, then broadcast
      Ten<T>* cq1 = mha(cq, nullptr, cq, nullptr, cq, K, K, K, 1, 32, 32, s2, &rc);
      Ten<T>* cq2 = ln(drop(lin(cq1, cls(l, C_SA_O), kC, 0, nullptr, 0, 0, &rc), l, 6, &rc), cq, cls(l, C_NORM1), 1e-5f, &rc);   // dropout3 :1062 (one mask for all instances: the rows are shared)
      Qin = mk(NK, kC);
      if (fwd()) CQ_TRY(broadcast_rows<T>(cq2->p, Qin->p, NK, K, st));
      if (rec()) {
        Ten<T>* Qi = Qin;
        tape.push_back([=]() -> int {
          if (!Qi->gi) return 0;
          return broadcast_rows_bwd<T>(Qi->g, cq2->g, beta(cq2), NK, K, st);
        });
      }
    } else {
      s2.q_bs = s2.k_bs = s2.v_bs = s2.o_bs = (long)K * kC;
      Ten<T>* saoc;
      if (use_tc) {
        saoc = mk(NK, kC);
        Ten<T>* Qp = Qprev;
        if (fwd()) {
          ProfScope ps(P_T_FWD_OTHER, st);
          int r = tc_sattn_fwd(Qp->p, xt, ldxt, saoc->p, N, K, K8, st);
          if (r == 1) r = set_error(CQVAD_E_UNSUPPORTED_SHAPE, "class self-attention: shape rejected by the tensor-core kernel");
          if (r != 0) return r;
        }
        if (rec()) {
          tape.push_back([=]() -> int {
            if (!saoc->gi) return 0;
            ProfScope ps(P_T_ATTN_BWD, st);
            int r = tc_sattn_bwd(Qp->p, saoc->g, Qp->g, beta(Qp), N, K, st);
            if (r == 1) r = set_error(CQVAD_E_UNSUPPORTED_SHAPE, "class self-attention backward: shape rejected");
            return r;
          });
        }
      } else {
        saoc = mha(Qprev, nullptr, Qprev, nullptr, Qprev, NK, K, K, (int)N, 32, 32, s2, &rc);
      }
      Qin = ln(drop(lin(saoc, cls(l, C_SA_O), kC, 0, nullptr, 0, 0, &rc), l, 6, &rc), Qprev, cls(l, C_NORM1), 1e-5f, &rc);   // dropout3 :1062
    }
    // class cross-attention :1067-1071
    Ten<T>* kx = lin(X3, cls(l, C_KPROJ), kC, 0, nullptr, 0, 0, &rc);     // on the padded layout
    Ten<T>* vx = lin(qm, cls(l, C_VPROJ), kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* cqp = lin(qse, cls(l, C_QPS), kC, 0, nullptr, 0, 0, &rc);
    StdStrides s3{};
    s3.q_ls = kC; s3.q_bs = (long)K * kC;
    s3.q2_ls = 0; s3.q2_bs = kC;
    s3.k_ls = kC; s3.k_bs = (long)Sp * kC;
    s3.k2_ls = (long)BT * kC; s3.k2_bs = kC; s3.k2_bmod = BT;
    s3.v_ls = kC; s3.v_bs = (long)Sq * kC;
    s3.o_ls = kC; s3.o_bs = (long)K * kC;
    Ten<T>* caoc;
    {
      // gradient buffers of kx / vx must be zero on the rows the attention never touches (y-pad separator rows of kx, pitch
      // padding of vx): zero them right before the attention backward writes the valid rows
      Ten<T>* O = mk(NK, kC);
      const int wv = cls(l, C_VPROJ);
      if (fwd()) {
        ProfScope ps(P_T_FWD_OTHER, st);
        int r = use_tc ? tc_xattn_fwd(Qin->p, cqp->p, kx->p, pos0->p, Wm(wv), qm->p, vt, ldvt, Wf(wv + 1), O->p, N, NSq, K, S, Sq, Sp,
                                      BT, st)
                       : mha_std<T>(Qin->p, cqp->p, kx->p, pos0->p, vx->p, nullptr, O->p, K, S, (int)N, kH, 64, 32, s3, st);
        if (r == 1) r = set_error(CQVAD_E_UNSUPPORTED_SHAPE, "class cross-attention: shape rejected by the tensor-core kernel");
        if (r != 0) return r;
      }
      if (rec()) {
        tape.push_back([=]() -> int {
          if (!O->gi) return 0;
          CQ_CUDA(cudaMemsetAsync(kx->g, 0, (size_t)kx->n() * sizeof(T), st));
          if (Sq != S) CQ_CUDA(cudaMemsetAsync(vx->g, 0, (size_t)vx->n() * sizeof(T), st));
          kx->gi = true; vx->gi = true;
          const float bq = beta(Qin), bq2 = beta(cqp);
          ProfScope ps(P_T_ATTN_BWD, st);
          // kx/vx: beta 0 on valid rows is equivalent to accumulate-after-memset; use overwrite
          if (use_tc) {
            int r = tc_xattn_bwd(Qin->p, cqp->p, kx->p, pos0->p, vx->p, O->g, Qin->g, bq, cqp->g, bq2, kx->g, vx->g, N, K, S, Sq, Sp, BT,
                                 st);
            if (r == 1) r = set_error(CQVAD_E_UNSUPPORTED_SHAPE, "class cross-attention backward: shape rejected");
            return r;
          }
          return mha_std_bwd<T>(Qin->p, cqp->p, kx->p, pos0->p, vx->p, nullptr, O->g, Qin->g, bq, cqp->g, bq2, kx->g, 0.f, vx->g,
                                0.f, K, S, (int)N, kH, 64, 32, s3, st);
        });
      }
      caoc = O;
    }
    Ten<T>* cls0 = lin(caoc, cls(l, C_CA_O), kC, 0, nullptr, 0, 0, &rc);
    Ten<T>* cf = drop(lin(drop(lin(cls0, cls(l, C_L1_), F, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), l, 7, &rc), cls(l, C_L2_), kC, 0, nullptr, 0, 0, &rc), l, 8, &rc);   // dropout1_, dropout2_ :1076-1077
    Ten<T>* cls_out = ln(cf, cls0, cls(l, C_NORM_), 1e-5f, &rc);
    Qprev = cls_out;

    // ---- outputs of this layer :826-827 ----
    if (fwd())   // still on the class-branch stream
      CQ_TRY(layernorm_permute<T>(cls_out->p, Wf(glob(G_CLSNORM2)), Wf(glob(G_CLSNORM2) + 1), 1e-5f,
                                  (char*)io.cls_hs + (size_t)l * NK * kC * osz, of32, NK, nq, BT, K, nullptr, st));
    if (rec()) {
      const int wc = glob(G_CLSNORM2);
      tape.push_back([=]() -> int {
        if (io.g_cls)
          CQ_TRY(ln_bwd<T>(cls_out->p, nullptr, Wf(wc), 1e-5f, (const char*)io.g_cls + (size_t)l * NK * kC * osz, of32, nq, BT, K,
                           cls_out->g, beta(cls_out), nullptr, 0.f, G(wc), G(wc + 1), NK, st));
        return 0;
      });
    }
    CQ_TRY(set_branch(0));
    if (fwd())
      CQ_TRY(layernorm_permute<T>(out2->p, Wf(glob(G_NORM)), Wf(glob(G_NORM) + 1), 1e-5f, (char*)io.hs + (size_t)l * N * kC * osz,
                                  of32, N, nq, BT, 1, nullptr, st));
    if (rec()) {
      const int wn = glob(G_NORM);
      tape.push_back([=]() -> int {
        if (io.g_hs)
          CQ_TRY(ln_bwd<T>(out2->p, nullptr, Wf(wn), 1e-5f, (const char*)io.g_hs + (size_t)l * N * kC * osz, of32, nq, BT, 1, out2->g,
                           beta(out2), nullptr, 0.f, G(wn), G(wn + 1), N, st));
        return 0;
      });
    }
    // ---- iterative box refinement :813-823 ----
    Ten<T>* bb2 = lin(lin(out2, glob(G_BB0), kC, CQVAD_ACT_RELU, nullptr, 0, 0, &rc), glob(G_BB1), kC, CQVAD_ACT_RELU, nullptr, 0, 0, &rc);
    float* r_next = rl[l + 1];
    if (fwd())
      CQ_TRY(box_refine<T>(bb2->p, Wf(glob(G_BB2)), Wf(glob(G_BB2) + 1), r_cur, r_next,
                           (l != Lr - 1) ? io.refs + (size_t)(l + 1) * N * 4 : nullptr, N, nq, BT, false, st));
    if (rec() && l != Lr - 1) {   // the last layer's refined points are not an output (:819-820): no gradient
      const int wi = glob(G_BB2);
      tape.push_back([=]() -> int {
        if (!io.g_refs) return 0;
        return box_refine_bwd<T>(bb2->p, Wf(wi), Wf(wi + 1), r_cur, io.g_refs + (size_t)(l + 1) * N * 4, bb2->g, beta(bb2), G(wi),
                                 G(wi + 1), dr_cur, N, nq, BT, st);
      });
    }
    out = out2;
  }
  if (rc != 0) return rc;
  if (a.overflow) return 0;   // PLAN mode / caller checks
  if (fwd()) CQ_TRY(link(1, 0));   // join: the caller's stream waits for the class branch
  if (rec()) {
    // ---- backward: zero what is accumulated, then run the tape in reverse ----
    st = streams[0];
    CQ_CUDA(cudaMemsetAsync(drl[0], 0, (size_t)N * 4 * sizeof(float), st));
    CQ_TRY(build_all_wt());       // one launch for every transposed weight copy (280 launches of ~3 us on the chains before)
    CQ_TRY(link(0, 1));            // fork: the class-branch stream starts after everything already on the caller's stream
    int ti = (int)tape.v.size(), prev = 0, done_layer = Lr;
    static const bool eager_link = getenv("CQVAD_TRAIN_EAGER_LINK") != nullptr;   // A/B switch: join at the first loc op (round-1 schedule)
    bool pending_link = false;
    for (auto it = tape.v.rbegin(); it != tape.v.rend(); ++it) {
      // every backward op of the layers above it->layer has been enqueued: signal their parameter gradients as final
      while (done_layer - 1 > it->layer && done_layer - 1 >= 0) { --done_layer; CQ_TRY(signal_layer(done_layer)); }
      const int b = it->branch;
      // the loc layer consumes the class branch's gradients (q_memory, qse) -- but only from its `class_sync` op downwards: the
      // join is deferred to that op, so the head of the loc layer's backward overlaps the class branch
      if (b == 0 && prev == 1) pending_link = true;
      if (pending_link && (b == 1 || it->class_sync || eager_link)) { CQ_TRY(link(1, 0)); pending_link = false; }
      prev = b;
      st = streams[two_streams ? b : 0];
      set_wgrad_scratch(wg_scr[b], wgrad_scratch_bytes());
      CQ_TRY(it->fn());
      CQ_TRY(dbg("tape", --ti));
    }
    while (done_layer > 0) { --done_layer; CQ_TRY(signal_layer(done_layer)); }
    CQ_TRY(link(1, 0));            // join
    CQ_TRY(wgrad_join());
    st = streams[0];
  }
  return 0;
}

template <typename T>
int run_train(const cqvad_decoder_desc* d, const void* const* weights, const TrainIO& io, void* ws, size_t ws_bytes,
              cudaStream_t st, Mode mode, size_t* need) {
  const size_t skew = ws ? ((1024 - (((uintptr_t)ws) & 1023)) & 1023) : 0;
  if (mode != PLAN && ws_bytes < skew) return set_error(CQVAD_E_WORKSPACE, "decoder training workspace too small");
  Arena a(mode == PLAN ? nullptr : (char*)ws + skew, mode == PLAN ? 0 : ws_bytes - skew);
  if (mode != PLAN) {   // size check first: never enqueue anything on a workspace that is too small
    Arena probe(nullptr, 0);
    Trainer<T> t(*d, weights, io, st, probe, PLAN);
    CQ_TRY(t.run());
    if (probe.off + 1024 > ws_bytes)
      return set_error(CQVAD_E_WORKSPACE, "decoder training workspace too small (%zu needed, %zu given)", probe.off + 1024, ws_bytes);
  }
  Trainer<T> t(*d, weights, io, st, a, mode);
  int r = t.run();
  if (r != 0) t.join_after_error();
  if (need) *need = a.off + 1024;
  return r;
}

int check_train_desc(const cqvad_decoder_desc* d) {
  CQ_CHECK_ARG(d != nullptr, "decoder: null descriptor");
  CQ_CHECK_ARG(d->dtype == CQVAD_F32 || d->dtype == CQVAD_BF16, "decoder: unknown dtype %d", d->dtype);
  CQ_CHECK_ARG(d->BT >= 1 && d->nq >= 1 && d->h >= 1 && d->w >= 1 && d->K >= 1 && d->layers >= 1, "decoder: bad extents");
  CQ_CHECK_SHAPE(d->F >= 8 && d->F % 8 == 0, "decoder: dim_feedforward must be a multiple of 8");
  CQ_CHECK_SHAPE(d->w <= 128, "decoder: feature-map width %d > 128 not supported", d->w);
  return 0;
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

extern "C" int cqvad_decoder_backward_layer_events(void* const* events, int n) {
  CQ_CHECK_ARG(n >= 0 && (n == 0 || events != nullptr), "decoder_backward_layer_events: bad arguments");
  g_layer_events = n > 0 ? events : nullptr;
  g_layer_events_n = n;
  return 0;
}

extern "C" size_t cqvad_decoder_train_workspace_bytes(const cqvad_decoder_desc* d) {
  if (check_train_desc(d) != 0) return 0;
  TrainIO io{};
  size_t need = 0;
  int r = d->dtype == CQVAD_F32 ? run_train<float>(d, nullptr, io, nullptr, 0, nullptr, PLAN, &need)
                                : run_train<bf16>(d, nullptr, io, nullptr, 0, nullptr, PLAN, &need);
  return r == 0 ? need : 0;
}

extern "C" int cqvad_decoder_train_forward(const cqvad_decoder_desc* d, const void* const* weights, const float* tgt,
                                           const float* memory, const float* pos, const uint8_t* mask,
                                           const float* refpoints_unsigmoid, void* hs, void* cls_hs, float* refs,
                                           void* workspace, size_t ws_bytes, void* stream) {
  CQ_TRY(check_train_desc(d));
  CQ_CHECK_ARG(weights && tgt && memory && pos && refpoints_unsigmoid && hs && cls_hs && refs && workspace, "decoder: null pointer");
  CQ_CHECK_ARG(d->dropout_p >= 0.f && d->dropout_p < 1.f, "decoder: dropout_p must be in [0, 1)");
  TrainIO io{};
  io.tgt = tgt; io.memory = memory; io.pos = pos; io.mask = mask; io.ref_u = refpoints_unsigmoid;
  io.hs = hs; io.cls_hs = cls_hs; io.refs = refs;
  reset_launch_count();
  if (d->dtype == CQVAD_F32) return run_train<float>(d, weights, io, workspace, ws_bytes, as_stream(stream), FWD, nullptr);
  return run_train<bf16>(d, weights, io, workspace, ws_bytes, as_stream(stream), FWD, nullptr);
}

extern "C" int cqvad_decoder_backward(const cqvad_decoder_desc* d, const void* const* weights, const uint8_t* mask,
                                      const void* grad_hs, const void* grad_cls_hs, const float* grad_refs,
                                      float* const* grad_weights, float* grad_memory, float* grad_tgt,
                                      float* grad_refpoints_unsigmoid, void* workspace, size_t ws_bytes, void* stream) {
  CQ_TRY(check_train_desc(d));
  CQ_CHECK_ARG(weights && grad_weights && grad_memory && workspace, "decoder backward: null pointer");
  TrainIO io{};
  io.mask = mask;
  io.g_hs = grad_hs; io.g_cls = grad_cls_hs; io.g_refs = grad_refs;
  io.gw = grad_weights; io.g_memory = grad_memory; io.g_tgt = grad_tgt; io.g_ref_u = grad_refpoints_unsigmoid;
  reset_launch_count();
  if (d->dtype == CQVAD_F32) return run_train<float>(d, weights, io, workspace, ws_bytes, as_stream(stream), REPLAY, nullptr);
  return run_train<bf16>(d, weights, io, workspace, ws_bytes, as_stream(stream), REPLAY, nullptr);
}

// Weight gradient of a Linear / 3x3 conv as a stand-alone op (used by the op-level tests and by nn.Module backward hooks).
extern "C" size_t cqvad_wgrad_workspace_bytes(void) { return wgrad_scratch_bytes() + 1024; }
extern "C" int cqvad_linear_wgrad(int dtype, const void* dY, const void* X, float* dW, float* db, long M, int N, int K,
                                  int conv_h, int conv_w, void* workspace, size_t ws_bytes, void* stream) {
  CQ_CHECK_ARG(dY && X && (dW || db) && M >= 0 && N >= 1 && K >= 1, "linear_wgrad: bad argument");
  ConvGeom cg; cg.h = conv_h; cg.w = conv_w;
  const ConvGeom* conv = conv_w > 0 ? &cg : nullptr;
  const long ldw = conv ? 9L * K : K;
  if (dtype == CQVAD_F32)
    return wgrad<float>((const float*)dY, N, (const float*)X, K, dW, ldw, db, M, N, K, conv, as_stream(stream));
  if (dtype != CQVAD_BF16) return set_error(CQVAD_E_INVALID_ARG, "linear_wgrad: unknown dtype %d", dtype);
  if (workspace) {
    const size_t skew = (1024 - (((uintptr_t)workspace) & 1023)) & 1023;
    if (ws_bytes >= skew + wgrad_scratch_bytes()) set_wgrad_scratch((float*)((char*)workspace + skew), wgrad_scratch_bytes());
    else set_wgrad_scratch(nullptr, 0);
  } else {
    set_wgrad_scratch(nullptr, 0);
  }
  return wgrad<bf16>((const bf16*)dY, N, (const bf16*)X, K, dW, ldw, db, M, N, K, conv, as_stream(stream));
}
