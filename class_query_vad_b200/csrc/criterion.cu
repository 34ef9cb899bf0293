// Loss / matcher / post-process on the device.
//   HungarianMatcherAVA.forward      models/detr/matcher.py:39-78   (the reference builds the cost on the GPU, copies it to the host
//                                    and calls scipy.optimize.linear_sum_assignment per clip: one D2H sync per decoder output)
//   SetCriterionAVA.loss_labels      models/detr/criterion.py:50-105 (person cross-entropy with eos weight; sigmoid focal loss with
//                                    label smoothing and a positive-row weight, models/detr/segmentation.py:200-229)
//   SetCriterionAVA.loss_boxes       models/detr/criterion.py:119-138 (L1 + GIoU on matched pairs, utils/box_ops.py:40-108)
//   PostProcessAVA.forward           models/detr/criterion.py:740-773
// Design: one CTA per clip.  The cost matrix (<= 64 queries x <= 64 targets) lives in shared memory; thread 0 solves the
// assignment exactly (shortest-augmenting-path Hungarian in fp64 on the fp32 costs -- scipy converts the same fp32 matrix to
// double), every thread then evaluates its share of the focal terms.  Per-clip partial sums are reduced in a fixed order by a
// one-CTA kernel (deterministic), and a third kernel writes the gradients of the weighted total with respect to the three
// prediction tensors, so that the training step needs no host round trip between the decoder forward and its backward.
// The total follows the reference training loop (train.py:148): sum over criterion.weight_dict = the LAST layer's four losses
// (the auxiliary per-layer losses are computed and logged by the reference, but their `_i` keys are not in weight_dict).
#include <float.h>
#include "common.cuh"

namespace cqvad {
namespace {

constexpr int kMaxQ = 64, kMaxT = 64;

struct CritCfg {
  float cost_class, cost_bbox, cost_giou;
  float w_ce, w_bbox, w_giou, w_ce_b;
  float pos_weight, eos_coef, alpha, gamma, smooth;
};

__device__ __forceinline__ void xyxy(const float* b, float& x0, float& y0, float& x1, float& y1) {   // box_cxcywh_to_xyxy
  x0 = b[0] - 0.5f * b[2]; y0 = b[1] - 0.5f * b[3]; x1 = b[0] + 0.5f * b[2]; y1 = b[1] + 0.5f * b[3];
}
// generalized_box_iou of one pair (utils/box_ops.py:83-108); optionally the gradient with respect to the FIRST box (cx,cy,w,h)
__device__ float giou_pair(const float* a, const float* t, float* grad /* [4] or nullptr */) {
  float ax0, ay0, ax1, ay1, tx0, ty0, tx1, ty1;
  xyxy(a, ax0, ay0, ax1, ay1); xyxy(t, tx0, ty0, tx1, ty1);
  const float area1 = (ax1 - ax0) * (ay1 - ay0), area2 = (tx1 - tx0) * (ty1 - ty0);
  const float ix0 = fmaxf(ax0, tx0), iy0 = fmaxf(ay0, ty0), ix1 = fminf(ax1, tx1), iy1 = fminf(ay1, ty1);
  const float iw = fmaxf(ix1 - ix0, 0.f), ih = fmaxf(iy1 - iy0, 0.f);
  const float inter = iw * ih, uni = area1 + area2 - inter, iou = inter / uni;
  const float ex0 = fminf(ax0, tx0), ey0 = fminf(ay0, ty0), ex1 = fmaxf(ax1, tx1), ey1 = fmaxf(ay1, ty1);
  const float ew = fmaxf(ex1 - ex0, 0.f), eh = fmaxf(ey1 - ey0, 0.f), earea = ew * eh;
  const float g = iou - (earea - uni) / earea;
  if (grad) {
    // reverse mode: g = inter/uni - 1 + uni/earea
    const float d_inter0 = 1.f / uni, d_uni = -inter / (uni * uni) + 1.f / earea, d_earea = -uni / (earea * earea);
    const float d_area1 = d_uni, d_inter = d_inter0 - d_uni;
    const float d_iw = d_inter * ih * (ix1 - ix0 >= 0.f ? 1.f : 0.f), d_ih = d_inter * iw * (iy1 - iy0 >= 0.f ? 1.f : 0.f);
    const float d_ew = d_earea * eh * (ex1 - ex0 >= 0.f ? 1.f : 0.f), d_eh = d_earea * ew * (ey1 - ey0 >= 0.f ? 1.f : 0.f);
    float gx0 = -d_area1 * (ay1 - ay0), gx1 = d_area1 * (ay1 - ay0), gy0 = -d_area1 * (ax1 - ax0), gy1 = d_area1 * (ax1 - ax0);
    if (ax0 > tx0) gx0 -= d_iw; else if (ax0 == tx0) gx0 -= 0.5f * d_iw;          // ix0 = max(ax0, tx0)
    if (ay0 > ty0) gy0 -= d_ih; else if (ay0 == ty0) gy0 -= 0.5f * d_ih;
    if (ax1 < tx1) gx1 += d_iw; else if (ax1 == tx1) gx1 += 0.5f * d_iw;          // ix1 = min(ax1, tx1)
    if (ay1 < ty1) gy1 += d_ih; else if (ay1 == ty1) gy1 += 0.5f * d_ih;
    if (ax0 < tx0) gx0 -= d_ew; else if (ax0 == tx0) gx0 -= 0.5f * d_ew;          // ex0 = min(ax0, tx0)
    if (ay0 < ty0) gy0 -= d_eh; else if (ay0 == ty0) gy0 -= 0.5f * d_eh;
    if (ax1 > tx1) gx1 += d_ew; else if (ax1 == tx1) gx1 += 0.5f * d_ew;          // ex1 = max(ax1, tx1)
    if (ay1 > ty1) gy1 += d_eh; else if (ay1 == ty1) gy1 += 0.5f * d_eh;
    grad[0] = gx0 + gx1; grad[1] = gy0 + gy1; grad[2] = 0.5f * (gx1 - gx0); grad[3] = 0.5f * (gy1 - gy0);
  }
  return g;
}

// Exact rectangular assignment: n rows (each gets a distinct column), m >= n columns.  cost(i, j) read through `at`.
// Shortest augmenting paths with potentials (O(n^2 m)); 1-based arrays as in the classical formulation.
template <typename F>
__device__ void hungarian(int n, int m, F at, int* col_of_row /* [n] */) {
  double u[kMaxT + 1], v[kMaxQ + 1], minv[kMaxQ + 1];
  int p[kMaxQ + 1], way[kMaxQ + 1];
  bool used[kMaxQ + 1];
  for (int j = 0; j <= m; ++j) { v[j] = 0; p[j] = 0; }
  for (int i = 0; i <= n; ++i) u[i] = 0;
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int j = 0; j <= m; ++j) { minv[j] = DBL_MAX; used[j] = false; }
    do {
      used[j0] = true;
      const int i0 = p[j0];
      double delta = DBL_MAX;
      int j1 = 0;
      for (int j = 1; j <= m; ++j) {
        if (used[j]) continue;
        const double cur = (double)at(i0 - 1, j - 1) - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= m; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
  }
  for (int j = 1; j <= m; ++j)
    if (p[j] > 0) col_of_row[p[j] - 1] = j - 1;
}

// ---- kernel A: cost matrix + assignment.  match[b, q] = target index or -1 ------------------------------------------------
__global__ void __launch_bounds__(128) crit_match_kernel(CritCfg c, const float* __restrict__ boxes, const float* __restrict__ logits_b,
                                                         const float* __restrict__ tboxes, const int* __restrict__ n_tgt,
                                                         int* __restrict__ match, int nq, int maxT) {
  __shared__ float sC[kMaxQ * kMaxT];
  __shared__ float sB[kMaxQ * 4], sT[kMaxT * 4], sP[kMaxQ];
  __shared__ int sM[kMaxQ > kMaxT ? kMaxQ : kMaxT];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nt = min(n_tgt[b], maxT);
  for (int i = tid; i < nq * 4; i += blockDim.x) sB[i] = boxes[(long)b * nq * 4 + i];
  for (int i = tid; i < nt * 4; i += blockDim.x) sT[i] = tboxes[(long)b * maxT * 4 + i];
  for (int q = tid; q < nq; q += blockDim.x) {     // softmax(pred_logits_b)[1]  (matcher.py:68-69)
    const float* z = logits_b + ((long)b * nq + q) * 3;
    const float mx = fmaxf(z[0], fmaxf(z[1], z[2]));
    const float e0 = expf(z[0] - mx), e1 = expf(z[1] - mx), e2 = expf(z[2] - mx);
    sP[q] = e1 / (e0 + e1 + e2);
    match[(long)b * nq + q] = -1;
  }
  __syncthreads();
  for (int i = tid; i < nq * nt; i += blockDim.x) {
    const int q = i / nt, j = i % nt;
    const float* a = sB + q * 4; const float* t = sT + j * 4;
    const float l1 = fabsf(a[0] - t[0]) + fabsf(a[1] - t[1]) + fabsf(a[2] - t[2]) + fabsf(a[3] - t[3]);   // cdist p=1 (:62)
    sC[q * kMaxT + j] = c.cost_bbox * l1 + c.cost_giou * (-giou_pair(a, t, nullptr)) + c.cost_class * (-sP[q]);   // :71
  }
  __syncthreads();
  if (tid == 0 && nt > 0) {
    if (nt <= nq) {          // every target gets a distinct query
      hungarian(nt, nq, [&](int j, int q) { return sC[q * kMaxT + j]; }, sM);
      for (int j = 0; j < nt; ++j) match[(long)b * nq + sM[j]] = j;
    } else {                 // more targets than queries: every query gets a distinct target
      hungarian(nq, nt, [&](int q, int j) { return sC[q * kMaxT + j]; }, sM);
      for (int q = 0; q < nq; ++q) match[(long)b * nq + q] = sM[q];
    }
  }
}

// per-clip partials: [0] focal sum (before / n_p), [1] L1 sum, [2] sum(1 - giou), [3] sum w*nll (person CE), [4] sum w (person CE),
// [5] sum of matched target labels (n_p), [6] matched pairs, [7] rows whose top-k class set equals the label set (class_error)
constexpr int kPart = 8;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

__device__ __forceinline__ float focal_term(float x, float t, float w, float alpha, float gamma, float* dx) {
  // sigmoid_focal_loss (models/detr/segmentation.py:216-228) for one element; dx = d term / d x
  const float p = 1.f / (1.f + expf(-x));
  const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));     // binary_cross_entropy_with_logits
  const float pt = p * t + (1.f - p) * (1.f - t);
  const float om = 1.f - pt;
  const float mod = (gamma == 2.f) ? om * om : powf(om, gamma);
  const float at = alpha >= 0.f ? alpha * t + (1.f - alpha) * (1.f - t) : 1.f;
  if (dx) {
    const float dmod = (gamma == 2.f) ? 2.f * om : gamma * powf(om, gamma - 1.f);
    const float dpt = p * (1.f - p) * (2.f * t - 1.f);
    *dx = at * w * ((p - t) * mod - ce * dmod * dpt);
  }
  return at * w * ce * mod;
}

// ---- kernel B: per-clip loss partials (GRAD = false) or gradients (GRAD = true, needs the global normalisers) ---------------
template <bool GRAD>
__global__ void __launch_bounds__(256) crit_loss_kernel(CritCfg c, const float* __restrict__ logits, const float* __restrict__ boxes,
                                                        const float* __restrict__ logits_b, const float* __restrict__ tboxes,
                                                        const float* __restrict__ tlabels, const int* __restrict__ match,
                                                        float* __restrict__ part, const float* __restrict__ losses,
                                                        float* __restrict__ g_logits, float* __restrict__ g_boxes,
                                                        float* __restrict__ g_logits_b, int nq, int K, int maxT) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float tl = (1.f - c.smooth) * 1.f + 0.5f * c.smooth, fl = 0.5f * c.smooth;   // criterion.py:70-77 (label smoothing)
  float n_p = 1.f, num_boxes = 1.f, wsum = 1.f;
  if (GRAD) { n_p = losses[8]; num_boxes = losses[9]; wsum = losses[10]; }
  float focal = 0.f;
  for (int i = tid; i < nq * K; i += blockDim.x) {
    const int q = i / K, k = i % K;
    const int j = match[(long)b * nq + q];
    float t = fl, w = 1.f;
    if (j >= 0) { const float lab = tlabels[((long)b * maxT + j) * K + k]; t = lab == 0.f ? fl : (lab == 1.f ? tl : lab); w = c.pos_weight; }
    float dx;
    focal += focal_term(logits[((long)b * nq + q) * K + k], t, w, c.alpha, c.gamma, GRAD ? &dx : nullptr);
    if (GRAD && g_logits) g_logits[((long)b * nq + q) * K + k] = c.w_ce * dx / ((float)K * n_p);
  }
  float l1 = 0.f, gi = 0.f, nll = 0.f, ws = 0.f, labs = 0.f, npair = 0.f, hit = 0.f;
  for (int q = tid; q < nq; q += blockDim.x) {
    const int j = match[(long)b * nq + q];
    const float* z = logits_b + ((long)b * nq + q) * 3;
    const int y = j >= 0 ? 1 : 2;                                           // criterion.py:60-62
    const float wy = y == 2 ? c.eos_coef : 1.f;
    const float mx = fmaxf(z[0], fmaxf(z[1], z[2]));
    const float e0 = expf(z[0] - mx), e1 = expf(z[1] - mx), e2 = expf(z[2] - mx), se = e0 + e1 + e2;
    nll += wy * (logf(se) + mx - z[y]);
    ws += wy;
    if (GRAD && g_logits_b) {
      float* g = g_logits_b + ((long)b * nq + q) * 3;
      const float s = c.w_ce_b * wy / wsum;
      g[0] = s * (e0 / se - (y == 0)); g[1] = s * (e1 / se - (y == 1)); g[2] = s * (e2 / se - (y == 2));
    }
    float gb[4] = {0.f, 0.f, 0.f, 0.f};
    if (j >= 0) {
      const float* a = boxes + ((long)b * nq + q) * 4; const float* t = tboxes + ((long)b * maxT + j) * 4;
      float gg[4];
      const float g = giou_pair(a, t, GRAD ? gg : nullptr);
      gi += 1.f - g;
      for (int e = 0; e < 4; ++e) {
        const float d = a[e] - t[e];
        l1 += fabsf(d);
        if (GRAD) gb[e] = c.w_bbox * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / num_boxes - c.w_giou * gg[e] / num_boxes;
      }
      npair += 1.f;
      if (!GRAD) {
        // n_p (:67) and accuracy_sigmoid (utils/misc.py:468-489): top-(number of labels) predicted classes == label set
        int nl = 0;
        for (int k = 0; k < K; ++k) { const float lab = tlabels[((long)b * maxT + j) * K + k]; labs += lab; nl += lab != 0.f; }
        bool ok = nl > 0;
        // reference quirk: accuracy_sigmoid receives the SMOOTHED targets (criterion.py:68,76-77 modify them in place), whose
        // every entry is non-zero, so with label smoothing every matched row counts as correct and class_error is 0
        if (c.smooth > 0.f) { hit += 1.f; continue; }
        const float* x = logits + ((long)b * nq + q) * K;
        for (int k = 0; k < K && ok; ++k) {
          if (tlabels[((long)b * maxT + j) * K + k] == 0.f) continue;
          int rank = 0;                                   // number of classes scoring strictly higher (ties: lower index first, as topk)
          for (int k2 = 0; k2 < K; ++k2) rank += (x[k2] > x[k]) || (x[k2] == x[k] && k2 < k);
          ok = rank < nl;
        }
        hit += ok ? 1.f : 0.f;
      }
    }
    if (GRAD && g_boxes) for (int e = 0; e < 4; ++e) g_boxes[((long)b * nq + q) * 4 + e] = gb[e];
  }
  if (GRAD) return;
  const float vals[kPart] = {focal / (float)K, l1, gi, nll, ws, labs, npair, hit};
  for (int i = 0; i < kPart; ++i) {
    const float s = block_sum(vals[i], red);
    if (tid == 0) part[(long)b * kPart + i] = s;
  }
}

// ---- kernel C: fixed-order reduction of the per-clip partials -> losses[0..7] and the normalisers [8..10] -------------------
// losses: [0] loss_ce [1] loss_bbox [2] loss_giou [3] loss_ce_b [4] weighted total [5] class_error [6] matched pairs [7] num_boxes
__global__ void crit_reduce_kernel(CritCfg c, const float* __restrict__ part, const int* __restrict__ n_tgt, float* __restrict__ losses,
                                   int B) {
  if (threadIdx.x != 0) return;
  double s[kPart] = {0, 0, 0, 0, 0, 0, 0, 0};
  double nb = 0;
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < kPart; ++i) s[i] += part[(long)b * kPart + i];
    nb += n_tgt[b];
  }
  const float n_p = fmaxf((float)s[5], 1.f);                 // criterion.py:67
  const float num_boxes = (float)nb;                         // :196-197 (no world-size normalisation, no clamp in the reference)
  losses[0] = (float)s[0] / n_p;
  losses[1] = (float)s[1] / num_boxes;
  losses[2] = (float)s[2] / num_boxes;
  losses[3] = (float)(s[3] / s[4]);
  losses[4] = c.w_ce * losses[0] + c.w_bbox * losses[1] + c.w_giou * losses[2] + c.w_ce_b * losses[3];
  losses[5] = s[6] > 0 ? 100.f - (float)(s[7] * (100.0 / s[6])) : 100.f;
  losses[6] = (float)s[6];
  losses[7] = num_boxes;
  losses[8] = n_p; losses[9] = num_boxes; losses[10] = (float)s[4];
}

// PostProcessAVA.forward (criterion.py:740-773): det[b, q, :] = [sigmoid(logits) (K) | box xyxy * (w, h, w, h) (4) | softmax(logits_b)[1]]
// ucf != 0: PostProcessUCF / PostProcessJHMDB (criterion.py:775-846): scores = sigmoid(inverse_sigmoid(sigmoid(logits) * person_prob))
__global__ void postprocess_ava_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const float* __restrict__ logits_b,
                                       const float* __restrict__ sizes /* [B,2] (h, w) */, float* __restrict__ det, long rows, int nq, int K,
                                       int ucf) {
  const long r = blockIdx.x;
  if (r >= rows) return;
  const int b = (int)(r / nq);
  float* o = det + r * (K + 5);
  float pb = 1.f;
  if (ucf) {
    const float* z = logits_b + r * 3;
    const float mx = fmaxf(z[0], fmaxf(z[1], z[2]));
    const float e0 = expf(z[0] - mx), e1 = expf(z[1] - mx), e2 = expf(z[2] - mx);
    pb = e1 / (e0 + e1 + e2);
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float p = 1.f / (1.f + expf(-logits[r * K + k]));
    if (ucf) {   // inverse_sigmoid (utils/misc.py:530-534, eps 1e-5) followed by sigmoid
      const float x = fminf(fmaxf(p * pb, 0.f), 1.f);
      const float y = logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
      p = 1.f / (1.f + expf(-y));
    }
    o[k] = p;
  }
  if (threadIdx.x == 0) {
    const float ih = sizes[b * 2], iw = sizes[b * 2 + 1];
    float x0, y0, x1, y1;
    xyxy(boxes + r * 4, x0, y0, x1, y1);
    o[K] = x0 * iw; o[K + 1] = y0 * ih; o[K + 2] = x1 * iw; o[K + 3] = y1 * ih;
    const float* z = logits_b + r * 3;
    const float mx = fmaxf(z[0], fmaxf(z[1], z[2]));
    const float e0 = expf(z[0] - mx), e1 = expf(z[1] - mx), e2 = expf(z[2] - mx);
    o[K + 4] = e1 / (e0 + e1 + e2);
  }
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

extern "C" size_t cqvad_criterion_ava_workspace_bytes(int B) { return (size_t)B * kPart * sizeof(float) + 256; }

extern "C" int cqvad_criterion_ava(const cqvad_criterion_cfg* cfg, const float* pred_logits, const float* pred_boxes,
                                   const float* pred_logits_b, const float* tgt_boxes, const float* tgt_labels, const int32_t* n_tgt,
                                   int B, int nq, int K, int maxT, int32_t* match, float* losses, float* grad_logits,
                                   float* grad_boxes, float* grad_logits_b, void* workspace, size_t ws_bytes, void* stream) {
  CQ_CHECK_ARG(cfg && pred_logits && pred_boxes && pred_logits_b && tgt_boxes && tgt_labels && n_tgt && match && losses && workspace,
               "criterion_ava: null pointer");
  CQ_CHECK_ARG(B >= 1 && nq >= 1 && K >= 1 && maxT >= 1, "criterion_ava: bad extents");
  CQ_CHECK_SHAPE(nq <= kMaxQ && maxT <= kMaxT, "criterion_ava: at most %d queries and %d targets per clip", kMaxQ, kMaxT);
  if (ws_bytes < cqvad_criterion_ava_workspace_bytes(B)) return set_error(CQVAD_E_WORKSPACE, "criterion_ava: workspace too small");
  CritCfg c{cfg->cost_class, cfg->cost_bbox, cfg->cost_giou, cfg->w_ce, cfg->w_bbox, cfg->w_giou, cfg->w_ce_b,
            cfg->pos_weight, cfg->eos_coef, cfg->focal_alpha, cfg->focal_gamma, cfg->label_smoothing};
  cudaStream_t st = as_stream(stream);
  float* part = (float*)workspace;
  crit_match_kernel<<<B, 128, 0, st>>>(c, pred_boxes, pred_logits_b, tgt_boxes, n_tgt, match, nq, maxT);
  CQ_LAUNCH_CHECK();
  crit_loss_kernel<false><<<B, 256, 0, st>>>(c, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, match, part, nullptr,
                                             nullptr, nullptr, nullptr, nq, K, maxT);
  CQ_LAUNCH_CHECK();
  crit_reduce_kernel<<<1, 32, 0, st>>>(c, part, n_tgt, losses, B);
  CQ_LAUNCH_CHECK();
  if (grad_logits || grad_boxes || grad_logits_b) {
    crit_loss_kernel<true><<<B, 256, 0, st>>>(c, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, match, part, losses,
                                              grad_logits, grad_boxes, grad_logits_b, nq, K, maxT);
    CQ_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int cqvad_postprocess_ava(const float* pred_logits, const float* pred_boxes, const float* pred_logits_b,
                                     const float* target_sizes, float* detections, int B, int nq, int K, void* stream) {
  CQ_CHECK_ARG(pred_logits && pred_boxes && pred_logits_b && target_sizes && detections && B >= 0 && nq >= 1 && K >= 1,
               "postprocess_ava: bad argument");
  if (B == 0) return 0;
  postprocess_ava_kernel<<<(unsigned)((long)B * nq), 128, 0, as_stream(stream)>>>(pred_logits, pred_boxes, pred_logits_b, target_sizes,
                                                                                  detections, (long)B * nq, nq, K, 0);
  CQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int cqvad_postprocess_ucf(const float* pred_logits, const float* pred_boxes, const float* pred_logits_b,
                                     const float* target_sizes, float* detections, int B, int nq, int K, void* stream) {
  CQ_CHECK_ARG(pred_logits && pred_boxes && pred_logits_b && target_sizes && detections && B >= 0 && nq >= 1 && K >= 1,
               "postprocess_ucf: bad argument");
  if (B == 0) return 0;
  postprocess_ava_kernel<<<(unsigned)((long)B * nq), 128, 0, as_stream(stream)>>>(pred_logits, pred_boxes, pred_logits_b, target_sizes,
                                                                                  detections, (long)B * nq, nq, K, 1);
  CQ_LAUNCH_CHECK();
  return 0;
}
