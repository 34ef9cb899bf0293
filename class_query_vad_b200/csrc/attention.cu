// Attention cores of models/detr/attention.py:190-422 (projection-free MHA; out_proj is a GEMM done by the caller).
//   mha_std_kernel  : mode A (bmm form, :336-341,377,409). K/V head slices staged in shared memory (reused by all L
//                     queries), one warp per query row, lane = key for QK^T, lane = channel for PV.
//   mha_qsk_kernel  : mode B (query_specific_key einsum form, :343-346,379,411) -- every query owns its keys: GEMV
//                     shaped, HBM-bound; one warp per (query, head), keys streamed with coalesced loads.
//   dec_qsk_kernel  : the decoder's localisation cross-attention (dab_transformer.py:951-988) without materialising
//                     the per-head [content | position] concatenations of :972-979.
#include "common.cuh"
#include "attention.cuh"

namespace cqvad {

namespace {


template <typename T>
__global__ void __launch_bounds__(128) mha_std_kernel(const T* __restrict__ q, const T* __restrict__ q2,
                                                      const T* __restrict__ k, const T* __restrict__ k2,
                                                      const T* __restrict__ v, const uint8_t* __restrict__ kpm,
                                                      T* __restrict__ o, int L, int S, int H, int hd, int vd,
                                                      StdStrides st, float scale) {
  extern __shared__ float smem[];
  const int nb = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kst = hd + 1, vst = vd + 1;
  float* Ks = smem;                    // [S][hd+1]
  float* Vs = Ks + (size_t)S * kst;    // [S][vd+1]
  float* sc = Vs + (size_t)S * vst;    // [4][S]
  float* qs = sc + 4 * (size_t)S;      // [4][hd]
  const bool second = (q2 != nullptr) && (h >= H / 2);
  const int hh = second ? h - H / 2 : h;
  const int nb2 = st.k2_bmod > 0 ? nb % st.k2_bmod : nb;
  const T* kb = second ? (k2 + (long)nb2 * st.k2_bs + (long)hh * hd) : (k + (long)nb * st.k_bs + (long)hh * hd);
  const long kls = second ? st.k2_ls : st.k_ls;
  for (int idx = threadIdx.x; idx < S * hd; idx += 128) {
    const int s = idx / hd, d = idx % hd;
    Ks[s * kst + d] = to_f(kb[(long)s * kls + d]);
  }
  const T* vb = v + (long)nb * st.v_bs + (long)h * vd;
  for (int idx = threadIdx.x; idx < S * vd; idx += 128) {
    const int s = idx / vd, d = idx % vd;
    Vs[s * vst + d] = to_f(vb[(long)s * st.v_ls + d]);
  }
  __syncthreads();
  const T* qb = second ? (q2 + (long)nb * st.q2_bs + (long)hh * hd) : (q + (long)nb * st.q_bs + (long)hh * hd);
  const long qls = second ? st.q2_ls : st.q_ls;
  float* my_sc = sc + warp * S;
  float* my_q = qs + warp * hd;
  for (int l = warp; l < L; l += 4) {
    for (int d = lane; d < hd; d += 32) my_q[d] = to_f(qb[(long)l * qls + d]) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int m = lane; m < S; m += 32) {
      float a = 0.f;
      const float* kr = Ks + m * kst;
      for (int d = 0; d < hd; ++d) a = fmaf(my_q[d], kr[d], a);
      if (kpm && kpm[(long)nb * S + m]) a = -INFINITY;
      my_sc[m] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int m = lane; m < S; m += 32) {
      const float e = expf(my_sc[m] - mx);
      my_sc[m] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < vd; d += 32) {
      float a = 0.f;
      for (int m = 0; m < S; ++m) a = fmaf(my_sc[m], Vs[m * vst + d], a);
      o[(long)l * st.o_ls + (long)nb * st.o_bs + (long)h * vd + d] = from_f<T>(a * inv);
    }
    __syncwarp();
  }
}

// mode B: grid (L, Nb), block = 32*min(H,8) threads; warp w handles heads w, w+nw, ...
template <typename T>
__global__ void __launch_bounds__(256) mha_qsk_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                      const T* __restrict__ v, const uint8_t* __restrict__ kpm,
                                                      T* __restrict__ o, int S, int H, int hd, int vd, long q_ls,
                                                      long q_bs, long k_qs, long k_ls, long k_bs, long v_qs, long v_ls,
                                                      long v_bs, long o_ls, long o_bs, float scale) {
  extern __shared__ float smem[];
  const int l = blockIdx.x, nb = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* my_sc = smem + (size_t)warp * S;
  for (int h = warp; h < H; h += nw) {
    float qr[4];  // hd <= 128
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = lane + 32 * j;
      qr[j] = d < hd ? to_f(q[(long)l * q_ls + (long)nb * q_bs + (long)h * hd + d]) * scale : 0.f;
    }
    const T* kb = k + (long)l * k_qs + (long)nb * k_bs + (long)h * hd;
    float mx = -INFINITY;
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        if (d < hd) a = fmaf(qr[j], to_f(kb[(long)s * k_ls + d]), a);
      }
      a = warp_sum(a);
      if (kpm && kpm[(long)nb * S + s]) a = -INFINITY;
      if (lane == 0) my_sc[s] = a;
      mx = fmaxf(mx, a);
    }
    __syncwarp();
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float e = expf(my_sc[s] - mx);
      my_sc[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    const T* vb = v + (long)l * v_qs + (long)nb * v_bs + (long)h * vd;
    for (int d = lane; d < vd; d += 32) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a = fmaf(my_sc[s], to_f(vb[(long)s * v_ls + d]), a);
      o[(long)l * o_ls + (long)nb * o_bs + (long)h * vd + d] = from_f<T>(a * inv);
    }
    __syncwarp();
  }
}

// Decoder localisation cross-attention.  Block per actor instance i = n*BT + b, warp = head (8 heads x 32 dims).
//   score[h,s] = ( qc_h . (kc_h[i,s] + first*kp_h[s,b]) + qs_h . kp_h[s,b] ) / sqrt(64);  masked by mask[b,s]
//   o_h = sum_s softmax_s(score)[s] * v_h[i,s]
// kc, v: rows (i*S + s) with row stride ldkv (k at column 0.., v given as its own pointer); kp rows (s*BT + b).
// Channel-vectorised: warp = head, lane = (key slot g = lane / 4, 8-channel chunk ch = lane % 4): eight keys per step, one
// 16-byte load per operand and lane, scores folded over the 4 lanes of a slot; the PV pass accumulates 8 channels per lane over
// its key slot and folds the 8 slots at the end (the lane-per-channel version issued 2-byte loads key by key: 73 us per launch).
template <typename T>
__global__ void __launch_bounds__(256) dec_qsk_kernel(const T* __restrict__ qc, const T* __restrict__ qs,
                                                      const T* __restrict__ kc, const T* __restrict__ v, long ldkv,
                                                      const T* __restrict__ kp, const uint8_t* __restrict__ mask,
                                                      T* __restrict__ o, int S, int Sq, int BT, int first) {
  extern __shared__ float smem[];
  const long i = blockIdx.x;
  const int bb = (int)(i % BT);
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, ch = lane & 3;
  float* my_sc = smem + (size_t)h * S;
  const int c = h * 32 + ch * 8;
  float qcv[8], qk[8];
  {
    float qsv[8];
    load8(qc + i * kC + c, qcv);
    load8(qs + i * kC + c, qsv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      qcv[e] *= 0.125f;
      qk[e] = first ? qcv[e] + qsv[e] * 0.125f : qsv[e] * 0.125f;  // coefficient of kp: the first layer adds kp to the content key too (:964-967)
    }
  }
  const T* kcb = kc + i * Sq * ldkv + c;
  const T* kpb = kp + (long)bb * kC + c;
  float mx = -INFINITY;
#pragma unroll 4
  for (int s0 = 0; s0 < S; s0 += 8) {   // unrolled: four key groups of loads in flight per lane
    const int s = s0 + g;
    float r = 0.f;
    if (s < S) {
      float a[8], b[8];
      load8(kcb + (long)s * ldkv, a);
      load8(kpb + (long)s * BT * kC, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) r = fmaf(qcv[e], a[e], fmaf(qk[e], b[e], r));
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    if (s < S) {
      if (mask && mask[(long)bb * S + s]) r = -INFINITY;
      if (ch == 0) my_sc[s] = r;
      mx = fmaxf(mx, r);
    }
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int m = lane; m < S; m += 32) {
    const float e = expf(my_sc[m] - mx);
    my_sc[m] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const T* vb = v + i * Sq * ldkv + c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
  for (int s0 = 0; s0 < S; s0 += 8) {   // unrolled: four key groups of loads in flight per lane
    const int s = s0 + g;
    if (s >= S) continue;
    float w[8];
    load8(vb + (long)s * ldkv, w);
    const float pr = my_sc[s];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, w[e], acc[e]);
  }
#pragma unroll
  for (int of = 4; of < 32; of <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], of);
  }
  if (g == 0) {
    const float inv = 1.0f / sum;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    store8(o + i * kC + c, acc);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    CQ_CHECK_SHAPE(bytes <= 227 * 1024, "attention: needs %zu bytes of shared memory (> 227 KB)", bytes);
    CQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return 0;
}

}  // namespace

template <typename T>
int mha_std(const T* q, const T* q2, const T* k, const T* k2, const T* v, const uint8_t* kpm, T* o, int L, int S, int Nb,
            int H, int hd, int vd, const StdStrides& st, cudaStream_t stm) {
  if (L == 0 || Nb == 0) return 0;
  CQ_CHECK_SHAPE(S >= 1, "mha: S must be >= 1");
  const size_t smem = ((size_t)S * (hd + 1) + (size_t)S * (vd + 1) + 4 * (size_t)S + 4 * (size_t)hd) * sizeof(float);
  CQ_TRY(set_smem(mha_std_kernel<T>, smem));
  mha_std_kernel<T><<<(unsigned)(Nb * H), 128, smem, stm>>>(q, q2, k, k2, v, kpm, o, L, S, H, hd, vd, st,
                                                             1.0f / sqrtf((float)hd));
  CQ_LAUNCH_CHECK();
  return 0;
}
template int mha_std<float>(const float*, const float*, const float*, const float*, const float*, const uint8_t*, float*, int, int, int, int, int, int, const StdStrides&, cudaStream_t);
template int mha_std<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, const bf16*, const uint8_t*, bf16*, int, int, int, int, int, int, const StdStrides&, cudaStream_t);

template <typename T>
int mha_core(int mode, const T* q, const T* k, const T* v, const uint8_t* kpm, T* o, int L, int S, int Nb, int H, int E,
             int Ev, long q_ls, long q_bs, long k_ls, long k_bs, long k_qs, long v_ls, long v_bs, long v_qs, long o_ls,
             long o_bs, cudaStream_t stm) {
  CQ_CHECK_SHAPE(H >= 1 && E % H == 0 && Ev % H == 0, "mha: E (%d) and Ev (%d) must be divisible by H (%d)", E, Ev, H);
  const int hd = E / H, vd = Ev / H;
  if (mode == 0) {
    StdStrides ss{};
    ss.q_ls = q_ls; ss.q_bs = q_bs; ss.k_ls = k_ls; ss.k_bs = k_bs; ss.v_ls = v_ls; ss.v_bs = v_bs; ss.o_ls = o_ls; ss.o_bs = o_bs;
    return mha_std<T>(q, nullptr, k, nullptr, v, kpm, o, L, S, Nb, H, hd, vd, ss, stm);
  }
  CQ_CHECK_SHAPE(hd <= 128, "mha(query_specific_key): head dim %d > 128", hd);
  if (L == 0 || Nb == 0) return 0;
  const int nw = H < 8 ? H : 8;
  const size_t smem = (size_t)nw * S * sizeof(float);
  CQ_TRY(set_smem(mha_qsk_kernel<T>, smem));
  dim3 grid((unsigned)L, (unsigned)Nb);
  mha_qsk_kernel<T><<<grid, nw * 32, smem, stm>>>(q, k, v, kpm, o, S, H, hd, vd, q_ls, q_bs, k_qs, k_ls, k_bs, v_qs, v_ls,
                                                  v_bs, o_ls, o_bs, 1.0f / sqrtf((float)hd));
  CQ_LAUNCH_CHECK();
  return 0;
}
template int mha_core<float>(int, const float*, const float*, const float*, const uint8_t*, float*, int, int, int, int, int, int, long, long, long, long, long, long, long, long, long, long, cudaStream_t);
template int mha_core<bf16>(int, const bf16*, const bf16*, const bf16*, const uint8_t*, bf16*, int, int, int, int, int, int, long, long, long, long, long, long, long, long, long, long, cudaStream_t);

template <typename T>
int dec_qsk_attn(const T* qc, const T* qs, const T* kc, const T* v, long ldkv, const T* kp, const uint8_t* mask, T* o,
                 long N, int S, int Sq, int BT, bool first, cudaStream_t stm) {
  if (N == 0) return 0;
  const size_t smem = (size_t)kH * S * sizeof(float);
  CQ_TRY(set_smem(dec_qsk_kernel<T>, smem));
  dec_qsk_kernel<T><<<(unsigned)N, 256, smem, stm>>>(qc, qs, kc, v, ldkv, kp, mask, o, S, Sq, BT, first ? 1 : 0);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int dec_qsk_attn<float>(const float*, const float*, const float*, const float*, long, const float*, const uint8_t*, float*, long, int, int, int, bool, cudaStream_t);
template int dec_qsk_attn<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, long, const bf16*, const uint8_t*, bf16*, long, int, int, int, bool, cudaStream_t);

}  // namespace cqvad
