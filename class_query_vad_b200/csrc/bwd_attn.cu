// Attention backward (models/detr/attention.py:190-422 under autograd).  The probabilities are recomputed from q/k (they
// are never stored by the forward), the softmax Jacobian is applied per row, and the three gradients are produced
// without atomics on q/k/v: the probability / score-gradient matrix of one (batch, head) lives in shared memory and the
// block switches its thread mapping between the phases (rows of P for dP/dS, (key, channel) pairs for dK/dV).
#include "common.cuh"
#include "bwd.cuh"

namespace cqvad {

namespace {

// block per (nb, h); 256 threads.
//   phase 0: stage K [S][hd], V [S][vd], Q*scale [Le][hd], dO [Le][vd]   (Le = 1 when the second query source is shared by
//            all L queries: the L identical rows collapse to one with dO summed -- exact)
//   phase 1: P = softmax(QK^T) rows -> PS [Le][S]
//   phase 2: dV[m][d] = sum_l P[l][m] dO[l][d]
//   phase 3: dS = P * (dP - rowsum(P*dP)),  dP = dO V^T               (PS overwritten)
//   phase 4: dK[m][d] = sum_l dS[l][m] Qs[l][d]
//   phase 5: dQ[l][d] = scale * sum_m dS[l][m] K[m][d]
template <typename T>
__global__ void __launch_bounds__(256) mha_std_bwd_kernel(const T* __restrict__ q, const T* __restrict__ q2,
                                                          const T* __restrict__ k, const T* __restrict__ k2,
                                                          const T* __restrict__ v, const uint8_t* __restrict__ kpm,
                                                          const T* __restrict__ dO, T* dq, float beta_q, T* dq2, float beta_q2,
                                                          T* dk, float beta_k, T* dv, float beta_v, int L, int S, int H, int hd,
                                                          int vd, StdStrides st, float scale) {
  extern __shared__ float smem[];
  const int nb = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const bool second = (q2 != nullptr) && (h >= H / 2);
  const bool bcast = second && st.q2_ls == 0;
  const int Le = bcast ? 1 : L;
  const int hh = second ? h - H / 2 : h;
  const int kst = hd + 1, vst = vd + 1;
  float* Ks = smem;                       // [S][hd+1]
  float* Vs = Ks + (size_t)S * kst;       // [S][vd+1]
  float* Qs = Vs + (size_t)S * vst;       // [Le][hd+1]
  float* Ds = Qs + (size_t)Le * kst;      // [Le][vd+1]
  float* PS = Ds + (size_t)Le * vst;      // [Le][S]
  const int nb2 = st.k2_bmod > 0 ? nb % st.k2_bmod : nb;
  const T* kb = second ? (k2 + (long)nb2 * st.k2_bs + (long)hh * hd) : (k + (long)nb * st.k_bs + (long)hh * hd);
  const long kls = second ? st.k2_ls : st.k_ls;
  for (int idx = tid; idx < S * hd; idx += blockDim.x) {
    const int s = idx / hd, d = idx % hd;
    Ks[s * kst + d] = to_f(kb[(long)s * kls + d]);
  }
  const T* vb = v + (long)nb * st.v_bs + (long)h * vd;
  for (int idx = tid; idx < S * vd; idx += blockDim.x) {
    const int s = idx / vd, d = idx % vd;
    Vs[s * vst + d] = to_f(vb[(long)s * st.v_ls + d]);
  }
  const T* qb = second ? (q2 + (long)nb * st.q2_bs + (long)hh * hd) : (q + (long)nb * st.q_bs + (long)hh * hd);
  const long qls = second ? st.q2_ls : st.q_ls;
  for (int idx = tid; idx < Le * hd; idx += blockDim.x) {
    const int l = idx / hd, d = idx % hd;
    Qs[l * kst + d] = to_f(qb[(long)l * qls + d]) * scale;
  }
  const T* dob = dO + (long)nb * st.o_bs + (long)h * vd;
  if (bcast) {
    for (int d = tid; d < vd; d += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < L; ++l) s += to_f(dob[(long)l * st.o_ls + d]);
      Ds[d] = s;
    }
  } else {
    for (int idx = tid; idx < L * vd; idx += blockDim.x) {
      const int l = idx / vd, d = idx % vd;
      Ds[l * vst + d] = to_f(dob[(long)l * st.o_ls + d]);
    }
  }
  __syncthreads();
  // phase 1
  for (int l = warp; l < Le; l += nwarps) {
    const float* qr = Qs + l * kst;
    float* pr = PS + (size_t)l * S;
    float mx = -INFINITY;
    for (int m = lane; m < S; m += 32) {
      float a = 0.f;
      const float* kr = Ks + m * kst;
      for (int d = 0; d < hd; ++d) a = fmaf(qr[d], kr[d], a);
      if (kpm && kpm[(long)nb * S + m]) a = -INFINITY;
      pr[m] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int m = lane; m < S; m += 32) {
      const float e = expf(pr[m] - mx);
      pr[m] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int m = lane; m < S; m += 32) pr[m] *= inv;
  }
  __syncthreads();
  // phase 2: dV
  if (dv) {
    T* dvb = dv + (long)nb * st.v_bs + (long)h * vd;
    for (int idx = tid; idx < S * vd; idx += blockDim.x) {
      const int m = idx / vd, d = idx % vd;
      float a = 0.f;
      for (int l = 0; l < Le; ++l) a = fmaf(PS[(size_t)l * S + m], Ds[l * vst + d], a);
      T* p = dvb + (long)m * st.v_ls + d;
      *p = from_f<T>(beta_v != 0.f ? fmaf(beta_v, to_f(*p), a) : a);
    }
  }
  __syncthreads();
  // phase 3: dS in place (dP recomputed in the second pass: each lane revisits its own keys, no exchange needed)
  for (int l = warp; l < Le; l += nwarps) {
    const float* dr = Ds + l * vst;
    float* pr = PS + (size_t)l * S;
    float dot = 0.f;
    for (int m = lane; m < S; m += 32) {
      float a = 0.f;
      const float* vr = Vs + m * vst;
      for (int d = 0; d < vd; ++d) a = fmaf(dr[d], vr[d], a);
      dot = fmaf(pr[m], a, dot);
    }
    dot = warp_sum(dot);
    for (int m = lane; m < S; m += 32) {
      float a = 0.f;
      const float* vr = Vs + m * vst;
      for (int d = 0; d < vd; ++d) a = fmaf(dr[d], vr[d], a);
      pr[m] = pr[m] * (a - dot);
    }
  }
  __syncthreads();
  // phase 4: dK (the positional keys k2 of the class cross-attention get no gradient)
  if (dk && !second) {
    T* dkb = dk + (long)nb * st.k_bs + (long)hh * hd;
    for (int idx = tid; idx < S * hd; idx += blockDim.x) {
      const int m = idx / hd, d = idx % hd;
      float a = 0.f;
      for (int l = 0; l < Le; ++l) a = fmaf(PS[(size_t)l * S + m], Qs[l * kst + d], a);
      T* p = dkb + (long)m * st.k_ls + d;
      *p = from_f<T>(beta_k != 0.f ? fmaf(beta_k, to_f(*p), a) : a);
    }
  }
  __syncthreads();
  // phase 5: dQ
  T* dqb = second ? (dq2 ? dq2 + (long)nb * st.q2_bs + (long)hh * hd : nullptr) : (dq ? dq + (long)nb * st.q_bs + (long)hh * hd : nullptr);
  const float bq = second ? beta_q2 : beta_q;
  if (dqb) {
    for (int idx = tid; idx < Le * hd; idx += blockDim.x) {
      const int l = idx / hd, d = idx % hd;
      float a = 0.f;
      const float* pr = PS + (size_t)l * S;
      for (int m = 0; m < S; ++m) a = fmaf(pr[m], Ks[m * kst + d], a);
      a *= scale;
      T* p = dqb + (long)l * qls + d;
      *p = from_f<T>(bq != 0.f ? fmaf(bq, to_f(*p), a) : a);
    }
  }
}

// Decoder localisation cross-attention backward: block per actor instance, warp = head.  Channel-vectorised: lane = (key slot
// g = lane / 4, 8-channel chunk ch = lane % 4) -- eight keys per step with one 16-byte load per operand and lane, dot products
// folded over the 4 lanes of a slot (2 shuffles instead of 5).  The lane-per-channel version walked the S keys one at a time
// with 2-byte loads: 133 us per launch for 192 MB of k / v / dk / dv traffic (one latency-bound wave of 480 blocks).
__device__ __forceinline__ void red_add_v4g(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <typename T>
__global__ void __launch_bounds__(256) dec_qsk_bwd_kernel(const T* __restrict__ qc, const T* __restrict__ qs,
                                                          const T* __restrict__ kc, const T* __restrict__ v, long ldkv,
                                                          const T* __restrict__ kp, const uint8_t* __restrict__ mask,
                                                          const T* __restrict__ dO, T* dqc, float beta_qc, T* dqs,
                                                          float beta_qs, T* __restrict__ dkc, T* __restrict__ dv,
                                                          float* __restrict__ dkp32, int S, int Sq, int BT, int first) {
  extern __shared__ float smem[];
  const long i = blockIdx.x;
  const int bb = (int)(i % BT);
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, ch = lane & 3;
  float* my_p = smem + (size_t)h * S;              // probabilities
  float* my_ds = smem + (size_t)(kH + h) * S;      // score gradients
  const int c = h * 32 + ch * 8;
  float qcv[8], qsv[8], qk[8], dov[8];
  load8(qc + i * kC + c, qcv);
  load8(qs + i * kC + c, qsv);
  load8(dO + i * kC + c, dov);
#pragma unroll
  for (int e = 0; e < 8; ++e) { qcv[e] *= 0.125f; qsv[e] *= 0.125f; qk[e] = first ? qcv[e] + qsv[e] : qsv[e]; }
  const T* kcb = kc + i * Sq * ldkv + c;
  const T* vb = v + i * Sq * ldkv + c;
  const T* kpb = kp + (long)bb * kC + c;
  float mx = -INFINITY;
#pragma unroll 4
  for (int s0 = 0; s0 < S; s0 += 8) {   // unrolled: four key groups of loads in flight per lane
    const int s = s0 + g;
    float r = 0.f, dp = 0.f;
    if (s < S) {
      float a[8], b[8], w[8];
      load8(kcb + (long)s * ldkv, a);
      load8(kpb + (long)s * BT * kC, b);
      load8(vb + (long)s * ldkv, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) { r = fmaf(qcv[e], a[e], fmaf(qk[e], b[e], r)); dp = fmaf(dov[e], w[e], dp); }
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1); r += __shfl_xor_sync(0xffffffffu, r, 2);
    dp += __shfl_xor_sync(0xffffffffu, dp, 1); dp += __shfl_xor_sync(0xffffffffu, dp, 2);
    if (s < S) {
      if (mask && mask[(long)bb * S + s]) r = -INFINITY;
      if (ch == 0) { my_p[s] = r; my_ds[s] = dp; }
      mx = fmaxf(mx, r);
    }
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int m = lane; m < S; m += 32) {
    const float e = expf(my_p[m] - mx);
    my_p[m] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  float dot = 0.f;
  for (int m = lane; m < S; m += 32) {
    const float p = my_p[m] * inv;
    my_p[m] = p;
    dot = fmaf(p, my_ds[m], dot);
  }
  dot = warp_sum(dot);
  for (int m = lane; m < S; m += 32) my_ds[m] = my_p[m] * (my_ds[m] - dot);
  __syncwarp();
  float aqc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, aqs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  T* dkcb = dkc + i * Sq * ldkv + c;
  T* dvb = dv + i * Sq * ldkv + c;
  float* dkpb = dkp32 + (long)bb * kC + c;
#pragma unroll 4
  for (int s0 = 0; s0 < S; s0 += 8) {   // unrolled: four key groups of loads in flight per lane
    const int s = s0 + g;
    if (s >= S) continue;
    const float ds = my_ds[s], p = my_p[s];
    float a[8], b[8], o1[8], o2[8];
    load8(kcb + (long)s * ldkv, a);
    load8(kpb + (long)s * BT * kC, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o1[e] = ds * qcv[e];
      o2[e] = p * dov[e];
      aqc[e] = fmaf(ds, first ? a[e] + b[e] : a[e], aqc[e]);
      aqs[e] = fmaf(ds, b[e], aqs[e]);
    }
    store8(dkcb + (long)s * ldkv, o1);
    store8(dvb + (long)s * ldkv, o2);
    float* dk = dkpb + (long)s * BT * kC;
    red_add_v4g(dk, ds * qk[0], ds * qk[1], ds * qk[2], ds * qk[3]);
    red_add_v4g(dk + 4, ds * qk[4], ds * qk[5], ds * qk[6], ds * qk[7]);
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { aqc[e] += __shfl_xor_sync(0xffffffffu, aqc[e], o); aqs[e] += __shfl_xor_sync(0xffffffffu, aqs[e], o); }
  }
  if (g == 0) {
    float oc[8], os[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { oc[e] = aqc[e] * 0.125f; os[e] = aqs[e] * 0.125f; }
    if (beta_qc != 0.f) { float t[8]; load8(dqc + i * kC + c, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) oc[e] = fmaf(beta_qc, t[e], oc[e]); }
    if (beta_qs != 0.f) { float t[8]; load8(dqs + i * kC + c, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) os[e] = fmaf(beta_qs, t[e], os[e]); }
    store8(dqc + i * kC + c, oc);
    store8(dqs + i * kC + c, os);
  }
}

template <typename K>
int set_smem_bwd(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    CQ_CHECK_SHAPE(bytes <= 227 * 1024, "attention backward: needs %zu bytes of shared memory (> 227 KB)", bytes);
    CQ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return 0;
}

}  // namespace

template <typename T>
int mha_std_bwd(const T* q, const T* q2, const T* k, const T* k2, const T* v, const uint8_t* kpm, const T* dO, T* dq,
                float beta_q, T* dq2, float beta_q2, T* dk, float beta_k, T* dv, float beta_v, int L, int S, int Nb, int H,
                int hd, int vd, const StdStrides& st, cudaStream_t stm) {
  if (L == 0 || Nb == 0) return 0;
  CQ_CHECK_SHAPE(S >= 1, "mha backward: S must be >= 1");
  const size_t smem = ((size_t)S * (hd + 1) + (size_t)S * (vd + 1) + (size_t)L * (hd + 1) + (size_t)L * (vd + 1) +
                       (size_t)L * S) * sizeof(float);
  CQ_TRY(set_smem_bwd(mha_std_bwd_kernel<T>, smem));
  mha_std_bwd_kernel<T><<<(unsigned)(Nb * H), 256, smem, stm>>>(q, q2, k, k2, v, kpm, dO, dq, beta_q, dq2, beta_q2, dk, beta_k,
                                                                dv, beta_v, L, S, H, hd, vd, st, 1.0f / sqrtf((float)hd));
  CQ_LAUNCH_CHECK();
  return 0;
}
template int mha_std_bwd<float>(const float*, const float*, const float*, const float*, const float*, const uint8_t*, const float*, float*, float, float*, float, float*, float, float*, float, int, int, int, int, int, int, const StdStrides&, cudaStream_t);
template int mha_std_bwd<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, const bf16*, const uint8_t*, const bf16*, bf16*, float, bf16*, float, bf16*, float, bf16*, float, int, int, int, int, int, int, const StdStrides&, cudaStream_t);

template <typename T>
int dec_qsk_bwd(const T* qc, const T* qs, const T* kc, const T* v, long ldkv, const T* kp, const uint8_t* mask, const T* dO,
                T* dqc, float beta_qc, T* dqs, float beta_qs, T* dkc, T* dv, float* dkp32, long N, int S, int Sq, int BT,
                bool first, cudaStream_t stm) {
  if (N == 0) return 0;
  const size_t smem = (size_t)2 * kH * S * sizeof(float);
  CQ_TRY(set_smem_bwd(dec_qsk_bwd_kernel<T>, smem));
  dec_qsk_bwd_kernel<T><<<(unsigned)N, 256, smem, stm>>>(qc, qs, kc, v, ldkv, kp, mask, dO, dqc, beta_qc, dqs, beta_qs, dkc, dv,
                                                         dkp32, S, Sq, BT, first ? 1 : 0);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int dec_qsk_bwd<float>(const float*, const float*, const float*, const float*, long, const float*, const uint8_t*, const float*, float*, float, float*, float, float*, float*, float*, long, int, int, int, bool, cudaStream_t);
template int dec_qsk_bwd<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, long, const bf16*, const uint8_t*, const bf16*, bf16*, float, bf16*, float, bf16*, bf16*, float*, long, int, int, int, bool, cudaStream_t);

}  // namespace cqvad
