// Backward-pass kernels of the decoder that are not dense contractions on the tensor pipe: LayerNorm backward (with the
// gamma/beta reductions), the level-mix / actor-add LayerNorm fusions, activation derivatives, the reference-point
// chain (sine embedding, modulated query_sine_embed, box refinement), small-N linears, and the CUDA-core weight-gradient
// GEMM (fp32 parity mode and shapes the tcgen05 wgrad kernel does not take).  Same layout conventions as kernels_mem.cu:
// one warp per 256-channel row, 8 channels per lane, 16-byte accesses, shuffle reductions.
//
// Gradient spec = autograd of the reference forward (models/detr/dab_transformer.py:722-852, 907-997, 1040-1079); parity is
// checked against torch autograd of the unmodified reference (tests/golden/grad_*.npz, oracle/make_golden_grads.py).
#include <algorithm>
#include "common.cuh"
#include "bwd.cuh"
#include "kernels_mem.cuh"
#include "tc_common.cuh"

namespace cqvad {

namespace {
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
inline unsigned row_grid(long rows) { return (unsigned)cdiv(rows, kWarps); }
inline unsigned persist_grid(long rows) {
  long b = cdiv(rows, kWarps);
  return (unsigned)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

__device__ __forceinline__ float half_sum_lo(float v, int lane) { return warp_sum(lane < 16 ? v : 0.f); }
__device__ __forceinline__ float half_sum_hi(float v, int lane) { return warp_sum(lane >= 16 ? v : 0.f); }

// LayerNorm backward of one row held as 8 values per lane.  z: pre-norm values, dy: upstream gradient.
// Returns dz; accumulates the per-lane gamma / beta gradient partials.
__device__ __forceinline__ void ln_bwd_row(const float (&z)[8], const float (&dy)[8], const float (&g)[8], float eps,
                                           float (&dz)[8], float (&ag)[8], float (&ab)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += z[i];
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = z[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + eps);
  float xh[8], dxh[8], m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xh[i] = (z[i] - mean) * rstd;
    dxh[i] = dy[i] * g[i];
    m1 += dxh[i];
    m2 = fmaf(dxh[i], xh[i], m2);
    ag[i] = fmaf(dy[i], xh[i], ag[i]);
    ab[i] += dy[i];
  }
  m1 = warp_sum(m1) * (1.0f / 256.0f);
  m2 = warp_sum(m2) * (1.0f / 256.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) dz[i] = rstd * (dxh[i] - m1 - xh[i] * m2);
}

// block-level reduction of the per-lane gamma/beta partials of all warps, then one atomicAdd per channel
__device__ __forceinline__ void flush_gb(float (&ag)[8], float (&ab)[8], float* __restrict__ dg, float* __restrict__ db,
                                         float (*red)[2][256]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[warp][0][lane * 8 + i] = ag[i]; red[warp][1][lane * 8 + i] = ab[i]; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 512; c += blockDim.x) {
    const int which = c >> 8, ch = c & 255;
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w][which][ch];
    float* dst = which ? db : dg;
    if (dst) atomicAdd(dst + ch, s);
  }
}

// block-level reduction of NR x 256 per-lane partials (small-N linear weight gradients), one atomicAdd per element
template <int NR>
__device__ __forceinline__ void flush_rows(const float (&acc)[NR][8], float* __restrict__ dst, float* red /*[kWarps][NR*256]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(warp * NR + r) * 256 + lane * 8 + i] = acc[r][i];
  __syncthreads();
  for (int c = threadIdx.x; c < NR * 256; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += red[w * NR * 256 + c];
    atomicAdd(dst + c, s);
  }
  __syncthreads();
}
inline unsigned small_grid(long rows) {
  long b = cdiv(rows, kWarps * 4);
  return (unsigned)(b < 1 ? 1 : (b > 64 ? 64 : b));
}

template <typename T>
__device__ __forceinline__ void store_beta(T* p, const float (&v)[8], float beta) {
  if (beta != 0.f) {
    float o[8];
    load8(p, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(beta, o[i], v[i]);
    store8(p, o);
  } else {
    store8(p, v);
  }
}

// ---- LayerNorm backward ------------------------------------------------------------------------------------------
template <typename T, typename DY>
__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                          const float* __restrict__ g, float eps, const DY* __restrict__ dy,
                                                          int nq, int BT, int K, T* dx, float beta_x, T* dres, float beta_r,
                                                          float* __restrict__ dg, float* __restrict__ db, long rows) {
  __shared__ float red[kWarps][2][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float gm[8];
  load8(g + lane * 8, gm);
  for (long row = (long)blockIdx.x * kWarps + warp; row < rows; row += (long)gridDim.x * kWarps) {
    float z[8];
    load8(x + row * kC + lane * 8, z);
    if (res) {
      float r[8];
      load8(res + row * kC + lane * 8, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] += r[i];
    }
    long drow = row;
    if (nq > 0) {   // caller layout [b][n][k]
      const long i = row / K;
      const int k = (int)(row % K);
      const int n = (int)(i / BT), bb = (int)(i % BT);
      drow = ((long)bb * nq + n) * K + k;
    }
    float d[8], dz[8];
    load8(dy + drow * kC + lane * 8, d);
    ln_bwd_row(z, d, gm, eps, dz, ag, ab);
    store_beta(dx + row * kC + lane * 8, dz, beta_x);
    if (dres) store_beta(dres + row * kC + lane * 8, dz, beta_r);
  }
  flush_gb(ag, ab, dg, db, red);
}

// ---- level mix + norm_ backward ----------------------------------------------------------------------------------
// warp per (s, b): the four memory rows are loaded once and reused by the nq actors that share them
template <typename T>
__global__ void __launch_bounds__(kThreads) lvlmix_ln_bwd_kernel(const T* __restrict__ mem, const float* __restrict__ lvlw,
                                                                 const float* __restrict__ g, const T* __restrict__ dqm,
                                                                 float* __restrict__ dmem32, float* __restrict__ dlvlw,
                                                                 float* __restrict__ dg, float* __restrict__ db, int nq,
                                                                 int S, int Sq, int BT) {
  __shared__ float red[kWarps][2][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float gm[8];
  load8(g + lane * 8, gm);
  const long rows = (long)S * BT;
  for (long row = (long)blockIdx.x * kWarps + warp; row < rows; row += (long)gridDim.x * kWarps) {
    const int s = (int)(row / BT), bb = (int)(row % BT);
    float m[kL][8], dm[kL][8];
#pragma unroll
    for (int l = 0; l < kL; ++l) {
      load8(mem + (((long)l * S + s) * BT + bb) * kC + lane * 8, m[l]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dm[l][j] = 0.f;
    }
    for (int n = 0; n < nq; ++n) {
      const long i = (long)n * BT + bb;
      const float4 w4 = *reinterpret_cast<const float4*>(lvlw + i * 4);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int l = 0; l < kL; ++l)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaf(w[l], m[l][j], z[j]);
      float d[8], dz[8];
      load8(dqm + (i * Sq + s) * kC + lane * 8, d);
      ln_bwd_row(z, d, gm, 1e-5f, dz, ag, ab);
#pragma unroll
      for (int l = 0; l < kL; ++l) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { dm[l][j] = fmaf(w[l], dz[j], dm[l][j]); a = fmaf(dz[j], m[l][j], a); }
        a = warp_sum(a);
        if (lane == 0) atomicAdd(dlvlw + i * 4 + l, a);
      }
    }
#pragma unroll
    for (int l = 0; l < kL; ++l) {
      float* p = dmem32 + (((long)l * S + s) * BT + bb) * kC + lane * 8;
      float o[8];
      load8(p, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += dm[l][j];
      store8(p, o);
    }
  }
  flush_gb(ag, ab, dg, db, red);
}

// ---- lvl_w = softmax(lvl_w_embed(x)) backward ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) lvlw_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ p, const float* __restrict__ dp, T* dx,
                                                            float beta, float* __restrict__ dW, float* __restrict__ dB,
                                                            long rows) {
  __shared__ float red[kWarps * 4 * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[4][8], abias[4] = {0, 0, 0, 0};
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[l][j] = 0.f;
  for (long row = (long)blockIdx.x * kWarps + warp; row < rows; row += (long)gridDim.x * kWarps) {
    const float4 p4 = *reinterpret_cast<const float4*>(p + row * 4), d4 = *reinterpret_cast<const float4*>(dp + row * 4);
    const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
    const float dot = pv[0] * dv[0] + pv[1] * dv[1] + pv[2] * dv[2] + pv[3] * dv[3];
    float xv[8], o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(x + row * kC + lane * 8, xv);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float dl = pv[l] * (dv[l] - dot);
      float wv[8];
      load8(w + l * kC + lane * 8, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = fmaf(dl, wv[j], o[j]); aw[l][j] = fmaf(dl, xv[j], aw[l][j]); }
      abias[l] += dl;
    }
    store_beta(dx + row * kC + lane * 8, o, beta);
  }
  flush_rows<4>(aw, dW, red);
  if (lane == 0)
    for (int l = 0; l < 4; ++l) atomicAdd(dB + l, abias[l]);
}

// ---- conv_norm(actor[i] + qm[i,s]) backward: block per actor instance ---------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) add_ln_pad_bwd_kernel(const T* __restrict__ actor, const T* __restrict__ qm,
                                                                  const float* __restrict__ g, const T* __restrict__ dxpad,
                                                                  T* dqm, float beta_qm, T* dactor, float beta_a,
                                                                  float* __restrict__ dg, float* __restrict__ db, int S,
                                                                  int Sq, int Sp) {
  __shared__ float red[kWarps][2][256];
  const long i = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0}, da[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float gm[8], a[8];
  load8(g + lane * 8, gm);
  load8(actor + i * kC + lane * 8, a);
  for (int s = warp; s < S; s += kWarps) {
    float z[8], d[8], dz[8];
    load8(qm + (i * Sq + s) * kC + lane * 8, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] += a[j];
    load8(dxpad + (i * Sp + s) * kC + lane * 8, d);
    ln_bwd_row(z, d, gm, 1e-5f, dz, ag, ab);
    store_beta(dqm + (i * Sq + s) * kC + lane * 8, dz, beta_qm);
#pragma unroll
    for (int j = 0; j < 8; ++j) da[j] += dz[j];
  }
  // actor gradient = sum over the S positions: reduce the warps' partials through shared memory
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][0][lane * 8 + j] = da[j];
  __syncthreads();
  {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += red[w][0][c];
    const float old = beta_a != 0.f ? beta_a * to_f(dactor[i * kC + c]) : 0.f;
    dactor[i * kC + c] = from_f<T>(old + s);
  }
  __syncthreads();
  flush_gb(ag, ab, dg, db, red);
}

// ---- query_sine_embed backward -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) qse_bwd_kernel(const float* __restrict__ ref, const T* __restrict__ scale,
                                                           const T* __restrict__ hidden, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const T* __restrict__ dqse,
                                                           T* dscale, float beta_s, T* dhidden, float beta_h,
                                                           float* __restrict__ dw1, float* __restrict__ db1,
                                                           float* __restrict__ dref, long rows) {
  __shared__ float red[kWarps * 2 * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[2][8], ab0 = 0.f, ab1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { aw[0][j] = 0.f; aw[1][j] = 0.f; }
  float wa[8], wb[8];
  load8(w1 + lane * 8, wa);
  load8(w1 + kC + lane * 8, wb);
  for (long row = (long)blockIdx.x * kWarps + warp; row < rows; row += (long)gridDim.x * kWarps) {
  float hdn[8];
  load8(hidden + row * kC + lane * 8, hdn);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0 = fmaf(hdn[j], wa[j], a0); a1 = fmaf(hdn[j], wb[j], a1); }
  a0 = warp_sum(a0) + b1[0];
  a1 = warp_sum(a1) + b1[1];
  a0 = 1.0f / (1.0f + expf(-a0));
  a1 = 1.0f / (1.0f + expf(-a1));
  const float4 r = *reinterpret_cast<const float4*>(ref + row * 4);
  const bool lo = lane < 16;
  const float mod = lo ? (a1 / r.w) : (a0 / r.z);
  const float coord = lo ? r.y : r.x;
  const float two_pi = 6.283185307179586f;
  float sc[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  if (scale) load8(scale + row * kC + lane * 8, sc);
  float dq[8];
  load8(dqse + row * kC + lane * 8, dq);
  float dsc[8], tsum = 0.f, csum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = (lane * 8 + j) & 127;
    const float dim_t = powf(10000.0f, (float)(2 * (d / 2)) / 128.0f);
    const float p = (coord * two_pi) / dim_t;
    float sn, cs;
    sincosf(p, &sn, &cs);
    const float e = (d & 1) ? cs : sn;
    const float dedp = (d & 1) ? -sn : cs;
    dsc[j] = dq[j] * e * mod;
    tsum = fmaf(dq[j] * e, sc[j], tsum);                               // d/d mod
    csum = fmaf(dq[j] * sc[j] * mod, dedp * two_pi / dim_t, csum);     // d/d coord
  }
  if (dscale) store_beta(dscale + row * kC + lane * 8, dsc, beta_s);
  const float t_lo = half_sum_lo(tsum, lane), t_hi = half_sum_hi(tsum, lane);
  const float c_lo = half_sum_lo(csum, lane), c_hi = half_sum_hi(csum, lane);
  const float da1 = t_lo / r.w, da0 = t_hi / r.z;
  const float dl0 = da0 * a0 * (1.f - a0), dl1 = da1 * a1 * (1.f - a1);
  float dh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dh[j] = dl0 * wa[j] + dl1 * wb[j];
    aw[0][j] = fmaf(dl0, hdn[j], aw[0][j]);
    aw[1][j] = fmaf(dl1, hdn[j], aw[1][j]);
  }
  store_beta(dhidden + row * kC + lane * 8, dh, beta_h);
  ab0 += dl0; ab1 += dl1;
  if (lane == 0) {
    if (dref) {
      dref[row * 4 + 0] += c_hi;                           // x
      dref[row * 4 + 1] += c_lo;                           // y
      dref[row * 4 + 2] += -t_hi * a0 / (r.z * r.z);       // w
      dref[row * 4 + 3] += -t_lo * a1 / (r.w * r.w);       // h
    }
  }
  }
  flush_rows<2>(aw, dw1, red);
  if (lane == 0) { atomicAdd(db1, ab0); atomicAdd(db1 + 1, ab1); }
}

template <typename T>
__global__ void __launch_bounds__(128) sine_embed_bwd_kernel(const float* __restrict__ ref, const T* __restrict__ de,
                                                             float* __restrict__ dref) {
  __shared__ float red[4][4];
  const long row = blockIdx.x;
  const int j = threadIdx.x;
  const float4 r = *reinterpret_cast<const float4*>(ref + row * 4);
  const float dim_t = powf(10000.0f, (float)(2 * (j / 2)) / 128.0f);
  const float two_pi = 6.283185307179586f;
  const float c[4] = {r.y, r.x, r.z, r.w};
  float part[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float p = (c[q] * two_pi) / dim_t;
    float sn, cs;
    sincosf(p, &sn, &cs);
    const float d = to_f(de[row * 512 + q * 128 + j]);
    part[q] = warp_sum(d * ((j & 1) ? -sn : cs) * two_pi / dim_t);
  }
  if ((j & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) red[j >> 5][q] = part[q];
  }
  __syncthreads();
  if (j < 4) {
    const float s = red[0][j] + red[1][j] + red[2][j] + red[3][j];
    const int idx = j == 0 ? 1 : (j == 1 ? 0 : j);   // embedding order (y,x,w,h) -> ref order (x,y,w,h)
    dref[row * 4 + idx] += s;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float inv_sigmoid(float x) {
  x = fminf(fmaxf(x, 0.f), 1.f);
  const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
  return logf(x1 / x2);
}
__device__ __forceinline__ float inv_sigmoid_grad(float x) {   // autograd of utils/misc.py:530-534 (clamps pass no gradient)
  if (x < 0.f || x > 1.f) return 0.f;
  return (x > 1e-5f ? 1.0f / x : 0.f) + ((1.f - x) > 1e-5f ? 1.0f / (1.f - x) : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) box_refine_bwd_kernel(const T* __restrict__ hidden, const float* __restrict__ w2,
                                                                  const float* __restrict__ b2, const float* __restrict__ ref,
                                                                  const float* __restrict__ dnew_perm, T* dhidden, float beta,
                                                                  float* __restrict__ dw2, float* __restrict__ db2,
                                                                  float* __restrict__ dref, long rows, int nq, int BT) {
  __shared__ float red[kWarps * 4 * 256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aw[4][8], abias[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[c][j] = 0.f;
  for (long row = (long)blockIdx.x * kWarps + warp; row < rows; row += (long)gridDim.x * kWarps) {
    float xv[8], dh[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(hidden + row * kC + lane * 8, xv);
    const int n = (int)(row / BT), bb = (int)(row % BT);
    const long prow = (long)bb * nq + n;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float wv[8];
      load8(w2 + c * kC + lane * 8, wv);
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(xv[j], wv[j], a);
      const float rc = ref[row * 4 + c];
      const float rn = sigmoidf_(warp_sum(a) + b2[c] + inv_sigmoid(rc));
      const float dt = dnew_perm[prow * 4 + c] * rn * (1.f - rn);
#pragma unroll
      for (int j = 0; j < 8; ++j) { dh[j] = fmaf(dt, wv[j], dh[j]); aw[c][j] = fmaf(dt, xv[j], aw[c][j]); }
      abias[c] += dt;
      if (lane == 0 && dref) dref[row * 4 + c] += dt * inv_sigmoid_grad(rc);
    }
    store_beta(dhidden + row * kC + lane * 8, dh, beta);
  }
  flush_rows<4>(aw, dw2, red);
  if (lane == 0)
    for (int c = 0; c < 4; ++c) atomicAdd(db2 + c, abias[c]);
}

__global__ void sigmoid4_bwd_kernel(const float* __restrict__ r, const float* __restrict__ dr, const float* __restrict__ dperm,
                                    float* __restrict__ dref_u, long rows, int nq, int BT) {
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int n = (int)(row / BT), bb = (int)(row % BT);
  const long prow = (long)bb * nq + n;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float rv = r[row * 4 + c];
    const float d = (dr ? dr[row * 4 + c] : 0.f) + (dperm ? dperm[prow * 4 + c] : 0.f);
    dref_u[row * 4 + c] += d * rv * (1.f - rv);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) broadcast_rows_bwd_kernel(const T* __restrict__ dout, T* dsrc, float beta,
                                                                      long n_inst, int K) {
  const int k = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long i = 0; i < n_inst; ++i) {
    float v[8];
    load8(dout + (i * K + k) * kC + lane * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j];
  }
  store_beta(dsrc + (long)k * kC + lane * 8, s, beta);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) zero_pad_rows_kernel(T* x, long rows, int S, int Sp) {
  const long row = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);   // over n_img * (Sp - S)
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int np = Sp - S;
  const long i = row / np;
  const int s = S + (int)(row % np);
  const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  store8(x + (i * Sp + s) * kC + lane * 8, z);
}

// ---- elementwise -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
constexpr int EW_UNROLL = 4;   // 16-byte vectors in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(T* dH, const T* __restrict__ ref, int act, long n8) {
  const long base = ((long)blockIdx.x * blockDim.x) * EW_UNROLL + threadIdx.x;
  float d[EW_UNROLL][8], r[EW_UNROLL][8];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const long i = base + (long)u * blockDim.x;
    if (i < n8) { load8(dH + i * 8, d[u]); load8(ref + i * 8, r[u]); }
  }
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const long i = base + (long)u * blockDim.x;
    if (i >= n8) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) d[u][j] = act == CQVAD_ACT_RELU ? (r[u][j] > 0.f ? d[u][j] : 0.f) : d[u][j] * (DT<T>::id == CQVAD_BF16 ? tc::gelu_grad_fast(r[u][j]) : gelu_grad(r[u][j]));
    store8(dH + i * 8, d[u]);
  }
}
// out = gelu(pre) and (training) dact = gelu'(pre): the derivative is stored next to the activation so that the backward
// needs no pass of its own -- the data-gradient GEMM multiplies by it in its epilogue (Epilogue::mul_mode 3)
template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ pre, T* __restrict__ out, T* __restrict__ dact, long n8) {
  const long base = ((long)blockIdx.x * blockDim.x) * EW_UNROLL + threadIdx.x;
  float v[EW_UNROLL][8];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const long i = base + (long)u * blockDim.x;
    if (i < n8) load8(pre + i * 8, v[u]);
  }
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const long i = base + (long)u * blockDim.x;
    if (i >= n8) continue;
    float a[8], d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = DT<T>::id == CQVAD_BF16 ? tc::gelu_fast(v[u][j]) : gelu_erf(v[u][j]);
      d[j] = DT<T>::id == CQVAD_BF16 ? tc::gelu_grad_fast(v[u][j]) : gelu_grad(v[u][j]);
    }
    store8(out + i * 8, a);
    if (dact) store8(dact + i * 8, d);
  }
}
template <typename D, typename S>
__global__ void axpby_kernel(D* dst, const S* __restrict__ src, float beta, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(src + i * 8, v);
  if (beta != 0.f) {
    float o[8];
    load8(dst + i * 8, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(beta, o[j], v[j]);
  }
  store8(dst + i * 8, v);
}

// ---- weight transposes ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ W, T* __restrict__ Wt, int out, int in) {
  __shared__ T tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (x < in && y0 + j < out) tile[j][threadIdx.x] = W[(long)(y0 + j) * in + x];
  __syncthreads();
  const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (ox < out && oy0 + j < in) Wt[(long)(oy0 + j) * out + ox] = tile[threadIdx.x][j];
}
template <typename T>
__global__ void conv_w_flip_kernel(const T* __restrict__ W, T* __restrict__ Wd) {
  // Wd[ci][8-tap][co] = W[co][tap][ci]; one thread per destination element (co fastest: coalesced writes)
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 256L * 9 * 256) return;
  const int co = (int)(idx & 255), t = (int)((idx >> 8) % 9), ci = (int)(idx / (9 * 256));
  Wd[idx] = W[((long)co * 9 + (8 - t)) * 256 + ci];
}

// ---- CUDA-core weight gradient: dW[n,k] += sum_m dY[m,n] X[m,k] (split over m, fp32 atomics) ---------------------------
constexpr int WG_BN = 64, WG_BK = 64, WG_BM = 16;
template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T, bool CONV>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const T* __restrict__ dY, long lddy, const T* __restrict__ X, long ldx,
                                                         float* __restrict__ dW, long ldw, float* __restrict__ db, long M,
                                                         int Nout, int Kin, int cw, long rows_per_split) {
  __shared__ float Ys[WG_BM][WG_BN + 4];
  __shared__ float Xs[WG_BM][WG_BK + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * WG_BN, k0 = blockIdx.x * WG_BK;
  const int tap = CONV ? (int)(blockIdx.z % 9) : 0;
  const long split = CONV ? blockIdx.z / 9 : blockIdx.z;
  const long m_begin = split * rows_per_split;
  const long m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  float acc[4][4], accb[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long m0 = m_begin; m0 < m_end; m0 += WG_BM) {
    const long m = m0 + lr;
    float yv[4] = {0, 0, 0, 0}, xv[4] = {0, 0, 0, 0};
    if (m < m_end) {
      if (n0 + lc < Nout) ld4<T>(dY + m * lddy + n0 + lc, yv);
      long src = m;
      bool ok = true;
      if (CONV) {
        src = m + (long)dy * cw + dx;
        const int xx = (int)(m % cw) + dx;
        ok = src >= 0 && src < M && xx >= 0 && xx < cw;
      }
      if (ok && k0 + lc < Kin) ld4<T>(X + src * ldx + k0 + lc, xv);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { Ys[lr][lc + i] = yv[i]; Xs[lr][lc + i] = xv[i]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_BM; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&Ys[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        accb[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= Nout) continue;
    if (dW) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + tx * 4 + j;
        if (k < Kin) atomicAdd(dW + (long)n * ldw + (long)tap * 256 + k, acc[i][j]);
      }
    }
    if (db && blockIdx.x == 0 && tx == 0 && tap == 0) atomicAdd(db + n, accb[i]);
  }
}

}  // namespace

// ---- host wrappers ---------------------------------------------------------------------------------------------------
template <typename T>
int wgrad_simt(const T* dY, long lddy, const T* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
               const ConvGeom* conv, cudaStream_t st) {
  if (M == 0) return 0;
  CQ_CHECK_SHAPE(Nout % 4 == 0 && Kin % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0, "wgrad: extents must be multiples of 4");
  const int gx = (int)cdiv(Kin, WG_BK), gy = (int)cdiv(Nout, WG_BN), taps = conv ? 9 : 1;
  long splits = (148L * 6) / ((long)gx * gy * taps);
  const long max_splits = cdiv(M, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long rps = cdiv(M, splits);
  rps = cdiv(rps, WG_BM) * WG_BM;
  splits = cdiv(M, rps);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(splits * taps));
  if (conv) {
    CQ_CHECK_SHAPE(Kin == kC, "conv wgrad: 256 input channels");
    wgrad_simt_kernel<T, true><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, dW, ldw, db, M, Nout, Kin, conv->w, rps);
  } else {
    wgrad_simt_kernel<T, false><<<grid, 256, 0, st>>>(dY, lddy, X, ldx, dW, ldw, db, M, Nout, Kin, 0, rps);
  }
  CQ_LAUNCH_CHECK();
  return 0;
}
template int wgrad_simt<float>(const float*, long, const float*, long, float*, long, float*, long, int, int, const ConvGeom*, cudaStream_t);
template int wgrad_simt<bf16>(const bf16*, long, const bf16*, long, float*, long, float*, long, int, int, const ConvGeom*, cudaStream_t);

template <>
int wgrad<float>(const float* dY, long lddy, const float* X, long ldx, float* dW, long ldw, float* db, long M, int Nout,
                 int Kin, const ConvGeom* conv, cudaStream_t st) {
  return wgrad_simt<float>(dY, lddy, X, ldx, dW, ldw, db, M, Nout, Kin, conv, st);
}
template <>
int wgrad<bf16>(const bf16* dY, long lddy, const bf16* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
                const ConvGeom* conv, cudaStream_t st) {
  if (!force_simt()) {
    int r = wgrad_tc(dY, lddy, X, ldx, dW, ldw, db, M, Nout, Kin, conv, st);
    if (r <= 0) return r;
  }
  return wgrad_simt<bf16>(dY, lddy, X, ldx, dW, ldw, db, M, Nout, Kin, conv, st);
}

template <typename T>
int transpose_w(const T* W, T* Wt, int out, int in, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(in, 32), (unsigned)cdiv(out, 32)), block(32, 8);
  transpose_kernel<T><<<grid, block, 0, st>>>(W, Wt, out, in);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int transpose_w<float>(const float*, float*, int, int, cudaStream_t);
template int transpose_w<bf16>(const bf16*, bf16*, int, int, cudaStream_t);

// All transposed weight copies of a backward in one launch: block -> (job, 32 x 32 tile) through the tile prefix table.
struct TransposeBatch {
  TransposeJob job[kMaxTransposeJobs];
  int tile_start[kMaxTransposeJobs + 1];
  int n;
};
template <typename T>
__global__ void __launch_bounds__(256) transpose_batch_kernel(const __grid_constant__ TransposeBatch b) {
  __shared__ T tile[32][33];
  int lo = 0, hi = b.n;                       // last job whose first tile is <= blockIdx.x
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.tile_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
  const TransposeJob& j = b.job[lo];
  const int t = blockIdx.x - b.tile_start[lo];
  const int tx_n = (j.in + 31) >> 5;
  const int bx = t % tx_n, by = t / tx_n;
  const T* W = (const T*)j.W; T* Wt = (T*)j.Wt;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = bx * 32 + tx, y0 = by * 32;
  for (int r = ty; r < 32; r += 8)
    if (x < j.in && y0 + r < j.out) tile[r][tx] = W[(long)(y0 + r) * j.in + x];
  __syncthreads();
  const int ox = by * 32 + tx, oy0 = bx * 32;
  for (int r = ty; r < 32; r += 8)
    if (ox < j.out && oy0 + r < j.in) Wt[(long)(oy0 + r) * j.out + ox] = tile[tx][r];
}
template <typename T>
int transpose_w_batch(const TransposeJob* jobs, int n, cudaStream_t st) {
  for (int base = 0; base < n; base += kMaxTransposeJobs) {
    TransposeBatch b;
    b.n = std::min(kMaxTransposeJobs, n - base);
    int tiles = 0;
    for (int i = 0; i < b.n; ++i) {
      b.job[i] = jobs[base + i];
      b.tile_start[i] = tiles;
      tiles += ((b.job[i].in + 31) / 32) * ((b.job[i].out + 31) / 32);
    }
    b.tile_start[b.n] = tiles;
    if (tiles == 0) continue;
    transpose_batch_kernel<T><<<(unsigned)tiles, 256, 0, st>>>(b);
    CQ_LAUNCH_CHECK();
  }
  return 0;
}
template int transpose_w_batch<float>(const TransposeJob*, int, cudaStream_t);
template int transpose_w_batch<bf16>(const TransposeJob*, int, cudaStream_t);

template <typename T>
int conv_w_flip(const T* W, T* Wd, cudaStream_t st) {
  conv_w_flip_kernel<T><<<(unsigned)cdiv(256L * 9 * 256, 256), 256, 0, st>>>(W, Wd);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int conv_w_flip<float>(const float*, float*, cudaStream_t);
template int conv_w_flip<bf16>(const bf16*, bf16*, cudaStream_t);

template <typename T>
int act_bwd(T* dH, const T* ref, int act, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "act_bwd: element count must be a multiple of 8");
  if (n == 0) return 0;
  act_bwd_kernel<T><<<(unsigned)cdiv(n / 8, 256 * EW_UNROLL), 256, 0, st>>>(dH, ref, act, n / 8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int act_bwd<float>(float*, const float*, int, long, cudaStream_t);
template int act_bwd<bf16>(bf16*, const bf16*, int, long, cudaStream_t);

template <typename T>
int gelu_fwd(const T* pre, T* out, T* dact, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "gelu: element count must be a multiple of 8");
  if (n == 0) return 0;
  gelu_fwd_kernel<T><<<(unsigned)cdiv(n / 8, 256 * EW_UNROLL), 256, 0, st>>>(pre, out, dact, n / 8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int gelu_fwd<float>(const float*, float*, float*, long, cudaStream_t);
template int gelu_fwd<bf16>(const bf16*, bf16*, bf16*, long, cudaStream_t);

template <typename T>
int axpby(T* dst, const T* src, float beta, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "axpby: element count must be a multiple of 8");
  if (n == 0) return 0;
  axpby_kernel<T, T><<<(unsigned)cdiv(n / 8, 256), 256, 0, st>>>(dst, src, beta, n / 8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int axpby<float>(float*, const float*, float, long, cudaStream_t);
template int axpby<bf16>(bf16*, const bf16*, float, long, cudaStream_t);

template <typename T>
int acc_to_f32(float* dst32, const T* src, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "acc_to_f32: element count must be a multiple of 8");
  if (n == 0) return 0;
  axpby_kernel<float, T><<<(unsigned)cdiv(n / 8, 256), 256, 0, st>>>(dst32, src, 1.0f, n / 8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int acc_to_f32<float>(float*, const float*, long, cudaStream_t);
template int acc_to_f32<bf16>(float*, const bf16*, long, cudaStream_t);

template <typename T>
int f32_to_t(const float* src, T* dst, float beta, long n, cudaStream_t st) {
  CQ_CHECK_SHAPE(n % 8 == 0, "f32_to_t: element count must be a multiple of 8");
  if (n == 0) return 0;
  axpby_kernel<T, float><<<(unsigned)cdiv(n / 8, 256), 256, 0, st>>>(dst, src, beta, n / 8);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int f32_to_t<float>(const float*, float*, float, long, cudaStream_t);
template int f32_to_t<bf16>(const float*, bf16*, float, long, cudaStream_t);

template <typename T>
int ln_bwd(const T* x, const T* res, const float* g, float eps, const void* dy, bool dy_f32, int perm_nq, int perm_BT,
           int perm_K, T* dx, float beta_x, T* dres, float beta_r, float* dg, float* db, long rows, cudaStream_t st) {
  if (rows == 0) return 0;
  if (dy_f32)
    ln_bwd_kernel<T, float><<<persist_grid(rows), kThreads, 0, st>>>(x, res, g, eps, (const float*)dy, perm_nq, perm_BT, perm_K,
                                                                     dx, beta_x, dres, beta_r, dg, db, rows);
  else
    ln_bwd_kernel<T, T><<<persist_grid(rows), kThreads, 0, st>>>(x, res, g, eps, (const T*)dy, perm_nq, perm_BT, perm_K, dx,
                                                                 beta_x, dres, beta_r, dg, db, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int ln_bwd<float>(const float*, const float*, const float*, float, const void*, bool, int, int, int, float*, float, float*, float, float*, float*, long, cudaStream_t);
template int ln_bwd<bf16>(const bf16*, const bf16*, const float*, float, const void*, bool, int, int, int, bf16*, float, bf16*, float, float*, float*, long, cudaStream_t);

template <typename T>
int lvlmix_ln_bwd(const T* mem, const float* lvlw, const float* g, const T* dqm, float* dmem32, float* dlvlw, float* dg,
                  float* db, long N, int nq, int S, int Sq, int BT, cudaStream_t st) {
  if (N == 0) return 0;
  lvlmix_ln_bwd_kernel<T><<<persist_grid((long)S * BT), kThreads, 0, st>>>(mem, lvlw, g, dqm, dmem32, dlvlw, dg, db, nq, S, Sq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int lvlmix_ln_bwd<float>(const float*, const float*, const float*, const float*, float*, float*, float*, float*, long, int, int, int, int, cudaStream_t);
template int lvlmix_ln_bwd<bf16>(const bf16*, const float*, const float*, const bf16*, float*, float*, float*, float*, long, int, int, int, int, cudaStream_t);

template <typename T>
int lvlw_bwd(const T* x, const float* w, const float* p, const float* dp, T* dx, float beta, float* dW, float* dB, long rows,
             cudaStream_t st) {
  if (rows == 0) return 0;
  lvlw_bwd_kernel<T><<<small_grid(rows), kThreads, 0, st>>>(x, w, p, dp, dx, beta, dW, dB, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int lvlw_bwd<float>(const float*, const float*, const float*, const float*, float*, float, float*, float*, long, cudaStream_t);
template int lvlw_bwd<bf16>(const bf16*, const float*, const float*, const float*, bf16*, float, float*, float*, long, cudaStream_t);

template <typename T>
int add_ln_pad_bwd(const T* actor, const T* qm, const float* g, const T* dxpad, T* dqm, float beta_qm, T* dactor,
                   float beta_a, float* dg, float* db, long N, int S, int Sq, int Sp, cudaStream_t st) {
  if (N == 0) return 0;
  add_ln_pad_bwd_kernel<T><<<(unsigned)N, kThreads, 0, st>>>(actor, qm, g, dxpad, dqm, beta_qm, dactor, beta_a, dg, db, S, Sq, Sp);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int add_ln_pad_bwd<float>(const float*, const float*, const float*, const float*, float*, float, float*, float, float*, float*, long, int, int, int, cudaStream_t);
template int add_ln_pad_bwd<bf16>(const bf16*, const bf16*, const float*, const bf16*, bf16*, float, bf16*, float, float*, float*, long, int, int, int, cudaStream_t);

template <typename T>
int qse_bwd(const float* ref, const T* scale, const T* hidden, const float* w1, const float* b1, const T* dqse, T* dscale,
            float beta_s, T* dhidden, float beta_h, float* dw1, float* db1, float* dref, long rows, cudaStream_t st) {
  if (rows == 0) return 0;
  qse_bwd_kernel<T><<<small_grid(rows), kThreads, 0, st>>>(ref, scale, hidden, w1, b1, dqse, dscale, beta_s, dhidden, beta_h, dw1,
                                                         db1, dref, rows);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int qse_bwd<float>(const float*, const float*, const float*, const float*, const float*, const float*, float*, float, float*, float, float*, float*, float*, long, cudaStream_t);
template int qse_bwd<bf16>(const float*, const bf16*, const bf16*, const float*, const float*, const bf16*, bf16*, float, bf16*, float, float*, float*, float*, long, cudaStream_t);

template <typename T>
int sine_embed_bwd(const float* ref, const T* de, float* dref, long rows, cudaStream_t st) {
  if (rows == 0) return 0;
  sine_embed_bwd_kernel<T><<<(unsigned)rows, 128, 0, st>>>(ref, de, dref);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int sine_embed_bwd<float>(const float*, const float*, float*, long, cudaStream_t);
template int sine_embed_bwd<bf16>(const float*, const bf16*, float*, long, cudaStream_t);

template <typename T>
int box_refine_bwd(const T* hidden, const float* w2, const float* b2, const float* ref, const float* dnew_perm, T* dhidden,
                   float beta, float* dw2, float* db2, float* dref, long rows, int nq, int BT, cudaStream_t st) {
  if (rows == 0) return 0;
  box_refine_bwd_kernel<T><<<small_grid(rows), kThreads, 0, st>>>(hidden, w2, b2, ref, dnew_perm, dhidden, beta, dw2, db2, dref,
                                                                rows, nq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int box_refine_bwd<float>(const float*, const float*, const float*, const float*, const float*, float*, float, float*, float*, float*, long, int, int, cudaStream_t);
template int box_refine_bwd<bf16>(const bf16*, const float*, const float*, const float*, const float*, bf16*, float, float*, float*, float*, long, int, int, cudaStream_t);

int sigmoid4_bwd(const float* r, const float* dr, const float* dperm, float* dref_u, long rows, int nq, int BT, cudaStream_t st) {
  if (rows == 0) return 0;
  sigmoid4_bwd_kernel<<<(unsigned)cdiv(rows, 128), 128, 0, st>>>(r, dr, dperm, dref_u, rows, nq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int broadcast_rows_bwd(const T* dout, T* dsrc, float beta, long rows, int K, cudaStream_t st) {
  if (rows == 0) return 0;
  broadcast_rows_bwd_kernel<T><<<(unsigned)cdiv(K, kWarps), kThreads, 0, st>>>(dout, dsrc, beta, rows / K, K);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int broadcast_rows_bwd<float>(const float*, float*, float, long, int, cudaStream_t);
template int broadcast_rows_bwd<bf16>(const bf16*, bf16*, float, long, int, cudaStream_t);

template <typename T>
int zero_pad_rows(T* x, long n_img, int S, int Sp, cudaStream_t st) {
  const long rows = n_img * (Sp - S);
  if (rows == 0) return 0;
  zero_pad_rows_kernel<T><<<row_grid(rows), kThreads, 0, st>>>(x, rows, S, Sp);
  CQ_LAUNCH_CHECK();
  return 0;
}
template int zero_pad_rows<float>(float*, long, int, int, cudaStream_t);
template int zero_pad_rows<bf16>(bf16*, long, int, int, cudaStream_t);

}  // namespace cqvad
