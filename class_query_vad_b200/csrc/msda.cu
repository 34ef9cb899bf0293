// 3-D multi-scale deformable attention sampling for sm_100a (HBM/L2-bound gather).
// Replaces ms_deformable_im2col_gpu_kernel / ms_deformable_col2im_* of the reference
// (ops/src/cuda/ms_deform_im2col_cuda_t.cuh:374-439, 441-549).
//
// Design (not a port): the reference runs one thread per output scalar, so every channel of a head redoes the same
// floor/weight/offset arithmetic for all L*P points and issues 8 scalar loads per point.  Here one WARP owns one
// (batch, query, head): lane p first computes the integer corner offsets, validity bits and trilinear weights of
// sampling point p (L*P points spread over the lanes, coalesced 12-byte/4-byte loads of loc/attn), then the warp
// walks the points, broadcasting them by shuffle, and each lane gathers its channel(s): a corner read is one
// coalesced D*sizeof(T) segment.  Index arithmetic follows the kernel contract bit-exactly:
//   x_im = fl32(loc * dim - 0.5)   (__fmaf_rn: ONE rounding.  The reference source reads `loc * dim - 0.5` with a double
//   literal, cuh:424-426, but nvcc narrows and contracts it: the compiled reference kernel executes FFMA R, loc, dim, -0.5 --
//   verified in its SASS and by the one-hot index probe of tests/test_msda_ref_gpu.py, which the round-1 "product rounded
//   first" reading failed on locations within half an ulp of a voxel centre)
//   low = (int)floorf(x_im); point used iff -1 < x_im < dim on all axes (cuh:428); corner validity cuh:63-109.
#include "common.cuh"
#include <stdlib.h>

namespace cqvad {

namespace {

__device__ __forceinline__ void point_geometry(float loc_x, float loc_y, float loc_t, int T, int H, int W, int& tl,
                                               int& hl, int& wl, unsigned& mask, float& lt, float& lh, float& lw) {
  const float t_im = __fmaf_rn(loc_t, (float)T, -0.5f);
  const float h_im = __fmaf_rn(loc_y, (float)H, -0.5f);
  const float w_im = __fmaf_rn(loc_x, (float)W, -0.5f);
  tl = (int)floorf(t_im); hl = (int)floorf(h_im); wl = (int)floorf(w_im);
  lt = t_im - (float)tl; lh = h_im - (float)hl; lw = w_im - (float)wl;
  const bool inside = t_im > -1.f && h_im > -1.f && w_im > -1.f && t_im < (float)T && h_im < (float)H && w_im < (float)W;
  mask = 0;
  if (inside) {
    const bool t0 = tl >= 0, t1 = tl + 1 <= T - 1, h0 = hl >= 0, h1 = hl + 1 <= H - 1, w0 = wl >= 0, w1 = wl + 1 <= W - 1;
    mask = (unsigned)(t0 && h0 && w0) | ((unsigned)(t0 && h0 && w1) << 1) | ((unsigned)(t0 && h1 && w0) << 2) |
           ((unsigned)(t0 && h1 && w1) << 3) | ((unsigned)(t1 && h0 && w0) << 4) | ((unsigned)(t1 && h0 && w1) << 5) |
           ((unsigned)(t1 && h1 && w0) << 6) | ((unsigned)(t1 && h1 && w1) << 7);
  }
}

__global__ void msda_indices_kernel(const int64_t* __restrict__ shapes, const float* __restrict__ loc,
                                    int32_t* __restrict__ tlo, int32_t* __restrict__ hlo, int32_t* __restrict__ wlo,
                                    uint8_t* __restrict__ cmask, long total, int L, int P) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)((idx / P) % L);
  const int T = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
  int tl, hl, wl; unsigned m; float a, b, c;
  point_geometry(loc[idx * 3], loc[idx * 3 + 1], loc[idx * 3 + 2], T, H, W, tl, hl, wl, m, a, b, c);
  tlo[idx] = tl; hlo[idx] = hl; wlo[idx] = wl; cmask[idx] = (uint8_t)m;
}

constexpr int kWarps = 8;

// Forward.  Warp per (b,q,m); D <= 32*CPL channels per head, lane owns channels lane + 32*j.
template <typename T, int CPL>
__global__ void __launch_bounds__(kWarps * 32) msda_fwd_kernel(const T* __restrict__ value,
                                                               const int64_t* __restrict__ shapes,
                                                               const int64_t* __restrict__ lsi,
                                                               const float* __restrict__ loc,
                                                               const float* __restrict__ attn, T* __restrict__ out,
                                                               long n_warps_total, int Len, int M, int D, int L, int Lq,
                                                               int P) {
  const long wid = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_warps_total) return;
  const int m = (int)(wid % M);
  const long bq = wid / M;
  const int b = (int)(bq / Lq);
  const int LP = L * P;
  const long row_stride = (long)M * D;
  const T* vbase = value + (long)b * Len * row_stride + (long)m * D;
  const float* locp = loc + wid * LP * 3;
  const float* attp = attn + wid * LP;
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = 0.f;

  for (int p0 = 0; p0 < LP; p0 += 32) {
    // phase 1: lane -> point p0+lane
    const int pt = p0 + lane;
    int g_base = 0, g_hs = 0, g_ts = 0; unsigned g_mask = 0; float g_lt = 0, g_lh = 0, g_lw = 0, g_a = 0;
    if (pt < LP) {
      const int l = pt / P;
      const int Tt = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
      int tl, hl, wl;
      point_geometry(locp[pt * 3], locp[pt * 3 + 1], locp[pt * 3 + 2], Tt, H, W, tl, hl, wl, g_mask, g_lt, g_lh, g_lw);
      g_hs = W; g_ts = H * W;
      g_base = (int)lsi[l] + (tl * H + hl) * W + wl;
      g_a = attp[pt];
    }
    // phase 2: walk the points
    const int np = min(32, LP - p0);
    for (int j = 0; j < np; ++j) {
      const unsigned mk = __shfl_sync(0xffffffffu, g_mask, j);
      if (mk == 0) continue;  // warp-uniform
      const int base = __shfl_sync(0xffffffffu, g_base, j);
      const int hs = __shfl_sync(0xffffffffu, g_hs, j), ts = __shfl_sync(0xffffffffu, g_ts, j);
      const float lt = __shfl_sync(0xffffffffu, g_lt, j), lh = __shfl_sync(0xffffffffu, g_lh, j);
      const float lw = __shfl_sync(0xffffffffu, g_lw, j), a = __shfl_sync(0xffffffffu, g_a, j);
      const float ht = 1.f - lt, hh = 1.f - lh, hw = 1.f - lw;
      const float wgt[8] = {ht * hh * hw, ht * hh * lw, ht * lh * hw, ht * lh * lw,
                            lt * hh * hw, lt * hh * lw, lt * lh * hw, lt * lh * lw};
      const int off[8] = {0, 1, hs, hs + 1, ts, ts + 1, ts + hs, ts + hs + 1};
#pragma unroll
      for (int cj = 0; cj < CPL; ++cj) {
        const int c = lane + 32 * cj;
        if (c < D) {
          float val = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (mk & (1u << k)) val = fmaf(wgt[k], to_f(vbase[(long)(base + off[k]) * row_stride + c]), val);
          acc[cj] = fmaf(val, a, acc[cj]);
        }
      }
    }
  }
#pragma unroll
  for (int cj = 0; cj < CPL; ++cj) {
    const int c = lane + 32 * cj;
    if (c < D) out[wid * D + c] = from_f<T>(acc[cj]);
  }
}

// Forward, channel-vectorised (D % 8 == 0, D/8 a power of two <= 32: every shipped configuration has D = 32).
// Same warp per (b,q,m) and the same phase 1 (lane -> point geometry, bit-exact index contract), but phase 2 walks
// G = 32 / (D/8) points at a time: lane = (g, ch) takes point j0+g and channels [8*ch, 8*ch+8) -- one 16-byte (bf16) load per
// corner instead of one 2-byte load per lane, i.e. 32 wide loads per warp where the scalar kernel issued 256 narrow ones --
// and the G partial sums are folded with xor-shuffles at the end.  ncu launch list of the encoder layer (B = 4, ViT-B/224
// pyramid): the scalar kernel took 6.5 ms of an 8.0 ms layer (17.5 GB of 64-byte gathers through L2, LSU-issue bound).
// FUSED = true (the encoder layer): `loc` holds the raw sampling OFFSETS and `attn` the raw LOGITS
// (fp32 outputs of the two query projections), `ref` the reference points [B*Lq, L, 3]; the softmax over the L*P logits and
// loc = ref + off / (T_l, W_l, H_l) (ops/modules/ms_deform_attn.py:187-192, reference normaliser order) happen in phase 1, so
// the 102 MB/clip sampling_locations and the attention weights never exist in memory.
template <typename T, bool FUSED>
__global__ void __launch_bounds__(kWarps * 32) msda_fwd_vec_kernel(const T* __restrict__ value,
                                                                   const int64_t* __restrict__ shapes,
                                                                   const int64_t* __restrict__ lsi,
                                                                   const float* __restrict__ loc,
                                                                   const float* __restrict__ attn,
                                                                   const float* __restrict__ ref, T* __restrict__ out,
                                                                   long n_warps_total, int Len, int M, int D, int L, int Lq,
                                                                   int P) {
  const long wid = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_warps_total) return;
  const int m = (int)(wid % M);
  const long bq = wid / M;
  const int b = (int)(bq / Lq);
  const int LP = L * P;
  const long row_stride = (long)M * D;
  const int lpp = D >> 3;                 // lanes per point
  const int G = 32 / lpp;                 // points per step
  const int g = lane / lpp, ch = lane % lpp;
  const T* vbase = value + (long)b * Len * row_stride + (long)m * D + ch * 8;
  const float* locp = loc + wid * LP * 3;
  const float* attp = attn + wid * LP;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float sm_max = 0.f, sm_sum = 1.f;
  if constexpr (FUSED) {   // softmax statistics over the head's L*P logits (same arithmetic as msda_prepare_kernel)
    sm_max = -INFINITY;
    for (int j = lane; j < LP; j += 32) sm_max = fmaxf(sm_max, attp[j]);
    sm_max = warp_max(sm_max);
    sm_sum = 0.f;
    for (int j = lane; j < LP; j += 32) sm_sum += expf(attp[j] - sm_max);
    sm_sum = warp_sum(sm_sum);
  }

  for (int p0 = 0; p0 < LP; p0 += 32) {
    const int pt = p0 + lane;
    int g_base = 0, g_hs = 0, g_ts = 0; unsigned g_mask = 0; float g_lt = 0, g_lh = 0, g_lw = 0, g_a = 0;
    if (pt < LP) {
      const int l = pt / P;
      const int Tt = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
      int tl, hl, wl;
      float lx = locp[pt * 3], ly = locp[pt * 3 + 1], lz = locp[pt * 3 + 2];
      if constexpr (FUSED) {
        const float* r = ref + (bq * L + l) * 3;
        lx = r[0] + __fdiv_rn(lx, (float)Tt); ly = r[1] + __fdiv_rn(ly, (float)W); lz = r[2] + __fdiv_rn(lz, (float)H);
      }
      point_geometry(lx, ly, lz, Tt, H, W, tl, hl, wl, g_mask, g_lt, g_lh, g_lw);
      g_hs = W; g_ts = H * W;
      g_base = (int)lsi[l] + (tl * H + hl) * W + wl;
      g_a = FUSED ? expf(attp[pt] - sm_max) / sm_sum : attp[pt];
    }
    const int np = min(32, LP - p0);
    for (int j0 = 0; j0 < np; j0 += G) {
      const int src = min(j0 + g, 31);
      unsigned mk = __shfl_sync(0xffffffffu, g_mask, src);
      const int base = __shfl_sync(0xffffffffu, g_base, src);
      const int hs = __shfl_sync(0xffffffffu, g_hs, src), ts = __shfl_sync(0xffffffffu, g_ts, src);
      const float lt = __shfl_sync(0xffffffffu, g_lt, src), lh = __shfl_sync(0xffffffffu, g_lh, src);
      const float lw = __shfl_sync(0xffffffffu, g_lw, src), a = __shfl_sync(0xffffffffu, g_a, src);
      if (j0 + g >= np) mk = 0;
      if (mk == 0) continue;
      const float ht = 1.f - lt, hh = 1.f - lh, hw = 1.f - lw;
      const float wgt[8] = {ht * hh * hw, ht * hh * lw, ht * lh * hw, ht * lh * lw,
                            lt * hh * hw, lt * hh * lw, lt * lh * hw, lt * lh * lw};
      const int off[8] = {0, 1, hs, hs + 1, ts, ts + 1, ts + hs, ts + hs + 1};
      float val[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (mk & (1u << k)) {
          float v[8];
          load8(vbase + (long)(base + off[k]) * row_stride, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) val[e] = fmaf(wgt[k], v[e], val[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(val[e], a, acc[e]);
    }
  }
  for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  }
  if (g == 0) store8(out + wid * D + ch * 8, acc);
}


// ---- Forward, query-tile kernel (round 2) -----------------------------------------------------------------------------
// ncu on msda_fwd_vec_kernel (B = 4, ViT-B/224 pyramid): 1.89 G warp instructions (1 770 per (b,q,m)), sm__throughput 71 %,
// lts__throughput 22 %, 10.7 GB of L2->L1 sectors for 0.68 GB of algorithmic bytes: issue-bound first, L1-miss traffic second.
// This kernel attacks both:
//  * a CTA owns TQ consecutive queries of ONE head (the vec kernel gave a CTA the 8 heads of one query, which share no value
//    bytes): neighbouring queries sample overlapping voxels of the same head, so the 8 warps -- and the CTA's later
//    iterations -- hit in L1 what the first toucher brought in;
//  * bf16 values are consumed by the mixed-precision FMA `fma.rn.f32.bf16` (SASS FHFMA.BF16: bf16 x bf16 + f32 -> f32, the
//    product is exact, one fp32 rounding per accumulate), straight from the packed registers of the 16-byte load: the
//    8 bf16->f32 conversions per corner are gone.  The trilinear corner weight is rounded to bf16 for that product (relative
//    2^-9 per term, unbiased; the attention weight stays fp32 and multiplies the per-point sum in fp32); fp32 values keep
//    fp32 weights;
//  * invalid corners are predicated loads into zeroed registers instead of branches: one straight-line body per step.
// Index contract (point_geometry) unchanged.
constexpr int kTileQ = 64;   // queries per CTA (8 per warp)

__device__ __forceinline__ float fma_bf16(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 ldg16_or_zero(const void* p, bool valid) {
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (valid) r = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}

constexpr int kMaxLevels = 16;

// 32-byte load (SASS LDG.E.ENL2.256): 16 bf16 channels per lane
struct U8 { unsigned v[8]; };
__device__ __forceinline__ U8 ldg32(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

// one step of the tile kernel: this lane's point (already broadcast) against its channel chunk (NW 32-bit words per corner:
// 4 = 8 bf16, 8 = 16 bf16 or 8 fp32).  ALLVALID: every corner of every lane's point is inside the volume (warp-uniform fast
// path, no predicates); otherwise invalid corners are predicated loads into zeroed registers.
template <typename T, int NW, bool ALLVALID>
__device__ __forceinline__ void tile_step(const char* __restrict__ ubase, unsigned b0, unsigned dw, unsigned dh, unsigned dt,
                                          unsigned mk, float lt, float lh, float lw, float a, float (&acc)[NW * (4 / sizeof(T)) * 1]) {
  constexpr int NC = NW * (4 / sizeof(T));          // channels per lane
  const float ht = 1.f - lt, hh = 1.f - lh, hw = 1.f - lw;
  const float thh = ht * hh, thl = ht * lh, tlh = lt * hh, tll = lt * lh;
  const float wgt[8] = {thh * hw, thh * lw, thl * hw, thl * lw, tlh * hw, tlh * lw, tll * hw, tll * lw};
  const unsigned off[8] = {b0, b0 + dw, b0 + dh, b0 + dh + dw, b0 + dt, b0 + dt + dw, b0 + dt + dh, b0 + dt + dh + dw};
  float val[NC];
#pragma unroll
  for (int e = 0; e < NC; ++e) val[e] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < 8; k0 += 4) {   // 4 corners in flight at a time
    unsigned r[4][NW];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      const bool ok = ALLVALID || ((mk >> k) & 1u);
#pragma unroll
      for (int e = 0; e < NW; ++e) r[kk][e] = 0u;
      if (ok) {
        if constexpr (NW == 8 && sizeof(T) == 2) {   // 16 bf16 channels: one 32-byte load (measured slower than 2 lanes x 16 B: opt-in)
          const U8 u = ldg32(ubase + off[k]);
#pragma unroll
          for (int e = 0; e < 8; ++e) r[kk][e] = u.v[e];
        } else if constexpr (NW == 8) {                // 8 fp32 channels: two 16-byte loads
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(ubase + off[k]));
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(ubase + off[k]) + 1);
          r[kk][0] = u.x; r[kk][1] = u.y; r[kk][2] = u.z; r[kk][3] = u.w; r[kk][4] = v.x; r[kk][5] = v.y; r[kk][6] = v.z; r[kk][7] = v.w;
        } else {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(ubase + off[k]));
          r[kk][0] = u.x; r[kk][1] = u.y; r[kk][2] = u.z; r[kk][3] = u.w;
        }
      }
    }
    if constexpr (sizeof(T) == 2) {
#pragma unroll
      for (int kk = 0; kk < 4; kk += 2) {
        const unsigned wp = pack_bf16x2(wgt[k0 + kk], wgt[k0 + kk + 1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const unsigned short wb = h ? (unsigned short)(wp >> 16) : (unsigned short)(wp & 0xffffu);
#pragma unroll
          for (int e = 0; e < NW; ++e) {
            val[2 * e] = fma_bf16((unsigned short)(r[kk + h][e] & 0xffffu), wb, val[2 * e]);
            val[2 * e + 1] = fma_bf16((unsigned short)(r[kk + h][e] >> 16), wb, val[2 * e + 1]);
          }
        }
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int e = 0; e < NW; ++e) val[e] = fmaf(wgt[k0 + kk], __uint_as_float(r[kk][e]), val[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < NC; ++e) acc[e] = fmaf(val[e], a, acc[e]);
}

// NW: 32-bit words per lane and corner.  bf16: NW = 8 (16 channels, D % 16 == 0) or 4 (8 channels); fp32: NW = 8 (8 channels).
// FUSED: `loc` holds the raw sampling OFFSETS and `attn` the raw attention LOGITS of the two query projections; the softmax over the
// head's L*P logits and loc = ref + off / (T_l, W_l, H_l) (ops/modules/ms_deform_attn.py:186-192, the reference's normaliser order;
// same arithmetic as msda_prepare_kernel: IEEE division, expf) happen in the geometry phase, so neither sampling_locations nor the
// attention weights (136 MB per clip in fp32) are ever written or re-read.
template <typename T, int NW, bool FUSED = false>
__global__ void __launch_bounds__(kWarps * 32) msda_fwd_tile_kernel(const T* __restrict__ value,
                                                                    const int64_t* __restrict__ shapes,
                                                                    const int64_t* __restrict__ lsi,
                                                                    const float* __restrict__ loc,
                                                                    const float* __restrict__ attn, T* __restrict__ out,
                                                                    int n_tiles, int Len, int M, int D, int L, int Lq, int P,
                                                                    const float* __restrict__ ref = nullptr) {
  constexpr int NC = NW * (4 / sizeof(T));   // channels per lane
  __shared__ int s_T[kMaxLevels], s_H[kMaxLevels], s_W[kMaxLevels], s_ls[kMaxLevels];
  if (threadIdx.x < L) {
    s_T[threadIdx.x] = (int)shapes[threadIdx.x * 3]; s_H[threadIdx.x] = (int)shapes[threadIdx.x * 3 + 1];
    s_W[threadIdx.x] = (int)shapes[threadIdx.x * 3 + 2]; s_ls[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x % n_tiles;
  const int bm = blockIdx.x / n_tiles;           // (batch, head)
  const int m = bm % M, b = bm / M;
  const int LP = L * P;
  const unsigned rsb = (unsigned)(M * D) * (unsigned)sizeof(T);   // bytes between consecutive tokens (Len*M*D*sizeof(T) < 2^32 checked)
  const int lpp = D / NC, G = 32 / lpp;          // lanes per point, points per step
  const int g = lane / lpp, ch = lane % lpp;
  // warp-uniform base of this (batch, head); per-lane 32-bit byte offsets below
  const char* ubase = reinterpret_cast<const char*>(value + (long)b * Len * ((long)M * D) + (long)m * D);
  const unsigned choff = (unsigned)ch * (unsigned)(NC * sizeof(T));
  const int q_end = min(Lq, (tile + 1) * kTileQ);
  // geometry of up to 32 sampling points of query q (one per lane): corner base offset, strides, validity mask, lerp weights and
  // the attention weight
  auto geometry = [&](int q, int p0, unsigned& g_b0, unsigned& g_dh, unsigned& g_dt, unsigned& g_mask, float& g_lt, float& g_lh,
                      float& g_lw, float& g_a) {
    const long wid = ((long)b * Lq + q) * M + m;
    const float* locp = loc + wid * LP * 3;
    const float* attp = attn + wid * LP;
    const int pt = p0 + lane;
    g_b0 = g_dh = g_dt = g_mask = 0; g_lt = g_lh = g_lw = g_a = 0.f;
    float sm_max = 0.f, sm_sum = 1.f, lg = 0.f;
    if constexpr (FUSED) {
      if (LP <= 32) {                       // one logit per lane: a single load feeds the max, the sum and the weight
        lg = pt < LP ? attp[pt] : -INFINITY;
        sm_max = warp_max(lg);
        sm_sum = warp_sum(pt < LP ? expf(lg - sm_max) : 0.f);
      } else {
        sm_max = -INFINITY;
        for (int j = lane; j < LP; j += 32) sm_max = fmaxf(sm_max, attp[j]);
        sm_max = warp_max(sm_max);
        sm_sum = 0.f;
        for (int j = lane; j < LP; j += 32) sm_sum += expf(attp[j] - sm_max);
        sm_sum = warp_sum(sm_sum);
        if (pt < LP) lg = attp[pt];
      }
    }
    if (pt < LP) {
      const int l = pt / P;
      const int Tt = s_T[l], H = s_H[l], W = s_W[l];
      int tl, hl, wl;
      float lx = locp[pt * 3], ly = locp[pt * 3 + 1], lz = locp[pt * 3 + 2];
      if constexpr (FUSED) {
        const float* r = ref + (((long)b * Lq + q) * L + l) * 3;
        lx = r[0] + __fdiv_rn(lx, (float)Tt); ly = r[1] + __fdiv_rn(ly, (float)W); lz = r[2] + __fdiv_rn(lz, (float)H);
        g_a = expf(lg - sm_max) / sm_sum;   // exactly msda_prepare_kernel's arithmetic
      } else {
        g_a = attp[pt];
      }
      point_geometry(lx, ly, lz, Tt, H, W, tl, hl, wl, g_mask, g_lt, g_lh, g_lw);
      g_dh = (unsigned)W * rsb; g_dt = (unsigned)(H * W) * rsb;
      g_b0 = (unsigned)(s_ls[l] + (tl * H + hl) * W + wl) * rsb;   // wraps for invalid low corners; those are never read
    }
  };
  auto gather = [&](int np, unsigned g_b0, unsigned g_dh, unsigned g_dt, unsigned g_mask, float g_lt, float g_lh, float g_lw, float g_a,
                    float (&acc)[NC]) {
    for (int j0 = 0; j0 < np; j0 += G) {
      const int src = min(j0 + g, 31);
      unsigned mk = __shfl_sync(0xffffffffu, g_mask, src);
      const unsigned b0 = __shfl_sync(0xffffffffu, g_b0, src) + choff;
      const unsigned dh = __shfl_sync(0xffffffffu, g_dh, src), dt = __shfl_sync(0xffffffffu, g_dt, src);
      const float lt = __shfl_sync(0xffffffffu, g_lt, src), lh = __shfl_sync(0xffffffffu, g_lh, src);
      const float lw = __shfl_sync(0xffffffffu, g_lw, src), a = __shfl_sync(0xffffffffu, g_a, src);
      if (j0 + g >= np) mk = 0;
      if (__all_sync(0xffffffffu, mk == 0xffu)) tile_step<T, NW, true>(ubase, b0, rsb, dh, dt, mk, lt, lh, lw, a, acc);
      else if (__any_sync(0xffffffffu, mk != 0u)) tile_step<T, NW, false>(ubase, b0, rsb, dh, dt, mk, lt, lh, lw, a, acc);
    }
  };
  auto finish = [&](int q, float (&acc)[NC]) {
    const long wid = ((long)b * Lq + q) * M + m;
    for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
      for (int e = 0; e < NC; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
    }
    if (g == 0) {
#pragma unroll
      for (int e0 = 0; e0 < NC; e0 += 8) {
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = acc[e0 + e];
        store8(out + wid * D + ch * NC + e0, o8);
      }
    }
  };
  int q = tile * kTileQ + warp;
  if (q >= q_end) return;
  if (LP <= 32) {
    // software pipeline over the warp's queries: the next query's geometry (its loc / logit loads, the softmax reductions) is
    // issued before the current query's gathers, so its latency hides under them
    unsigned c_b0, c_dh, c_dt, c_mask; float c_lt, c_lh, c_lw, c_a;
    geometry(q, 0, c_b0, c_dh, c_dt, c_mask, c_lt, c_lh, c_lw, c_a);
    for (; q < q_end; q += kWarps) {
      unsigned n_b0 = 0, n_dh = 0, n_dt = 0, n_mask = 0; float n_lt = 0, n_lh = 0, n_lw = 0, n_a = 0;
      if (q + kWarps < q_end) geometry(q + kWarps, 0, n_b0, n_dh, n_dt, n_mask, n_lt, n_lh, n_lw, n_a);
      float acc[NC];
#pragma unroll
      for (int e = 0; e < NC; ++e) acc[e] = 0.f;
      gather(LP, c_b0, c_dh, c_dt, c_mask, c_lt, c_lh, c_lw, c_a, acc);
      finish(q, acc);
      c_b0 = n_b0; c_dh = n_dh; c_dt = n_dt; c_mask = n_mask; c_lt = n_lt; c_lh = n_lh; c_lw = n_lw; c_a = n_a;
    }
    return;
  }
  for (; q < q_end; q += kWarps) {
    float acc[NC];
#pragma unroll
    for (int e = 0; e < NC; ++e) acc[e] = 0.f;
    for (int p0 = 0; p0 < LP; p0 += 32) {
      unsigned g_b0, g_dh, g_dt, g_mask; float g_lt, g_lh, g_lw, g_a;
      geometry(q, p0, g_b0, g_dh, g_dt, g_mask, g_lt, g_lh, g_lw, g_a);
      gather(min(32, LP - p0), g_b0, g_dh, g_dt, g_mask, g_lt, g_lh, g_lw, g_a, acc);
    }
    finish(q, acc);
  }
}

// Backward (mathematical gradient of the forward).  Same warp mapping; grad_value via red.global.add.f32 (one
// coalesced 128-byte reduction per corner), grad_loc / grad_attn reduced over the head's channels by shuffle and
// written once by the owning warp (no atomics).
template <typename T, int CPL>
__global__ void __launch_bounds__(kWarps * 32) msda_bwd_kernel(const T* __restrict__ value,
                                                               const int64_t* __restrict__ shapes,
                                                               const int64_t* __restrict__ lsi,
                                                               const float* __restrict__ loc,
                                                               const float* __restrict__ attn,
                                                               const T* __restrict__ grad_out,
                                                               float* __restrict__ grad_value,
                                                               float* __restrict__ grad_loc,
                                                               float* __restrict__ grad_attn, long n_warps_total,
                                                               int Len, int M, int D, int L, int Lq, int P) {
  const long wid = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_warps_total) return;
  const int m = (int)(wid % M);
  const long bq = wid / M;
  const int b = (int)(bq / Lq);
  const int LP = L * P;
  const long row_stride = (long)M * D;
  const long voff = (long)b * Len * row_stride + (long)m * D;
  const T* vbase = value + voff;
  float* gvbase = grad_value + voff;
  const float* locp = loc + wid * LP * 3;
  const float* attp = attn + wid * LP;
  float go[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    go[j] = c < D ? to_f(grad_out[wid * D + c]) : 0.f;
  }
  for (int pt = 0; pt < LP; ++pt) {
    const int l = pt / P;
    const int T_ = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
    int tl, hl, wl; unsigned mk; float lt, lh, lw;
    point_geometry(locp[pt * 3], locp[pt * 3 + 1], locp[pt * 3 + 2], T_, H, W, tl, hl, wl, mk, lt, lh, lw);
    float g_w = 0.f, g_h = 0.f, g_t = 0.f, g_a = 0.f;
    if (mk != 0) {
      const float a = attp[pt];
      const int base = (int)lsi[l] + (tl * H + hl) * W + wl;
      const int hs = W, ts = H * W;
      const float ht = 1.f - lt, hh = 1.f - lh, hw = 1.f - lw;
      const float ft[2] = {ht, lt}, fh[2] = {hh, lh}, fw[2] = {hw, lw};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!(mk & (1u << k))) continue;
        const int kt = k >> 2, kh = (k >> 1) & 1, kw = k & 1;
        const long row = (long)(base + kt * ts + kh * hs + kw) * row_stride;
        const float wgt = ft[kt] * fh[kh] * fw[kw];
        float dot = 0.f;
#pragma unroll
        for (int cj = 0; cj < CPL; ++cj) {
          const int c = lane + 32 * cj;
          if (c < D) {
            dot = fmaf(to_f(vbase[row + c]), go[cj], dot);
            atomicAdd(gvbase + row + c, a * wgt * go[cj]);
          }
        }
        g_a = fmaf(wgt, dot, g_a);
        g_w = fmaf((kw ? 1.f : -1.f) * ft[kt] * fh[kh], dot, g_w);
        g_h = fmaf((kh ? 1.f : -1.f) * ft[kt] * fw[kw], dot, g_h);
        g_t = fmaf((kt ? 1.f : -1.f) * fh[kh] * fw[kw], dot, g_t);
      }
      g_w *= a * (float)W; g_h *= a * (float)H; g_t *= a * (float)T_;
    }
    g_a = warp_sum(g_a); g_w = warp_sum(g_w); g_h = warp_sum(g_h); g_t = warp_sum(g_t);
    if (lane == 0) {
      grad_attn[wid * LP + pt] = g_a;
      float* gl = grad_loc + (wid * LP + pt) * 3;
      gl[0] = g_w; gl[1] = g_h; gl[2] = g_t;
    }
  }
}

// Backward, channel-vectorised (same conditions and lane mapping as msda_fwd_vec_kernel): lane = (point slot g, 8-channel
// chunk ch).  Per corner one 16-byte (bf16) value load, an 8-term dot with the lane's slice of grad_out, and two
// red.global.add.v4.f32 into grad_value; the four scalar gradients of a point are folded over the D/8 lanes of its slot only
// (2 shuffle steps at D = 32) and written by the slot's first lane.  The scalar kernel recomputed the point geometry on all 32
// lanes and paid four full warp reductions per point.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 4 consecutive elements at p and 4 at p + second -> v[0..3], v[4..7]
__device__ __forceinline__ void load4x2(const float* p, int second, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + second);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load4x2(const bf16* p, int second, float (&v)[8]) {
  const uint2 a = *reinterpret_cast<const uint2*>(p), b = *reinterpret_cast<const uint2*>(p + second);
  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
  float2 f;
  f = __bfloat1622float2(ha[0]); v[0] = f.x; v[1] = f.y;
  f = __bfloat1622float2(ha[1]); v[2] = f.x; v[3] = f.y;
  f = __bfloat1622float2(hb[0]); v[4] = f.x; v[5] = f.y;
  f = __bfloat1622float2(hb[1]); v[6] = f.x; v[7] = f.y;
}
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) msda_bwd_vec_kernel(const T* __restrict__ value,
                                                                   const int64_t* __restrict__ shapes,
                                                                   const int64_t* __restrict__ lsi,
                                                                   const float* __restrict__ loc,
                                                                   const float* __restrict__ attn,
                                                                   const T* __restrict__ grad_out,
                                                                   float* __restrict__ grad_value,
                                                                   float* __restrict__ grad_loc,
                                                                   float* __restrict__ grad_attn, long n_warps_total,
                                                                   int Len, int M, int D, int L, int Lq, int P) {
  const long wid = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n_warps_total) return;
  const int m = (int)(wid % M);
  const long bq = wid / M;
  const int b = (int)(bq / Lq);
  const int LP = L * P;
  const long row_stride = (long)M * D;
  const int lpp = D >> 3, G = 32 / lpp;
  const int g = lane / lpp, ch = lane % lpp;
  // Channel ownership is INTERLEAVED: lane `ch` of a point's lpp lanes owns channels [4 ch, 4 ch + 4) and [D/2 + 4 ch, D/2 + 4 ch + 4),
  // so that one red.global.add.v4.f32 of the lpp lanes covers D/2 * 4 CONTIGUOUS bytes (full 32-byte sectors) of the voxel's gradient row.
  // With 8 consecutive channels per lane every instruction left half of each sector it touched unwritten: 1.45 G sector operations for
  // 35 GB of payload (ncu, profiles/r02_ncu_kernels.txt), and the kernel is bound by exactly that rate (l1tex 81 %, lts 65 %).
  const int hD = D >> 1;
  const long voff = (long)b * Len * row_stride + (long)m * D + ch * 4;
  const T* vbase = value + voff;
  float* gvbase = grad_value + voff;
  const float* locp = loc + wid * LP * 3;
  const float* attp = attn + wid * LP;
  float go[8];
  load4x2(grad_out + wid * D + ch * 4, hD, go);
  for (int p0 = 0; p0 < LP; p0 += G) {
    const int pt = p0 + g;
    float g_w = 0.f, g_h = 0.f, g_t = 0.f, g_a = 0.f;
    if (pt < LP) {
      const int l = pt / P;
      const int T_ = (int)shapes[l * 3], H = (int)shapes[l * 3 + 1], W = (int)shapes[l * 3 + 2];
      int tl, hl, wl; unsigned mk; float lt, lh, lw;
      point_geometry(locp[pt * 3], locp[pt * 3 + 1], locp[pt * 3 + 2], T_, H, W, tl, hl, wl, mk, lt, lh, lw);
      if (mk != 0) {
        const float a = attp[pt];
        const int base = (int)lsi[l] + (tl * H + hl) * W + wl;
        const int hs = W, ts = H * W;
        const float ht = 1.f - lt, hh = 1.f - lh, hw = 1.f - lw;
        const float ft[2] = {ht, lt}, fh[2] = {hh, lh}, fw[2] = {hw, lw};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!(mk & (1u << k))) continue;
          const int kt = k >> 2, kh = (k >> 1) & 1, kw = k & 1;
          const long row = (long)(base + kt * ts + kh * hs + kw) * row_stride;
          const float wgt = ft[kt] * fh[kh] * fw[kw];
          float v[8];
          load4x2(vbase + row, hD, v);
          float dot = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) dot = fmaf(v[e], go[e], dot);
          const float s = a * wgt;
          red_add_v4(gvbase + row, s * go[0], s * go[1], s * go[2], s * go[3]);
          red_add_v4(gvbase + row + hD, s * go[4], s * go[5], s * go[6], s * go[7]);
          g_a = fmaf(wgt, dot, g_a);
          g_w = fmaf((kw ? 1.f : -1.f) * ft[kt] * fh[kh], dot, g_w);
          g_h = fmaf((kh ? 1.f : -1.f) * ft[kt] * fw[kw], dot, g_h);
          g_t = fmaf((kt ? 1.f : -1.f) * fh[kh] * fw[kw], dot, g_t);
        }
        g_w *= a * (float)W; g_h *= a * (float)H; g_t *= a * (float)T_;
      }
    }
    for (int o = 1; o < lpp; o <<= 1) {      // fold the slot's D/8 channel chunks
      g_a += __shfl_xor_sync(0xffffffffu, g_a, o); g_w += __shfl_xor_sync(0xffffffffu, g_w, o);
      g_h += __shfl_xor_sync(0xffffffffu, g_h, o); g_t += __shfl_xor_sync(0xffffffffu, g_t, o);
    }
    if (ch == 0 && pt < LP) {
      grad_attn[wid * LP + pt] = g_a;
      float* gl = grad_loc + (wid * LP + pt) * 3;
      gl[0] = g_w; gl[1] = g_h; gl[2] = g_t;
    }
  }
}

template <typename T>
int msda_fwd_t(const void* value, const int64_t* shapes, const int64_t* lsi, const float* loc, const float* attn, void* out,
               int N, int Len, int M, int D, int L, int Lq, int P, cudaStream_t st) {
  const long nw = (long)N * Lq * M;
  if (nw == 0) return 0;
  const unsigned grid = (unsigned)cdiv(nw, kWarps);
  static const bool no_vec = getenv("CQVAD_MSDA_NO_VEC") != nullptr;
  static const bool old_vec = getenv("CQVAD_MSDA_OLD_FWD") != nullptr;   // A/B switch: the round-1 warp-per-(b,q,m) kernel
  const int lpp = D >> 3;
  const bool vec_ok = !no_vec && D % 8 == 0 && lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0 && (((uintptr_t)value) & 15) == 0 &&
                      (((uintptr_t)out) & 15) == 0;
  if (vec_ok && !old_vec && L <= kMaxLevels && (long)Len * M * D * (long)sizeof(T) < (1L << 32) &&
      true) {
    const int n_tiles = (int)cdiv(Lq, kTileQ);
    const long n_cta = (long)n_tiles * N * M;
    CQ_CHECK_SHAPE(n_cta < (1L << 31), "msda3d: grid too large");
    const int lpp16 = D >> 4;
    const bool wide = sizeof(T) == 2 && D % 16 == 0 && lpp16 >= 1 && (lpp16 & (lpp16 - 1)) == 0 && (((uintptr_t)value) & 31) == 0 &&
                      ((size_t)M * D * sizeof(T)) % 32 == 0 && getenv("CQVAD_MSDA_WIDE") != nullptr;   // LDG.256 variant: 1.85 ms vs 1.40 ms (112 registers), opt-in
    if (sizeof(T) == 4)       // 8 fp32 channels = one 32-byte load
      msda_fwd_tile_kernel<T, 8><<<(unsigned)n_cta, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, n_tiles, Len, M, D, L, Lq, P);
    else if (wide)            // 16 bf16 channels per lane
      msda_fwd_tile_kernel<T, 8><<<(unsigned)n_cta, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, n_tiles, Len, M, D, L, Lq, P);
    else if constexpr (sizeof(T) == 2)
      msda_fwd_tile_kernel<T, 4><<<(unsigned)n_cta, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, n_tiles, Len, M, D, L, Lq, P);
    CQ_LAUNCH_CHECK();
    return 0;
  }
  if (vec_ok) {
    msda_fwd_vec_kernel<T, false><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, nullptr, (T*)out, nw, Len, M, D, L, Lq, P);
    CQ_LAUNCH_CHECK();
    return 0;
  }
  if (D <= 32)
    msda_fwd_kernel<T, 1><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, nw, Len, M, D, L, Lq, P);
  else if (D <= 64)
    msda_fwd_kernel<T, 2><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, nw, Len, M, D, L, Lq, P);
  else
    msda_fwd_kernel<T, 4><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (T*)out, nw, Len, M, D, L, Lq, P);
  CQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int msda_bwd_t(const void* value, const int64_t* shapes, const int64_t* lsi, const float* loc, const float* attn,
               const void* go, float* gv, float* gl, float* ga, int N, int Len, int M, int D, int L, int Lq, int P,
               cudaStream_t st) {
  const long nw = (long)N * Lq * M;
  if (nw == 0) return 0;
  const unsigned grid = (unsigned)cdiv(nw, kWarps);
  static const bool no_vec = getenv("CQVAD_MSDA_NO_VEC") != nullptr;
  const int lpp = D >> 3;
  if (!no_vec && D % 8 == 0 && lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0 && (((uintptr_t)value) & 15) == 0 &&
      (((uintptr_t)go) & 15) == 0 && (((uintptr_t)gv) & 15) == 0) {
    msda_bwd_vec_kernel<T><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (const T*)go, gv, gl, ga, nw, Len, M, D, L, Lq, P);
    CQ_LAUNCH_CHECK();
    return 0;
  }
  if (D <= 32)
    msda_bwd_kernel<T, 1><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (const T*)go, gv, gl, ga, nw, Len, M, D, L, Lq, P);
  else if (D <= 64)
    msda_bwd_kernel<T, 2><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (const T*)go, gv, gl, ga, nw, Len, M, D, L, Lq, P);
  else
    msda_bwd_kernel<T, 4><<<grid, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, loc, attn, (const T*)go, gv, gl, ga, nw, Len, M, D, L, Lq, P);
  CQ_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int msda_fwd_fused_t(const void* value, const int64_t* shapes, const int64_t* lsi, const float* off, const float* logit,
                     const float* ref, void* out, int N, int Len, int M, int D, int L, int Lq, int P, cudaStream_t st) {
  const long nw = (long)N * Lq * M;
  if (nw == 0) return 0;
  const int lpp = D >> 3;
  if (D % 8 != 0 || lpp < 1 || lpp > 32 || (lpp & (lpp - 1)) != 0 || (((uintptr_t)value) & 15) || (((uintptr_t)out) & 15)) return 1;
  if (L <= kMaxLevels && (long)Len * M * D * (long)sizeof(T) < (1L << 32)) {       // query-tile kernel (the default sampling kernel)
    const int n_tiles = (int)cdiv(Lq, kTileQ);
    const long n_cta = (long)n_tiles * N * M;
    if (n_cta < (1L << 31)) {
      if (sizeof(T) == 4)
        msda_fwd_tile_kernel<T, 8, true><<<(unsigned)n_cta, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, off, logit, (T*)out, n_tiles,
                                                                                  Len, M, D, L, Lq, P, ref);
      else if constexpr (sizeof(T) == 2)
        msda_fwd_tile_kernel<T, 4, true><<<(unsigned)n_cta, kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, off, logit, (T*)out, n_tiles,
                                                                                  Len, M, D, L, Lq, P, ref);
      CQ_LAUNCH_CHECK();
      return 0;
    }
  }
  msda_fwd_vec_kernel<T, true><<<(unsigned)cdiv(nw, kWarps), kWarps * 32, 0, st>>>((const T*)value, shapes, lsi, off, logit, ref,
                                                                                  (T*)out, nw, Len, M, D, L, Lq, P);
  CQ_LAUNCH_CHECK();
  return 0;
}

int check_dims(int N, int Len, int M, int D, int L, int Lq, int P) {
  CQ_CHECK_ARG(N >= 0 && Len >= 0 && M >= 1 && D >= 1 && L >= 1 && Lq >= 0 && P >= 1, "msda3d: bad dimensions");
  CQ_CHECK_SHAPE(D <= 128, "msda3d: head dim D=%d > 128 not supported", D);
  CQ_CHECK_SHAPE((long)Len * M * D < (1L << 31), "msda3d: Len*M*D must fit int32 (as in the reference, cuh:49-60)");
  return 0;
}

}  // namespace

// offsets + logits + reference points -> sampled values (encoder.cu); returns 1 when the vectorised kernel does not apply
int msda_fwd_fused(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi, const float* off, const float* logit,
                   const float* ref, void* out, int N, int Len, int M, int D, int L, int Lq, int P, cudaStream_t st) {
  CQ_TRY(check_dims(N, Len, M, D, L, Lq, P));
  if (dtype == CQVAD_F32) return msda_fwd_fused_t<float>(value, shapes, lsi, off, logit, ref, out, N, Len, M, D, L, Lq, P, st);
  return msda_fwd_fused_t<bf16>(value, shapes, lsi, off, logit, ref, out, N, Len, M, D, L, Lq, P, st);
}
}  // namespace cqvad

using namespace cqvad;

extern "C" int cqvad_msda3d_forward(int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                                    const float* loc, const float* attn, void* out, int N, int Len, int M, int D, int L,
                                    int Lq, int P, void* stream) {
  CQ_TRY(check_dims(N, Len, M, D, L, Lq, P));
  if ((long)N * Lq == 0) return 0;   // empty query set: nothing to write (pointers of empty tensors may be NULL)
  CQ_CHECK_ARG(value && shapes && level_start && loc && attn && out, "msda3d_forward: null pointer");
  if (dtype == CQVAD_F32) return msda_fwd_t<float>(value, shapes, level_start, loc, attn, out, N, Len, M, D, L, Lq, P, as_stream(stream));
  if (dtype == CQVAD_BF16) return msda_fwd_t<bf16>(value, shapes, level_start, loc, attn, out, N, Len, M, D, L, Lq, P, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "msda3d_forward: unknown dtype %d", dtype);
}

extern "C" int cqvad_msda3d_backward(int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                                     const float* loc, const float* attn, const void* grad_out, float* grad_value,
                                     float* grad_loc, float* grad_attn, int N, int Len, int M, int D, int L, int Lq, int P,
                                     void* stream) {
  CQ_TRY(check_dims(N, Len, M, D, L, Lq, P));
  if ((long)N * Lq == 0) return 0;
  CQ_CHECK_ARG(value && shapes && level_start && loc && attn && grad_out && grad_value && grad_loc && grad_attn,
               "msda3d_backward: null pointer");
  if (dtype == CQVAD_F32) return msda_bwd_t<float>(value, shapes, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, Len, M, D, L, Lq, P, as_stream(stream));
  if (dtype == CQVAD_BF16) return msda_bwd_t<bf16>(value, shapes, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, N, Len, M, D, L, Lq, P, as_stream(stream));
  return set_error(CQVAD_E_INVALID_ARG, "msda3d_backward: unknown dtype %d", dtype);
}

extern "C" int cqvad_msda3d_indices(const int64_t* shapes, const float* loc, int32_t* t_low, int32_t* h_low,
                                    int32_t* w_low, uint8_t* corner_mask, int N, int Lq, int M, int L, int P, void* stream) {
  CQ_CHECK_ARG(shapes && loc && t_low && h_low && w_low && corner_mask, "msda3d_indices: null pointer");
  const long total = (long)N * Lq * M * L * P;
  if (total == 0) return 0;
  msda_indices_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(shapes, loc, t_low, h_low, w_low, corner_mask, total, L, P);
  CQ_LAUNCH_CHECK();
  return 0;
}
