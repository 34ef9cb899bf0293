// Backward of the class-query attentions (models/detr/attention.py:336-341,377,400-409 under autograd) on tcgen05:
// class cross-attention heads 0-3 (hd 64) and class self-attention (8 heads of 32).  One CTA per (head, instance):
//
//   S  [128 x Sp] = Q_h K_h^T              dP [128 x Sp] = dO_h V_h^T                 (two accumulators: 512 TMEM columns)
//   one thread per query row:  P = softmax(S / sqrt(hd)),  dS = P * (dP - rowsum(P*dP)) / sqrt(hd)   -> bf16 tiles in smem
//   dV [Sp x 64]  = P^T  dO                 (A = P tile read MN-major: keys contiguous, classes = reduction)
//   dQ [128 x 64] = dS   K                  (A = dS tile K-major, B = K tile read MN-major)
//   dK [Sp x 64]  = dS^T Q                  (both MN-major)
//
// Every operand is the row-major activation as TMA delivered it ([rows x 64-column box], 128-byte swizzle): a tile is a
// K-major operand when its 64 columns are the reduction, and an MN-major operand when its rows are (same bytes, other
// descriptor), so no transposed copy of Q, K, V, dO, P or dS is ever made.  With hd / vd = 32 the head's columns are half of
// a 64-column swizzle atom: K-major uses select them by k-step, MN-major uses compute the whole atom (64 wide) and the
// epilogue keeps the head's half.  Query rows >= K and keys >= S are written as zeros into P / dS so that neighbouring
// instances' rows inside the 128-row boxes contribute nothing.
//
// The class-independent heads 4-7 of the cross-attention (actor sine position . spatial position) collapse to ONE query
// row per actor with dO summed over the classes (exact): cls_xattn_pos_bwd_kernel, CUDA cores.
#include "common.cuh"
#include "tc_common.cuh"
#include "bwd.cuh"
#include <stdlib.h>

namespace cqvad {

using namespace tc;
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int tc_num_sms();

namespace {

constexpr int AB_THREADS = 288;   // warp 0: TMA + MMA; warps 1-8: softmax (two per TMEM lane quarter, half of the keys each); warps 1-4: epilogue
constexpr int TILE16K = 16384;

struct AbParams {
  bf16 *dQ, *dK, *dV;
  float beta_q, beta_k, beta_v;
  int K, S, Sp16;
  int q_rows, k_rows, v_rows;     // row pitch per instance of the q / k / v matrices
  int hd;                         // 64 or 32
  int fused;                      // self-attention: q = k = v, one gradient tensor receives dq + dk + dv
  int halves;                     // 128-key halves (1 or 2)
  float scale, scale_log2;
  uint32_t off_k, off_v, off_do, off_p, off_ds, off_bar;
  long long* trace;               // dev tool (CQVAD_ATTN_TRACE=<device address>): globaltimer stamps of CTA (0,0)'s phases
};
#define AB_TRACE(k, tid) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == (tid)) { long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[k] = t_; } } while (0)

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld16b(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16b(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_tile8(uint32_t tile, int row, int col, const float (&v)[8]) {
  const uint32_t chunk = (uint32_t)((col & 63) >> 3) ^ (uint32_t)(row & 7);
  const uint32_t addr = tile + (uint32_t)(col >> 6) * TILE16K + (uint32_t)row * 128u + chunk * 16u;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk(v[0], v[1])), "r"(pk(v[2], v[3])),
               "r"(pk(v[4], v[5])), "r"(pk(v[6], v[7]))
               : "memory");
}
// 32 gradient values of one row -> global (beta accumulate), optionally summed with two more register sets
__device__ __forceinline__ void store32(bf16* dst, const uint32_t (&r)[32], float beta) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g8 * 8 + e]);
    if (beta != 0.f) {
      float o[8];
      load8(dst + g8 * 8, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(beta, o[e], v[e]);
    }
    store8(dst + g8 * 8, v);
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + p.off_k, sV = base + p.off_v, sDO = base + p.off_do, sP = base + p.off_p,
                 sDS = base + p.off_ds;
  const uint32_t bars = base + p.off_bar;
  const uint32_t bar_in = bars, bar_s = bars + 8, bar_p = bars + 16, bar_o = bars + 24, tmem_slot = bars + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x;
  const long i = blockIdx.y;
  AB_TRACE(0, 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(bar_in, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  AB_TRACE(1, 0);
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 256;
  const uint32_t t_dv = tmem_base, t_dq = tmem_base + 128, t_dk = tmem_base + 192;   // reuse once S / dP are consumed
  const int off32 = (h * 32) & 63;            // the head's 32 value columns inside their 64-column block
  const int qoff = (h * p.hd) & 63;           // the head's q/k columns inside their block (0 for hd 64)

  if (warp == 0) {
    if (lane == 0) {
      const int qcol0 = ((h * p.hd) >> 6) << 6, vcol0 = ((h * 32) >> 6) << 6;
      mbar_arrive_expect_tx(bar_in, (uint32_t)(2 * TILE16K + 2 * p.Sp16 * 128));
      tma_load_2d(sQ, &tmQ, bar_in, qcol0, (int)(i * p.q_rows));
      tma_load_2d(sK, &tmK, bar_in, qcol0, (int)(i * p.k_rows));
      tma_load_2d(sV, &tmV, bar_in, vcol0, (int)(i * p.v_rows));
      tma_load_2d(sDO, &tmDO, bar_in, vcol0, (int)(i * p.K));
      mbar_wait(bar_in, 0);
      AB_TRACE(2, 0);
      tc_fence_after();
      {
        const uint32_t idesc = make_idesc_bf16(128, p.Sp16);
        const uint64_t qd = make_smem_desc_sw128(sQ), kd = make_smem_desc_sw128(sK);
        const int k0 = qoff >> 4, nk = p.hd >> 4;
        for (int k = 0; k < nk; ++k) umma_bf16(t_s, qd + (uint64_t)(2 * (k0 + k)), kd + (uint64_t)(2 * (k0 + k)), idesc, k ? 1u : 0u);
        const uint64_t dd = make_smem_desc_sw128(sDO), vd = make_smem_desc_sw128(sV);
        const int v0 = off32 >> 4;
        for (int k = 0; k < 2; ++k) umma_bf16(t_dp, dd + (uint64_t)(2 * (v0 + k)), vd + (uint64_t)(2 * (v0 + k)), idesc, k ? 1u : 0u);
        umma_commit(bar_s);
      }
      mbar_wait(bar_p, 0);
      tc_fence_after();
      {
        const int csteps = (p.K + 15) >> 4;      // class rows >= K are zero in P / dS
        const uint32_t id_mm = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);   // A, B MN-major
        const uint32_t id_km = make_idesc_bf16(128, 64) | (1u << 16);                // A K-major, B MN-major
        for (int hf = 0; hf < p.halves; ++hf) {   // dV = P^T dO
          for (int j = 0; j < csteps; ++j)
            umma_bf16(t_dv + (uint32_t)(hf * 64), desc_mn(sP + hf * 2 * TILE16K + j * 2048, TILE16K), desc_mn(sDO + j * 2048, TILE16K),
                      id_mm, j ? 1u : 0u);
        }
        const int ksteps = p.Sp16 >> 4;           // dQ = dS K
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t a = make_smem_desc_sw128(sDS + (ks >> 2) * TILE16K) + (uint64_t)(2 * (ks & 3));
          umma_bf16(t_dq, a, desc_mn(sK + ks * 2048, TILE16K), id_km, ks ? 1u : 0u);
        }
        for (int hf = 0; hf < p.halves; ++hf) {   // dK = dS^T Q
          for (int j = 0; j < csteps; ++j)
            umma_bf16(t_dk + (uint32_t)(hf * 64), desc_mn(sDS + hf * 2 * TILE16K + j * 2048, TILE16K), desc_mn(sQ + j * 2048, TILE16K),
                      id_mm, j ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // Softmax / dS phase on 8 warps.  The phase trace (tools/trace_attn_bwd.py) of the 4-warp, 3-pass version showed 16 us of a
    // 23.7 us CTA here (exp2f twice per element, one exposed warp per TMEM lane quarter): now each lane quarter has two warps
    // (key halves), the exponentials are computed ONCE (ex2.approx) and parked in TMEM over S, and the row statistics of the
    // two halves are exchanged through shared memory.
    const int q = warp & 3;
    const int sh = (warp - 1) >> 2;                 // key half of this warp
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool vrow = row < p.K;
    const int S = p.S, Sp16 = p.Sp16;
    const int nch = Sp16 >> 4, ch0 = sh ? (nch + 1) >> 1 : 0, ch1 = sh ? nch : (nch + 1) >> 1;
    // [3][2][128] partials, aliased onto the (still unused) dS tile: S = 256 leaves no spare shared memory (224 KB of tiles)
    float* xch = reinterpret_cast<float*>(smem_raw + (sDS - smem_u32(smem_raw)));
    mbar_wait(bar_s, 0);
    AB_TRACE(4, 32);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = ch0 * 16; c < ch1 * 16; c += 16) {
      uint32_t r[16];
      tmem_ld16b(t_s + lane_off + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) if (c + e < S) mx = fmaxf(mx, __uint_as_float(r[e]));
    }
    xch[sh * 128 + row] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(mx, xch[(sh ^ 1) * 128 + row]);
    AB_TRACE(5, 32);
    const float sc = p.scale_log2, mxs = mx * sc;
    float sum = 0.f, sdp = 0.f;
    for (int c = ch0 * 16; c < ch1 * 16; c += 16) {
      uint32_t r[16], d[16];
      tmem_ld16b(t_s + lane_off + c, r);
      tmem_ld16b(t_dp + lane_off + c, d);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float x = (c + e < S) ? ex2_approx(fmaf(__uint_as_float(r[e]), sc, -mxs)) : 0.f;
        sum += x;
        sdp = fmaf(x, __uint_as_float(d[e]), sdp);
        r[e] = __float_as_uint(x);
      }
      tmem_st16b(t_s + lane_off + c, r);          // exp values parked over S: the last pass does not recompute them
    }
    tmem_st_wait();
    xch[256 + sh * 128 + row] = sum;
    xch[512 + sh * 128 + row] = sdp;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum += xch[256 + (sh ^ 1) * 128 + row];
    sdp += xch[512 + (sh ^ 1) * 128 + row];
    asm volatile("bar.sync 1, 256;" ::: "memory");   // every partial has been read: the dS tile may be written
    AB_TRACE(6, 32);
    const float inv = 1.0f / sum;
    const float D = sdp * inv;
    for (int c = ch0 * 16; c < ch1 * 16; c += 16) {
      uint32_t r[16], d[16];
      tmem_ld16b(t_s + lane_off + c, r);
      tmem_ld16b(t_dp + lane_off + c, d);
      tmem_ld_wait();
      float pv[16], dsv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float x = vrow ? __uint_as_float(r[e]) * inv : 0.f;   // 0 for keys >= S (exp parked as 0) and class rows >= K
        pv[e] = x;
        dsv[e] = vrow ? p.scale * x * (__uint_as_float(d[e]) - D) : 0.f;
      }
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        float a[8], b[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] = pv[g8 * 8 + e]; b[e] = dsv[g8 * 8 + e]; }
        st_tile8(sP, row, c + g8 * 8, a);
        st_tile8(sDS, row, c + g8 * 8, b);
      }
    }
    AB_TRACE(7, 32);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (sh == 0) {
    mbar_wait(bar_o, 0);
    AB_TRACE(8, 32);
    tc_fence_after();
    // ---- epilogue (warps 1-4: one TMEM lane quarter each) ----
    uint32_t rq[32], rk[32], rv[32];
    if (p.fused) {   // self-attention: keys == queries, one output row per thread
      tmem_ld32(t_dq + lane_off + qoff, rq);
      tmem_ld32(t_dk + lane_off + qoff, rk);
      tmem_ld32(t_dv + lane_off + off32, rv);
      tmem_ld_wait();
      if (vrow) {
#pragma unroll
        for (int e = 0; e < 32; ++e) rq[e] = __float_as_uint(__uint_as_float(rq[e]) + __uint_as_float(rk[e]) + __uint_as_float(rv[e]));
        store32(p.dQ + (i * p.q_rows + row) * kC + h * 32, rq, p.beta_q);
      }
    } else {
      const int nq32 = p.hd >> 5;   // 32-column groups of the head's q/k slice
      for (int g = 0; g < nq32; ++g) {
        tmem_ld32(t_dq + lane_off + qoff + g * 32, rq);
        tmem_ld_wait();
        if (vrow) store32(p.dQ + (i * p.q_rows + row) * kC + h * p.hd + g * 32, rq, p.beta_q);
      }
      for (int hf = 0; hf < p.halves; ++hf) {
        const int key = hf * 128 + row;
        tmem_ld32(t_dv + lane_off + hf * 64 + off32, rv);
        tmem_ld_wait();
        if (key < S) store32(p.dV + (i * p.v_rows + key) * kC + h * 32, rv, p.beta_v);
        for (int g = 0; g < nq32; ++g) {
          tmem_ld32(t_dk + lane_off + hf * 64 + qoff + g * 32, rk);
          tmem_ld_wait();
          if (key < S) store32(p.dK + (i * p.k_rows + key) * kC + h * p.hd + g * 32, rk, p.beta_k);
        }
      }
    }
    }   // sh == 0: epilogue warps
  }
  AB_TRACE(9, 32);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  AB_TRACE(10, 0);
}

// heads 4-7 of the class cross-attention: one block per actor instance, warp = head
__global__ void __launch_bounds__(128) cls_xattn_pos_bwd_kernel(const bf16* __restrict__ cqp, const bf16* __restrict__ pos0,
                                                                const bf16* __restrict__ vx, const bf16* __restrict__ dO,
                                                                bf16* dcqp, float beta_q, bf16* __restrict__ dvx, int K, int S,
                                                                int Sq, int BT) {
  extern __shared__ float sm[];
  const long i = blockIdx.x;
  const int bb = (int)(i % BT);
  const int hh = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = sm + hh * (2 * S + 32);   // probabilities
  float* ds = pr + S;                   // score gradients
  float* dos = ds + S;                  // dO summed over the classes [32]
  const int c0 = 128 + hh * 32;
  {
    float a = 0.f;
    for (int k = 0; k < K; ++k) a += __bfloat162float(dO[(i * K + k) * kC + c0 + lane]);
    dos[lane] = a;
  }
  float qv[64];
  {
    const bf16* qp = cqp + i * kC + hh * 64;
#pragma unroll
    for (int d8 = 0; d8 < 8; ++d8) {
      float t[8];
      load8(qp + d8 * 8, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) qv[d8 * 8 + e] = t[e] * 0.125f;
    }
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int s = lane; s < S; s += 32) {
    const bf16* kp = pos0 + ((long)s * BT + bb) * kC + hh * 64;
    float a = 0.f;
#pragma unroll
    for (int d8 = 0; d8 < 8; ++d8) {
      float t[8];
      load8(kp + d8 * 8, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(qv[d8 * 8 + e], t[e], a);
    }
    pr[s] = a;
    mx = fmaxf(mx, a);
    // dP[s] = dOsum . V[s]
    const bf16* vr = vx + (i * Sq + s) * kC + c0;
    float b = 0.f;
#pragma unroll
    for (int d8 = 0; d8 < 4; ++d8) {
      float t[8];
      load8(vr + d8 * 8, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) b = fmaf(dos[d8 * 8 + e], t[e], b);
    }
    ds[s] = b;
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float e = expf(pr[s] - mx);
    pr[s] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  float dot = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float pp = pr[s] * inv;
    pr[s] = pp;
    dot = fmaf(pp, ds[s], dot);
  }
  dot = warp_sum(dot);
  for (int s = lane; s < S; s += 32) ds[s] = pr[s] * (ds[s] - dot);
  __syncwarp();
  // dV[s, c0 + lane] = p[s] * dOsum[lane];  dcqp[c] = 0.125 * sum_s dS[s] * pos0[s, c]
  const float my_do = dos[lane];
  float a0 = 0.f, a1 = 0.f;
  for (int s = 0; s < S; ++s) {
    dvx[(i * Sq + s) * kC + c0 + lane] = __float2bfloat16_rn(pr[s] * my_do);
    const bf16* kp = pos0 + ((long)s * BT + bb) * kC + hh * 64;
    const float d = ds[s];
    a0 = fmaf(d, __bfloat162float(kp[lane]), a0);
    a1 = fmaf(d, __bfloat162float(kp[lane + 32]), a1);
  }
  bf16* dq = dcqp + i * kC + hh * 64;
  a0 *= 0.125f; a1 *= 0.125f;
  if (beta_q != 0.f) { a0 = fmaf(beta_q, __bfloat162float(dq[lane]), a0); a1 = fmaf(beta_q, __bfloat162float(dq[lane + 32]), a1); }
  dq[lane] = __float2bfloat16_rn(a0);
  dq[lane + 32] = __float2bfloat16_rn(a1);
}

// xt[c][i*K8 + k] = x[i*K + k][c]   (V^T operand of the forward self-attention kernel)
__global__ void __launch_bounds__(256) transpose_tokens_kernel(const bf16* __restrict__ x, bf16* __restrict__ xt, long ldxt,
                                                               int K, int K8) {
  __shared__ bf16 tile[32][33];
  const long i = blockIdx.z;
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    tile[j][tx] = (k0 + j < K) ? x[(i * K + k0 + j) * kC + c0 + tx] : __float2bfloat16_rn(0.f);
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (k0 + tx < K8) xt[(long)(c0 + j) * ldxt + i * K8 + k0 + tx] = tile[tx][j];
}

int g_ab_attr = 0;

int launch_attn_bwd(const bf16* q, long q_total, const bf16* kmat, long k_total, const bf16* v, long v_total, const bf16* dO,
                    long N, int heads, int hd, int K, int S, int q_rows, int k_rows, int v_rows, bf16* dQ, float bq, bf16* dK,
                    float bk, bf16* dV, float bv, int fused, cudaStream_t st) {
  const int Sp16 = (S + 15) & ~15;
  if (K > 128 || K < 1 || Sp16 > 256 || (hd != 64 && hd != 32)) return 1;
  if (fused && (hd != 32 || K != S)) return 1;
  if (tc_num_sms() <= 0) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed");
  AbParams p{};
  p.halves = Sp16 > 128 ? 2 : 1;
  const uint32_t kbytes = ((uint32_t)Sp16 * 128u + 1023u) & ~1023u;
  const uint32_t pbytes = (uint32_t)p.halves * 2u * TILE16K;
  p.off_k = TILE16K; p.off_v = p.off_k + kbytes; p.off_do = p.off_v + kbytes; p.off_p = p.off_do + TILE16K;
  p.off_ds = p.off_p + pbytes; p.off_bar = p.off_ds + pbytes;
  const int smem_bytes = (int)p.off_bar + 64 + 1024;
  if (smem_bytes > 227 * 1024) return 1;
  if (smem_bytes > g_ab_attr) {
    CQ_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    g_ab_attr = smem_bytes;
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  const cuuint64_t strides[1] = {(cuuint64_t)kC * 2};
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)q_total};
    const cuuint32_t box[2] = {64, 128};
    CQ_TRY(make_tmap_bf16(&tmQ, q, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)k_total};
    const cuuint32_t box[2] = {64, (cuuint32_t)Sp16};
    CQ_TRY(make_tmap_bf16(&tmK, kmat, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)v_total};
    const cuuint32_t box[2] = {64, (cuuint32_t)Sp16};
    CQ_TRY(make_tmap_bf16(&tmV, v, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)(N * K)};
    const cuuint32_t box[2] = {64, 128};
    CQ_TRY(make_tmap_bf16(&tmDO, dO, 2, dims, strides, box));
  }
  p.dQ = dQ; p.dK = dK; p.dV = dV; p.beta_q = bq; p.beta_k = bk; p.beta_v = bv;
  p.K = K; p.S = S; p.Sp16 = Sp16; p.q_rows = q_rows; p.k_rows = k_rows; p.v_rows = v_rows; p.hd = hd; p.fused = fused;
  if (const char* tr = getenv("CQVAD_ATTN_TRACE")) p.trace = (long long*)strtoull(tr, nullptr, 0) + (fused ? 16 : 0);
  p.scale = 1.0f / sqrtf((float)hd); p.scale_log2 = 1.4426950408889634f * p.scale;
  dim3 grid((unsigned)heads, (unsigned)N);
  attn_bwd_tc_kernel<<<grid, AB_THREADS, smem_bytes, st>>>(tmQ, tmK, tmV, tmDO, p);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// class cross-attention backward (heads 0-3 on tcgen05, heads 4-7 collapsed).  kx rows on the padded layout (pitch Sp_rows),
// vx rows with pitch Sq.  dkx / dvx: valid rows overwritten (caller zero-fills pad rows), dQin / dcqp with beta.
int cls_xattn_bwd_tc(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vx, const bf16* dO, bf16* dQin,
                     float beta_q, bf16* dcqp, float beta_q2, bf16* dkx, bf16* dvx, long N, int K, int S, int Sq, int Sp_rows, int BT,
                     cudaStream_t st) {
  int r = launch_attn_bwd(Qin, N * K, kx, N * Sp_rows, vx, N * Sq, dO, N, 4, 64, K, S, K, Sp_rows, Sq, dQin, beta_q, dkx, 0.f, dvx,
                          0.f, 0, st);
  if (r != 0) return r;
  cls_xattn_pos_bwd_kernel<<<(unsigned)N, 128, 4 * (2 * S + 32) * sizeof(float), st>>>(cqp, pos0, vx, dO, dcqp, beta_q2, dvx, K, S,
                                                                                     Sq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}

// class self-attention backward: q = k = v = x [N*K,256]; dx = beta*dx + dq + dk + dv
int cls_sattn_bwd_tc(const bf16* x, const bf16* dO, bf16* dx, float beta, long N, int K, cudaStream_t st) {
  return launch_attn_bwd(x, N * K, x, N * K, x, N * K, dO, N, 8, 32, K, K, K, K, K, dx, beta, nullptr, 0.f, nullptr, 0.f, 1, st);
}

int transpose_tokens(const bf16* x, bf16* xt, long ldxt, long N, int K, int K8, cudaStream_t st) {
  dim3 grid((unsigned)cdiv(K8, 32), kC / 32, (unsigned)N);
  transpose_tokens_kernel<<<grid, 256, 0, st>>>(x, xt, ldxt, K, K8);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace cqvad
