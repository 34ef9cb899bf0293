// Class-specific cross-attention of TransformerClassDecoderLayer (models/detr/dab_transformer.py:1067-1071 through
// models/detr/attention.py:336-341,377,400-409) on tcgen05.
//
// The 512-wide query/key are split CONTIGUOUSLY into 8 heads of 64 (attention.py:336,339), so
//   heads 0-3:  class-query content (K classes x 64)  .  k_proj(conv feature) (S keys x 64)     -> per class
//   heads 4-7:  actor sine-position (same for all K)  .  spatial position pos[0] (S x 64)       -> class independent
// Heads 0-3 run here, one CTA per (head, actor instance):
//   S[128 x Sp]   = Q_h[128(K valid) x 64] . K_h[Sp x 64]^T        4 x UMMA 128 x Sp x 16, accumulator in TMEM
//   P             = exp((S - rowmax)/8)  (one thread per row, straight out of TMEM) -> bf16, 128-byte-swizzled smem tile
//   O[128 x 32]   = P[128 x Sp] . V_h^T[32 x Sp]^T                 Sp/16 x UMMA 128 x 32 x 16
//   out           = O / rowsum + b_v
// V_h^T comes from vt = W_v . q_memory^T (the v_proj GEMM with swapped operands), so both UMMA operands are K-major and
// no transpose kernel exists; keys s >= S of the 16-padded tile get P = 0.  Heads 4-7 are computed once per actor by
// cls_xattn_pos_kernel (GEMV-shaped, CUDA cores) and broadcast over the K class rows (exact: identical query rows).
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace cqvad {

using namespace tc;
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int tc_num_sms();

namespace {

constexpr int QT_BYTES = 128 * 128;        // Q tile [128 x 64] bf16
constexpr int NUM_THREADS = 160;

// Generic projection-free attention tile on tcgen05 (one CTA per (head, instance)):
//   class cross-attention heads 0-3 : hd 64, keys = k_proj rows of the padded map, V^T = W_v . q_memory^T
//   class self-attention  heads 0-7 : hd 32 (two heads share one 64-column swizzle atom: the head's 32 columns are
//                                      selected by issuing only its two UMMA k-steps), keys = queries = class tokens,
//                                      V^T = transposed class tokens written by the class-FFN epilogue
// Shared memory: [Q | K] then V^T; the P tile ALIASES [Q | K] (dead once S = QK^T has completed), which is what lets
// two (cross) / four (self) CTAs share an SM.
struct XaParams {
  bf16* out;              // [N*K, 256]; head h writes columns [h*32, h*32+32)
  const float* bv;        // bias added to the output (v_proj bias) or nullptr
  int K, S, Sp16;         // valid query rows per instance, valid keys, keys padded to a multiple of 16
  int q_rows;             // query-row pitch per instance (= K)
  int k_rows;             // key-row pitch per instance ((h+1)*w for the padded map, K for self-attention)
  int v_pitch;            // per-instance column pitch of V^T (multiple of 8: TMA needs a 16-byte aligned inner start)
  int hd;                 // 64 or 32
  float scale_log2;       // hd^-0.5 * log2(e)
  int o_col, tmem_cols;
  uint32_t off_k, off_v, off_bar;   // byte offsets from the 1024-aligned base (Q and P at 0)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const XaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + p.off_k, sV = base + p.off_v, sP = base;
  const uint32_t bars = base + p.off_bar;
  const uint32_t bar_qk = bars, bar_v = bars + 8, bar_s = bars + 16, bar_p = bars + 24, bar_o = bars + 32, tmem_slot = bars + 40;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x;
  const long i = blockIdx.y;
  const int nkb = (p.Sp16 + 63) >> 6;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); }   // a warp with no divergent prologue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t t_s = tmem_base, t_o = tmem_base + (uint32_t)p.o_col;

  if (warp == 0) {
    if (lane == 0) {
      const int col0 = ((h * p.hd) >> 6) << 6;          // 64-column block holding this head's q/k columns
      const int k0 = ((h * p.hd) & 63) >> 4;            // first UMMA k-step of the head inside the block
      const int nk = p.hd >> 4;                         // k-steps per head (4 for hd 64, 2 for hd 32)
      mbar_arrive_expect_tx(bar_qk, (uint32_t)(QT_BYTES + p.Sp16 * 128));
      tma_load_2d(sQ, &tmQ, bar_qk, col0, (int)(i * p.q_rows));
      tma_load_2d(sK, &tmK, bar_qk, col0, (int)(i * p.k_rows));
      mbar_arrive_expect_tx(bar_v, (uint32_t)(nkb * 32 * 128));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sV + kb * 4096, &tmV, bar_v, (int)(i * p.v_pitch) + kb * 64, h * 32);
      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      {
        const uint32_t idesc = make_idesc_bf16(128, p.Sp16);
        const uint64_t a_desc = make_smem_desc_sw128(sQ), b_desc = make_smem_desc_sw128(sK);
        for (int k = 0; k < nk; ++k)
          umma_bf16(t_s, a_desc + (uint64_t)(2 * (k0 + k)), b_desc + (uint64_t)(2 * (k0 + k)), idesc, k ? 1u : 0u);
        umma_commit(bar_s);
      }
      // ---- O = P V ----
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      {
        const uint32_t idesc = make_idesc_bf16(128, 32);
        const int ksteps = p.Sp16 >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          const int kb = ks >> 2, k = ks & 3;
          const uint64_t a_desc = make_smem_desc_sw128(sP + kb * 16384) + (uint64_t)(2 * k);
          const uint64_t b_desc = make_smem_desc_sw128(sV + kb * 4096) + (uint64_t)(2 * k);
          umma_bf16(t_o, a_desc, b_desc, idesc, ks ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ---- softmax / epilogue warps 1..4: TMEM lane quarter q = warp % 4, one thread per query row ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    mbar_wait(bar_s, 0);       // S complete => the tensor core has finished reading Q and K: P may overwrite them
    tc_fence_after();
    const int S = p.S, Sp16 = p.Sp16;
    // pass 1: row max of the raw scores over the valid keys
    float mx = -INFINITY;
    for (int c = 0; c < Sp16; c += 32) {
      if (c + 32 <= Sp16) {
        uint32_t r[32];
        tmem_ld32(t_s + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) if (c + e < S) mx = fmaxf(mx, __uint_as_float(r[e]));
      } else {
        uint32_t r[16];
        tmem_ld16(t_s + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) if (c + e < S) mx = fmaxf(mx, __uint_as_float(r[e]));
      }
    }
    // pass 2: p = exp((s - max) * hd^-0.5); unnormalised bf16 P into the swizzled A tile; fp32 row sum
    const float sc = p.scale_log2;
    const float mxs = mx * sc;
    float sum = 0.f;
    const uint32_t p_row = sP + (uint32_t)row * 128u;
    for (int c = 0; c < Sp16; c += 16) {
      uint32_t r[16];
      tmem_ld16(t_s + lane_off + c, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float x = (c + e < S) ? ex2_approx(fmaf(__uint_as_float(r[e]), sc, -mxs)) : 0.f;   // one MUFU.EX2 (exp2f: ~5 instructions)
        // the bf16-rounded value is what the tensor core multiplies: accumulate the same value in the denominator
        x = __bfloat162float(__float2bfloat16_rn(x));
        v[e] = x;
        sum += x;
      }
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        const int cc = c + g8 * 8;
        const uint32_t chunk = (uint32_t)((cc & 63) >> 3) ^ (uint32_t)(row & 7);
        const uint32_t addr = p_row + (uint32_t)(cc >> 6) * 16384u + chunk * 16u;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(v[g8 * 8], v[g8 * 8 + 1])),
                     "r"(pack_bf16(v[g8 * 8 + 2], v[g8 * 8 + 3])), "r"(pack_bf16(v[g8 * 8 + 4], v[g8 * 8 + 5])),
                     "r"(pack_bf16(v[g8 * 8 + 6], v[g8 * 8 + 7]))
                     : "memory");
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // epilogue
    mbar_wait(bar_o, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(t_o + lane_off, r);
    tmem_ld_wait();
    if (row < p.K) {
      const float inv = 1.0f / sum;
      bf16* orow = p.out + (i * p.K + row) * kC + h * 32;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        float bs[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v[8];
        if (p.bv) load8(p.bv + h * 32 + g8 * 8, bs);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaf(__uint_as_float(r[g8 * 8 + e]), inv, bs[e]);
        store8(orow + g8 * 8, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// heads 4-7: one block per actor instance, warp = head.  score[s] = cqp_h[i] . pos0_h[s,b] / 8; o = softmax . V_h
__global__ void __launch_bounds__(128) cls_xattn_pos_kernel(const bf16* __restrict__ cqp, const bf16* __restrict__ pos0,
                                                            const bf16* __restrict__ vt, long ldvt,
                                                            const float* __restrict__ bv, bf16* __restrict__ out, int K,
                                                            int S, int Sq, int BT) {
  extern __shared__ float sm[];
  const long i = blockIdx.x;
  const int bb = (int)(i % BT);
  const int hh = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = sm + hh * S;
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
  float mx = -INFINITY;
#pragma unroll 2
  for (int s = lane; s < S; s += 32) {      // two keys (16 x 16-byte loads) in flight per lane
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
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float e = expf(pr[s] - mx);
    pr[s] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  const int c0 = 128 + hh * 32;
  float my_o = 0.f;   // lane d keeps o[d]
  // eight value channels per step: their (independent) key loops and shuffle reductions overlap instead of paying one full
  // load latency + reduction chain per channel
#pragma unroll 1
  for (int d0 = 0; d0 < 32; d0 += 8) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = lane; s < S; s += 32) {
      const float pv = pr[s];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(pv, __bfloat162float(vt[(long)(c0 + d0 + j) * ldvt + i * Sq + s]), a[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (lane == d0 + j) my_o = a[j] * inv + bv[c0 + d0 + j];
  }
  const bf16 ob = __float2bfloat16_rn(my_o);
#pragma unroll 8
  for (int k = 0; k < K; ++k) out[(i * K + k) * kC + c0 + lane] = ob;
}

int g_attr_bytes = 0;

// Launch attn_tc_kernel.  q: [N*q_rows, 256] rows; kmat: [N*k_rows(+), 256]; vt: [256, ldvt] with per-instance pitch.
int launch_attn_tc(const bf16* q, long q_total_rows, const bf16* kmat, long k_total_rows, const bf16* vt, long ldvt,
                   long v_total_cols, const float* bv, bf16* out, long N, int heads, int hd, int K, int S, int q_rows,
                   int k_rows, int v_pitch, cudaStream_t st) {
  const int Sp16 = (S + 15) & ~15;
  if (K > 128 || K < 1 || Sp16 > 256 || ldvt % 8 != 0 || v_pitch % 8 != 0 || (hd != 64 && hd != 32)) return 1;
  if (tc_num_sms() <= 0) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed");
  const int nkb = (Sp16 + 63) >> 6;
  XaParams p{};
  const uint32_t k_bytes = ((uint32_t)Sp16 * 128u + 1023u) & ~1023u;
  const uint32_t p_bytes = (uint32_t)nkb * 16384u;
  const uint32_t region_a = (QT_BYTES + k_bytes) > p_bytes ? (QT_BYTES + k_bytes) : p_bytes;
  p.off_k = QT_BYTES; p.off_v = region_a; p.off_bar = region_a + (uint32_t)nkb * 4096u;
  const int smem_bytes = (int)p.off_bar + 64 + 1024;
  if (smem_bytes > g_attr_bytes) {
    CQ_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    g_attr_bytes = smem_bytes;
  }
  CUtensorMap tmQ, tmK, tmV;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)q_total_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kC * 2};
    const cuuint32_t box[2] = {64, 128};
    CQ_TRY(make_tmap_bf16(&tmQ, q, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)k_total_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kC * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)Sp16};
    CQ_TRY(make_tmap_bf16(&tmK, kmat, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)v_total_cols, (cuuint64_t)kC};
    const cuuint64_t strides[1] = {(cuuint64_t)ldvt * 2};
    const cuuint32_t box[2] = {64, 32};
    CQ_TRY(make_tmap_bf16(&tmV, vt, 2, dims, strides, box));
  }
  p.out = out; p.bv = bv; p.K = K; p.S = S; p.Sp16 = Sp16; p.q_rows = q_rows; p.k_rows = k_rows; p.v_pitch = v_pitch;
  p.hd = hd; p.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int cols = 32;
  while (cols < Sp16 + 32) cols <<= 1;
  p.tmem_cols = cols; p.o_col = cols - 32;
  dim3 grid((unsigned)heads, (unsigned)N);
  attn_tc_kernel<<<grid, NUM_THREADS, smem_bytes, st>>>(tmQ, tmK, tmV, p);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Class cross-attention (dab_transformer.py:1067-1071).  Returns 1 when the shape is outside what the tensor-core kernel
// takes (caller uses the CUDA-core kernel).
int cls_xattn_tc(const bf16* Qin, const bf16* cqp, const bf16* kx, const bf16* pos0, const bf16* vt, long ldvt,
                 const float* bv, bf16* out, long N, int K, int S, int Sq, int Sp_rows, int BT, cudaStream_t st) {
  int r = launch_attn_tc(Qin, N * K, kx, N * Sp_rows, vt, ldvt, N * Sq, bv, out, N, 4, 64, K, S, K, Sp_rows, Sq, st);
  if (r != 0) return r;
  cls_xattn_pos_kernel<<<(unsigned)N, 128, 4 * S * sizeof(float), st>>>(cqp, pos0, vt, ldvt, bv, out, K, S, Sq, BT);
  CQ_LAUNCH_CHECK();
  return 0;
}

// Class-query self-attention (dab_transformer.py:1063): q = k = v = class tokens x [N*K,256], 8 heads of 32;
// xt = x^T [256, ldxt] with per-instance column pitch K8 (written by the class-FFN epilogue).  Output [N*K,256].
int cls_sattn_tc(const bf16* x, const bf16* xt, long ldxt, bf16* out, long N, int K, int K8, cudaStream_t st) {
  return launch_attn_tc(x, N * K, x, N * K, xt, ldxt, N * K8, nullptr, out, N, 8, 32, K, K, K, K, K8, st);
}

}  // namespace cqvad
