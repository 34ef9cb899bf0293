// Fused two-layer MLP on tcgen05 (sm_100a):   Y = LN?( res + act(X.W1^T + b1).W2^T + b2 )
//   X [M,256] bf16, W1 [F,256], W2 [256,F], Y [M,256];  F % 128 == 0.
// Used for ConvBlock.conv2/GELU/conv3 (+input) (dab_transformer.py:93-97; F = 1024, M = N*(h+1)*w) and for every FFN
// of the decoder (dab_transformer.py:994-996, 1043-1045, 1074-1076; F = 2048).  The 4x/8x-wide hidden activation
// never touches HBM: per 128-row tile the hidden dimension is walked in 128-column chunks
//     GEMM1(j):  Hacc[j%2] (TMEM, 128 cols)  = X(smem) . W1_j^T            16 x UMMA 128x128x16
//     epi(j)  :  Hs[j%2] (smem, 128-B swizzle) = bf16(act(Hacc + b1_j))     epilogue group j%2 (4 warps)
//     GEMM2(j):  Yacc (TMEM, 256 cols)      += Hs[j%2] . W2_j^T             8 x UMMA 128x256x16
// with GEMM2(j-1) issued after GEMM1(j) so the tensor pipe works while the epilogue of chunk j runs.
// Warps: 0 = TMA producer (X tile + 3-slot x 32 KB weight ring), 1 = MMA issuer + TMEM allocator, 2-5 = epilogue
// group 0, 6-9 = epilogue group 1.  Both groups share the final Y epilogue (128 columns each; LayerNorm statistics
// are exchanged through shared memory).
#include "common.cuh"
#include "tc_common.cuh"

namespace cqvad {

using namespace tc;
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int tc_num_sms();

namespace {

constexpr int BM = 128, CH = 128 /*hidden chunk*/, BK = 64, C = 256;
constexpr int X_BYTES = BM * C * 2;            // 64 KB: 4 k-blocks of [128 x 64]
constexpr int HS_BYTES = BM * CH * 2;          // 32 KB: 2 k-blocks of [128 x 64]
constexpr int SLOT_BYTES = 32 * 1024;          // W1: 2 k-blocks of [128 x 64]; W2: 1 k-block of [256 x 64]
constexpr int SLOTS = 3;
constexpr int SMEM_DATA = X_BYTES + 2 * HS_BYTES + SLOTS * SLOT_BYTES;   // 229 376
constexpr int SMEM_BYTES = SMEM_DATA + 1024 + 1280;
constexpr int NUM_THREADS = 320;

struct MlpParams {
  bf16* Y; long M; int F;
  const float* b1; const float* b2; int act;
  const bf16* res;
  const float* ln_g; const float* ln_b; float ln_eps;
  int zero_period, zero_valid;
  int m_tiles;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base, sH = base + X_BYTES, sW = sH + 2 * HS_BYTES;
  const uint32_t bars = base + SMEM_DATA;
  // barriers (8 bytes each)
  const uint32_t w_full = bars, w_empty = bars + 8 * SLOTS;          // weight ring
  const uint32_t x_full = bars + 16 * SLOTS, x_empty = x_full + 8;   // X tile
  const uint32_t hacc_full = x_empty + 8, hacc_empty = hacc_full + 16;   // TMEM H accumulators [2]
  const uint32_t hs_full = hacc_empty + 16, hs_empty = hs_full + 16;     // smem H buffers [2]
  const uint32_t y_full = hs_empty + 16, y_empty = y_full + 8;
  const uint32_t tmem_slot = y_empty + 8;
  const uint32_t stats = tmem_slot + 8;                                   // float[2][128][2] LN partial sums (2 KB.. uses 1024 B)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = p.F / CH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int s = 0; s < SLOTS; ++s) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 1); }
    mbar_init(x_full, 1); mbar_init(x_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(hacc_full + 8 * g, 1); mbar_init(hacc_empty + 8 * g, 4);
      mbar_init(hs_full + 8 * g, 4); mbar_init(hs_empty + 8 * g, 1);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, 8);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t t_y = tmem_base, t_h0 = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int slot = 0; uint32_t wphase = 0; uint32_t xphase = 0;
      auto load_w1 = [&](int j, int half) {   // k-blocks 2*half, 2*half+1 of W1 chunk j: two boxes [64 x 128 rows]
        mbar_wait(w_empty + 8 * slot, wphase ^ 1);
        const uint32_t fb = w_full + 8 * slot;
        mbar_arrive_expect_tx(fb, SLOT_BYTES);
        tma_load_2d(sW + slot * SLOT_BYTES, &tmW1, fb, (2 * half) * BK, j * CH);
        tma_load_2d(sW + slot * SLOT_BYTES + 16384, &tmW1, fb, (2 * half + 1) * BK, j * CH);
        if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
      };
      auto load_w2 = [&](int j, int kk) {     // W2[:, j*128 + kk*64 .. +64): one box [64 x 256 rows]
        mbar_wait(w_empty + 8 * slot, wphase ^ 1);
        const uint32_t fb = w_full + 8 * slot;
        mbar_arrive_expect_tx(fb, SLOT_BYTES);
        tma_load_2d(sW + slot * SLOT_BYTES, &tmW2, fb, j * CH + kk * BK, 0);
        if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * 16384, &tmX, x_full, kb * BK, tile * BM);
        xphase ^= 1;
        for (int j = 0; j < nch; ++j) {
          load_w1(j, 0); load_w1(j, 1);
          if (j >= 1) { load_w2(j - 1, 0); load_w2(j - 1, 1); }
        }
        load_w2(nch - 1, 0); load_w2(nch - 1, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc_bf16(BM, CH);    // 128 x 128
      constexpr uint32_t idesc2 = make_idesc_bf16(BM, C);     // 128 x 256
      int slot = 0; uint32_t wphase = 0, xphase = 0, yphase = 0;
      uint32_t hacc_ph[2] = {0, 0}, hs_ph[2] = {0, 0};
      bool y_started = false;
      auto gemm2 = [&](int c) {
        const int g = c & 1;
        if (!y_started) {   // first GEMM2 of the tile: the previous tile's Y epilogue must have drained TMEM
          mbar_wait(y_empty, yphase ^ 1);
          tc_fence_after();
        }
        mbar_wait(hs_full + 8 * g, hs_ph[g]);
        hs_ph[g] ^= 1;
        tc_fence_after();
        for (int kk = 0; kk < 2; ++kk) {
          mbar_wait(w_full + 8 * slot, wphase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(sH + g * HS_BYTES + kk * 16384);
          const uint64_t b_desc = make_smem_desc_sw128(sW + slot * SLOT_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(t_y, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc2, (y_started || kk || k) ? 1u : 0u);
          umma_commit(w_empty + 8 * slot);
          if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
        }
        y_started = true;
        umma_commit(hs_empty + 8 * g);
      };
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(x_full, xphase);
        xphase ^= 1;
        tc_fence_after();
        y_started = false;
        for (int j = 0; j < nch; ++j) {
          const int g = j & 1;
          mbar_wait(hacc_empty + 8 * g, hacc_ph[g] ^ 1);
          hacc_ph[g] ^= 1;
          tc_fence_after();
          const uint32_t t_h = t_h0 + g * CH;
          for (int half = 0; half < 2; ++half) {
            mbar_wait(w_full + 8 * slot, wphase);
            tc_fence_after();
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
              const int kb = 2 * half + t2;
              const uint64_t a_desc = make_smem_desc_sw128(sX + kb * 16384);
              const uint64_t b_desc = make_smem_desc_sw128(sW + slot * SLOT_BYTES + t2 * 16384);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(t_h, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc1, (kb || k) ? 1u : 0u);
            }
            umma_commit(w_empty + 8 * slot);
            if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
          }
          umma_commit(hacc_full + 8 * g);
          if (j == nch - 1) umma_commit(x_empty);   // all GEMM1 of this tile issued: X may be overwritten when they finish
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(nch - 1);
        umma_commit(y_full);
        yphase ^= 1;
      }
    }
  } else {
    // ===== epilogue groups: g = 0 (warps 2-5), g = 1 (warps 6-9) =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t hacc_ph = 0, hs_ph = 0, yphase = 0;
    float* stats_f = reinterpret_cast<float*>(smem_raw + (stats - smem_u32(smem_raw)));
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int j = g; j < nch; j += 2) {
        mbar_wait(hacc_full + 8 * g, hacc_ph);
        hacc_ph ^= 1;
        tc_fence_after();
        mbar_wait(hs_empty + 8 * g, hs_ph ^ 1);   // GEMM2 of chunk j-2 has finished reading Hs[g]
        hs_ph ^= 1;
        const uint32_t t_h = t_h0 + g * CH + lane_off;
        const uint32_t hs_row = sH + g * HS_BYTES + row_in_tile * 128;
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_h + c, r);
          tmem_ld_wait();
          const float* b1 = p.b1 + j * CH + c;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float bs[8];
            load8(b1 + g8 * 8, bs);
            uint32_t pk[4];
            if (p.act == CQVAD_ACT_GELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint64_t x2 = f2add(f2pack(__uint_as_float(r[g8 * 8 + 2 * e]), __uint_as_float(r[g8 * 8 + 2 * e + 1])),
                                          f2pack(bs[2 * e], bs[2 * e + 1]));
                pk[e] = f2_to_bf16x2(gelu2(x2));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                pk[e] = pack_bf16(fmaxf(__uint_as_float(r[g8 * 8 + 2 * e]) + bs[2 * e], 0.f),
                                  fmaxf(__uint_as_float(r[g8 * 8 + 2 * e + 1]) + bs[2 * e + 1], 0.f));
            }
            // hidden column cc = c + g8*8 -> k-block cc/64, 16-byte chunk (cc%64)/8, swizzled with (row & 7)
            const int cc = c + g8 * 8;
            const uint32_t chunk = (uint32_t)((cc & 63) >> 3) ^ (uint32_t)(row_in_tile & 7);
            const uint32_t addr = hs_row + (uint32_t)(cc >> 6) * 16384u + chunk * 16u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();   // make the st.shared above visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(hacc_empty + 8 * g); mbar_arrive(hs_full + 8 * g); }
      }
      // ---- final epilogue of the tile: group g owns output columns [g*128, g*128+128) ----
      mbar_wait(y_full, yphase);
      yphase ^= 1;
      tc_fence_after();
      const long grow = (long)tile * BM + row_in_tile;
      const bool row_ok = grow < p.M;
      const bool zero_row = p.zero_period > 0 && (int)(grow % p.zero_period) >= p.zero_valid;
      const uint32_t t_yrow = t_y + lane_off + g * 128;
      const int col0 = g * 128;
      const bool do_ln = p.ln_g != nullptr;
      float mean = 0.f, rstd = 1.f;
      if (do_ln) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_yrow + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float bs[8], rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            load8(p.b2 + col0 + c + g8 * 8, bs);
            if (p.res && row_ok) load8(p.res + grow * C + col0 + c + g8 * 8, rs);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float v = __uint_as_float(r[g8 * 8 + e]) + bs[e] + rs[e];
              s1 += v; s2 = fmaf(v, v, s2);
              r[g8 * 8 + e] = __float_as_uint(v);
            }
          }
          tmem_st32(t_yrow + c, r);
        }
        tmem_st_wait();
        stats_f[(g * 128 + row_in_tile) * 2] = s1;
        stats_f[(g * 128 + row_in_tile) * 2 + 1] = s2;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the two epilogue groups (8 warps)
        const float o1 = stats_f[((g ^ 1) * 128 + row_in_tile) * 2], o2 = stats_f[((g ^ 1) * 128 + row_in_tile) * 2 + 1];
        mean = (s1 + o1) * (1.0f / C);
        const float var = fmaxf((s2 + o2) * (1.0f / C) - mean * mean, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
      }
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_yrow + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int cb = col0 + c + g8 * 8;
          float v[8];
          if (do_ln) {
            float gm[8], bt[8];
            load8(p.ln_g + cb, gm);
            load8(p.ln_b + cb, bt);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (__uint_as_float(r[g8 * 8 + e]) - mean) * rstd * gm[e] + bt[e];
          } else {
            float bs[8], rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            load8(p.b2 + cb, bs);
            if (p.res && row_ok) load8(p.res + grow * C + cb, rs);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g8 * 8 + e]) + bs[e] + rs[e];
          }
          if (zero_row) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
          }
          if (row_ok) store8(p.Y + grow * C + cb, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_empty);
      if (do_ln) asm volatile("bar.sync 1, 256;" ::: "memory");   // stats buffer reuse across tiles
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool g_attr_set = false;

}  // namespace

int mlp_tc(const bf16* X, const bf16* W1, const float* b1, const bf16* W2, const float* b2, int act, const bf16* res,
           const float* ln_g, const float* ln_b, float ln_eps, bf16* Y, long M, int Cc, int F, int zero_period,
           int zero_valid, cudaStream_t st) {
  if (Cc != C || F % CH != 0 || F < 2 * CH || M < 1) return 1;
  if (act != CQVAD_ACT_RELU && act != CQVAD_ACT_GELU) return 1;
  if ((((uintptr_t)X) & 15) || (((uintptr_t)Y) & 15) || (((uintptr_t)W1) & 15) || (((uintptr_t)W2) & 15) ||
      (res && (((uintptr_t)res) & 15)))
    return 1;
  const int sms = tc_num_sms();
  if (sms <= 0) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed");
  if (!g_attr_set) {
    CQ_CUDA(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_attr_set = true;
  }
  CUtensorMap tmX, tmW1, tmW2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {BK, BM};
    CQ_TRY(make_tmap_bf16(&tmX, X, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)F};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {BK, CH};
    CQ_TRY(make_tmap_bf16(&tmW1, W1, 2, dims, strides, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)C};
    const cuuint64_t strides[1] = {(cuuint64_t)F * 2};
    const cuuint32_t box[2] = {BK, C};
    CQ_TRY(make_tmap_bf16(&tmW2, W2, 2, dims, strides, box));
  }
  MlpParams p{};
  p.Y = Y; p.M = M; p.F = F; p.b1 = b1; p.b2 = b2; p.act = act; p.res = res;
  p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps; p.zero_period = zero_period; p.zero_valid = zero_valid;
  p.m_tiles = (int)((M + BM - 1) / BM);
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  mlp_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmX, tmW1, tmW2, p);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace cqvad
