// Fused two-layer MLP on tcgen05 (sm_100a):   Y = LN?( res + act(X.W1^T + b1).W2^T + b2 )
//   X [M,256] bf16, W1 [F,256], W2 [256,F], Y [M,256];  F % 128 == 0.
// Used for ConvBlock.conv2/GELU/conv3 (+input) (dab_transformer.py:93-97; F = 1024, M = N*(h+1)*w) and for every FFN
// of the decoder (dab_transformer.py:994-996, 1043-1045, 1074-1076; F = 2048).  The 4x/8x-wide hidden activation
// never touches HBM: per 128-row tile the hidden dimension is walked in 128-column chunks
//     GEMM1(j):  Hacc[j%2] (TMEM, 128 cols)  = X(smem) . W1_j^T            16 x UMMA 128x128x16
//     epi(j)  :  Hs[j%2] (smem, 128-B swizzle) = bf16(act(Hacc + b1_j))     epilogue group j%2 (4 warps)
//     GEMM2(j):  Yacc (TMEM, 256 cols)      += Hs[j%2] . W2_j^T             8 x UMMA 128x256x16
// with GEMM2(j-2) issued after GEMM1(j) (four H buffers in flight) so the tensor pipe has independent work queued while
// the epilogue of a chunk runs.
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

constexpr int BM = 128, CH = 64 /*hidden chunk*/, BK = 64, C = 256;
constexpr int NB = 4;                          // H buffers in flight (TMEM accumulators and smem tiles)
constexpr int LA = 3;                          // GEMM2(j-LA) is issued after GEMM1(j): three independent GEMM1s stay queued
constexpr int X_BYTES = BM * C * 2;            // 64 KB: 4 k-blocks of [128 x 64]
constexpr int HS_BYTES = BM * CH * 2;          // 16 KB: one k-block [128 x 64]
constexpr int SLOT_BYTES = 32 * 1024;          // W1 chunk: 4 k-blocks of [64 x 64]; W2 chunk: one k-block of [256 x 64]
constexpr int SLOTS = 3;
constexpr int SMEM_DATA = X_BYTES + NB * HS_BYTES + SLOTS * SLOT_BYTES;   // 229 376
constexpr int SMEM_AUX = 256 /*barriers*/ + 2048 /*b1 chunk, double buffered: [2 groups][2][64] f32; the LN statistics [2][128][2] f32 of the
                                                      Y epilogue alias it (the chunk epilogues are over by then)*/ + 64 /*residual barriers*/;
constexpr int SMEM_BYTES = SMEM_DATA + SMEM_AUX;   // <= 227 KB; the dynamic buffer is declared 1024-aligned (no slack)
constexpr int NUM_THREADS = 320;

struct MlpParams {
  bf16* Y; long M; int F;
  const float* b1; const float* b2; int act;
  const bf16* res; const float* res32; float* Y32;
  const float* ln_g; const float* ln_b; float ln_eps;
  int zero_period, zero_valid;
  int m_tiles;
  int ts_res;         // with ts_out: the bf16 residual arrives as TMA boxes in the same staging (in-place epilogue)
  int ts_out;         // Y leaves through shared-memory staging + TMA stores (plain bf16 output: no Y32 / YT copies)
  long long* trace;   // dev tool (CQVAD_MLP_TRACE = device pointer to int64[4096]): pipeline timestamps of CTA 0, tools/trace_mlp.py
  bf16* YT; long ldyt; int yt_rows, yt_pitch;   // optional transposed copy: YT[c][(row/yt_rows)*yt_pitch + row%yt_rows]
};

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ long long mlp_gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MLP_TRACE(slot, idx, lim) do { if (p.trace && blockIdx.x == 0 && (idx) < (lim)) p.trace[(slot) + (idx)] = mlp_gtimer(); } while (0)

// SPEC: 0 = every option at run time (fp32 residual / fp32 + transposed copies of the class FFN, direct stores), 1 = GELU, no LN,
// TMA output (ConvBlock), 2 = ReLU + LN, TMA output (the FFNs).  The all-options kernel is 5 200 SASS instructions (84 KB); what a
// launch does not use is compiled out of its specialisation (the same lesson as the GEMM's lean / dual-GELU instantiations).
template <int SPEC>
__device__ __forceinline__ void mlp_tc_body(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmY,
                                            const CUtensorMap& tmR, const MlpParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);   // 1024-byte aligned (128-B swizzle atoms)
  const uint32_t sX = base, sH = base + X_BYTES, sW = sH + NB * HS_BYTES;
  const uint32_t bars = base + SMEM_DATA;
  // barriers (8 bytes each)
  const uint32_t w_full = bars, w_empty = bars + 8 * SLOTS;            // weight ring            [0,48)
  const uint32_t x_full = bars + 48, x_empty = bars + 56;              // X tile                 [48,64)
  const uint32_t hacc_full = bars + 64, hacc_empty = bars + 96;        // TMEM H accumulators [4] [64,128)
  const uint32_t hs_full = bars + 128, hs_empty = bars + 160;          // smem H tiles [4]        [128,192)
  const uint32_t y_full = bars + 192, y_empty = bars + 200;
  const uint32_t tmem_slot = bars + 208;
  const uint32_t b1s = bars + 256;                                     // float[2 groups][2][64]: b1 of the chunk in flight
  // LN partial sums float[2][128][2] alias the b1 staging buffers (+1 KB): no chunk epilogue runs during the Y epilogue
  const uint32_t stats = b1s;
  const uint32_t rbar_base = bars + 256 + 2048;                        // one per epilogue warp: the residual boxes have landed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = p.F / CH;
  // (a per-CTA rotation of the chunk order was tried against L2 hot-spotting: no gain on B200, and it makes the fp32
  //  summation order depend on the CTA index, breaking bit-exact batch independence -> chunks are walked in order)
  constexpr int rot = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int s = 0; s < SLOTS; ++s) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 1); }
    mbar_init(x_full, 1); mbar_init(x_empty, 1);
    for (int b = 0; b < NB; ++b) {
      mbar_init(hacc_full + 8 * b, 1); mbar_init(hacc_empty + 8 * b, 4);
      mbar_init(hs_full + 8 * b, 4); mbar_init(hs_empty + 8 * b, 1);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, 8);
    for (int i = 0; i < 8; ++i) mbar_init(rbar_base + 8 * i, 1);
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
      // ===== TMA producer: loads in exactly the order the MMA warp consumes them =====
      int slot = 0; uint32_t wphase = 0; uint32_t xphase = 0;
      int tev = 0;
      auto load_w1 = [&](int j) {   // W1 rows [j*64, j*64+64): four k-block boxes [64 k x 64 rows] of 8 KB
        mbar_wait(w_empty + 8 * slot, wphase ^ 1);
        MLP_TRACE(0, tev, 256); ++tev;
        const uint32_t fb = w_full + 8 * slot;
        mbar_arrive_expect_tx(fb, SLOT_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sW + slot * SLOT_BYTES + kb * 8192, &tmW1, fb, kb * BK, j * CH);
        if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
      };
      auto load_w2 = [&](int j) {   // W2[:, j*64 .. j*64+64): one box [64 k x 256 rows]
        mbar_wait(w_empty + 8 * slot, wphase ^ 1);
        const uint32_t fb = w_full + 8 * slot;
        mbar_arrive_expect_tx(fb, SLOT_BYTES);
        tma_load_2d(sW + slot * SLOT_BYTES, &tmW2, fb, j * CH, 0);
        if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, X_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * 16384, &tmX, x_full, kb * BK, tile * BM);
        xphase ^= 1;
        for (int j = 0; j < nch; ++j) {
          load_w1((j + rot) % nch);
          if (j >= LA) load_w2((j - LA + rot) % nch);
        }
        for (int c = nch - LA; c < nch; ++c) load_w2((c + rot) % nch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc_bf16(BM, CH);    // 128 x 64
      constexpr uint32_t idesc2 = make_idesc_bf16(BM, C);     // 128 x 256
      int slot = 0; uint32_t wphase = 0, xphase = 0, yphase = 0;
      uint32_t hacc_ph = 0, hs_ph = 0;    // one phase bit per buffer
      bool y_started = false;
      int e1 = 0, e2 = 0;
      auto gemm2 = [&](int c) {
        const int b = c & (NB - 1);
        if (!y_started) {   // first GEMM2 of the tile: the previous tile's Y epilogue must have drained TMEM
          mbar_wait(y_empty, yphase ^ 1);
          tc_fence_after();
        }
        mbar_wait(hs_full + 8 * b, (hs_ph >> b) & 1u);
        MLP_TRACE(1024, e2, 256);
        hs_ph ^= 1u << b;
        mbar_wait(w_full + 8 * slot, wphase);
        MLP_TRACE(1280, e2, 256); ++e2;
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(sH + b * HS_BYTES);
        const uint64_t b_desc = make_smem_desc_sw128(sW + slot * SLOT_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_y, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc2, (y_started || k) ? 1u : 0u);
        umma_commit(w_empty + 8 * slot);
        if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
        y_started = true;
        umma_commit(hs_empty + 8 * b);
      };
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(x_full, xphase);
        xphase ^= 1;
        tc_fence_after();
        y_started = false;
        for (int j = 0; j < nch; ++j) {
          const int b = j & (NB - 1);
          MLP_TRACE(256, e1, 256);
          mbar_wait(hacc_empty + 8 * b, ((hacc_ph >> b) & 1u) ^ 1u);   // epilogue of chunk j-4 has drained TMEM buffer b
          MLP_TRACE(512, e1, 256);
          hacc_ph ^= 1u << b;
          mbar_wait(w_full + 8 * slot, wphase);
          MLP_TRACE(768, e1, 256); ++e1;
          tc_fence_after();
          const uint32_t t_h = t_h0 + b * CH;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t a_desc = make_smem_desc_sw128(sX + kb * 16384);
            const uint64_t b_desc = make_smem_desc_sw128(sW + slot * SLOT_BYTES + kb * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(t_h, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc1, (kb || k) ? 1u : 0u);
          }
          umma_commit(w_empty + 8 * slot);
          if (++slot == SLOTS) { slot = 0; wphase ^= 1; }
          umma_commit(hacc_full + 8 * b);
          if (j == nch - 1) umma_commit(x_empty);   // all GEMM1 of this tile issued: X may be overwritten when they finish
          if (j >= LA) gemm2(j - LA);
        }
        for (int c = nch - LA; c < nch; ++c) gemm2(c);
        umma_commit(y_full);
        yphase ^= 1;
      }
    }
  } else {
    // ===== epilogue groups: g = 0 (warps 2-5) takes even chunks, g = 1 (warps 6-9) odd chunks =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t hacc_ph = 0, hs_ph = 0, yphase = 0;  // bit b = phase of buffer b (this group touches b = g and g+2)
    float* stats_f = reinterpret_cast<float*>(smem_raw + (stats - base));
    float* b1s_f = reinterpret_cast<float*>(smem_raw + (b1s - base)) + g * 128;   // this group's two 64-float buffers
    const int tg = ((warp - 2) & 3) * 32 + lane;                                   // thread index inside the group
    // b1 of the group's next chunk is prefetched into a register one chunk ahead, published through shared memory and
    // read back as broadcast LDS.128 (a global load per 8 columns left the epilogue exposed to the full L2 latency:
    // ncu source page of the first version, profiles/)
    float nb = tg < CH ? p.b1[((g + rot) % nch) * CH + tg] : 0.f;
    uint32_t bpar = 0, rphase = 0;
    int ee = 0;
    const bool trw = lane == 0 && (warp == 2 || warp == 6);
    const int tbase = 1536 + g * 1024;          // group 0: [1536, 2560), group 1: [2560, 3584)
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int j = g; j < nch; j += 2) {
        const int b = j & (NB - 1);
        float* bcur = b1s_f + bpar * 64;
        if (tg < CH) bcur[tg] = nb;
        bpar ^= 1;
        asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
        if (tg < CH) nb = p.b1[((((j + 2 < nch) ? (j + 2) : g) + rot) % nch) * CH + tg];
        if (trw) MLP_TRACE(tbase, ee, 256);
        mbar_wait(hacc_full + 8 * b, (hacc_ph >> b) & 1u);
        if (trw) MLP_TRACE(tbase + 256, ee, 256);
        hacc_ph ^= 1u << b;
        tc_fence_after();
        mbar_wait(hs_empty + 8 * b, ((hs_ph >> b) & 1u) ^ 1u);   // GEMM2 of chunk j-4 has finished reading Hs[b]
        hs_ph ^= 1u << b;
        const uint32_t t_h = t_h0 + b * CH + lane_off;
        const uint32_t hs_row = sH + b * HS_BYTES + row_in_tile * 128;
#pragma unroll
        for (int c = 0; c < CH; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_h + c, r);
          tmem_ld_wait();
          const float* b1 = bcur + c;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float bs[8];
            load8(b1 + g8 * 8, bs);
            uint32_t pk[4];
            if (SPEC == 1 || (SPEC == 0 && p.act == CQVAD_ACT_GELU)) {
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
            // hidden column cc = c + g8*8 of this 64-wide k-block: 16-byte chunk cc/8, swizzled with (row & 7)
            const int cc = c + g8 * 8;
            const uint32_t chunk = (uint32_t)(cc >> 3) ^ (uint32_t)(row_in_tile & 7);
            const uint32_t addr = hs_row + chunk * 16u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();   // make the st.shared above visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(hacc_empty + 8 * b); mbar_arrive(hs_full + 8 * b); }
        if (trw) { MLP_TRACE(tbase + 512, ee, 256); ++ee; }
      }
      const int tix = tile / (int)gridDim.x * 8;
      if (trw) { MLP_TRACE(tbase + 768, tix, 256); }
      // ---- final epilogue of the tile: group g owns output columns [g*128, g*128+128) ----
      const long grow = (long)tile * BM + row_in_tile;
      const bool row_ok = grow < p.M;
      // The bf16 residual is fetched one 32-column step ahead (4 x 16 bytes in flight per thread): one exposed L2 round
      // trip per step instead of one per 8 columns (ncu: 18% of all stall samples sat on these loads while Y blocked the
      // next tile).  Hoisting all 16 loads above the y_full wait was tried and lost to register spills.
      const bool res_bf16 = SPEC == 0 && p.res != nullptr && p.res32 == nullptr && row_ok && !p.ts_res;
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + (res_bf16 ? grow * C + g * 128 : 0));
      uint4 rnext[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rnext[u] = res_bf16 ? rp[u] : make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(y_full, yphase);
      yphase ^= 1;
      tc_fence_after();
      if (trw) { MLP_TRACE(tbase + 768, tix + 1, 256); }
      // TMA-store epilogue: this warp's rows [q*32, +32) x its group's 128 columns = two [32 x 64] bf16 boxes (128-byte rows, 128-byte
      // swizzle) staged in the H tiles, which are idle from y_full until the next tile's first chunk epilogue; the bf16 residual is
      // fetched into the same boxes (one exposed TMA round trip per tile) and the result overwrites it in place.  One row per
      // thread straight from / to global memory cost 32 L1 wavefronts per 16-byte instruction and one L2 round trip per 32-column
      // step (measured: 8.9 us of a 31 us tile).
      const uint32_t ystg = sH + (uint32_t)((warp - 2) * 8192);
      const uint32_t ymy = ystg + (uint32_t)(lane * 128);
      const int ysw = lane & 7;
      if (p.ts_res) {
        const uint32_t rb = rbar_base + (uint32_t)((warp - 2) * 8);
        if (lane == 0) {
          mbar_arrive_expect_tx(rb, 8192);
          tma_load_2d(ystg, &tmR, rb, g * 128, tile * BM + q * 32);
          tma_load_2d(ystg + 4096, &tmR, rb, g * 128 + 64, tile * BM + q * 32);
        }
        mbar_wait(rb, rphase);
        rphase ^= 1;
        if (trw) { MLP_TRACE(tbase + 768, tix + 5, 256); }
      }
      const bool zero_row = p.zero_period > 0 && (int)(grow % p.zero_period) >= p.zero_valid;
      const uint32_t t_yrow = t_y + lane_off + g * 128;
      const int col0 = g * 128;
      const bool do_ln = SPEC == 2 || (SPEC == 0 && p.ln_g != nullptr);
      float mean = 0.f, rstd = 1.f;
      // address of the 16-byte chunk holding columns [c + 8 g8, +8) of this thread's row inside the warp's two staging boxes
      auto ychunk = [&](int c, int g8) { return ymy + (uint32_t)((c >> 6) * 4096) + (uint32_t)(((((c & 63) >> 3) + g8) ^ ysw) << 4); };
      if (do_ln) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint4 rcur[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) rcur[u] = rnext[u];
          if (c + 32 < 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rnext[u] = res_bf16 ? rp[((c + 32) >> 3) + u] : make_uint4(0u, 0u, 0u, 0u);
          }
          uint32_t r[32];
          tmem_ld32(t_yrow + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float bs[8], rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            load8(p.b2 + col0 + c + g8 * 8, bs);
            if (p.ts_res) unpack8(lds128(ychunk(c, g8)), rs);
            else if (SPEC == 0 && p.res32) { if (row_ok) load8(p.res32 + grow * C + col0 + c + g8 * 8, rs); }
            else if (SPEC == 0) unpack8(rcur[g8], rs);
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
        if (trw) { MLP_TRACE(tbase + 768, tix + 2, 256); }
        stats_f[(g * 128 + row_in_tile) * 2] = s1;
        stats_f[(g * 128 + row_in_tile) * 2 + 1] = s2;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the two epilogue groups (8 warps)
        const float o1 = stats_f[((g ^ 1) * 128 + row_in_tile) * 2], o2 = stats_f[((g ^ 1) * 128 + row_in_tile) * 2 + 1];
        mean = (s1 + o1) * (1.0f / C);
        const float var = fmaxf((s2 + o2) * (1.0f / C) - mean * mean, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
        if (trw) { MLP_TRACE(tbase + 768, tix + 3, 256); }
      }
      if (SPEC != 0 || p.ts_out) {
        // lean store pass (plain bf16 output): no per-output branches or 64-bit address arithmetic in the inner loop
        const float zm = zero_row ? 0.f : 1.f;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t r[32];
          tmem_ld32(t_yrow + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int cb = col0 + c + g8 * 8;
            const uint32_t addr = ychunk(c, g8);
            float v[8];
            if (do_ln) {
              float gm[8], bt[8];
              load8(p.ln_g + cb, gm);
              load8(p.ln_b + cb, bt);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = ((__uint_as_float(r[g8 * 8 + e]) - mean) * rstd * gm[e] + bt[e]) * zm;
            } else {
              float bs[8], rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              load8(p.b2 + cb, bs);
              if (p.ts_res) unpack8(lds128(addr), rs);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = (__uint_as_float(r[g8 * 8 + e]) + bs[e] + rs[e]) * zm;
            }
            uint4 o;
            o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
            sts128(addr, o);
          }
          if (c & 32) {     // a 64-column box is complete
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&tmY, ystg + (uint32_t)((c >> 6) * 4096), col0 + (c >> 6) * 64, tile * BM + q * 32); bulk_commit(); }
          }
        }
      } else if constexpr (SPEC == 0) {
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint4 rcur[4];
        if (!do_ln) {
#pragma unroll
          for (int u = 0; u < 4; ++u) rcur[u] = rnext[u];
          if (c + 32 < 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rnext[u] = res_bf16 ? rp[((c + 32) >> 3) + u] : make_uint4(0u, 0u, 0u, 0u);
          }
        }
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
            if (p.res32) { if (row_ok) load8(p.res32 + grow * C + cb, rs); }
            else unpack8(rcur[g8], rs);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g8 * 8 + e]) + bs[e] + rs[e];
          }
          if (zero_row) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
          }
          if (row_ok) {
            store8(p.Y + grow * C + cb, v);
            if (p.Y32) store8(p.Y32 + grow * C + cb, v);
            if (p.YT) {   // lanes = consecutive rows -> consecutive columns of YT: coalesced 2-byte stores
              const long col = (grow / p.yt_rows) * p.yt_pitch + grow % p.yt_rows;
#pragma unroll
              for (int e = 0; e < 8; ++e) p.YT[(long)(cb + e) * p.ldyt + col] = __float2bfloat16_rn(v[e]);
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_empty);
      if (trw) { MLP_TRACE(tbase + 768, tix + 4, 256); }
      if (SPEC != 0 || p.ts_out) {   // the stores have read the staging boxes before any warp's next chunk epilogue writes the H tiles
        if (lane == 0) bulk_wait_read<0>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if (do_ln) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // stats buffer reuse across tiles
      }
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

template <int SPEC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
              const __grid_constant__ CUtensorMap tmR, const MlpParams p) {
  mlp_tc_body<SPEC>(tmX, tmW1, tmW2, tmY, tmR, p);
}

bool g_attr_set = false;

}  // namespace

int mlp_tc(const bf16* X, const bf16* W1, const float* b1, const bf16* W2, const float* b2, int act, const bf16* res,
           const float* ln_g, const float* ln_b, float ln_eps, bf16* Y, long M, int Cc, int F, int zero_period,
           int zero_valid, cudaStream_t st, bf16* YT, long ldyt, int yt_rows, int yt_pitch, const float* res32, float* Y32) {
  if (Cc != C || F % CH != 0 || F < 256 || M < 1) return 1;
  if (act != CQVAD_ACT_RELU && act != CQVAD_ACT_GELU) return 1;
  if ((((uintptr_t)X) & 15) || (((uintptr_t)Y) & 15) || (((uintptr_t)W1) & 15) || (((uintptr_t)W2) & 15) ||
      (res && (((uintptr_t)res) & 15)))
    return 1;
  const int sms = tc_num_sms();
  if (sms <= 0) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed");
  if (!g_attr_set) {
    CQ_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CQ_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CQ_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    g_attr_set = true;
  }
  CUtensorMap tmX, tmW1, tmW2, tmY, tmR;
  static const bool no_ts_out = getenv("CQVAD_MLP_NO_TS") != nullptr;
  const bool ts_out = !no_ts_out && !YT && !Y32;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {64, 32};
    CQ_TRY(make_tmap_bf16(&tmY, Y, 2, dims, strides, box));
    tmR = tmY;
    if (ts_out && res && !res32) CQ_TRY(make_tmap_bf16(&tmR, res, 2, dims, strides, box));
  }
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
  if (const char* tr = getenv("CQVAD_MLP_TRACE")) p.trace = (long long*)strtoull(tr, nullptr, 0);
  p.res32 = res32; p.Y32 = Y32; p.ts_out = ts_out ? 1 : 0; p.ts_res = (ts_out && res && !res32) ? 1 : 0;
  p.YT = YT; p.ldyt = ldyt; p.yt_rows = yt_rows > 0 ? yt_rows : 1; p.yt_pitch = yt_pitch;
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  static const bool no_spec = getenv("CQVAD_MLP_NO_SPEC") != nullptr;
  if (ts_out && !no_spec && act == CQVAD_ACT_GELU && !ln_g) mlp_tc_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmX, tmW1, tmW2, tmY, tmR, p);
  else if (ts_out && !no_spec && act == CQVAD_ACT_RELU && ln_g) mlp_tc_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmX, tmW1, tmW2, tmY, tmR, p);
  else mlp_tc_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmX, tmW1, tmW2, tmY, tmR, p);
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace cqvad
