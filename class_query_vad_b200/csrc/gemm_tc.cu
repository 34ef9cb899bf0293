// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] . W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2-9 =
//     epilogue (two per TMEM lane quarter, 128 columns each).  Three pipelines: smem ring (TMA <-> MMA, 128-byte swizzle),
//     TMEM accumulator double buffer (MMA <-> epilogue, 2 x 256 columns), and the static tile scheduler.
//   * two launch forms of the same body: single CTA (UMMA 128 x 256 x 16, cta_group::1; ring of 4 x (128x64 A + 256x64 B),
//     TS: 3) for the small-row GEMMs, and CTA PAIRS (gemm_tc2_kernel: cluster of 2 on one TPC, UMMA 256 x 256 x 16,
//     cta_group::2) whenever there are >= 148 m-tiles: each CTA loads its own A tile and HALF of the weight tile, the leader
//     issues the MMAs for both, commits are multicast to both CTAs' barriers, both CTAs' TMA bytes complete on the leader's
//     full barrier; ring of 6 x 32 KB (TS: 4, or 5 with single-buffered output staging for K >= 512).
//   * the A operand is either a 2-D row-major matrix or, for the 3x3 convolution of ConvBlock
//     (models/detr/dab_transformer.py:81,90), a 3-D view [256 ch, w, N*(h+1) rows] of the y-padded NHWC activation:
//     every k-block of a filter tap is ONE TMA box (64 ch x w x RT rows) shifted by (dx,dy); the horizontal halo is
//     the TMA out-of-bounds zero fill, the vertical halo the zero separator rows.  No im2col buffer exists.
//   * epilogue straight out of TMEM (tcgen05.ld 32x32b: one thread = one output row): bias, ReLU / erf-GELU, residual,
//     zero-row masking, and row LayerNorm when N == 256 (pre-norm values stashed back into TMEM with tcgen05.st so
//     the residual is read once).
//   * TS variant (every non-conv GEMM with a bf16-only epilogue): a 64 KB output staging area next to the ring.  The epilogue
//     warps write bf16 rows into 128-byte-swizzled shared memory (conflict-free 16-byte stores) and ONE lane per warp issues
//     a TMA store of a [32 rows x 64 cols] box; the bf16 residual (or the activation-derivative operand) arrives the same way
//     through a TMA load into the staging buffer.  Evidence (ncu, profiles/): with one thread per output row every global
//     store / residual load instruction touched 32 different 128-byte lines -- the LSU data pipe was the top unit at 55%,
//     L1->L2 write traffic was 2x the output (half-filled sectors) and the K = 256 GEMMs ran at 2.1 TB/s of 6.5.
//     Lean (branch-free) and dual-GELU forms of that epilogue: ts_lean_half / ts_dual_gelu_half below.
//   * every mbarrier wait carries a watchdog (tc_common.cuh::mbar_wait_slow): a stuck pipeline traps instead of hanging.
#include "common.cuh"
#include "tc_common.cuh"
#include <mutex>
#include <stdlib.h>

namespace cqvad {

using namespace tc;

namespace {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, STAGES_DIRECT = 4, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_FULL = BLOCK_N * BLOCK_K * 2;  // 32 KB (a CTA pair holds half of it per CTA)
constexpr int SMEM_TILES = STAGES_DIRECT * (A_STAGE_BYTES + B_STAGE_FULL);
// aux: barriers 256 | bias [2][256] f32 (per TMEM buffer) | gamma [256] | beta [256] | LN partial sums [2][2][128][2] f32
constexpr int AUX_BIAS = 320, AUX_GAM = AUX_BIAS + 2048, AUX_BET = AUX_GAM + 1024, AUX_STATS = AUX_BET + 1024, AUX_BYTES = AUX_STATS + 4096;
constexpr int SMEM_BYTES = SMEM_TILES + 1024 /*align slack*/ + AUX_BYTES;
constexpr int STAGES_TS = 3;
constexpr int STG_BOX_BYTES = 32 * 64 * 2;                      // one [32 rows x 64 cols] bf16 box, 128 B per row
constexpr int STG_BYTES = 8 * 2 * STG_BOX_BYTES;                // 8 epilogue warps x 2 column halves of their 128 columns
constexpr int SMEM_BYTES_TS = STAGES_TS * (A_STAGE_BYTES + B_STAGE_FULL) + STG_BYTES + 1024 + AUX_BYTES;
constexpr int AUX_RBAR = 160;                                   // 8 warps x 2 mbarriers for the TMA-loaded side input
constexpr int NUM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quarter: 128 columns each)
constexpr int TMEM_COLS = 512;

struct TcParams {
  bf16* C; long ldc; long M; int N; int K;
  const float* bias; int act; const bf16* res; long ldr; const float* res32; float* c32;
  int rowop; const float* ro_ref; const float* ro_norm;   // Epilogue::rowop
  const float* ln_g; const float* ln_b; float ln_eps;
  int zero_period, zero_valid;
  bf16* c2; int c2_act; const bf16* mul_aux; int mul_mode; float mul_scale;
  // conv
  int conv; int cw; int rt;  // image width, image rows per tile
  int m_tiles, n_tiles;
  long long* trace;   // dev tool (CQVAD_GEMM_TRACE=<device address>): globaltimer stamps of CTA 0's pipeline events
  int dual;         // TS: C = gelu(v), C2 (through tmR) = gelu'(v)
  int stg_single;   // TS pair without side input: one staging box per warp (two sequential stores) buys a 5th ring stage
  int lean;   // TS: bias + none/ReLU (+ bf16 residual) only -> branch-free epilogue
  int side;   // TS: 1 = bf16 residual, 2 = activation-derivative operand (mul_aux) arrives through tmR
};

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace slots: [0,256) producer stage issue, [256,512) MMA full-wait done, [512,640) MMA tempty-wait done,
// [640,768) epilogue warp 2 tfull-wait done, [768,896) epilogue warp 2 tile done
#define CQ_TRACE(slot, idx, lim) do { if (p.trace && blockIdx.x == 0 && (idx) < (lim)) p.trace[(slot) + (idx)] = gtimer(); } while (0)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// EXTRA = true: the training-path epilogue options (second activated output, activation-derivative mask); a separate
// instantiation so that the inference epilogue carries none of their instructions.
// Lean TS epilogue step (bias + optional ReLU + optional bf16 residual already sitting in the staging buffer): 64 columns of one
// row per thread, branch-free, packed fp32x2 adds.  The general epilogue spends ~210 warp instructions per 32 columns on
// runtime option checks; the pipeline trace (tools/trace_gemm.py) showed the epilogue, at 2.8 us per 128 x 256 tile, as the
// stage that paces the K = 256 GEMMs (MMA 1.1 us, loads hidden).
// SIDE: 0 none, 1 += side (residual), 2 ReLU mask (side > 0), 3 *= side (stored activation derivative)
template <int SIDE>
__device__ __forceinline__ void ts_lean_half(uint32_t t_addr, const float* bias, uint32_t row_base, int sw, float lo, bool gelu, float msc = 1.f) {
#pragma unroll
  for (int cc = 0; cc < 64; cc += 32) {
    uint32_t r[32];
    tmem_ld32(t_addr + cc, r);
    tmem_ld_wait();
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const uint32_t saddr = row_base + (uint32_t)((((cc >> 3) + g8) ^ sw) << 4);
      const float4 b0 = *reinterpret_cast<const float4*>(bias + cc + g8 * 8), b1 = *reinterpret_cast<const float4*>(bias + cc + g8 * 8 + 4);
      uint64_t x[4];
      x[0] = f2add(f2pack(__uint_as_float(r[g8 * 8 + 0]), __uint_as_float(r[g8 * 8 + 1])), f2pack(b0.x, b0.y));
      x[1] = f2add(f2pack(__uint_as_float(r[g8 * 8 + 2]), __uint_as_float(r[g8 * 8 + 3])), f2pack(b0.z, b0.w));
      x[2] = f2add(f2pack(__uint_as_float(r[g8 * 8 + 4]), __uint_as_float(r[g8 * 8 + 5])), f2pack(b1.x, b1.y));
      x[3] = f2add(f2pack(__uint_as_float(r[g8 * 8 + 6]), __uint_as_float(r[g8 * 8 + 7])), f2pack(b1.z, b1.w));
      if (gelu) {   // erf-form GELU on the packed polynomial (warp-uniform branch)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = gelu2(x[j]);
      }
      uint4 sv;
      if constexpr (SIDE != 0) sv = lds128(saddr);
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      const uint32_t* sw32 = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, b;
        f2unpack(x[j], a, b);
        a = fmaxf(a, lo); b = fmaxf(b, lo);
        if constexpr (SIDE != 0) {
          const float sa = __uint_as_float(sw32[j] << 16), sb = __uint_as_float(sw32[j] & 0xffff0000u);   // bf16 pair -> fp32
          if constexpr (SIDE == 1) { a += sa; b += sb; }
          else if constexpr (SIDE == 2) { a = sa > 0.f ? a * msc : 0.f; b = sb > 0.f ? b * msc : 0.f; }
          else { a *= sa; b *= sb; }
        }
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ow[j]) : "f"(b), "f"(a));
      }
      sts128(saddr, o);
    }
  }
}

// Dual GELU step (training forward): A = gelu(acc + bias) -> the C box, G = gelu'(acc + bias) -> the C2 box
__device__ __forceinline__ void ts_dual_gelu_half(uint32_t t_addr, const float* bias, uint32_t row_a, uint32_t row_g, int sw) {
#pragma unroll 1
  for (int cc = 0; cc < 64; cc += 32) {
    uint32_t r[32];
    tmem_ld32(t_addr + cc, r);
    tmem_ld_wait();
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const uint32_t off = (uint32_t)((((cc >> 3) + g8) ^ sw) << 4);
      const float4 b0 = *reinterpret_cast<const float4*>(bias + cc + g8 * 8), b1 = *reinterpret_cast<const float4*>(bias + cc + g8 * 8 + 4);
      const uint64_t bb[4] = {f2pack(b0.x, b0.y), f2pack(b0.z, b0.w), f2pack(b1.x, b1.y), f2pack(b1.z, b1.w)};
      uint4 oa, og;
      uint32_t* wa = reinterpret_cast<uint32_t*>(&oa);
      uint32_t* wg = reinterpret_cast<uint32_t*>(&og);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t x = f2add(f2pack(__uint_as_float(r[g8 * 8 + 2 * j]), __uint_as_float(r[g8 * 8 + 2 * j + 1])), bb[j]);
        uint64_t a2, g2;
        gelu_dual2(x, a2, g2);
        wa[j] = f2_to_bf16x2(a2);
        wg[j] = f2_to_bf16x2(g2);
      }
      sts128(row_a + off, oa);
      sts128(row_g + off, og);
    }
  }
}

// Correctly rounded x / n from r = RN(1 / n) without the division instruction sequence: q0 = RN(x r), then two residual
// corrections q <- RN(q + RN(x - n q) r) (the residual is exact in an FMA).  The first makes q faithful, the second rounds it
// correctly (Markstein's theorem; quotients in the normal range -- offsets / grid sizes are).  5 FMA-pipe instructions against
// ~11 for div.rn.f32; checked against exact rational arithmetic in tests/test_exact_division_cpu.py.
__device__ __forceinline__ float div_rn_by(float x, float n, float r) {
  float q = x * r;
  float e = fmaf(-q, n, x);
  q = fmaf(e, r, q);
  e = fmaf(-q, n, x);
  return fmaf(e, r, q);
}

// Epilogue::rowop 2, one 32-column step whose first column is PH (mod 96): column n = ((m*4 + l)*8 + p)*3 + i of the offsets
// projection -> sampling location  ref[l*3 + i] + x / norm[l*3 + i]  (the IEEE quotient, as msda_prepare_kernel); l and i are
// compile-time per unrolled column, so the row's 12 reference values, the 12 normalisers and their reciprocals stay in registers
template <int PH>
__device__ __forceinline__ void rowop_loc_step(float (&x)[32], const float (&ref)[12], const float (&nrm)[12], const float (&rcp)[12]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int n = PH + j, l = (n / 24) % 4, i = n % 3;
    x[j] = ref[l * 3 + i] + div_rn_by(x[j], nrm[l * 3 + i], rcp[l * 3 + i]);
  }
}

// CTA2 = true: CTA pair (cluster of 2 on one TPC, tcgen05 cta_group::2).  The pair computes two vertically adjacent
// 128 x 256 output tiles with ONE M = 256 UMMA stream issued by the leader; each CTA loads its own A tile and only HALF of
// the weight tile, so the weight traffic L2 -> SM halves (the conv / large-K GEMMs were bound by the ~6300 B/clk L2 slice
// throughput, not by the tensor pipe: B re-reads were 2/3 of the conv's L2 traffic) and the ring holds 6 (TS: 4) k-blocks.
// ROWOP = true: the Epilogue::rowop instantiation (softmax / sampling-location epilogues of the encoder's query projections), kept
// out of the other kernels so that their epilogues carry none of its registers (154 instead of 126 with it compiled in).
// SPEC (TMA-store kernels): 0 = every epilogue option, 1 = the lean half-tile epilogues only (bias, ReLU / GELU, bf16 residual or
// activation-derivative operand: most launches of a training step), 2 = the dual-GELU epilogue only, 3 = the LayerNorm epilogue only
// (out_proj + norm GEMMs of the encoder / decoder layers) -- the all-options kernel is
// ~9 000 SASS instructions (146 KB, far beyond the instruction cache), its specialisations a fraction of that.
template <bool EXTRA, bool TS, bool CTA2, bool ROWOP = false, int SPEC = 0>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                             const CUtensorMap& tmR, const TcParams& p) {
  // ring depth: pair 6 (TS: 4, or 5 with single-buffered output staging when there is no side input), single CTA 4 (TS: 3)
  constexpr int MAX_STAGES = 6;
  const int STAGES = CTA2 ? (TS ? (p.stg_single ? 5 : 4) : 6) : (TS ? STAGES_TS : STAGES_DIRECT);
  const int stg_stride = (TS && p.stg_single) ? 0 : STG_BOX_BYTES;   // byte distance between a warp's two staging boxes
  constexpr int B_STAGE_BYTES = CTA2 ? B_STAGE_FULL / 2 : B_STAGE_FULL;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int ncta = CTA2 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t stg_base = smem_base + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);   // TS only
  const uint32_t bars = stg_base + (TS ? (p.stg_single ? STG_BYTES / 2 : STG_BYTES) : 0);
  const uint32_t full_bar = bars, empty_bar = bars + 8 * MAX_STAGES, tfull_bar = bars + 16 * MAX_STAGES,
                 tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16, rbar_base = bars + AUX_RBAR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work unit = one tile (a vertical PAIR of tiles for a CTA pair: this CTA takes m-tile 2 * mpair + rank)
  const int num_tiles = ((p.m_tiles + ncta - 1) / ncta) * p.n_tiles;
  const int first_tile = blockIdx.x / ncta, tile_step = gridDim.x / ncta;
  const int kblocks = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + 8 * a, 1); mbar_init(tempty_bar + 8 * a, 8 * ncta); }
    if constexpr (TS) {
      tma_prefetch_desc(&tmC);
      if (p.side) tma_prefetch_desc(&tmR);
      for (int i = 0; i < 16; ++i) mbar_init(rbar_base + 8 * i, 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) { if constexpr (CTA2) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint32_t a_bytes = p.conv ? (uint32_t)(p.cw * p.rt * BLOCK_K * 2) : (uint32_t)A_STAGE_BYTES;
      int stage = 0; uint32_t phase = 0;
      int ev = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int mt = (tile / p.n_tiles) * ncta + (int)rank, nt = tile % p.n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          CQ_TRACE(0, ev, 256); ++ev;
          if constexpr (CTA2) {
            // both CTAs' bytes land on the LEADER's full barrier (the leader arms it with the pair's total)
            const uint32_t fb = mapa_shared(full_bar + 8 * stage, 0);
            if (rank == 0) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * (a_bytes + (uint32_t)B_STAGE_BYTES));
            if (p.conv) {
              const int tap = kb >> 2, c0 = (kb & 3) * BLOCK_K;
              tma_load_3d_pair(sA + stage * A_STAGE_BYTES, &tmA, fb, c0, tap % 3 - 1, mt * p.rt + tap / 3 - 1);
            } else {
              tma_load_2d_pair(sA + stage * A_STAGE_BYTES, &tmA, fb, kb * BLOCK_K, mt * BLOCK_M);
            }
            tma_load_2d_pair(sB + stage * B_STAGE_BYTES, &tmB, fb, kb * BLOCK_K, nt * BLOCK_N + (int)rank * (BLOCK_N / 2));
          } else {
            const uint32_t fb = full_bar + 8 * stage;
            mbar_arrive_expect_tx(fb, a_bytes + (uint32_t)B_STAGE_BYTES);
            if (p.conv) {
              const int tap = kb >> 2, c0 = (kb & 3) * BLOCK_K;
              tma_load_3d(sA + stage * A_STAGE_BYTES, &tmA, fb, c0, tap % 3 - 1, mt * p.rt + tap / 3 - 1);
            } else {
              tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, fb, kb * BLOCK_K, mt * BLOCK_M);
            }
            tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, fb, kb * BLOCK_K, nt * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (CTA2) {
        // tail: the leader's multicast commits keep arriving on THIS CTA's empty barriers until the last MMA has retired;
        // wait for all of them so that the CTA cannot exit (and its shared memory be re-used) underneath a remote arrival
        for (int s = 0; s < STAGES; ++s) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (the leader CTA of a pair) =====
      constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * BLOCK_M : BLOCK_M, BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int ev = 0, evt = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        CQ_TRACE(512, evt, 128); ++evt;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          CQ_TRACE(256, ev, 256); ++ev;
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(sA + stage * A_STAGE_BYTES);
          const uint64_t b_desc = make_smem_desc_sw128(sB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
            if constexpr (CTA2) umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) when the MMAs above have read it
          if constexpr (CTA2) umma_commit_pair(empty_bar + 8 * stage); else umma_commit(empty_bar + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTA2) umma_commit_pair(tfull_bar + 8 * acc); else umma_commit(tfull_bar + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q = warp % 4, column half g = (warp-2)/4 =====
    // Two warps per SM sub-partition (latency hiding) and every per-column vector (bias, gamma, beta) staged in shared
    // memory: the ncu source page of the first version showed 42% of all stall samples on FADDs waiting for the
    // global bias loads (long scoreboard) with a single exposed epilogue warp per sub-partition.
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int et = (warp - 2) * 32 + lane;          // 0..255
    const int row_in_tile = q * 32 + lane;
    float* aux = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)));
    float* bias_s = aux + AUX_BIAS / 4; float* gam_s = aux + AUX_GAM / 4; float* bet_s = aux + AUX_BET / 4;
    float* stats_s = aux + AUX_STATS / 4;
    const bool do_ln = SPEC == 3 || (SPEC == 0 && p.ln_g != nullptr);
    if (do_ln) { gam_s[et] = p.ln_g[et]; bet_s[et] = p.ln_b[et]; }   // N == 256, one n-tile
    int acc = 0; uint32_t acc_phase = 0;
    int etile = 0;
    uint32_t rph0 = 0, rph1 = 0;      // TS: parities of this warp's two side-input barriers
    const int tile_rows = p.conv ? p.cw * p.rt : BLOCK_M;
    const int c0 = g * 128;
    const uint32_t tempty_leader = CTA2 ? mapa_shared(tempty_bar, 0) : 0u;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int mt = (tile / p.n_tiles) * ncta + (int)rank, nt = tile % p.n_tiles;
      const int n0 = nt * BLOCK_N;
      float* bcur = bias_s + acc * 256;
      bcur[et] = (p.bias && n0 + et < p.N) ? p.bias[n0 + et] : 0.f;
      if constexpr (TS) {
        // ---- TMA-store epilogue (non-conv): this warp owns rows [row0, row0+32) x columns [colw, colw+128) of C as two
        // [32 x 64] boxes; buffer hf is reused one whole tile later, so the store of the previous tile has long been read ----
        const int row0 = mt * BLOCK_M + q * 32;
        const int colw = n0 + c0;
        const uint32_t stg = stg_base + (uint32_t)((warp - 2) * (p.stg_single ? 1 : 2) * STG_BOX_BYTES);
        const uint32_t rbar = rbar_base + (uint32_t)((warp - 2) * 16);
        const uint32_t my_row = stg + (uint32_t)(lane * 128);
        const int sw = lane & 7;                                  // 128-byte swizzle: 16-byte chunk j of row r sits at j ^ (r & 7)
        const bool act0 = colw < p.N, act1 = colw + 64 < p.N;     // column halves inside the matrix (warp-uniform)
        // side-input boxes (bf16 residual / activation-derivative operand) -> the staging buffers themselves.  The boxes of a
        // tile are requested at the END of the previous tile (right after its stores have read the buffers), so that their
        // DRAM latency overlaps the wait for the next accumulator; the first tile's are requested here.
        auto side_request = [&](int t) {
          const int mt_ = (t / p.n_tiles) * ncta + (int)rank, colw_ = (t % p.n_tiles) * BLOCK_N + c0, row0_ = mt_ * BLOCK_M + q * 32;
          bulk_wait_read<0>();
          if (colw_ < p.N) { mbar_arrive_expect_tx(rbar, STG_BOX_BYTES); tma_load_2d(stg, &tmR, rbar, colw_, row0_); }
          if (colw_ + 64 < p.N) { mbar_arrive_expect_tx(rbar + 8, STG_BOX_BYTES); tma_load_2d(stg + STG_BOX_BYTES, &tmR, rbar + 8, colw_ + 64, row0_); }
        };
        if (p.side && lane == 0 && tile == first_tile) side_request(tile);
        if (!p.side && lane == 0) bulk_wait_read<0>();   // the stores of the previous tile have read both buffers (ordered by bar.sync below)
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        if (warp == 2 && lane == 0) { CQ_TRACE(640, etile, 128); }
        tc_fence_after();
        asm volatile("bar.sync 1, 256;" ::: "memory");   // bias (and gamma/beta) visible to the 8 epilogue warps
#define CQ_TRACE_E(k) do { if (warp == 2 && lane == 0 && etile < 8) { CQ_TRACE(896 + etile * 16, (k), 16); } } while (0)
        CQ_TRACE_E(0);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
        const long grow = (long)row0 + lane;
        const bool zero_row = p.zero_period > 0 && (int)(grow % p.zero_period) >= p.zero_valid;
        if constexpr (ROWOP) {
          // MSDA "prepare" folded into the two query projections (Epilogue::rowop): each thread owns one row's 32 consecutive
          // columns per step = one head's 32 attention logits (softmax in registers) or 32 offset components (location
          // arithmetic).  fp32 output through the staging boxes: a [32 rows x 32 fp32] box has the 128-byte rows of the bf16
          // boxes, so the same swizzle and TMA store apply (tmC is an fp32 map here); 4 steps alternate the warp's two boxes.
          float ref12[12], nrm12[12], rcp12[12];
          if (p.rowop == 2) {
            const float4* rr = reinterpret_cast<const float4*>(p.ro_ref + (grow < p.M ? grow : 0) * 12);
            const float4 a0 = rr[0], a1 = rr[1], a2 = rr[2];
            ref12[0] = a0.x; ref12[1] = a0.y; ref12[2] = a0.z; ref12[3] = a0.w; ref12[4] = a1.x; ref12[5] = a1.y;
            ref12[6] = a1.z; ref12[7] = a1.w; ref12[8] = a2.x; ref12[9] = a2.y; ref12[10] = a2.z; ref12[11] = a2.w;
#pragma unroll
            for (int j = 0; j < 12; ++j) { nrm12[j] = p.ro_norm[j]; rcp12[j] = __frcp_rn(nrm12[j]); }
          }
#pragma unroll 1
          for (int s4 = 0; s4 < 4; ++s4) {
            const int c = s4 * 32;
            if (colw + c < p.N) {   // warp-uniform
              if (s4 >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }   // the store two steps back has read this box
              uint32_t r[32];
              tmem_ld32(t_addr + c, r);
              tmem_ld_wait();
              float x[32];
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                float bs[8];
                load8(bcur + c0 + c + g8 * 8, bs);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[g8 * 8 + j] = __uint_as_float(r[g8 * 8 + j]) + bs[j];
              }
              if (p.rowop == 1) {
                float mx = x[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) mx = fmaxf(mx, x[j]);
                float sp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 32; ++j) { x[j] = __expf(x[j] - mx); sp[j & 3] += x[j]; }
                const float inv = __frcp_rn((sp[0] + sp[1]) + (sp[2] + sp[3]));   // weights to ~2 ulp: they feed bf16 arithmetic
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] *= inv;
              } else {
                const int ph = (colw + c) % 96;      // warp-uniform
                if (ph == 0) rowop_loc_step<0>(x, ref12, nrm12, rcp12);
                else if (ph == 32) rowop_loc_step<32>(x, ref12, nrm12, rcp12);
                else rowop_loc_step<64>(x, ref12, nrm12, rcp12);
              }
              const uint32_t buf = my_row + (uint32_t)((s4 & 1) * STG_BOX_BYTES);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                uint4 o;
                o.x = __float_as_uint(x[j4 * 4]); o.y = __float_as_uint(x[j4 * 4 + 1]);
                o.z = __float_as_uint(x[j4 * 4 + 2]); o.w = __float_as_uint(x[j4 * 4 + 3]);
                sts128(buf + (uint32_t)((j4 ^ sw) << 4), o);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_2d(&tmC, stg + (uint32_t)((s4 & 1) * STG_BOX_BYTES), colw + c, row0);
            }
            if (lane == 0) bulk_commit();
          }
        } else {
          (void)zero_row;
        float mean = 0.f, rstd = 1.f;
        if (do_ln) {
          // pass 1 (N == 256): v = act(acc + bias) (+res); stash v in TMEM; partial row statistics over this warp's 128 columns
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            const int hf = c >> 6;
            const uint32_t buf = my_row + (uint32_t)(hf * stg_stride);
            if (p.side == 1 && (c & 63) == 0) mbar_wait(rbar + 8 * hf, hf ? rph1 : rph0);
            uint32_t r[32];
            tmem_ld32(t_addr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              if (p.side == 1) unpack8(lds128(buf + (uint32_t)(((((c & 63) >> 3) + g8) ^ sw) << 4)), rs);
              float bs[8];
              load8(bcur + c0 + c + g8 * 8, bs);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float v = __uint_as_float(r[g8 * 8 + j]) + bs[j];
                if (p.act == CQVAD_ACT_RELU) v = fmaxf(v, 0.f);
                else if (p.act == CQVAD_ACT_GELU) v = gelu_erf(v);
                v += rs[j];
                s1 += v; s2 = fmaf(v, v, s2);
                r[g8 * 8 + j] = __float_as_uint(v);
              }
            }
            tmem_st32(t_addr + c, r);
          }
          tmem_st_wait();
          float* st = stats_s + acc * 512;
          st[(g * 128 + row_in_tile) * 2] = s1;
          st[(g * 128 + row_in_tile) * 2 + 1] = s2;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float o1 = st[((g ^ 1) * 128 + row_in_tile) * 2], o2 = st[((g ^ 1) * 128 + row_in_tile) * 2 + 1];
          mean = (s1 + o1) * (1.0f / BLOCK_N);
          const float var = fmaxf((s2 + o2) * (1.0f / BLOCK_N) - mean * mean, 0.f);
          rstd = rsqrtf(var + p.ln_eps);
        }
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          const bool active = hf ? act1 : act0;
          const uint32_t buf = my_row + (uint32_t)(hf * stg_stride);
          if (active) {
            // single staging box per warp: the store of the first half must have read it before the second half is written
            if (p.stg_single && hf == 1) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
            if (p.side && !do_ln) mbar_wait(rbar + 8 * hf, hf ? rph1 : rph0);
            CQ_TRACE_E(1 + hf * 6);
            if (SPEC == 2 || (SPEC == 0 && p.dual)) {
              // both boxes of this warp are reused by the second half: its stores must have read them
              if (hf == 1) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
              ts_dual_gelu_half(t_addr + hf * 64, bcur + c0 + hf * 64, my_row, my_row + STG_BOX_BYTES, sw);
            } else if (SPEC == 1 || (SPEC == 0 && p.lean)) {
              const float lo = p.act == CQVAD_ACT_RELU ? 0.f : -INFINITY;
              const bool ge = p.act == CQVAD_ACT_GELU;
              // specialised kernels: the side-input variants (0 / 1) live in the EXTRA = false instantiation, the multiply variants in EXTRA = true
              if ((SPEC == 0 || !EXTRA) && p.side == 0) ts_lean_half<0>(t_addr + hf * 64, bcur + c0 + hf * 64, buf, sw, lo, ge);
              else if ((SPEC == 0 || !EXTRA) && p.side == 1) ts_lean_half<1>(t_addr + hf * 64, bcur + c0 + hf * 64, buf, sw, lo, ge);
              else if ((SPEC == 0 || EXTRA) && p.mul_mode == 1) ts_lean_half<2>(t_addr + hf * 64, bcur + c0 + hf * 64, buf, sw, lo, ge, p.mul_scale);
              else if (SPEC == 0 || EXTRA) ts_lean_half<3>(t_addr + hf * 64, bcur + c0 + hf * 64, buf, sw, lo, ge);
            } else if constexpr (SPEC == 0 || SPEC == 3) {
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 32) {
              const int c = hf * 64 + cc;
              uint32_t r[32];
              tmem_ld32(t_addr + c, r);
              tmem_ld_wait();
              CQ_TRACE_E(2 + hf * 6 + (cc >> 5) * 2);
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                const int cb = c + g8 * 8;
                const uint32_t saddr = buf + (uint32_t)((((cc >> 3) + g8) ^ sw) << 4);
                float v[8];
                if (do_ln) {
                  float gm[8], bt[8];
                  load8(gam_s + c0 + cb, gm);
                  load8(bet_s + c0 + cb, bt);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = (__uint_as_float(r[g8 * 8 + j]) - mean) * rstd * gm[j] + bt[j];
                } else {
                  float bs[8];
                  load8(bcur + c0 + cb, bs);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    float x = __uint_as_float(r[g8 * 8 + j]) + bs[j];
                    if (p.act == CQVAD_ACT_RELU) x = fmaxf(x, 0.f);
                    else if (p.act == CQVAD_ACT_GELU) x = gelu_erf(x);
                    v[j] = x;
                  }
                  if (p.side) {
                    float sx[8];
                    unpack8(lds128(saddr), sx);
                    if (p.side == 1) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) v[j] += sx[j];
                    } else if constexpr (EXTRA) {
                      if (p.mul_mode == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = sx[j] > 0.f ? v[j] * p.mul_scale : 0.f;
                      } else if (p.mul_mode == 3) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] *= sx[j];
                      } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast(sx[j]);
                      }
                    }
                  }
                }
                if (zero_row) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
                uint4 o;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                sts128(saddr, o);
              }
            }
            }   // general path
            CQ_TRACE_E(5 + hf * 6);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (SPEC == 2 || (SPEC == 0 && p.dual)) {
                tma_store_2d(&tmC, stg, colw + hf * 64, row0);
                tma_store_2d(&tmR, stg + STG_BOX_BYTES, colw + hf * 64, row0);   // tmR carries the second output here
              } else {
                tma_store_2d(&tmC, stg + (uint32_t)(hf * stg_stride), colw + hf * 64, row0);
              }
            }
            CQ_TRACE_E(6 + hf * 6);
          }
          if (lane == 0) bulk_commit();     // always two groups per tile (an empty group for a half outside the matrix)
        }
        }   // !rowop
        if (p.side) {
          if (act0) rph0 ^= 1;
          if (act1) rph1 ^= 1;
          if (lane == 0 && tile + tile_step < num_tiles) side_request(tile + tile_step);
        }
      } else {
      const long grow = (long)mt * tile_rows + row_in_tile;
      const bool row_ok = row_in_tile < tile_rows && grow < p.M;
      // bf16 residual fetched one 32-column step ahead (4 x 16 bytes in flight per thread)
      const bool res_bf16 = p.res != nullptr && p.res32 == nullptr && row_ok;
      const bf16* rbase = p.res + (res_bf16 ? grow * p.ldr + n0 + c0 : 0);
      auto ld_res = [&](int col) -> uint4 {   // col relative to c0, multiple of 8
        return (res_bf16 && n0 + c0 + col < p.N) ? *reinterpret_cast<const uint4*>(rbase + col) : make_uint4(0u, 0u, 0u, 0u);
      };
      uint4 rnext[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) rnext[u] = ld_res(u * 8);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // bias (and gamma/beta) visible to the 8 epilogue warps
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
      const bool zero_row = p.zero_period > 0 && (int)(grow % p.zero_period) >= p.zero_valid;
      bf16* crow = p.C + grow * p.ldc + n0 + c0;
      const float* rrow32 = p.res32 ? p.res32 + grow * p.ldr + n0 + c0 : nullptr;
      float* crow32 = p.c32 ? p.c32 + grow * p.ldc + n0 + c0 : nullptr;
      bf16* c2row = (EXTRA && p.c2) ? p.c2 + grow * p.ldc + n0 + c0 : nullptr;
      const bf16* arow = (EXTRA && p.mul_mode) ? p.mul_aux + grow * p.ldc + n0 + c0 : nullptr;
      float mean = 0.f, rstd = 1.f;
      if (do_ln) {
        // pass 1: v = act(acc + bias) (+res); stash v in TMEM; partial row statistics over this warp's 128 columns
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint4 rcur[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) rcur[u] = rnext[u];
          if (c + 32 < 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rnext[u] = ld_res(c + 32 + u * 8);
          }
          uint32_t r[32];
          tmem_ld32(t_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (rrow32) { if (row_ok) load8(rrow32 + c + g8 * 8, rs); }
            else unpack8(rcur[g8], rs);
            float bs[8];
            load8(bcur + c0 + c + g8 * 8, bs);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = __uint_as_float(r[g8 * 8 + j]) + bs[j];
              if (p.act == CQVAD_ACT_RELU) v = fmaxf(v, 0.f);
              else if (p.act == CQVAD_ACT_GELU) v = gelu_erf(v);
              v += rs[j];
              s1 += v; s2 = fmaf(v, v, s2);
              r[g8 * 8 + j] = __float_as_uint(v);
            }
          }
          tmem_st32(t_addr + c, r);
        }
        tmem_st_wait();
        float* st = stats_s + acc * 512;
        st[(g * 128 + row_in_tile) * 2] = s1;
        st[(g * 128 + row_in_tile) * 2 + 1] = s2;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float o1 = st[((g ^ 1) * 128 + row_in_tile) * 2], o2 = st[((g ^ 1) * 128 + row_in_tile) * 2 + 1];
        mean = (s1 + o1) * (1.0f / BLOCK_N);
        const float var = fmaxf((s2 + o2) * (1.0f / BLOCK_N) - mean * mean, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
      }
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        if (n0 + c0 + c >= p.N) break;  // warp-uniform
        uint4 rcur[4];
        if (!do_ln) {
#pragma unroll
          for (int u = 0; u < 4; ++u) rcur[u] = rnext[u];
          if (c + 32 < 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rnext[u] = ld_res(c + 32 + u * 8);
          }
        }
        uint4 anext[4];
        if constexpr (EXTRA) {
          if (p.mul_mode != 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              anext[u] = (row_ok && n0 + c0 + c + u * 8 < p.N) ? *reinterpret_cast<const uint4*>(arow + c + u * 8) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        uint32_t r[32];
        tmem_ld32(t_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int cb = c + g8 * 8;
          if (n0 + c0 + cb >= p.N) continue;
          float v[8];
          if (do_ln) {
            float gm[8], bt[8];
            load8(gam_s + c0 + cb, gm);
            load8(bet_s + c0 + cb, bt);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (__uint_as_float(r[g8 * 8 + j]) - mean) * rstd * gm[j] + bt[j];
          } else {
            float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (rrow32) { if (row_ok) load8(rrow32 + cb, rs); }
            else unpack8(rcur[g8], rs);
            float bs[8];
            load8(bcur + c0 + cb, bs);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = __uint_as_float(r[g8 * 8 + j]) + bs[j];
              if (p.act == CQVAD_ACT_RELU) x = fmaxf(x, 0.f);
              else if (p.act == CQVAD_ACT_GELU) x = gelu_erf(x);
              v[j] = x;
            }
            if constexpr (EXTRA) {
              if (p.mul_mode != 0) {
                float ax[8];
                unpack8(anext[g8], ax);
                if (p.mul_mode == 1) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = ax[j] > 0.f ? v[j] * p.mul_scale : 0.f;
                } else if (p.mul_mode == 3) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= ax[j];
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast(ax[j]);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += rs[j];
          }
          if (zero_row) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
          if (row_ok) {
            store8(crow + cb, v);
            if (crow32) store8(crow32 + cb, v);
            if constexpr (EXTRA) {
              if (c2row) {
                float v2[8];
                if (p.c2_act == CQVAD_ACT_GELU) {
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    float a, b;
                    f2unpack(gelu2(f2pack(v[j], v[j + 1])), a, b);
                    v2[j] = a; v2[j + 1] = b;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v2[j] = p.c2_act == CQVAD_ACT_RELU ? fmaxf(v[j], 0.f) : v[j];
                }
                store8(c2row + cb, v2);
              }
            }
          }
        }
      }
      }   // direct-store epilogue
      // release the accumulator buffer to the MMA warp
      if (warp == 2 && lane == 0) { CQ_TRACE(768, etile, 128); }
      ++etile;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CTA2) mbar_arrive_cluster(tempty_leader + 8 * acc); else mbar_arrive(tempty_bar + 8 * acc); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (TS) { if (lane == 0) bulk_wait_read<0>(); }   // shared memory stays valid until the last stores have read it
    (void)rph0; (void)rph1;
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer's MMAs / remote arrivals no longer touch this CTA's shared memory
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool EXTRA, bool TS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<EXTRA, TS, false>(tmA, tmB, tmC, tmR, p);
}
template <bool EXTRA, bool TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<EXTRA, TS, true>(tmA, tmB, tmC, tmR, p);
}
template <bool EXTRA, int SPEC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_spec_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<EXTRA, true, false, false, SPEC>(tmA, tmB, tmC, tmR, p);
}
template <bool EXTRA, int SPEC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_spec_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<EXTRA, true, true, false, SPEC>(tmA, tmB, tmC, tmR, p);
}
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_rowop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<false, true, false, true>(tmA, tmB, tmC, tmR, p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_rowop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  gemm_tc_body<false, true, true, true>(tmA, tmB, tmC, tmR, p);
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::once_flag g_once;
int g_init_err = 0;

void init_once() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) { g_init_err = 1; return; }
  g_encode = (EncodeTiledFn)fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_rowop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_spec_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_spec_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_spec_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_spec_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc_spec_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_spec_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_spec_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_spec_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
  if (cudaFuncSetAttribute(gemm_tc2_rowop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_TS) != cudaSuccess) g_init_err = 2;
}

}  // namespace

int tc_num_sms() {
  std::call_once(g_once, init_once);
  return g_num_sms;
}

// rank-2 / rank-3 tensor map with 128-byte swizzle.  dims/strides innermost first; strides in bytes for dims >= 1.
static int make_tmap_any(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                         const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  std::call_once(g_once, init_once);
  if (g_init_err) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed (%d)", g_init_err);
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(CQVAD_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  return make_tmap_any(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  return make_tmap_any(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

int gemm_tc(const bf16* A, long lda, const bf16* W, bf16* C, long ldc, long M, int N, int K, const Epilogue& epi,
            const ConvGeom* conv, cudaStream_t st) {
  // shapes the kernel takes; anything else goes to the CUDA-core kernel (return 1)
  if (K % BLOCK_K != 0 || lda % 8 != 0 || ldc % 8 != 0 || M < 1) return 1;
  // N % 8 != 0 is taken only for a bare GEMM whose row pitch covers the rounded-up width (8-wide vector stores)
  if (N % 8 != 0 && (ldc < ((N + 7) & ~7) || epi.bias || epi.res || epi.ln_g)) return 1;
  if ((((uintptr_t)A) & 15) || (((uintptr_t)W) & 15) || (((uintptr_t)C) & 15)) return 1;
  if (epi.res && ((((uintptr_t)epi.res) & 15) || epi.ldr % 8 != 0)) return 1;
  if ((epi.res32 && ((((uintptr_t)epi.res32) & 15) || epi.ldr % 8 != 0)) || (epi.c32 && (((uintptr_t)epi.c32) & 15))) return 1;
  if (epi.ln_g && N != BLOCK_N) return 1;
  if (epi.rowop && (N % 32 != 0 || !epi.c32 || ldc % 4 != 0 || epi.res || epi.res32 || epi.ln_g || epi.c2 || epi.mul_mode ||
                    epi.act != CQVAD_ACT_NONE || epi.zero_period || conv || epi.dual_gelu ||
                    (epi.rowop == 2 && (!epi.ro_ref || !epi.ro_norm || N % 96 != 0 || (((uintptr_t)epi.ro_ref) & 15)))))
    return 1;
  if ((epi.c2 || epi.mul_mode) && (epi.ln_g || N % 8 != 0)) return 1;
  if ((epi.c2 && (((uintptr_t)epi.c2) & 15)) || (epi.mul_aux && (((uintptr_t)epi.mul_aux) & 15))) return 1;
  if (conv && (conv->w > 128 || lda != kC || K != 9 * kC)) return 1;
  std::call_once(g_once, init_once);
  if (g_init_err) return set_error(CQVAD_E_CUDA, "tcgen05 path: initialisation failed (%d)", g_init_err);

  TcParams p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = epi.bias; p.act = epi.act; p.res = (const bf16*)epi.res; p.ldr = epi.ldr; p.res32 = epi.res32; p.c32 = epi.c32;
  p.rowop = epi.rowop; p.ro_ref = epi.ro_ref; p.ro_norm = epi.ro_norm;
  p.ln_g = epi.ln_g; p.ln_b = epi.ln_b; p.ln_eps = epi.ln_eps;
  p.zero_period = epi.zero_period; p.zero_valid = epi.zero_valid;
  p.c2 = (bf16*)epi.c2; p.c2_act = epi.c2_act; p.mul_aux = (const bf16*)epi.mul_aux; p.mul_mode = epi.mul_mode; p.mul_scale = epi.mul_scale;
  p.n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  if (const char* tr = getenv("CQVAD_GEMM_TRACE")) p.trace = (long long*)strtoull(tr, nullptr, 0);
  CUtensorMap tmA, tmB;
  if (conv) {
    const int w = conv->w;
    const int rt = BLOCK_M / w;                 // image rows per tile
    const long ny = M / w;                      // rows of the padded layout: n_img * (h+1)
    if (M % w != 0) return 1;
    p.conv = 1; p.cw = w; p.rt = rt;
    p.m_tiles = (int)((ny + rt - 1) / rt);
    const cuuint64_t dims[3] = {(cuuint64_t)kC, (cuuint64_t)w, (cuuint64_t)ny};
    const cuuint64_t strides[2] = {(cuuint64_t)kC * 2, (cuuint64_t)w * kC * 2};
    const cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)w, (cuuint32_t)rt};
    CQ_TRY(make_tmap_bf16(&tmA, A, 3, dims, strides, box));
  } else {
    p.conv = 0;
    p.m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M);
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
    const cuuint32_t box[2] = {BLOCK_K, BLOCK_M};
    CQ_TRY(make_tmap_bf16(&tmA, A, 2, dims, strides, box));
  }
  // CTA pairs (cta_group::2) for every GEMM with enough m-tiles to fill the machine pairwise
  static const bool no_pair = getenv("CQVAD_GEMM_NO_PAIR") != nullptr;
  const bool pair = !no_pair && p.m_tiles >= 2 * (g_num_sms / 2) && g_num_sms >= 2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(pair ? BLOCK_N / 2 : BLOCK_N)};
    CQ_TRY(make_tmap_bf16(&tmB, W, 2, dims, strides, box));
  }
  const long tiles = pair ? (long)((p.m_tiles + 1) / 2) * p.n_tiles * 2 : (long)p.m_tiles * p.n_tiles;
  const int sms = pair ? g_num_sms / 2 * 2 : g_num_sms;
  const int grid = (int)(tiles < sms ? tiles : sms);
  // TMA-store epilogue: every non-conv GEMM whose epilogue reads / writes bf16 only (at most one [M,N] side input)
  static const bool no_ts = getenv("CQVAD_GEMM_NO_TS") != nullptr;
  const bool dual = epi.dual_gelu;
  if (dual && (!epi.c2 || epi.res || epi.mul_mode || epi.ln_g || epi.zero_period || conv)) return 1;
  const bool ts = epi.rowop || (!no_ts && !conv && !epi.c32 && !epi.res32 && (!epi.c2 || dual) && N % 8 == 0 && !(epi.res && epi.mul_mode));
  if (dual && !ts) return 1;
  const bool extra = p.c2 || p.mul_mode;
  CUtensorMap tmC = tmA, tmR = tmA;
  if (epi.rowop) {     // fp32 output only: [32 rows x 32 fp32] boxes of the same 128-byte rows
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint32_t box[2] = {32, 32};
    const cuuint64_t sc[1] = {(cuuint64_t)ldc * 4};
    CQ_TRY(make_tmap_f32(&tmC, epi.c32, 2, dims, sc, box));
    tmR = tmC;
  } else if (ts) {
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint32_t box[2] = {64, 32};
    const cuuint64_t sc[1] = {(cuuint64_t)ldc * 2};
    CQ_TRY(make_tmap_bf16(&tmC, C, 2, dims, sc, box));
    tmR = tmC;
    if (epi.res) {
      const cuuint64_t sr[1] = {(cuuint64_t)epi.ldr * 2};
      CQ_TRY(make_tmap_bf16(&tmR, epi.res, 2, dims, sr, box));
      p.side = 1;
    } else if (epi.mul_mode) {
      CQ_TRY(make_tmap_bf16(&tmR, epi.mul_aux, 2, dims, sc, box));
      p.side = 2;
    } else if (dual) {
      CQ_TRY(make_tmap_bf16(&tmR, epi.c2, 2, dims, sc, box));
      p.dual = 1;
    }
  }
  static const bool no_lean = getenv("CQVAD_GEMM_NO_LEAN") != nullptr, no_stg1 = getenv("CQVAD_GEMM_NO_STG1") != nullptr;
  p.lean = ts && !epi.rowop && !epi.ln_g && epi.zero_period == 0 && epi.mul_mode != 2 && !no_lean;
  p.stg_single = ts && !epi.rowop && pair && !p.side && !dual && K >= 8 * BLOCK_K && !no_stg1;   // load-bound shapes only: measured +7..10% at K >= 512, -12% at K = 256 (epilogue-bound)
  const size_t smem = ts ? SMEM_BYTES_TS : SMEM_BYTES;
#define CQ_LAUNCH_TC(KERN)                                                                            \
  do {                                                                                                \
    if (extra && ts) KERN<true, true><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);         \
    else if (extra) KERN<true, false><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);         \
    else if (ts) KERN<false, true><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);            \
    else KERN<false, false><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);                   \
  } while (0)
  static const bool no_spec = getenv("CQVAD_GEMM_NO_SPEC") != nullptr;
  if (epi.rowop) {
    if (pair) gemm_tc2_rowop_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
    else gemm_tc_rowop_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
  } else if (ts && !no_spec && p.dual) {          // dual-GELU epilogue only
    if (pair) gemm_tc2_spec_kernel<true, 2><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
    else gemm_tc_spec_kernel<true, 2><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
  } else if (ts && !no_spec && !extra && epi.ln_g && getenv("CQVAD_GEMM_NO_SPEC_LN") == nullptr) {   // LayerNorm epilogue only
    if (pair) gemm_tc2_spec_kernel<false, 3><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
    else gemm_tc_spec_kernel<false, 3><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
  } else if (ts && !no_spec && p.lean) {          // lean half-tile epilogues only
    if (pair) { if (extra) gemm_tc2_spec_kernel<true, 1><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
                else gemm_tc2_spec_kernel<false, 1><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p); }
    else { if (extra) gemm_tc_spec_kernel<true, 1><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p);
           else gemm_tc_spec_kernel<false, 1><<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, tmC, tmR, p); }
  } else if (pair) CQ_LAUNCH_TC(gemm_tc2_kernel);
  else CQ_LAUNCH_TC(gemm_tc_kernel);
#undef CQ_LAUNCH_TC
  CQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace cqvad
