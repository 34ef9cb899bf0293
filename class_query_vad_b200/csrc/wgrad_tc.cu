// tcgen05 weight-gradient GEMM for sm_100a:  dW[n,k] += sum_m dY[m,n] * X[m,k]   (bf16 operands, fp32 accumulation in TMEM)
//
// The reduction runs over the ROW index of two row-major activations, i.e. both UMMA operands are "MN-major" (the
// non-reduced index is the contiguous one).  No transposed copy is made: TMA boxes of [64 rows x 64 channels] land in
// shared memory with the 128-byte swizzle exactly in the canonical MN-major SWIZZLE_128B layout
// ((8 x 16 B, n atoms) x (8 rows, k groups)), described to the tensor core by an MN-major shared-memory descriptor
// (LBO = byte distance between 64-channel atoms = one box, SBO = 8 rows x 128 B) and a_major = b_major = 1 in the
// instruction descriptor.  One UMMA is 128(n) x 256(k) x 16(rows); a CTA owns a 256x256 tile of dW (two accumulators =
// all 512 TMEM columns) over a contiguous range of rows, warp-specialised (TMA producer / MMA issuer / 4 epilogue warps)
// with a 3-stage 64 KB ring.
//
// 3x3 conv (ConvBlock, models/detr/dab_transformer.py:81,90): one job per (filter tap, row range); the X operand is the
// same shifted 3-D TMA box (64 ch x w x rt image rows) the forward implicit GEMM uses, so the horizontal halo is the TMA
// out-of-bounds zero fill and the vertical halo the zero separator rows; stage rows beyond w*rt stay zero (zeroed once).
//
// Split-K: every CTA adds its 256x256 fp32 partial tile straight into dW with TMA reduce-stores
// (cp.reduce.async.bulk.tensor .add): the accumulators go TMEM -> registers -> 128-byte-swizzled [32 x 32] fp32 boxes staged
// in the (by then idle) operand ring -> L2 reduction.  No scratch buffer, no second kernel (the former partial-tile write
// + reduce pass cost 1.7 ms of a 35 ms training step), edge tiles clipped by the tensor map.
#include "common.cuh"
#include "tc_common.cuh"
#include "bwd.cuh"
#include <mutex>

namespace cqvad {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box);
int tc_num_sms();

using namespace tc;

namespace {

constexpr int WROWS = 64, WSTAGES = 3;
constexpr int BOX_BYTES = WROWS * 128;                 // one [64 rows x 64 ch] box
constexpr int OPER_BYTES = 4 * BOX_BYTES;              // 256 channels of one operand per stage (32 KB)
constexpr int STAGE_BYTES = 2 * OPER_BYTES;            // dY tile + X tile
constexpr int WG_SMEM = WSTAGES * STAGE_BYTES + 1024 + 256;
constexpr int WG_THREADS = 192;
constexpr int TILE = 256;

struct WgParams {
  long M; int Nout, Kin;
  int n_tiles, k_tiles;
  long rows_per_split;    // non-conv: rows (multiple of 64); conv: image rows of the padded layout (multiple of rt)
  int conv, cw, rt; long ny;
  float* db;              // bias gradient (column sums of dY) accumulated by the otherwise idle epilogue warps, or nullptr
};

// MN-major, SWIZZLE_128B: start>>4 | LBO>>4 (one box) | SBO>>4 (8 rows x 128 B) | version 1 | layout 2
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(BOX_BYTES >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmW, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + WSTAGES * STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * WSTAGES, tfull_bar = bars + 16 * WSTAGES, tmem_slot = tfull_bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- job decode ----
  const int job = blockIdx.x;
  int nt = 0, kt = 0, tap = 0; long split;
  if (p.conv) { tap = job % 9; split = job / 9; }
  else { const int t = job % (p.n_tiles * p.k_tiles); split = job / (p.n_tiles * p.k_tiles); nt = t / p.k_tiles; kt = t % p.k_tiles; }
  const long total = p.conv ? p.ny : p.M;
  const long r0 = split * p.rows_per_split;
  const long r1 = r0 + p.rows_per_split < total ? r0 + p.rows_per_split : total;
  const int step = p.conv ? p.rt : WROWS;
  const int iters = (int)((r1 - r0 + step - 1) / step);
  const int valid_rows = p.conv ? p.cw * p.rt : WROWS;
  const int ksteps = (valid_rows + 15) / 16;
  const int n_slabs = (p.Nout - nt * TILE) > 128 ? 2 : 1;
  const bool do_colsum = p.db != nullptr && kt == 0 && tap == 0;   // one (k-tile, tap) per n-tile and row range

  if (p.conv && valid_rows < WROWS) {   // rows the TMA boxes never write must read as zeros
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < WSTAGES * STAGE_BYTES / 16; i += WG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    // a stage is released by the MMA commit and, when this job also sums dY's columns, by the column-sum warps
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, do_colsum ? 2 : 1); }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t box_tx = (uint32_t)(p.conv ? p.cw * p.rt * 128 : BOX_BYTES);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        const uint32_t fb = full_bar + 8 * stage;
        mbar_arrive_expect_tx(fb, 8 * box_tx);
        const uint32_t sY = smem_base + stage * STAGE_BYTES, sX = sY + OPER_BYTES;
        const long r = r0 + (long)it * step;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p.conv) {
            tma_load_3d(sY + j * BOX_BYTES, &tmY, fb, j * 64, 0, (int)r);
            tma_load_3d(sX + j * BOX_BYTES, &tmX, fb, j * 64, tap % 3 - 1, (int)r + tap / 3 - 1);
          } else {
            tma_load_2d(sY + j * BOX_BYTES, &tmY, fb, nt * TILE + j * 64, (int)r);
            tma_load_2d(sX + j * BOX_BYTES, &tmX, fb, kt * TILE + j * 64, (int)r);
          }
        }
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);   // A and B MN-major
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sY = smem_base + stage * STAGE_BYTES, sX = sY + OPER_BYTES;
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t b_desc = make_desc_mn_sw128(sX + j * 2048);
          for (int s = 0; s < n_slabs; ++s) {
            const uint64_t a_desc = make_desc_mn_sw128(sY + s * 2 * BOX_BYTES + j * 2048);
            umma_bf16(tmem_base + (uint32_t)(s * 256), a_desc, b_desc, idesc, (it | j) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty_bar + 8 * stage);
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter q = warp % 4 =====
    const int q = warp & 3;
    if (do_colsum) {
      // db[n] += sum_rows dY[row, n]: thread e owns columns 2e, 2e+1 of the 256-column dY tile (box 2e/64, 16-byte chunk
      // swizzled by the row): a warp reads one 128-byte box row per instruction -- bank-conflict free
      const int e = (warp - 2) * 32 + lane;            // 0..127
      const int col = 2 * e;
      const uint32_t boxo = (uint32_t)(col >> 6) * BOX_BYTES, chunk = (uint32_t)((col & 63) >> 3), inner = (uint32_t)(col & 7) * 2u;
      float s0 = 0.f, s1 = 0.f;
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full_bar + 8 * stage, phase);
        const uint32_t sY = smem_base + stage * STAGE_BYTES + boxo + inner;
#pragma unroll 8
        for (int r = 0; r < WROWS; ++r) {
          uint32_t v;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sY + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
          const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
          s0 += f.x; s1 += f.y;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // all four warps are done with this stage
        if (e == 0) mbar_arrive(empty_bar + 8 * stage);
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
      const int n = nt * TILE + col;
      if (n < p.Nout) atomicAdd(p.db + n, s0);
      if (n + 1 < p.Nout) atomicAdd(p.db + n + 1, s1);
    }
    if (iters > 0) {
      mbar_wait(tfull_bar, 0);      // every MMA has retired: the accumulators are final and the operand ring is idle
      tc_fence_after();
      // this warp: rows [q*32, +32) of each 128-row slab, 256 fp32 columns = 8 boxes of [32 rows x 32 cols] (128 B per row),
      // staged in 8 x 4 KB of the ring and added into dW by TMA
      const uint32_t stg = smem_base + (uint32_t)((warp - 2) * 8 * 4096);
      const uint32_t my_row = stg + (uint32_t)(lane * 128);
      const int sw = lane & 7;
      const int col0 = (p.conv ? tap * 256 : kt * TILE);
      for (int s = 0; s < n_slabs; ++s) {
        if (s > 0) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
        const int row0 = nt * TILE + s * 128 + q * 32;
#pragma unroll 1
        for (int c = 0; c < TILE; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 256 + c), r);
          tmem_ld_wait();
          const uint32_t buf = my_row + (uint32_t)((c >> 5) * 4096);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(buf + (uint32_t)((j ^ sw) << 4), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_reduce_add_2d(&tmW, stg + (uint32_t)((c >> 5) * 4096), col0 + c, row0); bulk_commit(); }
        }
      }
      if (lane == 0) bulk_wait_read<0>();
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

// db[n] += sum_m dY[m,n]: warp per row slice of 256 columns, 8 per lane; block partials through shared memory
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dY, long lddy, float* __restrict__ db, long M,
                                                     int Nout) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < Nout) {
    for (long row = (long)blockIdx.x * 8 + warp; row < M; row += (long)gridDim.x * 8) {
      float v[8];
      load8(dY + row * lddy + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c < Nout) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(db + c, s);
  }
}

std::once_flag g_wg_once;
int g_wg_err = 0;
float* g_scratch = nullptr;
size_t g_scratch_bytes = 0;

}  // namespace

// kept for the callers' sake: the split-K partials no longer need a scratch buffer (TMA reduce-stores straight into dW)
void set_wgrad_scratch(float* p, size_t bytes) { g_scratch = p; g_scratch_bytes = bytes; }
size_t wgrad_scratch_bytes() { return 256; }

int wgrad_tc(const bf16* dY, long lddy, const bf16* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
             const ConvGeom* conv, cudaStream_t st) {
  if (M < 1024 || lddy % 8 != 0 || ldx % 8 != 0 || Nout % 8 != 0 || Kin % 8 != 0 || ldw % 4 != 0) return 1;
  if ((((uintptr_t)dY) & 15) || (((uintptr_t)X) & 15) || (dW && (((uintptr_t)dW) & 15))) return 1;
  const int sms = tc_num_sms();
  if (sms <= 0) return 1;
  std::call_once(g_wg_once, [] {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) != cudaSuccess) g_wg_err = 1;
  });
  if (g_wg_err) return set_error(CQVAD_E_CUDA, "wgrad_tc: cannot reserve %d bytes of shared memory", WG_SMEM);

  WgParams p{};
  p.M = M; p.Nout = Nout; p.Kin = Kin; p.db = dW ? db : nullptr;
  CUtensorMap tmY, tmX;
  int tiles, splits;
  if (conv) {
    const int w = conv->w;
    if (w > WROWS || lddy != kC || ldx != kC || Nout != kC || Kin != kC || M % w != 0) return 1;
    const int rt = WROWS / w;
    const long ny = M / w;
    p.conv = 1; p.cw = w; p.rt = rt; p.ny = ny; p.n_tiles = p.k_tiles = 1;
    tiles = 9;
    splits = sms / 9 < 1 ? 1 : sms / 9;
    long yps = cdiv(ny, splits);
    yps = cdiv(yps, rt) * rt;
    splits = (int)cdiv(ny, yps);
    p.rows_per_split = yps;
    const cuuint64_t dims[3] = {(cuuint64_t)kC, (cuuint64_t)w, (cuuint64_t)ny};
    const cuuint64_t strides[2] = {(cuuint64_t)kC * 2, (cuuint64_t)w * kC * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)w, (cuuint32_t)rt};
    CQ_TRY(make_tmap_bf16(&tmY, dY, 3, dims, strides, box));
    CQ_TRY(make_tmap_bf16(&tmX, X, 3, dims, strides, box));
  } else {
    p.conv = 0;
    p.n_tiles = (Nout + TILE - 1) / TILE; p.k_tiles = (Kin + TILE - 1) / TILE;
    tiles = p.n_tiles * p.k_tiles;
    if (tiles > sms) return 1;
    splits = sms / tiles;
    long rps = cdiv(M, splits);
    rps = cdiv(rps, WROWS) * WROWS;
    splits = (int)cdiv(M, rps);
    p.rows_per_split = rps;
    {
      const cuuint64_t dims[2] = {(cuuint64_t)Nout, (cuuint64_t)M};
      const cuuint64_t strides[1] = {(cuuint64_t)lddy * 2};
      const cuuint32_t box[2] = {64, WROWS};
      CQ_TRY(make_tmap_bf16(&tmY, dY, 2, dims, strides, box));
    }
    {
      const cuuint64_t dims[2] = {(cuuint64_t)Kin, (cuuint64_t)M};
      const cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};
      const cuuint32_t box[2] = {64, WROWS};
      CQ_TRY(make_tmap_bf16(&tmX, X, 2, dims, strides, box));
    }
  }
  const int jobs = tiles * splits;
  if (dW) {
    CUtensorMap tmW;
    const cuuint64_t dims[2] = {(cuuint64_t)(conv ? 9 * kC : Kin), (cuuint64_t)Nout};
    const cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
    const cuuint32_t box[2] = {32, 32};
    CQ_TRY(make_tmap_f32(&tmW, dW, 2, dims, strides, box));
    wgrad_tc_kernel<<<jobs, WG_THREADS, WG_SMEM, st>>>(tmY, tmX, tmW, p);
    CQ_LAUNCH_CHECK();
  }
  if (db && !dW) {   // bias gradient alone (otherwise summed inside wgrad_tc_kernel)
    long gx = cdiv(M, 8 * 16);
    if (gx > 4 * sms) gx = 4 * sms;
    dim3 grid((unsigned)gx, (unsigned)cdiv(Nout, 256));
    colsum_kernel<<<grid, 256, 0, st>>>(dY, lddy, db, M, Nout);
    CQ_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace cqvad
