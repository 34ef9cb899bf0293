// ViT simple-feature-pyramid neck (models/backbone_3d_builder.py:133-182 `lateral_convs`, applied by
// space_forward :190-200): one pyramid level from a ViT feature map [B, C_in, T, H, W], written straight into the encoder's token
// sequence [B, Len, 256].  Level kinds (the reference's scales 4 / 2 / 1 / 0.5):
//   0: ConvT(C,C/2,(1,2,2)) -> LN(C/2) -> GELU -> ConvT(C/2,C/4,(1,2,2)) -> Conv1x1x1(C/4,256) -> LN(256) -> Conv3x3x3(256,256)
//   1: ConvT(C,C/2,(1,2,2))                                              -> Conv1x1x1(C/2,256) -> LN(256) -> Conv3x3x3
//   2:                                                                      Conv1x1x1(C,256)   -> LN(256) -> Conv3x3x3
//   3: MaxPool3d((1,2,2))                                                -> Conv1x1x1(C,256)   -> LN(256) -> Conv3x3x3
// Everything runs token-major (one row = one voxel, channels contiguous):
//   * ConvTranspose3d with kernel = stride = (1,2,2) is a GEMM [voxels, C] x [C, 4*C_out] (columns ordered (dy, dx, c_out)) followed by
//     a pixel shuffle of 2x2 column groups into rows -- the tcgen05 GEMM of gemm_tc.cu does the arithmetic;
//   * the channel-first LayerNorm of the reference (:20-40, eps 1e-6) is a row LayerNorm here (fused into the 1x1x1 GEMM epilogue
//     for the 256-channel one);
//   * Conv3d 3x3x3 = three time taps of the CTA-pair implicit-GEMM 3x3 conv (the decoder ConvBlock's kernel) on a layout padded in y
//     AND in time: [1 + B*(T+1) + 1 images][H+1 rows][W][256] with one zero separator row per image and one zero image per clip, so that
//     both the vertical and the temporal halo read zeros; the partial sums of the taps chain through the residual input of the epilogue.
#include <algorithm>
#include "common.cuh"

namespace cqvad {
namespace {

// [B, C, N] channel-first -> rows [B*N, C]
template <typename T>
__global__ void __launch_bounds__(256) cf_to_rows_kernel(const T* __restrict__ x, T* __restrict__ out, long N, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long n0 = (long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* xb = x + (long)b * C * N;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const long n = n0 + tx;
    tile[r][tx] = n < N ? to_f(xb[(long)(c0 + r) * N + n]) : 0.f;
  }
  __syncthreads();
  T* ob = out + (long)b * N * C;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const long n = n0 + r;
    if (n < N) ob[n * C + c0 + tx] = from_f<T>(tile[tx][r]);
  }
}

// rows [BT*H*W, 4*Co] (columns (dy, dx, co)) -> rows [BT*2H*2W, Co]; one thread = 8 channels
template <typename T>
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const T* __restrict__ in, T* __restrict__ out, long BT, int H, int W, int Co) {
  const int c8 = Co / 8;
  const long total = BT * 4L * H * W * c8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    long r = i / c8;                                   // output row ((bt*2H + oy)*2W + ox)
    const int ox = (int)(r % (2 * W)); r /= 2 * W;
    const int oy = (int)(r % (2 * H)); const long bt = r / (2 * H);
    const long src = ((bt * H + (oy >> 1)) * W + (ox >> 1)) * (4L * Co) + (long)((oy & 1) * 2 + (ox & 1)) * Co + c * 8;
    float v[8];
    load8(in + src, v);
    store8(out + (i / c8) * Co + c * 8, v);
  }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// row LayerNorm over C channels (C % 32 == 0), optional exact GELU; warp per row, three passes over an L1-resident row
template <typename T>
__global__ void __launch_bounds__(256) ln_rows_kernel(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                      float eps, int gelu, T* __restrict__ out, long rows, int C) {
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = to_f(xr[c]) - mean; q = fmaf(d, d, q); }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)C + eps);
  T* orow = out + row * C;
  for (int c = lane; c < C; c += 32) {
    float v = (to_f(xr[c]) - mean) * rstd * g[c] + b[c];
    if (gelu) v = gelu_erf(v);
    orow[c] = from_f<T>(v);
  }
}

// MaxPool3d(kernel = stride = (1,2,2)) on rows [BT*H*W, C] -> [BT*(H/2)*(W/2), C] (floor); one thread = 8 channels
template <typename T>
__global__ void __launch_bounds__(256) maxpool122_kernel(const T* __restrict__ in, T* __restrict__ out, long BT, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, c8 = C / 8;
  const long total = BT * (long)Ho * Wo * c8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    long r = i / c8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long bt = r / Ho;
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        load8(in + ((bt * H + 2 * oy + dy) * W + 2 * ox + dx) * (long)C + c * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    store8(out + (i / c8) * C + c * 8, m);
  }
}

// dense rows [B*T*H*W, 256] <-> the (y, t)-padded conv layout: image index 1 + b*(T+1) + t, (H+1)*W rows per image
template <typename T, bool TO_PADDED>
__global__ void __launch_bounds__(256) pad_copy_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Tn, int H, int W,
                                                       long Len, long level_start) {
  const long rows = (long)B * Tn * H * W;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= rows) return;
  const long hw = (long)H * W;
  const long s = wid % hw;
  const int t = (int)((wid / hw) % Tn);
  const int b = (int)(wid / (hw * Tn));
  const long img = TO_PADDED ? 1 + (long)b * (Tn + 1) + t : (long)b * (Tn + 1) + t;   // the conv output has no leading zero image
  const long prow = img * (long)(H + 1) * W + s;
  float v[8];
  if (TO_PADDED) {
    load8(in + wid * kC + lane * 8, v);
    store8(out + prow * kC + lane * 8, v);
  } else {   // -> tokens [B, Len, 256] at level_start
    load8(in + prow * kC + lane * 8, v);
    store8(out + ((long)b * Len + level_start + (long)t * hw + s) * kC + lane * 8, v);
  }
}

struct Bump {
  char* p; size_t off = 0, cap;
  Bump(void* base, size_t c) : p((char*)base), cap(c) {}
  template <typename U> U* take(size_t n) {
    U* r = p ? (U*)(p + off) : nullptr;
    off = (off + n * sizeof(U) + 255) & ~(size_t)255;
    return r;
  }
};

inline unsigned grid_for(long work_items) { return (unsigned)std::min<long>(cdiv(work_items, 256), 148L * 16); }

// weights (host-packed, modules/neck.py): kind 0: [ct1_w [4*C/2, C], ct1_b4 [4*C/2] f32, ln1_g, ln1_b [C/2] f32, ct2_w [4*C/4, C/2],
//   ct2_b4 [4*C/4] f32, conv1_w [256, C/4], ln_g, ln_b [256] f32, conv3_w [3][256][9*256]];  kind 1: [ct1_w, ct1_b4, conv1_w [256, C/2], ln_g,
//   ln_b, conv3_w];  kinds 2 / 3: [conv1_w [256, C], ln_g, ln_b, conv3_w]
template <typename T>
int neck_level(int kind, const T* x, const void* const* w, T* tokens, long Len, long level_start, void* ws, size_t ws_bytes, int B,
               int Cin, int Tn, int H, int W, cudaStream_t st, size_t* need) {
  const long BT = (long)B * Tn;
  int Ho = H, Wo = W;
  if (kind == 0) { Ho = 4 * H; Wo = 4 * W; } else if (kind == 1) { Ho = 2 * H; Wo = 2 * W; } else if (kind == 3) { Ho = H / 2; Wo = W / 2; }
  const long rows_in = BT * H * W, rows_out = BT * Ho * Wo;
  const long img_rows = (long)(Ho + 1) * Wo, n_img = (long)B * (Tn + 1);
  Bump a(need ? nullptr : ws, ws_bytes);
  T* xr = a.take<T>(rows_in * Cin);
  T* t1 = nullptr; T* t2 = nullptr; T* t3 = nullptr; T* t4 = nullptr;
  if (kind == 0) {
    t1 = a.take<T>(rows_in * 2 * Cin);                 // ConvT1 output before the shuffle [rows_in, 4*C/2]
    t2 = a.take<T>(rows_in * 4 * (Cin / 2));           // shuffled [4*rows_in, C/2] (LN + GELU in place)
    t3 = a.take<T>(rows_in * 4 * Cin);                 // ConvT2 output before the shuffle [4*rows_in, 4*C/4]
    t4 = a.take<T>(rows_out * (Cin / 4));              // shuffled [16*rows_in, C/4]
  } else if (kind == 1) {
    t1 = a.take<T>(rows_in * 2 * Cin);
    t4 = a.take<T>(rows_out * (Cin / 2));
  } else if (kind == 3) {
    t4 = a.take<T>(rows_out * Cin);
  }
  T* y1 = a.take<T>(rows_out * kC);                    // Conv1x1x1 + LN(256), dense rows
  T* P = a.take<T>((n_img + 2) * img_rows * kC);       // padded conv input (leading + trailing zero image)
  T* acc0 = a.take<T>(n_img * img_rows * kC);
  T* acc1 = a.take<T>(n_img * img_rows * kC);
  if (need) { *need = a.off + 256; return 0; }
  if (a.off > ws_bytes) return set_error(CQVAD_E_WORKSPACE, "vit_neck_level: workspace too small (%zu < %zu)", ws_bytes, a.off);

  {
    const dim3 grid((unsigned)cdiv((long)Tn * H * W, 32), Cin / 32, (unsigned)B);
    cf_to_rows_kernel<T><<<grid, 256, 0, st>>>(x, xr, (long)Tn * H * W, Cin);
    CQ_LAUNCH_CHECK();
  }
  const T* feat = xr; int Cf = Cin; int wi = 0;
  if (kind == 0 || kind == 1) {
    const int Co = Cin / 2;
    Epilogue e; e.bias = (const float*)w[1];
    CQ_TRY(gemm<T>(xr, Cin, (const T*)w[0], t1, 4L * Co, rows_in, 4 * Co, Cin, e, nullptr, st));
    T* sh = kind == 0 ? t2 : t4;
    pixel_shuffle_kernel<T><<<grid_for(rows_in * 4 * (Co / 8)), 256, 0, st>>>(t1, sh, BT, H, W, Co);
    CQ_LAUNCH_CHECK();
    feat = sh; Cf = Co; wi = 2;
    if (kind == 0) {
      const long r2 = rows_in * 4;
      ln_rows_kernel<T><<<(unsigned)cdiv(r2 * 32, 256), 256, 0, st>>>(t2, (const float*)w[2], (const float*)w[3], 1e-6f, 1, t2, r2, Co);
      CQ_LAUNCH_CHECK();
      const int Co2 = Cin / 4;
      Epilogue e2; e2.bias = (const float*)w[5];
      CQ_TRY(gemm<T>(t2, Co, (const T*)w[4], t3, 4L * Co2, r2, 4 * Co2, Co, e2, nullptr, st));
      pixel_shuffle_kernel<T><<<grid_for(r2 * 4 * (Co2 / 8)), 256, 0, st>>>(t3, t4, BT, 2 * H, 2 * W, Co2);
      CQ_LAUNCH_CHECK();
      feat = t4; Cf = Co2; wi = 6;
    }
  } else if (kind == 3) {
    maxpool122_kernel<T><<<grid_for(rows_out * (Cin / 8)), 256, 0, st>>>(xr, t4, BT, H, W, Cin);
    CQ_LAUNCH_CHECK();
    feat = t4;
  }
  if (rows_out == 0) return 0;
  {   // Conv3d 1x1x1 (no bias) + channel LayerNorm(256, eps 1e-6) in the GEMM epilogue
    Epilogue e; e.ln_g = (const float*)w[wi + 1]; e.ln_b = (const float*)w[wi + 2]; e.ln_eps = 1e-6f;
    CQ_TRY(gemm<T>(feat, Cf, (const T*)w[wi], y1, kC, rows_out, kC, Cf, e, nullptr, st));
  }
  CQ_CUDA(cudaMemsetAsync(P, 0, (size_t)(n_img + 2) * img_rows * kC * sizeof(T), st));
  pad_copy_kernel<T, true><<<(unsigned)cdiv(rows_out * 32, 256), 256, 0, st>>>(y1, P, B, Tn, Ho, Wo, 0, 0);
  CQ_LAUNCH_CHECK();
  {   // three time taps of the implicit-GEMM 3x3 conv; tap kt reads the image kt - 1 steps away
    const T* w3 = (const T*)w[wi + 3];
    ConvGeom geo; geo.h = Ho; geo.w = Wo;
    const long M = n_img * img_rows;
    T* bufs[2] = {acc0, acc1};
    for (int kt = 0; kt < 3; ++kt) {
      Epilogue e;
      e.zero_period = (int)img_rows; e.zero_valid = Ho * Wo;
      if (kt > 0) { e.res = bufs[(kt - 1) & 1]; e.ldr = kC; }
      CQ_TRY(gemm<T>(P + (long)kt * img_rows * kC, kC, w3 + (long)kt * kC * 9 * kC, bufs[kt & 1], kC, M, kC, 9 * kC, e, &geo, st));
    }
  }
  pad_copy_kernel<T, false><<<(unsigned)cdiv(rows_out * 32, 256), 256, 0, st>>>(acc0, tokens, B, Tn, Ho, Wo, Len, level_start);
  CQ_LAUNCH_CHECK();
  return 0;
}

int check_neck(int dtype, int kind, int B, int Cin, int Tn, int H, int W) {
  CQ_CHECK_ARG(dtype == CQVAD_F32 || dtype == CQVAD_BF16, "vit_neck_level: unknown dtype %d", dtype);
  CQ_CHECK_ARG(kind >= 0 && kind <= 3 && B >= 1 && Tn >= 1 && H >= 1 && W >= 1, "vit_neck_level: bad arguments");
  CQ_CHECK_SHAPE(Cin % 256 == 0 && Cin <= 2048, "vit_neck_level: C_in %d must be a multiple of 256 (<= 2048)", Cin);
  CQ_CHECK_SHAPE(B <= 65535 && 4 * W <= 128, "vit_neck_level: batch <= 65535, feature width <= 32 (the x4 level must fit the 128-wide conv tile)");
  return 0;
}

}  // namespace
}  // namespace cqvad

using namespace cqvad;

extern "C" size_t cqvad_vit_neck_workspace_bytes(int dtype, int kind, int B, int Cin, int T, int H, int W) {
  if (check_neck(dtype, kind, B, Cin, T, H, W) != 0) return 0;
  size_t need = 0;
  if (dtype == CQVAD_F32) neck_level<float>(kind, nullptr, nullptr, nullptr, 0, 0, nullptr, 0, B, Cin, T, H, W, nullptr, &need);
  else neck_level<bf16>(kind, nullptr, nullptr, nullptr, 0, 0, nullptr, 0, B, Cin, T, H, W, nullptr, &need);
  return need;
}

extern "C" int cqvad_vit_neck_level(int dtype, int kind, const void* x, const void* const* weights, void* tokens, long Len,
                                    long level_start, void* workspace, size_t ws_bytes, int B, int Cin, int T, int H, int W,
                                    void* stream) {
  CQ_TRY(check_neck(dtype, kind, B, Cin, T, H, W));
  CQ_CHECK_ARG(x && weights && tokens && workspace, "vit_neck_level: null pointer");
  int Ho = H, Wo = W;
  if (kind == 0) { Ho = 4 * H; Wo = 4 * W; } else if (kind == 1) { Ho = 2 * H; Wo = 2 * W; } else if (kind == 3) { Ho = H / 2; Wo = W / 2; }
  CQ_CHECK_ARG(level_start >= 0 && level_start + (long)T * Ho * Wo <= Len, "vit_neck_level: level does not fit the token sequence");
  const int nw = kind == 0 ? 10 : (kind == 1 ? 6 : 4);
  for (int i = 0; i < nw; ++i) CQ_CHECK_ARG(weights[i] != nullptr, "vit_neck_level: weights[%d] is NULL", i);
  if (dtype == CQVAD_F32)
    return neck_level<float>(kind, (const float*)x, weights, (float*)tokens, Len, level_start, workspace, ws_bytes, B, Cin, T, H, W,
                             as_stream(stream), nullptr);
  return neck_level<bf16>(kind, (const bf16*)x, weights, (bf16*)tokens, Len, level_start, workspace, ws_bytes, B, Cin, T, H, W,
                          as_stream(stream), nullptr);
}
