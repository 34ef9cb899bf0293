/*
 * cqvad.h -- C ABI of libcqvad.so: the B200 (sm_100a) implementation of the class-query decoder hot path of
 * dlrudco/class-query-vad.  Plain pointers and sizes only; no torch types.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative CQVAD_E_* code otherwise; cqvad_last_error() gives the thread-local message.
 *     Nothing throws.  (The reference raises C++ exceptions through pybind: ops/src/ms_deform_attn.h:20-61; its
 *     kernel-launch errors are only printf'ed, ops/src/cuda/ms_deform_im2col_cuda_t.cuh:1123-1127 -- here they
 *     are returned.)
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are contiguous, row-major.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation, no device allocation:
 *     the caller owns outputs and passes an explicit workspace where one is needed.
 *   - `dtype` selects the storage/compute type of activations and matrix weights: CQVAD_F32 (CUDA-core FFMA path,
 *     the "rel 1e-3, TF32 off" parity mode) or CQVAD_BF16 (tcgen05 tensor-core path, fp32 accumulation).
 *     Biases, LayerNorm affine parameters, reference points and the four "small-N" linears are always fp32.
 */
#ifndef CQVAD_H_
#define CQVAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CQVAD_VERSION 1

enum { CQVAD_F32 = 0, CQVAD_BF16 = 1 };
enum { CQVAD_OK = 0, CQVAD_E_INVALID_ARG = -1, CQVAD_E_UNSUPPORTED_SHAPE = -2, CQVAD_E_CUDA = -3, CQVAD_E_WORKSPACE = -4 };
enum { CQVAD_ACT_NONE = 0, CQVAD_ACT_RELU = 1, CQVAD_ACT_GELU = 2 };

int cqvad_version(void);
const char* cqvad_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * ops/ path: 3-D multi-scale deformable attention.
 * Replaces the pybind functions `ms_deform_attn_forward` / `ms_deform_attn_backward` of the reference module
 * `MultiScaleDeformableAttention` (ops/src/vision.cpp:13-16, ops/src/ms_deform_attn.h:20-61,
 * ops/src/cuda/ms_deform_attn_cuda_t.cu:20-80 and :83-153), i.e. what
 * `MSDeformAttnFunction.forward/backward` bind (ops/functions/ms_deform_attn_func.py:26,42).
 *   value   [N, Len, M, D]  (dtype)          shapes      [L,3] int64 (T,H,W)     level_start [L] int64
 *   loc     [N, Lq, M, L, P, 3] fp32 (x,y,t) in [0,1]     attn   [N, Lq, M, L, P] fp32
 *   out     [N, Lq, M*D]    (dtype)
 * No im2col_step: the whole batch is one launch (the reference chunks the batch, ms_deform_attn_cuda_t.cu:61-75).
 * backward: grad_value [N,Len,M,D] fp32 must be ZERO-FILLED by the caller (accumulated with red.global.add);
 * grad_loc / grad_attn fp32 are fully overwritten.  The gradient is the mathematical gradient of the forward
 * (the reference backward kernel is not: SURVEY.md section 8a, DESIGN.md "Deliberate divergences").
 */
int cqvad_msda3d_forward(int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                         const float* loc, const float* attn, void* out,
                         int N, int Len, int M, int D, int L, int Lq, int P, void* stream);
int cqvad_msda3d_backward(int dtype, const void* value, const int64_t* shapes, const int64_t* level_start,
                          const float* loc, const float* attn, const void* grad_out,
                          float* grad_value, float* grad_loc, float* grad_attn,
                          int N, int Len, int M, int D, int L, int Lq, int P, void* stream);
/* Integer part of the sampling (the bit-exact contract, cuh:38-60,424-428): per (n,q,m,l,p) the int32 low corner
 * and an 8-bit corner-validity mask (bit k = corner v(k+1) of cuh:63-109 is read; 0 = point skipped). */
int cqvad_msda3d_indices(const int64_t* shapes, const float* loc, int32_t* t_low, int32_t* h_low, int32_t* w_low,
                         uint8_t* corner_mask, int N, int Lq, int M, int L, int P, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Building blocks (used by the drop-in nn.Modules; each is also a step of cqvad_decoder_forward).
 */
/* y = LayerNorm(x (+ res)) over the last dim C (C == 256).  Replaces F.layer_norm call sites
 * models/detr/dab_transformer.py:92,826-827,938,946,992,996,1045,1054,1065,1076.  `out_f32` != 0 writes fp32. */
int cqvad_layernorm(int dtype, const void* x, const void* res, const float* gamma, const float* beta, float eps,
                    void* out, int out_f32, long rows, int C, void* stream);
/* C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ res).  nn.Linear / 1x1 conv (dab_transformer.py:47, 1067-1070).
 * BF16: tcgen05 kernel when K % 64 == 0 and N % 8 == 0, CUDA-core kernel otherwise. */
int cqvad_linear(int dtype, const void* A, const void* W, const float* bias, const void* res, void* C,
                 long M, int N, int K, int act, void* stream);
/* Training forward of a GELU linear layer (ConvBlock conv2 + nn.GELU, dab_transformer.py:84,93-94 under autograd):
 * act_out[M,N] = gelu(A . W^T + bias) and dact_out[M,N] = gelu'(A . W^T + bias), both written by the GEMM epilogue; the
 * pre-activation is never stored (autograd keeps it, the backward here needs only gelu').  BF16: tcgen05 kernel. */
int cqvad_linear_gelu_train(int dtype, const void* A, const void* W, const float* bias, void* act_out, void* dact_out,
                            long M, int N, int K, void* stream);
/* Data gradient through an activation in one epilogue: dX[M,N] = (dY[M,K] . Wt[N,K]^T) * aux[M,N] (mode 3: aux holds the
 * stored derivative) or masked by aux > 0 (mode 1: aux is the ReLU output).  Wt is the transposed weight [in, out]. */
int cqvad_linear_dgrad_act(int dtype, const void* dY, const void* Wt, const void* aux, int mode, void* dX, long M, int N,
                           int K, void* stream);
/* attention_weights = softmax over the L*P logits of a head; sampling_locations = reference_points + offsets / (T_l, W_l, H_l)
 * (ops/modules/ms_deform_attn.py:187-192, the reference's normaliser order against (x, y, t) offsets), 8 heads:
 * offsets [rows, 8, L, P, 3], logits [rows, 8, L*P], reference_points [rows, L, 3] (all fp32), shapes [L,3] int64 on the device
 * -> loc [rows, 8, L, P, 3], attn [rows, 8, L, P] (fp32), the inputs of cqvad_msda3d_forward.  rows = N * Len_q. */
int cqvad_msda3d_prepare(const float* offsets, const float* logits, const float* reference_points, const int64_t* shapes,
                         float* loc, float* attn, long rows, int L, int P, void* stream);
/* One deformable encoder layer around the MSDA-3D op:
 * DeformableTransformerEncoderLayer.forward (models/detr/dab_transformer.py:513-523) with MSDeformAttn3D.forward
 * (ops/modules/ms_deform_attn.py:167-203) inlined, eval semantics (dropout = identity), 8 heads, d_model 256.
 *   src, pos, out [B, Len, 256] (dtype); reference_points [B, Len, L, 3] fp32 ((x,y,t), get_reference_points :433-452);
 *   shapes [L,3] (T,H,W) / level_start [L] int64 ON THE DEVICE (as for cqvad_msda3d_forward); padding_mask [B, Len] u8 or NULL;
 *   weights: 16 device pointers in state_dict order -- self_attn.{sampling_offsets,attention_weights,value_proj,output_proj}
 *   .{weight,bias}, norm1.{weight,bias}, linear1.{weight,bias}, linear2.{weight,bias}, norm2.{weight,bias}; matrices in
 *   `dtype`, biases / LayerNorm vectors fp32.  attn_out (optional, [B, Len, 256]) receives the attention module's own output
 *   (output_proj of the sampled values, before the residual).  The offsets / logits GEMMs hand fp32 values to the location /
 *   softmax kernel in both dtypes. */
int cqvad_deform_encoder_layer_num_weights(void);
size_t cqvad_deform_encoder_layer_workspace_bytes(int dtype, int B, long Len, int L, int P, int F);
int cqvad_deform_encoder_layer_forward(int dtype, const void* const* weights, const void* src, const void* pos,
                                       const float* reference_points, const int64_t* shapes, const int64_t* level_start,
                                       const uint8_t* padding_mask, void* out, void* attn_out, void* workspace,
                                       size_t workspace_bytes, int B, long Len, int L, int P, int F, void* stream);
/* Training pair of the encoder layer (the reference differentiates DeformableTransformerEncoderLayer.forward with autograd,
 * train.py:151; the sampling gradient is the mathematical one, see cqvad_msda3d_backward).  train_forward = the same result as
 * cqvad_deform_encoder_layer_forward with unfused LayerNorms / FFN, keeping q, value, locations, attention weights, the
 * sampled values, both pre-norm sums and the FFN hidden in `workspace`; backward consumes that workspace (same pointer, same
 * size, untouched in between), writes grad_src / grad_pos [B, Len, 256] (dtype) and ACCUMULATES the 16 parameter gradients
 * (fp32, state_dict order, caller zero-fills).  dropout_p > 0: dropout1 / dropout2 / dropout3 of dab_transformer.py:499-519 with
 * Philox masks keyed by `seed` (the backward must receive the forward's p and seed; 0 = eval semantics).  reference_points get no
 * gradient (derived from shapes). */
size_t cqvad_deform_encoder_layer_train_workspace_bytes(int dtype, int B, long Len, int L, int P, int F);
int cqvad_deform_encoder_layer_train_forward(int dtype, const void* const* weights, const void* src, const void* pos,
                                             const float* reference_points, const int64_t* shapes,
                                             const int64_t* level_start, const uint8_t* padding_mask, void* out,
                                             void* workspace, size_t workspace_bytes, int B, long Len, int L, int P, int F,
                                             float dropout_p, uint64_t seed, void* stream);
int cqvad_deform_encoder_layer_backward(int dtype, const void* const* weights, const void* src, const int64_t* shapes,
                                        const int64_t* level_start, const uint8_t* padding_mask, const void* grad_out,
                                        void* grad_src, void* grad_pos, float* const* grad_weights, void* workspace,
                                        size_t workspace_bytes, int B, long Len, int L, int P, int F, float dropout_p, uint64_t seed,
                                        void* stream);
/* Input projection of one backbone level (CSN configurations, models/model.py:64-71,162-164):
 * tokens[b, level_start + n, :] = GroupNorm(32, 256)( Conv3d(C_in, 256, kernel_size = 1)(x) )[b, :, n] for x [B, C_in, N = T*H*W]
 * channel-first (dtype): transpose to token-major, tcgen05 GEMM with W [256, C_in] (dtype) + bias, per-clip group statistics
 * over (8 channels x N), normalise + affine (gn_weight / gn_bias fp32 [256]) straight into the encoder's token sequence
 * [B, Len, 256] (conv + norm + flatten in one call).  C_in % 32 == 0 (bf16 tensor-core path: % 64).  The extra stride-2 level
 * (Conv3d kernel 3, models/model.py:72-76) is cqvad_input_proj_3x3s2_gn. */
size_t cqvad_input_proj_workspace_bytes(int dtype, int B, int Cin, long N);
int cqvad_input_proj_1x1_gn(int dtype, const void* x, const void* weight, const float* bias, const float* gn_weight,
                            const float* gn_bias, float eps, void* tokens, void* workspace, size_t workspace_bytes, int B, int Cin,
                            long N, long Len, long level_start, void* stream);
/* The extra pyramid level of the non-ViT configurations (models/model.py:72-76,166-170): Conv3d(C_in, 256, kernel_size = 3,
 * stride = (1, 2, 2), padding = 1) -> GroupNorm(32, 256) -> flatten, written at level_start of the token sequence.
 * x [B, C_in, T, H, W] channel-first (dtype); weight_taps [256, 27, C_in] (dtype) = conv.weight.permute(0, 2, 3, 4, 1) flattened
 * (tap = (kt * 3 + ky) * 3 + kx).  Output positions: T x ((H-1)/2+1) x ((W-1)/2+1).  im2col gather + tcgen05 GEMM + group norm. */
size_t cqvad_input_proj_3x3s2_workspace_bytes(int dtype, int B, int Cin, int T, int H, int W);
int cqvad_input_proj_3x3s2_gn(int dtype, const void* x, const void* weight_taps, const float* bias, const float* gn_weight,
                              const float* gn_bias, float eps, void* tokens, void* workspace, size_t workspace_bytes, int B, int Cin,
                              int T, int H, int W, long Len, long level_start, void* stream);
/* ViT simple-feature-pyramid neck: one `lateral_convs[kind]` of the reference Backbone
 * (models/backbone_3d_builder.py:133-182; space_forward :190-200) applied to a ViT feature map x [B, C_in, T, H, W] (dtype,
 * channel-first) and written into the encoder's token sequence tokens [B, Len, 256] at level_start (conv + norm + flatten).
 * kind 0 = scale 4 (ConvTranspose3d x2 with channel LayerNorm + GELU between; output T x 4H x 4W), 1 = scale 2 (one ConvTranspose3d;
 * T x 2H x 2W), 2 = scale 1, 3 = scale 0.5 (MaxPool3d (1,2,2); T x H/2 x W/2); each followed by Conv3d 1x1x1 (no bias) -> channel
 * LayerNorm(256, eps 1e-6) -> Conv3d 3x3x3 (padding 1, no bias).  weights: the host-packed table documented in csrc/neck.cu
 * (class_query_vad_b200/modules/neck.py::pack_neck_weights): ConvTranspose weights as [(dy,dx,c_out), c_in] (dtype) with the bias
 * tiled 4x (fp32), 1x1x1 weight [256, C] (dtype), LayerNorm affine fp32, 3x3x3 weight as [kt][256][(ky,kx), 256] (dtype).
 * C_in % 256 == 0, 4*W <= 128. */
size_t cqvad_vit_neck_workspace_bytes(int dtype, int kind, int B, int Cin, int T, int H, int W);
int cqvad_vit_neck_level(int dtype, int kind, const void* x, const void* const* weights, void* tokens, long Len, long level_start,
                         void* workspace, size_t ws_bytes, int B, int Cin, int T, int H, int W, void* stream);
/* One pyramid level into the encoder's token sequence (Transformer.forward, models/detr/dab_transformer.py:310-327):
 * tokens[b, level_start + n, c] = x[b, c, n] (+ add[c]) for x [B, 256, N = T*H*W] channel-first (dtype), add = level_embed[lvl]
 * (fp32, for the position embedding; NULL for the features), tokens [B, Len, 256]. */
int cqvad_level_to_tokens(int dtype, const void* x, const float* add, void* tokens, int B, long N, long Len, long level_start,
                          void* stream);
/* Backward of cqvad_level_to_tokens (autograd of dab_transformer.py:316-321): grad_x [B, 256, N] channel-first (dtype; may be
 * NULL) = the level's rows of grad_tokens [B, Len, 256] (dtype) transposed back; grad_add [256] fp32 (may be NULL) ACCUMULATES
 * sum_{b,n} grad_tokens[b, level_start + n, :] = d level_embed[lvl] when called on the gradient of lvl_pos_embed_flatten. */
int cqvad_level_to_tokens_backward(int dtype, const void* grad_tokens, void* grad_x, float* grad_add, int B, long N, long Len,
                                   long level_start, void* stream);
/* Encoder output -> decoder memory: the part of Transformer.forward between the two
 * (models/detr/dab_transformer.py:349-393): per-level un-flatten, make_interpolated_features (:239-294, grid_sample with
 * align_corners = False and zeros padding onto the (num_frames, H, W) grid of level -2, including the reference's (meshy, meshx)
 * grid order in the T == num_frames branch), key-frame slice when `eff`, and the "L (H W) (B T) C" rearrange, in ONE pass that
 * computes only the consumed frames.  tokens / pos_tokens [B, Len, 256] (dtype; pos_tokens = lvl_pos_embed_flatten, may be
 * NULL with pos0); shapes [L,3] (T,H,W) / level_start [L] int64 on the device; (Tt, H, W) = shapes[L-2] passed by value;
 * memory [L, H*W, B*T', 256], pos0 [H*W, B*T', 256] (T' = 1 if eff else num_frames). */
int cqvad_encoder_to_decoder_memory(int dtype, const void* tokens, const void* pos_tokens, const int64_t* shapes,
                                    const int64_t* level_start, int L, int B, long Len, int Tt, int H, int W, int num_frames,
                                    int eff, void* memory, void* pos0, void* stream);
/* Backward of cqvad_encoder_to_decoder_memory (autograd of F.grid_sample + the slicing / rearrange around it): grad_memory fp32
 * [L, H*W, B*T', 256] (what cqvad_decoder_backward writes) is scattered with the forward's trilinear weights into grad_tokens
 * [B, Len, 256] (dtype), which is fully overwritten (tokens no consumed frame touches get zero).  workspace: fp32 [B*Len*256]
 * accumulator, required for bf16 (NULL allowed for fp32).  pos0 is a copy of pos_tokens rows of level -2; the decoder uses it only
 * on the key side of its softmax attentions, see INTEGRATION.md section 5 for why level_embed receives no gradient through it. */
int cqvad_encoder_to_decoder_memory_backward(int dtype, const float* grad_memory, const int64_t* shapes, const int64_t* level_start,
                                             int L, int B, long Len, int Tt, int H, int W, int num_frames, int eff, void* grad_tokens,
                                             float* workspace, void* stream);
/* Y[M,256] = LN?( res + W2 . act(W1 . X + b1) + b2 ): the FFN blocks of the decoder (dab_transformer.py:994-996,
 * 1043-1045, 1074-1076).  X [M,256], W1 [F,256], W2 [256,F] (dtype); ln_g/ln_b may be NULL (no LayerNorm); res may be
 * NULL.  hidden [M,F] (dtype) is scratch used only when the fused tensor-core kernel does not apply (fp32, or F % 128). */
int cqvad_mlp(int dtype, const void* X, const void* W1, const float* b1, const void* W2, const float* b2, int act,
              const void* res, const float* ln_g, const float* ln_b, float ln_eps, void* Y, void* hidden, long M, int F,
              void* stream);
/* y = x + W3.gelu(W2.LN(conv3x3(x)+b1)+b2)+b3 : ConvBlock.forward (dab_transformer.py:88-98) on NHWC input
 * x [Nimg, h, w, 256] (dtype) -> y same shape.  w1 is [256 out][9 taps (ky*3+kx)][256 in] (dtype).
 * workspace: cqvad_convblock_workspace_bytes(). */
size_t cqvad_convblock_workspace_bytes(int dtype, long n_img, int h, int w);
int cqvad_convblock_forward(int dtype, const void* x, void* y, const void* w1, const float* b1,
                            const float* ln_g, const float* ln_b, const void* w2, const float* b2,
                            const void* w3, const float* b3, long n_img, int h, int w,
                            void* workspace, size_t ws_bytes, void* stream);
/* Projection-free multi-head attention core of models/detr/attention.py:190-422 up to (not including) out_proj:
 *   mode 0 (standard, :336-341,377,409): q [L,Nb,E], k [S,Nb,E], v [S,Nb,Ev] -> o [L,Nb,Ev]
 *   mode 1 (query_specific_key, :343-346,379,411): q [L,Nb,E], k [L,S,Nb,E], v [L,S,Nb,Ev] -> o [L,Nb,Ev]
 * q is scaled by (E/H)^-0.5 (:291-293); key_padding_mask [Nb,S] uint8 (non-zero = ignore, :390-396) or NULL;
 * softmax after explicit max subtraction (:400-401); dropout is identity (eval). */
int cqvad_mha_core(int dtype, int mode, const void* q, const void* k, const void* v, const uint8_t* key_padding_mask,
                   void* o, int L, int S, int Nb, int H, int E, int Ev, void* stream);
/* PositionEmbeddingSine_3D.forward (models/position_encoding.py:32-73, normalize=True): mask [B,T,H,W] uint8
 * -> pos [B, num_pos_feats, T, H, W] fp32. */
int cqvad_posenc3d(const uint8_t* mask, float* pos, int B, int T, int H, int W, int num_pos_feats, void* stream);
/* gen_sineembed_for_position (dab_transformer.py:50-76): ref [rows,4] fp32 (x,y,w,h) -> [rows,512] fp32. */
int cqvad_sine_embed(const float* ref, float* out, long rows, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * The decoder: TransformerDecoder.forward (dab_transformer.py:722-852) + DETR heads (models/model.py:191-221),
 * eval semantics (dropout = identity).  One call enqueues the whole forward.
 */
typedef struct cqvad_decoder_desc {
  int dtype;        /* CQVAD_F32 | CQVAD_BF16                                                        */
  int BT;           /* B*T' frame-batch (dab_transformer.py:391-393)                                   */
  int nq;           /* actor queries (MODEL.QUERY_NUM)                                                 */
  int h, w;         /* orig_res, S = h*w keys                                                          */
  int K;            /* class queries (DATA.NUM_CLASSES)                                                */
  int F;            /* dim_feedforward                                                                 */
  int layers;       /* DEC_LAYERS                                                                      */
  int out_f32;      /* 1: hs/cls_hs written as fp32 (reference dtype); 0: written in `dtype`           */
  int flags;        /* CQVAD_DEC_* below                                                               */
  float dropout_p;  /* TRAINING calls only: p of the decoder's nn.Dropout modules (0 = identity, eval semantics)        */
  unsigned seed_lo, seed_hi;   /* Philox key of this step's masks; cqvad_decoder_backward must receive the forward's   */
} cqvad_decoder_desc;
enum { CQVAD_DEC_SKIP_CLS_HS = 1 /* do not materialise cls_hs (only pred_logits) */,
       CQVAD_DEC_FP32_CLS_STREAM = 2 /* bf16 path: keep fp32 side copies of the class-token residual/output stream */ };

/* Weight table: an array of device pointers ordered as cqvad_decoder_weight_name(i, layers) enumerates them
 * (reference state_dict names, SURVEY.md App. C, + "heads.class_embed_b.*").  kind 0 = matrix stored in `dtype`
 * ([out,in] row-major; conv1 as [out][ky*3+kx][in]); kind 1 = fp32 (biases, LayerNorm, small-N linears,
 * class_queries).  NULL is allowed only for layers.{i>0}.ca_qpos_proj.* (dab_transformer.py:711-713) and for the
 * synthesised entries whose name contains ".__" (layers.{i}.__ca_kv.* = [ca_kcontent_proj ; ca_v_proj] stacked to
 * [512,256] / [512]: when present the two projections of q_memory run as one GEMM). */
int cqvad_decoder_num_weights(int layers);
const char* cqvad_decoder_weight_name(int idx, int layers);
int cqvad_decoder_weight_kind(int idx, int layers);

size_t cqvad_decoder_workspace_bytes(const cqvad_decoder_desc* d);
/* tgt [nq,BT,256] fp32, memory/pos [4,S,BT,256] fp32, mask [BT,S] uint8, refpoints_unsigmoid [nq,BT,4] fp32.
 * Outputs: hs [layers,BT,nq,256], cls_hs [layers,BT,nq,K,256] (fp32 or dtype per out_f32; cls_hs may be NULL with
 * CQVAD_DEC_SKIP_CLS_HS), refs [layers,BT,nq,4] fp32; heads (any may be NULL): pred_logits [layers,BT,nq,K],
 * pred_boxes [layers,BT,nq,4], pred_logits_b [layers,BT,nq,3] fp32. */
int cqvad_decoder_forward(const cqvad_decoder_desc* d, const void* const* weights,
                          const float* tgt, const float* memory, const float* pos, const uint8_t* mask,
                          const float* refpoints_unsigmoid,
                          void* hs, void* cls_hs, float* refs,
                          float* pred_logits, float* pred_boxes, float* pred_logits_b,
                          void* workspace, size_t ws_bytes, void* stream);
/* Number of kernels the last cqvad_decoder_forward on this thread launched (bench.py's gpu_launches). */
long cqvad_last_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Training step of the decoder (BASELINE.json configs[1] "full decoder fwd+bwd"): what torch autograd does to
 * TransformerDecoder.forward in the reference training loop (train.py:126-182 -> loss.backward()).  Replaces, on the
 * decoder, the autograd graph of dab_transformer.py:722-852 (there is no hand-written backward in the reference except
 * the MSDA op, ops/functions/ms_deform_attn_func.py:36-45).
 *   cqvad_decoder_train_forward : same inputs/outputs as cqvad_decoder_forward without the heads (hs, cls_hs, refs are
 *       what TransformerDecoder.forward returns) and keeps every intermediate the backward needs in `workspace`
 *       (cqvad_decoder_train_workspace_bytes; 25.9 GiB at 32 AVA clips in bf16).  desc.dropout_p > 0 applies nn.Dropout at the
 *       nine residual-branch / FFN-hidden sites of every layer pair (Philox masks keyed by desc.seed_*; 0 = eval semantics).
 *   cqvad_decoder_backward : given dL/d(hs), dL/d(cls_hs) (element type of the forward outputs: fp32 when out_f32, else
 *       `dtype`) and dL/d(refs) (fp32); any may be NULL = zero.  MUST be called with the same descriptor, weights and
 *       workspace as the preceding train_forward, before anything else touches the workspace.
 *       Every gradient output is ACCUMULATED (+=) -- zero-fill for a plain backward, or keep across micro-batches:
 *         grad_weights[i]  fp32, shape of weight i of the weight table (conv1 as [out][ky*3+kx][in]; the stacked
 *                          __ca_kv entry receives d[ca_kcontent_proj ; ca_v_proj]); required for every non-NULL weight
 *         grad_memory      [4,S,BT,256] fp32      grad_tgt [nq,BT,256] fp32 (may be NULL)
 *         grad_refpoints_unsigmoid [nq,BT,4] fp32 (may be NULL)
 *       Autograd semantics of the reference: reference points detached between layers (:823), actor feature detached on
 *       entry to the class branch (:810); `pos` gets no gradient: it enters only on the key side of softmax attentions, so the
 *       per-level constant level_embed (Transformer) has an identically zero gradient through the decoder. */
/* Gradient-bucket signalling for data-parallel training (the reference wraps the model in DDP, utils/model_utils.py:113-121, whose
 * buckets all-reduce while the backward still runs): events[l] (cudaEvent_t, l = 0..n-1 = decoder layer; NULL entries skipped) is
 * recorded by every later cqvad_decoder_backward ON THIS THREAD as soon as all parameter gradients of layers.l.* / cls_layers.l.*
 * are final (behind the three streams of the backward), layers-1 first.  The gradients of the shared modules (ref_point_head,
 * query_scale, ref_anchor_head, bbox_embed, norm, cls_norm2, class_queries) are final when the call's stream is.  n = 0 disables.
 * The array must stay valid until it is replaced. */
int cqvad_decoder_backward_layer_events(void* const* events, int n);
size_t cqvad_decoder_train_workspace_bytes(const cqvad_decoder_desc* d);
int cqvad_decoder_train_forward(const cqvad_decoder_desc* d, const void* const* weights,
                                const float* tgt, const float* memory, const float* pos, const uint8_t* mask,
                                const float* refpoints_unsigmoid, void* hs, void* cls_hs, float* refs,
                                void* workspace, size_t ws_bytes, void* stream);
int cqvad_decoder_backward(const cqvad_decoder_desc* d, const void* const* weights, const uint8_t* mask,
                           const void* grad_hs, const void* grad_cls_hs, const float* grad_refs,
                           float* const* grad_weights, float* grad_memory, float* grad_tgt,
                           float* grad_refpoints_unsigmoid, void* workspace, size_t ws_bytes, void* stream);

/* Weight gradient of nn.Linear / the 3x3 conv of ConvBlock (what autograd's AddmmBackward / ConvolutionBackward produce
 * for dab_transformer.py:47,90): dW[n,k] += sum_m dY[m,n] X[m,k], db[n] += sum_m dY[m,n]   (fp32, ACCUMULATED; either
 * may be NULL).  conv_w > 0: X and dY are y-padded NHWC maps [n_img,(conv_h+1),conv_w,256] flattened to M rows (separator
 * rows zero) and dW is [N][ky*3+kx][K].  BF16 runs on tcgen05 (MN-major operands straight from the row-major activations)
 * when M >= 1024 and `workspace` (cqvad_wgrad_workspace_bytes) is given, on CUDA cores otherwise. */
size_t cqvad_wgrad_workspace_bytes(void);
int cqvad_linear_wgrad(int dtype, const void* dY, const void* X, float* dW, float* db, long M, int N, int K,
                       int conv_h, int conv_w, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss / matcher / post-process on the device.  Replaces, for the AVA configurations,
 *   HungarianMatcherAVA.forward   (models/detr/matcher.py:39-78: cost = cost_bbox * L1 + cost_giou * (-GIoU) + cost_class *
 *                                  (-softmax(pred_logits_b)[:, 1]); the reference copies the cost matrix to the host and calls
 *                                  scipy.optimize.linear_sum_assignment per clip),
 *   SetCriterionAVA.forward       (models/detr/criterion.py:50-138,184-224: loss_ce = sigmoid focal loss (alpha, gamma,
 *                                  models/detr/segmentation.py:200-229) on label-smoothed multi-hot targets with weight
 *                                  `pos_weight` on matched rows, / n_p; loss_ce_b = cross-entropy over {.., person, no-object}
 *                                  with weight eos_coef on the last class; loss_bbox = L1 / num_boxes; loss_giou),
 *   the weighted total of train.py:148 (sum over criterion.weight_dict: the LAST layer's four losses -- the reference computes
 *   the auxiliary per-layer losses for logging only, their keys are not in weight_dict) and its gradient with respect to the
 *   three prediction tensors, and PostProcessAVA.forward (criterion.py:740-773).
 *   pred_logits [B,nq,K], pred_boxes [B,nq,4] (cx,cy,w,h), pred_logits_b [B,nq,3]  fp32 -- one decoder layer's head outputs;
 *   tgt_boxes [B,maxT,4] (cx,cy,w,h: columns 1..4 of the reference's target["boxes"]), tgt_labels [B,maxT,K] multi-hot fp32,
 *   n_tgt [B] int32 (targets of clip b; rows beyond it are ignored).  nq <= 64, maxT <= 64.
 *   match [B,nq] int32 out: matched target index or -1 (the reference's `indices`);
 *   losses [16] fp32 out: [0] loss_ce [1] loss_bbox [2] loss_giou [3] loss_ce_b [4] weighted total [5] class_error
 *                         [6] matched pairs [7] num_boxes [8..10] normalisers (n_p, num_boxes, sum of CE weights);
 *   grad_* (may be NULL): d total / d prediction, fully overwritten.  No host synchronisation: the assignment is solved by
 *   the device (exact shortest-augmenting-path Hungarian in fp64 on the fp32 costs). */
typedef struct cqvad_criterion_cfg {
  float cost_class, cost_bbox, cost_giou;        /* MATCHER.COST_CLASS / COST_BBOX / COST_GIOU                         */
  float w_ce, w_bbox, w_giou, w_ce_b;            /* weight_dict: LOSS_COFS.DICE_COF / BBOX_COF / GIOU_COF / PERSON_COF */
  float pos_weight;                              /* LOSS_COFS.WEIGHT (criterion.py:84)                                 */
  float eos_coef;                                /* LOSS_COFS.EOS_COF                                                  */
  float focal_alpha, focal_gamma;                /* 0.25, 2 (criterion.py:46-47)                                       */
  float label_smoothing;                         /* MODEL.LABEL_SMOOTHING_ALPHA (criterion.py:48: 0.1)                 */
} cqvad_criterion_cfg;
size_t cqvad_criterion_ava_workspace_bytes(int B);
int cqvad_criterion_ava(const cqvad_criterion_cfg* cfg, const float* pred_logits, const float* pred_boxes,
                        const float* pred_logits_b, const float* tgt_boxes, const float* tgt_labels, const int32_t* n_tgt,
                        int B, int nq, int K, int maxT, int32_t* match, float* losses, float* grad_logits, float* grad_boxes,
                        float* grad_logits_b, void* workspace, size_t ws_bytes, void* stream);
/* detections [B, nq, K + 5] = [sigmoid(pred_logits) | box as (x0,y0,x1,y1) * (w,h,w,h) | softmax(pred_logits_b)[1]];
 * target_sizes [B,2] fp32 (h, w) as in the reference. */
int cqvad_postprocess_ava(const float* pred_logits, const float* pred_boxes, const float* pred_logits_b,
                          const float* target_sizes, float* detections, int B, int nq, int K, void* stream);
/* PostProcessUCF / PostProcessJHMDB (criterion.py:775-846; BASELINE configs[3]): same layout, scores = sigmoid(inverse_sigmoid(
 * sigmoid(pred_logits) * softmax(pred_logits_b)[1])) (utils/misc.py:530-534).  B = number of frames (clips x T'). */
int cqvad_postprocess_ucf(const float* pred_logits, const float* pred_boxes, const float* pred_logits_b,
                          const float* target_sizes, float* detections, int B, int nq, int K, void* stream);

/* Dropout pass of the native training path: out = (res ? res : 0) + keep * x / (1 - p), keep ~ Bernoulli(1 - p) from a
 * Philox4x32-10 counter keyed by (seed, site, element index / 8) -- nothing is stored; calling it on a gradient with the same
 * (seed, site) applies the forward's mask (nn.Dropout backward).  n % 8 == 0; in place when out == x; p is quantised to 1/65536. */
int cqvad_dropout(int dtype, const void* x, const void* res, void* out, long n, float p, uint64_t seed, uint32_t site, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * DETR heads of the reference model in the TRAINING step (models/model.py:191-236), fp32 like the reference (autocast off):
 *   pred_logits_b = class_embed_b(hs);  pred_boxes = sigmoid(bbox_embed(hs) + inverse_sigmoid(reference))  (utils/misc.py:530-534);
 *   pred_logits = Dropout(p_drop)(cls_hs).mean(-1)  (models/model.py:103,219: p = 0.5 in training, 0 in eval).
 * Rows R = Lr*BT*nq in the order of the decoder outputs: hs [R,256], refs [R,4], cls_hs [R,K,256] (all fp32).
 * weights[8] = bbox_embed.layers.{0,1,2}.{weight,bias}, class_embed_b.{weight,bias} (fp32, PyTorch [out,in] layout).
 * The dropout mask is a Philox4x32-10 stream keyed by `seed` (counter = element index): the backward regenerates it, nothing is
 * stored.  The forward keeps the MLP hidden activations in `workspace`, which the backward of the SAME rows must receive.
 * backward: grad_hs [R,256], grad_refs [R,4] (may be NULL), grad_cls_hs [R,K,256] (may be NULL) are overwritten;
 * grad_weights[8] (may be NULL) are ACCUMULATED (bbox_embed is shared with the decoder's box refinement, models/model.py:100-101). */
size_t cqvad_heads_train_workspace_bytes(long R);
int cqvad_heads_train_forward(const float* const* weights, const float* hs, const float* cls_hs, const float* refs, long R, int K,
                              float p_drop, uint64_t seed, float* pred_logits, float* pred_boxes, float* pred_logits_b,
                              void* workspace, size_t ws_bytes, void* stream);
int cqvad_heads_train_backward(const float* const* weights, const float* hs, const float* refs, const float* grad_logits,
                               const float* grad_boxes, const float* grad_logits_b, long R, int K, float p_drop, uint64_t seed,
                               float* grad_hs, float* grad_cls_hs, float* grad_refs, float* const* grad_weights, void* workspace,
                               size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer step of the reference training loop (train.py:83,158-167) over ONE flat fp32 buffer of n parameters:
 *   torch.nn.utils.clip_grad_norm_(parameters, max_norm) -- global L2 norm of all gradients, coef = min(1, max_norm/(norm+1e-6)),
 *   torch.optim.AdamW(lr, (beta1, beta2), eps, weight_decay).step() -- decoupled decay, bias-corrected moments.
 * grads are first multiplied by grad_scale (1 / world size after a SUM all-reduce; 1 otherwise): the norm is that of the scaled
 * gradient.  step = 1-based count of this update.  max_norm <= 0 disables clipping.  params_bf16 (may be NULL): a bf16 copy of
 * the updated parameters written in the same pass (tensor-core operand).  zero_grad != 0 zero-fills grads afterwards
 * (optimizer.zero_grad()).  grad_norm_out (may be NULL): the pre-clip norm, device scalar.  Two launches, no host sync. */
size_t cqvad_adamw_workspace_bytes(void);
int cqvad_adamw_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, long n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, long step, float max_norm, float grad_scale,
                          int zero_grad, float* grad_norm_out, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py).  cqvad_profile_enable(1) makes cqvad_decoder_forward bracket each kernel class with
 * CUDA events on the launch stream; cqvad_profile_read() synchronises and returns the accumulated milliseconds, the
 * number of timed scopes and of kernel launches inside them since the enable call.  Replaces the reference's host
 * wall-clock meters (utils/video_action_recognition.py:31-32,64-65,176-186). */
void cqvad_profile_enable(int on);
int cqvad_profile_num_classes(void);
const char* cqvad_profile_class_name(int cls);
int cqvad_profile_read(int cls, double* total_ms_host, long* scopes_host, long* launches_host);
/* Algorithmic work of the kernels timed in a class since the enable call: FLOPs (2 per multiply-add) and bytes (every operand
 * read once, the result written once) -- the numerators of bench.py's per-class tensor-pipe / HBM roofline fractions. */
int cqvad_profile_read_work(int cls, double* flops_host, double* bytes_host);
/* Timeline of the timed scopes since cqvad_profile_enable(1), in host issue order: class, stream id (0, 1, 2 ... by first
 * appearance), start / end in ms relative to the earliest start.  Returns the number of scopes; at most `max` are written. */
long cqvad_profile_timeline(int* cls, int* stream_id, double* start_ms, double* end_ms, long max);
/* Debug switch: route bf16 GEMMs through the CUDA-core kernel instead of tcgen05 (used by the parity tests to
 * cross-check the two implementations; not a fallback -- both are CUDA kernels of this library). */
void cqvad_debug_force_simt(int on);

#ifdef __cplusplus
}
#endif
#endif /* CQVAD_H_ */
