// Declarations of the backward-pass kernels' host wrappers (bwd_kernels.cu, bwd_attn.cu, wgrad_tc.cu).
// Gradient-accumulation convention: an output gradient written with `beta` = 0 is overwritten, with beta = 1 the
// contribution is added to what is already there (several consumers of one activation).  Parameter gradients are fp32
// and always accumulated (the driver zero-fills them once per backward).
#pragma once
#include "common.cuh"
#include "attention.cuh"
namespace cqvad {

// Wt[in][out] = W[out][in]
template <typename T> int transpose_w(const T* W, T* Wt, int out, int in, cudaStream_t st);
// the same for up to kMaxTransposeJobs matrices in ONE launch (the job table travels as a kernel parameter)
constexpr int kMaxTransposeJobs = 96;
struct TransposeJob { const void* W; void* Wt; int out, in; };
template <typename T> int transpose_w_batch(const TransposeJob* jobs, int n, cudaStream_t st);
// conv dgrad weights: Wd[ci][8-tap][co] = W[co][tap][ci]  (both [256][9][256])
template <typename T> int conv_w_flip(const T* W, T* Wd, cudaStream_t st);

// dW[n*ldw + k] += sum_m dY[m,n] * X[m(+tap shift),k] ; db[n] += sum_m dY[m,n]   (db may be NULL; dW may be NULL)
// conv != nullptr: X is the y-padded NHWC activation and dW is [Nout][9][256] (ldw = 9*256).
template <typename T>
int wgrad(const T* dY, long lddy, const T* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
          const ConvGeom* conv, cudaStream_t st);
template <typename T>
int wgrad_simt(const T* dY, long lddy, const T* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
               const ConvGeom* conv, cudaStream_t st);
// tcgen05 weight-gradient kernel (bf16); returns 1 when the shape is not supported
int wgrad_tc(const bf16* dY, long lddy, const bf16* X, long ldx, float* dW, long ldw, float* db, long M, int Nout, int Kin,
             const ConvGeom* conv, cudaStream_t st);

// scratch for the split-K partial tiles of wgrad_tc (owned by the training workspace)
void set_wgrad_scratch(float* p, size_t bytes);
size_t wgrad_scratch_bytes();

// dH *= act'(.)   ReLU: ref = post-activation H (mask H > 0);  GELU: ref = pre-activation (exact erf derivative)
template <typename T> int act_bwd(T* dH, const T* ref, int act, long n, cudaStream_t st);
template <typename T> int gelu_fwd(const T* pre, T* out, T* dact, long n, cudaStream_t st);   // dact (optional) = gelu'(pre)
// dst = beta*dst + src
template <typename T> int axpby(T* dst, const T* src, float beta, long n, cudaStream_t st);
// fp32 accumulate of a T tensor: dst32 += src
template <typename T> int acc_to_f32(float* dst32, const T* src, long n, cudaStream_t st);
template <typename T> int f32_to_t(const float* src, T* dst, float beta, long n, cudaStream_t st);

// y = LN(x (+res)) backward.  dy: T rows in internal order, or (perm_nq > 0) the caller's output layout
// [b][n][k] with element type float (dy_f32) or T.  dx = beta_x*dx + dz ; dres (optional) = beta_r*dres + dz.
template <typename T>
int ln_bwd(const T* x, const T* res, const float* g, float eps, const void* dy, bool dy_f32, int perm_nq, int perm_BT,
           int perm_K, T* dx, float beta_x, T* dres, float beta_r, float* dg, float* db, long rows, cudaStream_t st);

// q_memory = norm_(sum_l lvlw*mem_l) backward: dmem32 [4,S,BT,256] fp32 += ; dlvlw [N,4] fp32 += (pre-zeroed)
template <typename T>
int lvlmix_ln_bwd(const T* mem, const float* lvlw, const float* g, const T* dqm, float* dmem32, float* dlvlw, float* dg,
                  float* db, long N, int nq, int S, int Sq, int BT, cudaStream_t st);
// lvl_w = softmax(lvl_w_embed(x)): given dlvlw (w.r.t. the probabilities) -> dx (beta), dW [4,256] +=, dB [4] +=
template <typename T>
int lvlw_bwd(const T* x, const float* w, const float* p, const float* dp, T* dx, float beta, float* dW, float* dB, long rows,
             cudaStream_t st);
// XA = conv_norm(acls[i] + qm[i,s]) backward: dqm (beta), dacls (beta), dg/db +=
template <typename T>
int add_ln_pad_bwd(const T* actor, const T* qm, const float* g, const T* dxpad, T* dqm, float beta_qm, T* dactor,
                   float beta_a, float* dg, float* db, long N, int S, int Sq, int Sp, cudaStream_t st);
// query_sine_embed backward (dab_transformer.py:757-763): dscale (beta; may be NULL), dhidden (beta), dw1/db1 +=,
// dref [rows,4] fp32 += (may be NULL: reference points are detached for layers >= 1)
template <typename T>
int qse_bwd(const float* ref, const T* scale, const T* hidden, const float* w1, const float* b1, const T* dqse, T* dscale,
            float beta_s, T* dhidden, float beta_h, float* dw1, float* db1, float* dref, long rows, cudaStream_t st);
// gen_sineembed_for_position backward: dref [rows,4] += from de512 [rows,512]
template <typename T> int sine_embed_bwd(const float* ref, const T* de, float* dref, long rows, cudaStream_t st);
// box refinement backward: r_new = sigmoid(hidden.w2 + b2 + inverse_sigmoid(ref)); dnew_perm [b][n][4] fp32 (may be NULL)
// dhidden (beta), dw2/db2 +=, dref += (may be NULL)
template <typename T>
int box_refine_bwd(const T* hidden, const float* w2, const float* b2, const float* ref, const float* dnew_perm, T* dhidden,
                   float beta, float* dw2, float* db2, float* dref, long rows, int nq, int BT, cudaStream_t st);
// refpoints: dref_u[row] += (dref[row] + drefs0_perm[b][n]) * r (1-r)
int sigmoid4_bwd(const float* r, const float* dr, const float* dperm, float* dref_u, long rows, int nq, int BT, cudaStream_t st);
// dsrc[k] = beta*dsrc[k] + sum_i dout[i*K+k]
template <typename T> int broadcast_rows_bwd(const T* dout, T* dsrc, float beta, long rows, int K, cudaStream_t st);
// zero the separator rows of a y-padded [n_img, Sp, 256] tensor
template <typename T> int zero_pad_rows(T* x, long n_img, int S, int Sp, cudaStream_t st);

// attention backward (recomputes the probabilities).  Gradients have the layouts / strides of their tensors.
// k2 (positional keys of the class cross-attention) gets no gradient.  q2 with q2_ls == 0 is summed over the L queries.
template <typename T>
int mha_std_bwd(const T* q, const T* q2, const T* k, const T* k2, const T* v, const uint8_t* kpm, const T* dO, T* dq,
                float beta_q, T* dq2, float beta_q2, T* dk, float beta_k, T* dv, float beta_v, int L, int S, int Nb, int H,
                int hd, int vd, const StdStrides& st, cudaStream_t stm);
// localisation cross-attention backward.  dkv [N*Sq, ldkv] (k | v halves) overwritten for s < S; dkp32 [S*BT,256] fp32 +=
template <typename T>
int dec_qsk_bwd(const T* qc, const T* qs, const T* kc, const T* v, long ldkv, const T* kp, const uint8_t* mask, const T* dO,
                T* dqc, float beta_qc, T* dqs, float beta_qs, T* dkc, T* dv, float* dkp32, long N, int S, int Sq, int BT,
                bool first, cudaStream_t stm);
}  // namespace cqvad
