// Declarations of the HBM-bound kernels' host wrappers (kernels_mem.cu).
#pragma once
#include "common.cuh"
namespace cqvad {
template <typename T> int layernorm_permute(const T* x, const float* g, const float* b, float eps, void* out, bool out_f32,
                                            long rows, int nq, int BT, int K, float* row_mean_out, cudaStream_t st);
int layernorm_permute_f32in(const float* x, const float* g, const float* b, float eps, void* out, bool out_f32, long rows,
                            int nq, int BT, int K, float* row_mean_out, cudaStream_t st);
template <typename T> int lvlmix_ln(const T* mem, const float* lvlw, const float* g, const float* b, T* qm, long N, int S,
                                    int Sq, int BT, cudaStream_t st);
template <typename T> int add_ln_pad(const T* actor, const T* qm, const float* g, const float* b, T* xpad, long N, int S,
                                     int Sq, int Sp, cudaStream_t st);
template <typename T> int pad_copy(const T* src, T* dst, long n_img, int S, int Sp, bool to_padded, cudaStream_t st);
template <typename T> int convert_f32(const float* in, T* out, long n, cudaStream_t st);
template <typename O> int sine_embed(const float* ref, O* out, long rows, cudaStream_t st);
template <typename T> int qse_modulate(const float* ref, const T* scale, const T* hidden, const float* w1, const float* b1,
                                       T* qse, long rows, cudaStream_t st);
template <typename T> int linear_smalln(const T* x, const float* w, const float* b, float* out, long rows, int n,
                                        bool softmax, cudaStream_t st);
template <typename T> int box_refine(const T* hidden, const float* w2, const float* b2, const float* ref, float* ref_new,
                                     float* out_perm, long rows, int nq, int BT, bool ref_is_perm, cudaStream_t st);
int sigmoid4(const float* in, float* out, float* out_perm, long rows, int nq, int BT, cudaStream_t st);
template <typename T> int broadcast_rows(const T* src, T* out, long rows, int K, cudaStream_t st);
int posenc3d(const uint8_t* mask, float* pos, int B, int T, int H, int W, int npf, cudaStream_t st);
template <typename T> int logits_b(const T* x, const float* w, const float* b, float* out, long rows, int nq, int BT,
                                   cudaStream_t st);
void set_force_simt(bool v);
bool force_simt();
}  // namespace cqvad
