// Declarations of the attention host wrappers (attention.cu).
#pragma once
#include "common.cuh"
namespace cqvad {
struct StdStrides {
  long q_ls, q_bs, k_ls, k_bs, v_ls, v_bs, o_ls, o_bs;   // element strides: row (l or s) and batch
  long q2_ls, q2_bs, k2_ls, k2_bs;                       // second source for heads >= H/2 (class cross-attention)
  int k2_bmod;                                           // > 0: the second key source is indexed by batch % k2_bmod
};
// heads h < H/2 read q/k, heads >= H/2 read q2/k2 (when q2 != nullptr); v/o use all H heads.
template <typename T>
int mha_std(const T* q, const T* q2, const T* k, const T* k2, const T* v, const uint8_t* kpm, T* o, int L, int S, int Nb,
            int H, int hd, int vd, const StdStrides& st, cudaStream_t stm);
template <typename T>
int dec_qsk_attn(const T* qc, const T* qs, const T* kc, const T* v, long ldkv, const T* kp, const uint8_t* mask, T* o,
                 long N, int S, int Sq, int BT, bool first, cudaStream_t stm);
}  // namespace cqvad
