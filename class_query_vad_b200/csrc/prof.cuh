// Optional per-kernel-class timing with CUDA events recorded on the launch stream (bench.py's roofline object and the
// per-step breakdown in profiles/).  Disabled by default: zero cost beyond one branch per scope.
#pragma once
#include <cuda_runtime.h>
namespace cqvad {
enum ProfClass { P_CONV = 0, P_CONV_MLP, P_CLS_FFN, P_BIG_PROJ, P_CLS_XATTN, P_CLS_SATTN, P_CLS_OPROJ, P_LOC_QSK, P_LVLMIX,
                 P_ADDLN, P_OUT_LN, P_SMALL, P_INPUT,
                 P_T_FWD_GEMM, P_T_FWD_OTHER, P_T_DGRAD, P_T_WGRAD, P_T_ACT_BWD, P_T_LN_BWD, P_T_ATTN_BWD, P_T_MISC_BWD,
                 P_T_CONV_FWD, P_T_CONV_DGRAD, P_T_CONV_WGRAD,
                 P_T_FWD_GEMM_GELU, P_T_DGRAD_ACT, P_T_FWD_GEMM_SMALL, P_T_DGRAD_SMALL, P_T_WGRAD_SMALL, P_COUNT };
bool prof_enabled();
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
// algorithmic work of the kernels inside the scopes of a class (bench.py: per-class tensor-pipe / HBM roofline fractions)
void prof_work(int cls, double flops, double bytes);
struct ProfScope {
  int cls; cudaStream_t st; bool on;
  ProfScope(int c, cudaStream_t s) : cls(c), st(s), on(prof_enabled()) { if (on) prof_begin(cls, st); }
  ProfScope(int c, cudaStream_t s, double flops, double bytes) : cls(c), st(s), on(prof_enabled()) {
    if (on) { prof_begin(cls, st); prof_work(cls, flops, bytes); }
  }
  ~ProfScope() { if (on) prof_end(cls, st); }
};
}  // namespace cqvad
