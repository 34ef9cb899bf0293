// Weight-table layout of the decoder (reference state_dict names) and the workspace arena; shared by the
// inference driver (decoder.cu) and the training driver (train.cu).
#pragma once
#include <string>
#include "common.cuh"

namespace cqvad {

// ---- weight table --------------------------------------------------------------------------------------------
// kind: 0 = matrix stored in the activation dtype, 1 = fp32
struct Slot { const char* name; int kind; };
#define LIN(n) {n ".weight", 0}, {n ".bias", 1}
#define LNP(n) {n ".weight", 1}, {n ".bias", 1}
static const Slot kLocSlots[] = {
    LIN("sa_qcontent_proj"), LIN("sa_qpos_proj"), LIN("sa_kcontent_proj"), LIN("sa_kpos_proj"), LIN("sa_v_proj"),
    LIN("self_attn.out_proj"), LNP("norm1"), {"lvl_w_embed.weight", 1}, {"lvl_w_embed.bias", 1},
    LIN("ca_qcontent_proj"), LIN("ca_qpos_proj"), LIN("ca_kcontent_proj"), LIN("ca_kpos_proj"), LIN("ca_v_proj"),
    LIN("ca_qpos_sine_proj"), LIN("cross_attn.out_proj"), LIN("linear1"), LIN("linear2"), LNP("norm2"), LNP("norm3"),
    LNP("norm_"),
    // synthesised at pack time (not a reference parameter): [ca_kcontent_proj ; ca_v_proj] stacked to [512,256] so that the
    // two projections of q_memory are ONE GEMM (q_memory is read once)
    LIN("__ca_kv")};
enum LocIdx { SA_QC = 0, SA_QP = 2, SA_KC = 4, SA_KP = 6, SA_V = 8, SA_O = 10, NORM1 = 12, LVLW = 14, CA_QC = 16,
              CA_QP = 18, CA_KC = 20, CA_KP = 22, CA_V = 24, CA_QS = 26, CA_O = 28, LIN1 = 30, LIN2 = 32, NORM2 = 34,
              NORM3 = 36, NORMU = 38, CA_KV = 40, LOC_COUNT = 42 };
static const Slot kClsSlots[] = {
    LIN("cls_linear1"), LIN("cls_linear2"), LNP("cls_norm"), LNP("conv_norm"), LIN("conv_blocks.0.conv1"),
    LNP("conv_blocks.0.norm"), LIN("conv_blocks.0.conv2"), LIN("conv_blocks.0.conv3"), LIN("self_attn.out_proj"),
    LNP("norm1"), LIN("k_proj"), LIN("v_proj"), LIN("cls_qpos_sine_proj"), LIN("cross_attn.out_proj"),
    LIN("cls_linear1_"), LIN("cls_linear2_"), LNP("cls_norm_")};
enum ClsIdx { C_L1 = 0, C_L2 = 2, C_NORM = 4, C_CONVNORM = 6, C_CONV1 = 8, C_CBNORM = 10, C_CONV2 = 12, C_CONV3 = 14,
              C_SA_O = 16, C_NORM1 = 18, C_KPROJ = 20, C_VPROJ = 22, C_QPS = 24, C_CA_O = 26, C_L1_ = 28, C_L2_ = 30,
              C_NORM_ = 32, CLS_COUNT = 34 };
static const Slot kGlobSlots[] = {
    LNP("norm"), LNP("cls_norm2"), LIN("query_scale.layers.0"), LIN("query_scale.layers.1"),
    LIN("ref_point_head.layers.0"), LIN("ref_point_head.layers.1"), LIN("ref_anchor_head.layers.0"),
    {"ref_anchor_head.layers.1.weight", 1}, {"ref_anchor_head.layers.1.bias", 1}, {"class_queries.weight", 0},
    LIN("bbox_embed.layers.0"), LIN("bbox_embed.layers.1"), {"bbox_embed.layers.2.weight", 1},
    {"bbox_embed.layers.2.bias", 1}, {"heads.class_embed_b.weight", 1}, {"heads.class_embed_b.bias", 1}};
enum GlobIdx { G_NORM = 0, G_CLSNORM2 = 2, G_QS0 = 4, G_QS1 = 6, G_RPH0 = 8, G_RPH1 = 10, G_RAH0 = 12, G_RAH1 = 14,
               G_CQ = 16, G_BB0 = 17, G_BB1 = 19, G_BB2 = 21, G_CEB = 23, GLOB_COUNT = 25 };
static_assert(sizeof(kLocSlots) / sizeof(Slot) == LOC_COUNT, "loc slots");
static_assert(sizeof(kClsSlots) / sizeof(Slot) == CLS_COUNT, "cls slots");
static_assert(sizeof(kGlobSlots) / sizeof(Slot) == GLOB_COUNT, "glob slots");

static thread_local std::string g_name;
static bool weight_slot(int idx, int layers, std::string* name, int* kind) {
  if (idx < 0) return false;
  if (idx < layers * LOC_COUNT) {
    const Slot& s = kLocSlots[idx % LOC_COUNT];
    if (name) *name = "layers." + std::to_string(idx / LOC_COUNT) + "." + s.name;
    if (kind) *kind = s.kind;
    return true;
  }
  idx -= layers * LOC_COUNT;
  if (idx < layers * CLS_COUNT) {
    const Slot& s = kClsSlots[idx % CLS_COUNT];
    if (name) *name = "cls_layers." + std::to_string(idx / CLS_COUNT) + "." + s.name;
    if (kind) *kind = s.kind;
    return true;
  }
  idx -= layers * CLS_COUNT;
  if (idx < GLOB_COUNT) {
    if (name) *name = kGlobSlots[idx].name;
    if (kind) *kind = kGlobSlots[idx].kind;
    return true;
  }
  return false;
}

// ---- workspace arena ---------------------------------------------------------------------------------------------
struct Arena {
  char* base; size_t cap; size_t off = 0; bool overflow = false;
  Arena(void* b, size_t c) : base((char*)b), cap(c) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (base && off > cap) overflow = true;
    return p;
  }
};

}  // namespace cqvad
