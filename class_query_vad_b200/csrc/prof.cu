#include <vector>
#include "common.cuh"
#include "prof.cuh"

namespace cqvad {
namespace {
struct Pair { cudaEvent_t a, b; };
struct ClassState { std::vector<Pair> pool; size_t used = 0; long launches_at_begin = 0; long launches = 0; double flops = 0, bytes = 0; };
ClassState g_cls[P_COUNT];
bool g_on = false;
struct Scope { int cls; size_t idx; cudaStream_t st; };
std::vector<Scope> g_order;       // every scope since the enable call, in host issue order (timeline dump)
const char* kNames[P_COUNT] = {"conv3x3_ln (tcgen05 implicit GEMM)", "convblock_mlp (tcgen05 fused MLP)",
                               "class_ffn (tcgen05 fused MLP + LN)", "kv/k/v projections (tcgen05 GEMM)",
                               "class cross-attention", "class self-attention", "class out_proj GEMMs",
                               "loc query-specific-key attention", "level mix + LN", "actor add + conv_norm",
                               "output LN / heads", "small-row ops (prologue, loc SA, FFNs, box head)", "input conversion",
                               "train fwd: GEMM / conv", "train fwd: LN, attention, elementwise", "train bwd: dgrad GEMM / conv",
                               "train bwd: wgrad", "train bwd: activation / residual", "train bwd: LayerNorm", "train bwd: attention",
                               "train bwd: misc", "train fwd: conv3x3 (tcgen05 implicit GEMM)", "train bwd: conv3x3 dgrad (tcgen05 implicit GEMM)",
                               "train bwd: conv3x3 wgrad (tcgen05 MN-major)",
                               "train fwd: GELU GEMM (conv2, gelu + gelu' epilogue)", "train bwd: dgrad GEMM x act' (epilogue multiply)",
                               "train fwd: small-row GEMMs (< 8192 rows)", "train bwd: small-row dgrad GEMMs", "train bwd: small-row wgrad"};
}  // namespace
long launch_count_now();
bool prof_enabled() { return g_on; }
void prof_begin(int cls, cudaStream_t st) {
  ClassState& c = g_cls[cls];
  if (c.used == c.pool.size()) {
    Pair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    c.pool.push_back(p);
  }
  c.launches_at_begin = launch_count_now();
  g_order.push_back(Scope{cls, c.used, st});
  cudaEventRecord(c.pool[c.used].a, st);
}
void prof_work(int cls, double flops, double bytes) { g_cls[cls].flops += flops; g_cls[cls].bytes += bytes; }
void prof_end(int cls, cudaStream_t st) {
  ClassState& c = g_cls[cls];
  cudaEventRecord(c.pool[c.used].b, st);
  c.launches += launch_count_now() - c.launches_at_begin;
  c.used++;
}
}  // namespace cqvad

using namespace cqvad;
extern "C" void cqvad_profile_enable(int on) {
  g_on = on != 0;
  for (int i = 0; i < P_COUNT; ++i) { g_cls[i].used = 0; g_cls[i].launches = 0; g_cls[i].flops = 0; g_cls[i].bytes = 0; }
  g_order.clear();
}
// Timeline of the scopes since cqvad_profile_enable(1): for scope i (host issue order) its class, an id of the stream it ran on
// (0, 1, 2, ... in order of first appearance) and its start / end in ms relative to the earliest start.  Synchronises.  Returns the
// number of scopes (at most `max` are written).  Tool: tools/timeline_train.py.
extern "C" long cqvad_profile_timeline(int* cls, int* stream_id, double* start_ms, double* end_ms, long max) {
  std::vector<cudaStream_t> streams;
  const long n = (long)g_order.size();
  if (n == 0) return 0;
  const Scope& s0 = g_order[0];
  cudaEvent_t base = g_cls[s0.cls].pool[s0.idx].a;
  double mn = 0;
  for (long i = 0; i < n && i < max; ++i) {
    const Scope& s = g_order[i];
    const Pair& pr = g_cls[s.cls].pool[s.idx];
    cudaEventSynchronize(pr.b);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, base, pr.a);
    cudaEventElapsedTime(&b, base, pr.b);
    int sid = -1;
    for (size_t k = 0; k < streams.size(); ++k) if (streams[k] == s.st) sid = (int)k;
    if (sid < 0) { sid = (int)streams.size(); streams.push_back(s.st); }
    cls[i] = s.cls; stream_id[i] = sid; start_ms[i] = a; end_ms[i] = b;
    if (a < mn) mn = a;
  }
  for (long i = 0; i < n && i < max; ++i) { start_ms[i] -= mn; end_ms[i] -= mn; }
  return n;
}
extern "C" int cqvad_profile_num_classes(void) { return P_COUNT; }
extern "C" const char* cqvad_profile_class_name(int cls) { return (cls >= 0 && cls < P_COUNT) ? kNames[cls] : nullptr; }
// total elapsed ms, number of timed scopes and kernel launches inside them since cqvad_profile_enable(1); synchronises.
extern "C" int cqvad_profile_read(int cls, double* total_ms, long* scopes, long* launches) {
  if (cls < 0 || cls >= P_COUNT) return CQVAD_E_INVALID_ARG;
  ClassState& c = g_cls[cls];
  double tot = 0;
  for (size_t i = 0; i < c.used; ++i) {
    cudaEventSynchronize(c.pool[i].b);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c.pool[i].a, c.pool[i].b) == cudaSuccess) tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (scopes) *scopes = (long)c.used;
  if (launches) *launches = c.launches;
  return 0;
}

// algorithmic FLOPs and bytes (operands read once + result written once) of the kernels timed in a class since the enable call
extern "C" int cqvad_profile_read_work(int cls, double* flops, double* bytes) {
  if (cls < 0 || cls >= P_COUNT) return CQVAD_E_INVALID_ARG;
  if (flops) *flops = g_cls[cls].flops;
  if (bytes) *bytes = g_cls[cls].bytes;
  return 0;
}
