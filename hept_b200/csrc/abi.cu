// Error plumbing, launch accounting and the one-call forward (a3..a12) of the C ABI.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace hept {

static thread_local char g_error[512] = "";
static std::atomic<int> g_launches{0};   // process-wide: autograd runs the backward on its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// process-wide selectors (profiling / test aids); atomics: autograd calls the backward from its own thread
static std::atomic<int> g_bwd_mask{7};
int bwd_stage_mask() { return g_bwd_mask.load(std::memory_order_relaxed); }
static std::atomic<int> g_engine{1};  // tcgen05 tiles by default (faster, same tolerance contract); 0 = fp32 CUDA-core tiles
int engine() { return g_engine.load(std::memory_order_relaxed); }
static std::atomic<int> g_bwd_variant{3};  // tcgen05 backward tiles by default; 1 = fp32 CUDA-core tiles
int bwd_variant() { return g_bwd_variant.load(std::memory_order_relaxed); }

struct FwdPlan {
  size_t ext_bytes, span_bytes, hat_bytes, proj_bytes, keys_bytes, sort_bytes, stage_bytes, total;
};
static FwdPlan plan_fwd(const hept_shape* s) {
  FwdPlan p;
  const size_t th = (size_t)s->T * s->H, thn = th * s->N;
  p.ext_bytes = align_up(hept_hash_workspace_bytes(s), 256);
  p.span_bytes = align_up(sizeof(float) * th, 256);
  p.hat_bytes = align_up(sizeof(float) * (size_t)s->N * s->H * 8, 256);
  p.proj_bytes = align_up(sizeof(float) * 2 * thn, 256);
  p.keys_bytes = align_up(sizeof(float) * 2 * thn, 256);
  p.sort_bytes = align_up(hept_argsort_workspace_bytes((int32_t)(2 * th), s->N), 256);
  p.stage_bytes = align_up(sizeof(float) * (size_t)s->H * s->N * s->T * kStageRow, 256);
  // proj/keys/sort scratch is dead once the permutations exist; the staging rows reuse that space.
  size_t front = p.proj_bytes + p.keys_bytes + p.sort_bytes;
  p.total = p.ext_bytes + p.span_bytes + p.hat_bytes + (front > p.stage_bytes ? front : p.stage_bytes);
  return p;
}

}  // namespace hept

using namespace hept;

extern "C" int hept_abi_version(void) { return 1; }
extern "C" const char* hept_last_error(void) { return g_error; }
extern "C" int hept_launch_count(int reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

extern "C" void hept_set_bwd_stage_mask(int mask) { g_bwd_mask.store(mask & 7, std::memory_order_relaxed); }
extern "C" void hept_set_engine(int engine) { g_engine.store(engine ? 1 : 0, std::memory_order_relaxed); }
extern "C" int hept_get_engine(void) { return engine(); }
extern "C" void hept_set_bwd_variant(int variant) {
  g_bwd_variant.store((variant >= 3 && variant <= 5) ? variant : 1, std::memory_order_relaxed);
}
extern "C" int hept_get_bwd_variant(void) { return bwd_variant(); }

extern "C" size_t hept_attention_fwd_workspace_bytes(const hept_shape* s) {
  if (!s || s->N <= 0 || s->H <= 0 || s->T <= 0) return 0;
  return plan_fwd(s).total;
}

static int attention_fwd_impl(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                              const float* w_rpe_weight, int32_t K, const float* alpha, const int64_t* combined_shifts,
                              const int32_t* combined_shifts32, const float* region_eta, const float* region_phi,
                              const float* regions_h, float* scale, int32_t* positions, float* out_pre, float* den_sum,
                              void* workspace, size_t workspace_bytes, void* stream);

extern "C" int hept_attention_fwd(const hept_shape* s, const float* q, const float* k, const float* v,
                                  const float* coords, const float* w_rpe_weight, int32_t K, const float* alpha,
                                  const int64_t* combined_shifts, const float* region_eta, const float* region_phi,
                                  const float* regions_h, float* scale, int32_t* positions, float* out_pre,
                                  float* den_sum, void* workspace, size_t workspace_bytes, void* stream) {
  return attention_fwd_impl(s, q, k, v, coords, w_rpe_weight, K, alpha, combined_shifts, nullptr, region_eta, region_phi, regions_h,
                            scale, positions, out_pre, den_sum, workspace, workspace_bytes, stream);
}

extern "C" int hept_attention_fwd_shifts32(const hept_shape* s, const float* q, const float* k, const float* v,
                                           const float* coords, const float* w_rpe_weight, int32_t K, const float* alpha,
                                           const int32_t* combined_shifts32, float* scale, int32_t* positions,
                                           float* out_pre, float* den_sum, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  HEPT_REQUIRE(combined_shifts32, HEPT_EINVAL, "attention_fwd_shifts32: null pointer");
  return attention_fwd_impl(s, q, k, v, coords, w_rpe_weight, K, alpha, nullptr, combined_shifts32, nullptr, nullptr, nullptr, scale,
                            positions, out_pre, den_sum, workspace, workspace_bytes, stream);
}

static int attention_fwd_impl(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                              const float* w_rpe_weight, int32_t K, const float* alpha, const int64_t* combined_shifts,
                              const int32_t* combined_shifts32, const float* region_eta, const float* region_phi,
                              const float* regions_h, float* scale, int32_t* positions, float* out_pre, float* den_sum,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(q && k && v && coords && w_rpe_weight && alpha && scale && positions && out_pre && den_sum && workspace,
               HEPT_EINVAL, "attention_fwd: null pointer");
  const bool packed = combined_shifts != nullptr || combined_shifts32 != nullptr;
  const bool regions = region_eta && region_phi && regions_h;
  HEPT_REQUIRE(packed != regions, HEPT_EINVAL,
               "attention_fwd: pass either combined_shifts or (region_eta, region_phi, regions_h)");
  HEPT_REQUIRE(hept_shape_supported(s->D, s->C, s->B), HEPT_EUNSUPPORTED,
               "attention_fwd: (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  FwdPlan p = plan_fwd(s);
  HEPT_REQUIRE(workspace_bytes >= p.total, HEPT_EWORKSPACE, "attention_fwd: workspace needs %zu bytes, got %zu", p.total,
               workspace_bytes);
  char* w = (char*)workspace;
  void* ext = w;                       w += p.ext_bytes;
  float* span = (float*)w;             w += p.span_bytes;
  float* hat = (float*)w;              w += p.hat_bytes;
  float* stage = (float*)w;            // aliases proj/keys/sort scratch (dead by then)
  float* proj = (float*)w;             w += p.proj_bytes;
  float* keys = (float*)w;             w += p.keys_bytes;
  void* sort_ws = w;
  int rc;
  if ((rc = hept_coord_scale_fwd(w_rpe_weight, s->H, s->D, s->C - 1, K, scale, stream))) return rc;
  bool hat_done = false;   // the fused projection kernel emits the scaled coordinates as a by-product
  if ((rc = hash_project_impl(s, q, k, coords, scale, alpha, proj, span, ext, p.ext_bytes, hat, &hat_done, stream))) return rc;
  if (combined_shifts32) rc = hept_keys_from_packed_shifts32(s, proj, span, combined_shifts32, keys, stream);
  else if (packed) rc = hept_keys_from_packed_shifts(s, proj, span, combined_shifts, keys, stream);
  else rc = hept_keys_from_region_indices(s, proj, span, region_eta, region_phi, regions_h, keys, stream);
  if (rc) return rc;
  if ((rc = hept_segmented_argsort(keys, 2 * s->T * s->H, s->N, positions, sort_ws, p.sort_bytes, stream))) return rc;
  if (!hat_done && (rc = hept_hat_coords(s, coords, scale, hat, stream))) return rc;
  if ((rc = hept_block_attention_fwd(s, q, k, v, coords, scale, hat, positions, stage, stream))) return rc;
  return hept_or_combine(s, stage, out_pre, den_sum, stream);
}
