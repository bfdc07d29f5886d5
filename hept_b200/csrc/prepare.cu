// a13-a17: the per-forward preparation of the AND-hash inputs as CUDA kernels.
//
//   example/ flavour (example/transformer.py:10-63, example/hept_utils.py:6-14): per-event rank of eta / phi ->
//   quantile region -> bit-pack phi over eta and the batch index over both -> pad every event to a block multiple by
//   repeating real points -> gather coords and codes into padded order.
//   src/ flavour (HEPT branch of src/models/baselines/transformer.py:43-57): one event, +inf padding, float region indices.
//
// The reference walks the events in a Python loop and calls argsort / fancy indexing per event.  Here the ranks of all
// events come from ONE segmented stable argsort (sort.cu; every event is a fixed-length segment padded with +inf), the
// regions are recomputed from the ranks wherever they are needed (a rank is 4 bytes per hit, the 2 x T*H region tables the
// reference materialises are 192), and the padding plan is index arithmetic on the host-provided event offsets.
//
// Arithmetic that must match the reference bit for bit:
//   width  = ceil(reciprocal(num_regions) * n_event)     float32, two roundings: `n / tensor` is Tensor.__rtruediv__,
//                                                         i.e. reciprocal() * n (example/hept_utils.py:8)
//   region = floor(rank / width) + 1                      exact in float32 for rank < 2^24 (both are integers)
//   bits   = ceil(log2(max + 1))                          (example/transformer.py:11-12) computed in integers; equal to
//                                                         the float32 formula for max < 2^21 (region codes are < 2^16)
// The padding order (argsort of the (table 0, head 0) code) sorts the integer codes themselves: ceil(code_bits / 8) radix
// passes instead of the four a float key needs.
// Sort tie-break: stable (lowest index first); the reference's argsort is unstable, so among points with EQUAL eta (or phi,
// or packed code) it may pick another order — the documented deviation of DESIGN.md section 1.
#include "common.cuh"

namespace hept {

int segmented_argsort_launch(const void* keys, int32_t num_segments, int32_t n, int32_t* positions, void* workspace,
                             size_t workspace_bytes, cudaStream_t st, int key_bits = 0);

constexpr int kPrepThreads = 256;
constexpr int kPrepMaxTH = 64;

__device__ __forceinline__ float region_width(float num_regions, int n_event) {
  return ceilf(__fmul_rn(__frcp_rn(num_regions), (float)n_event));
}
// floor(rank / width) + 1 with width an integer-valued float >= 1: integer division is the exact result
__device__ __forceinline__ int region_of(int rank, float width) { return rank / (int)width + 1; }
__device__ __forceinline__ int ceil_log2(long long m) {   // smallest b with 2^b >= m, m >= 1
  int b = 0;
  while ((1ll << b) < m) ++b;
  return b;
}
// event of padded row / raw point i: last e with start[e] <= i (start has E + 1 entries, ascending; empty events allowed)
__device__ __forceinline__ int event_of(const int32_t* __restrict__ start, int E, int i) {
  int lo = 0, hi = E;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(start + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// keys (2 * E, L): segment (a, e) holds coords[start_e .. start_e + n_e, a], padded with +inf up to L
// The first TH threads also start the per-(table, head) bookkeeping: meta[2 th] = bits1 = ceil(log2(max eta region + 1)),
// known from the event sizes alone (the largest region of an event is that of its last rank); meta[2 th + 1] = 0, the running
// maximum of code1 = (phi << bits1) | eta that prep_code_max_kernel raises.
__global__ void __launch_bounds__(kPrepThreads) prep_keys_kernel(const float* __restrict__ coords, int C,
                                                                 const int32_t* __restrict__ ev_start, int E, int L,
                                                                 float* __restrict__ keys, const float* __restrict__ regions_h,
                                                                 int TH, int32_t* __restrict__ meta) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)TH) {
    const float r_eta = __ldg(regions_h + idx);
    long long mx = 0;
    for (int e = 0; e < E; ++e) {
      const int n = __ldg(ev_start + e + 1) - __ldg(ev_start + e);
      if (n > 0) mx = max(mx, (long long)region_of(n - 1, region_width(r_eta, n)));
    }
    meta[2 * idx] = ceil_log2(mx + 1);
    meta[2 * idx + 1] = 0;
  }
  if (idx >= (size_t)2 * E * L) return;
  const int i = (int)(idx % L);
  const int seg = (int)(idx / L);
  const int a = seg / E, e = seg - a * E;
  const int s = __ldg(ev_start + e), n = __ldg(ev_start + e + 1) - s;
  keys[idx] = i < n ? __ldg(coords + (size_t)(s + i) * C + a) : __int_as_float(0x7f800000);
}

// rank[a][start_e + pos[(a, e), r]] = r for r < n_e
__global__ void __launch_bounds__(kPrepThreads) prep_rank_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ ev_start,
                                                                 int E, int L, int n_raw, int32_t* __restrict__ rank) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)2 * E * L) return;
  const int r = (int)(idx % L);
  const int seg = (int)(idx / L);
  const int a = seg / E, e = seg - a * E;
  const int s = __ldg(ev_start + e), n = __ldg(ev_start + e + 1) - s;
  if (r < n) rank[(size_t)a * n_raw + s + __ldg(pos + idx)] = r;
}

// meta[2 th + 1] = max over all points of code1 = (phi << bits1) | eta for (table, head) th = blockIdx.y (integer atomicMax:
// order-independent, deterministic); bits2 = ceil(log2(that + 1)) is taken where it is used.
__global__ void __launch_bounds__(kPrepThreads) prep_code_max_kernel(const int32_t* __restrict__ rank, const int32_t* __restrict__ ev_start,
                                                                     int E, int n_raw, const float* __restrict__ regions_h, int TH,
                                                                     int32_t* __restrict__ meta) {
  __shared__ int red[kPrepThreads / 32];
  const int th = blockIdx.y;
  const int p = blockIdx.x * kPrepThreads + threadIdx.x;
  int code = 0;
  if (p < n_raw) {
    const int e = event_of(ev_start, E, p);
    const int n = __ldg(ev_start + e + 1) - __ldg(ev_start + e);
    const int eta = region_of(__ldg(rank + p), region_width(__ldg(regions_h + th), n));
    const int phi = region_of(__ldg(rank + n_raw + p), region_width(__ldg(regions_h + TH + th), n));
    code = (phi << __ldg(meta + 2 * th)) | eta;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) code = max(code, __shfl_xor_sync(0xffffffffu, code, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = code;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = red[0];
#pragma unroll
    for (int w = 1; w < kPrepThreads / 32; ++w) m = max(m, red[w]);
    atomicMax(meta + 2 * th + 1, m);
  }
}

// the packed code of raw point p for (table, head) th (example/transformer.py:55-56)
__device__ __forceinline__ long long packed_code(int rank_eta, int rank_phi, int n_event, long long batch_id, float r_eta,
                                                 float r_phi, int bits1, int bits2) {
  const long long eta = region_of(rank_eta, region_width(r_eta, n_event));
  const long long phi = region_of(rank_phi, region_width(r_phi, n_event));
  return (batch_id << bits2) | ((phi << bits1) | eta);
}

// key00[p] = code of (table 0, head 0) as an integer sort key: the order the padding rows are drawn from
// (example/transformer.py:59,23)
__global__ void __launch_bounds__(kPrepThreads) prep_key00_kernel(const int32_t* __restrict__ rank, const int64_t* __restrict__ batch,
                                                                  const int32_t* __restrict__ ev_start, int E, int n_raw,
                                                                  const float* __restrict__ regions_h, int TH,
                                                                  const int32_t* __restrict__ meta, uint32_t* __restrict__ key00) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_raw) return;
  const int e = event_of(ev_start, E, p);
  const int n = __ldg(ev_start + e + 1) - __ldg(ev_start + e);
  const long long code = packed_code(__ldg(rank + p), __ldg(rank + n_raw + p), n, __ldg(batch + p), __ldg(regions_h),
                                     __ldg(regions_h + TH), __ldg(meta), ceil_log2((long long)__ldg(meta + 1) + 1));
  key00[p] = (uint32_t)code;   // below 2^code_bits (the caller's bound, checked on the host side of the call)
}

// Padded row i: which raw point it shows (take), whether it is real, its coordinates and its codes for every (table, head).
// Event e owns padded rows [pad_start[e], pad_start[e + 1]): the first n_e are its own points in order, the rest repeat
// order[ev_end[e] - block + j], j = 0 .. pad_e - 1 (negative indices wrap like Python's; an event shorter than a block
// reaches into the previous event, exactly as the reference does).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) prep_emit_kernel(const float* __restrict__ coords, int C, const int64_t* __restrict__ batch,
                                                            const int32_t* __restrict__ rank, const int32_t* __restrict__ order,
                                                            const int32_t* __restrict__ ev_start, const int32_t* __restrict__ pad_start,
                                                            int E, int n_raw, int n_pad, int block, const float* __restrict__ regions_h,
                                                            int TH, const int32_t* __restrict__ meta, int64_t* __restrict__ shifts,
                                                            int32_t* __restrict__ shifts32, int64_t* __restrict__ take,
                                                            uint8_t* __restrict__ is_real, float* __restrict__ coords_pad) {
  __shared__ float s_reg[2 * kPrepMaxTH];
  __shared__ int s_meta[2 * kPrepMaxTH];
  for (int j = threadIdx.x; j < 2 * TH; j += THREADS) {
    s_reg[j] = __ldg(regions_h + j);
    s_meta[j] = (j & 1) ? ceil_log2((long long)__ldg(meta + j) + 1) : __ldg(meta + j);    // (bits1, bits2) per (table, head)
  }
  __syncthreads();
  const int i = blockIdx.x * THREADS + threadIdx.x;
  if (i >= n_pad) return;
  const int e = event_of(pad_start, E, i);
  const int local = i - __ldg(pad_start + e);
  const int s = __ldg(ev_start + e), n = __ldg(ev_start + e + 1) - s;
  int src;
  const bool real = local < n;
  if (real) {
    src = s + local;
  } else {
    int o = s + n - block + (local - n);
    if (o < 0) o += n_raw;
    src = __ldg(order + o);
  }
  take[i] = src;
  is_real[i] = real ? 1 : 0;
  for (int c = 0; c < C; ++c) coords_pad[(size_t)i * C + c] = __ldg(coords + (size_t)src * C + c);
  // the codes of the point shown, computed with ITS event's size (a borrowed point keeps its own event's code)
  const int es = event_of(ev_start, E, src);
  const int ns = __ldg(ev_start + es + 1) - __ldg(ev_start + es);
  const int r_eta = __ldg(rank + src), r_phi = __ldg(rank + n_raw + src);
  const long long b = __ldg(batch + src);
  for (int th = 0; th < TH; ++th) {
    const long long code = packed_code(r_eta, r_phi, ns, b, s_reg[th], s_reg[TH + th], s_meta[2 * th], s_meta[2 * th + 1]);
    shifts[(size_t)th * n_pad + i] = code;
    if (shifts32) shifts32[(size_t)th * n_pad + i] = (int32_t)code;
  }
}

// ---- src/ flavour -----------------------------------------------------------------------------------------------
// keys (2, n_pad): coords[:, a] for real rows, +inf for the padding rows (pad_to_multiple(..., value=inf))
__global__ void __launch_bounds__(kPrepThreads) prep_single_keys_kernel(const float* __restrict__ coords, int C, int n_raw, int n_pad,
                                                                        float* __restrict__ keys, float* __restrict__ coords_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const bool real = i < n_raw;
  keys[i] = real ? __ldg(coords + (size_t)i * C) : __int_as_float(0x7f800000);
  keys[n_pad + i] = real ? __ldg(coords + (size_t)i * C + 1) : __int_as_float(0x7f800000);
  // coords of the padding rows are set to zero after the region indices are taken (src/.../transformer.py:57)
  for (int c = 0; c < C; ++c) coords_pad[(size_t)i * C + c] = real ? __ldg(coords + (size_t)i * C + c) : 0.f;
}

__global__ void __launch_bounds__(kPrepThreads) prep_single_rank_kernel(const int32_t* __restrict__ pos, int n_pad, int32_t* __restrict__ rank) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n_pad) return;
  const int a = idx / n_pad, r = idx - a * n_pad;
  rank[(size_t)a * n_pad + __ldg(pos + idx)] = r;
}

// region_eta / region_phi (TH, n_pad) float32 = floor(rank / width) + 1, width from the PADDED size (the reference pads first)
__global__ void __launch_bounds__(kPrepThreads) prep_single_regions_kernel(const int32_t* __restrict__ rank, int n_pad,
                                                                           const float* __restrict__ regions_h, int TH,
                                                                           float* __restrict__ region_eta, float* __restrict__ region_phi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int th = blockIdx.y;
  if (i >= n_pad) return;
  region_eta[(size_t)th * n_pad + i] = (float)region_of(__ldg(rank + i), region_width(__ldg(regions_h + th), n_pad));
  region_phi[(size_t)th * n_pad + i] = (float)region_of(__ldg(rank + n_pad + i), region_width(__ldg(regions_h + TH + th), n_pad));
}

struct PrepPlan {
  size_t keys_bytes, pos_bytes, rank_bytes, meta_bytes, key00_bytes, order_bytes, sort_bytes, total;
};
static PrepPlan plan_prepare(int n_raw, int E, int L) {
  PrepPlan p;
  p.keys_bytes = align_up(sizeof(float) * (size_t)2 * E * L, 256);
  p.pos_bytes = align_up(sizeof(int32_t) * (size_t)2 * E * L, 256);
  p.rank_bytes = align_up(sizeof(int32_t) * (size_t)2 * n_raw, 256);
  p.meta_bytes = align_up(sizeof(int32_t) * 2 * kPrepMaxTH, 256);
  p.key00_bytes = align_up(sizeof(float) * (size_t)n_raw, 256);
  p.order_bytes = align_up(sizeof(int32_t) * (size_t)n_raw, 256);
  const size_t s1 = hept_argsort_workspace_bytes(2 * E, L), s2 = hept_argsort_workspace_bytes(1, n_raw);
  p.sort_bytes = align_up(s1 > s2 ? s1 : s2, 256);
  p.total = p.keys_bytes + p.pos_bytes + p.rank_bytes + p.meta_bytes + p.key00_bytes + p.order_bytes + p.sort_bytes;
  return p;
}

}  // namespace hept

using namespace hept;

extern "C" size_t hept_prepare_batched_workspace_bytes(int32_t n_raw, int32_t num_events, int32_t max_event) {
  if (n_raw <= 0 || num_events <= 0 || max_event <= 0) return 0;
  return plan_prepare(n_raw, num_events, max_event).total;
}

extern "C" int hept_prepare_batched(const float* coords, int32_t C, const int64_t* batch, const int32_t* event_start,
                                    const int32_t* pad_start, int32_t num_events, int32_t n_raw, int32_t n_pad,
                                    int32_t max_event, const float* regions_h, int32_t TH, int32_t block_size,
                                    int32_t code_bits, int64_t* combined_shifts, int32_t* combined_shifts32, int64_t* take, uint8_t* is_real,
                                    float* coords_pad, void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(coords && batch && event_start && pad_start && regions_h && combined_shifts && take && is_real && coords_pad &&
                   workspace,
               HEPT_EINVAL, "prepare_batched: null pointer");
  HEPT_REQUIRE(C >= 2 && num_events > 0 && n_raw > 0 && n_pad >= n_raw && max_event > 0 && max_event <= n_raw && block_size > 0,
               HEPT_EINVAL, "prepare_batched: bad sizes (C=%d E=%d n_raw=%d n_pad=%d max_event=%d B=%d)", C, num_events, n_raw,
               n_pad, max_event, block_size);
  HEPT_REQUIRE(TH > 0 && TH <= kPrepMaxTH, HEPT_EUNSUPPORTED, "prepare_batched: T*H=%d outside [1, %d]", TH, kPrepMaxTH);
  HEPT_REQUIRE(n_raw < (1 << 24), HEPT_EUNSUPPORTED, "prepare_batched: %d points: ranks are no longer exact in float32", n_raw);
  // code_bits: the caller's upper bound on the bits of the (table 0, head 0) code (batch index over both region indices);
  // it decides how many 8-bit passes the padding-order sort runs.  0 = unknown: all 32.
  HEPT_REQUIRE(code_bits >= 0 && code_bits <= 32, HEPT_EINVAL, "prepare_batched: code_bits=%d outside [0, 32]", code_bits);
  if (code_bits == 0) code_bits = 32;
  PrepPlan p = plan_prepare(n_raw, num_events, max_event);
  HEPT_REQUIRE(workspace_bytes >= p.total, HEPT_EWORKSPACE, "prepare_batched: workspace needs %zu bytes, got %zu", p.total,
               workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  float* keys = (float*)w;            w += p.keys_bytes;
  int32_t* pos = (int32_t*)w;         w += p.pos_bytes;
  int32_t* rank = (int32_t*)w;        w += p.rank_bytes;
  int32_t* meta = (int32_t*)w;        w += p.meta_bytes;
  uint32_t* key00 = (uint32_t*)w;     w += p.key00_bytes;
  int32_t* order = (int32_t*)w;       w += p.order_bytes;
  void* sort_ws = w;
  const int E = num_events, L = max_event;
  const size_t seg_items = (size_t)2 * E * L;
  const unsigned seg_grid = (unsigned)((seg_items + kPrepThreads - 1) / kPrepThreads);
  prep_keys_kernel<<<seg_grid, kPrepThreads, 0, st>>>(coords, C, event_start, E, L, keys, regions_h, TH, meta);
  HEPT_CHECK_LAUNCH("prep_keys");
  if (int rc = segmented_argsort_launch(keys, 2 * E, L, pos, sort_ws, p.sort_bytes, st)) return rc;
  prep_rank_kernel<<<seg_grid, kPrepThreads, 0, st>>>(pos, event_start, E, L, n_raw, rank);
  HEPT_CHECK_LAUNCH("prep_rank");
  prep_code_max_kernel<<<dim3((n_raw + kPrepThreads - 1) / kPrepThreads, TH), kPrepThreads, 0, st>>>(rank, event_start, E, n_raw,
                                                                                                     regions_h, TH, meta);
  HEPT_CHECK_LAUNCH("prep_code_max");
  prep_key00_kernel<<<(n_raw + kPrepThreads - 1) / kPrepThreads, kPrepThreads, 0, st>>>(rank, batch, event_start, E, n_raw, regions_h,
                                                                                         TH, meta, key00);
  HEPT_CHECK_LAUNCH("prep_key00");
  if (int rc = segmented_argsort_launch(key00, 1, n_raw, order, sort_ws, p.sort_bytes, st, code_bits)) return rc;
  prep_emit_kernel<128><<<(n_pad + 127) / 128, 128, 0, st>>>(coords, C, batch, rank, order, event_start, pad_start, E, n_raw, n_pad,
                                                             block_size, regions_h, TH, meta, combined_shifts, combined_shifts32, take,
                                                             is_real, coords_pad);
  HEPT_CHECK_LAUNCH("prep_emit");
  return HEPT_OK;
}

extern "C" size_t hept_prepare_single_workspace_bytes(int32_t n_pad) {
  if (n_pad <= 0) return 0;
  return align_up(sizeof(float) * (size_t)2 * n_pad, 256) + 2 * align_up(sizeof(int32_t) * (size_t)2 * n_pad, 256) +
         align_up(hept_argsort_workspace_bytes(2, n_pad), 256);
}

extern "C" int hept_prepare_single(const float* coords, int32_t C, int32_t n_raw, int32_t n_pad, const float* regions_h,
                                   int32_t TH, float* coords_pad, float* region_eta, float* region_phi, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(coords && regions_h && coords_pad && region_eta && region_phi && workspace, HEPT_EINVAL, "prepare_single: null pointer");
  HEPT_REQUIRE(C >= 2 && n_raw > 0 && n_pad >= n_raw && TH > 0, HEPT_EINVAL, "prepare_single: bad sizes (C=%d n_raw=%d n_pad=%d TH=%d)", C,
               n_raw, n_pad, TH);
  HEPT_REQUIRE(n_pad < (1 << 24), HEPT_EUNSUPPORTED, "prepare_single: %d points: ranks are no longer exact in float32", n_pad);
  HEPT_REQUIRE(workspace_bytes >= hept_prepare_single_workspace_bytes(n_pad), HEPT_EWORKSPACE, "prepare_single: workspace needs %zu bytes",
               hept_prepare_single_workspace_bytes(n_pad));
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  float* keys = (float*)w;         w += align_up(sizeof(float) * (size_t)2 * n_pad, 256);
  int32_t* pos = (int32_t*)w;      w += align_up(sizeof(int32_t) * (size_t)2 * n_pad, 256);
  int32_t* rank = (int32_t*)w;     w += align_up(sizeof(int32_t) * (size_t)2 * n_pad, 256);
  void* sort_ws = w;
  const size_t sort_bytes = align_up(hept_argsort_workspace_bytes(2, n_pad), 256);
  prep_single_keys_kernel<<<(n_pad + kPrepThreads - 1) / kPrepThreads, kPrepThreads, 0, st>>>(coords, C, n_raw, n_pad, keys, coords_pad);
  HEPT_CHECK_LAUNCH("prep_single_keys");
  if (int rc = segmented_argsort_launch(keys, 2, n_pad, pos, sort_ws, sort_bytes, st)) return rc;
  prep_single_rank_kernel<<<(2 * n_pad + kPrepThreads - 1) / kPrepThreads, kPrepThreads, 0, st>>>(pos, n_pad, rank);
  HEPT_CHECK_LAUNCH("prep_single_rank");
  prep_single_regions_kernel<<<dim3((n_pad + kPrepThreads - 1) / kPrepThreads, TH), kPrepThreads, 0, st>>>(rank, n_pad, regions_h, TH,
                                                                                                        region_eta, region_phi);
  HEPT_CHECK_LAUNCH("prep_single_regions");
  return HEPT_OK;
}
