// a7: segmented stable argsort of float32 keys (example/hept.py:67-68), as an LSD radix sort over
// order-preserving uint32 keys: 4 passes of 8 bits, each pass = per-tile digit histogram + stable
// scatter.  All segments are processed by the same launches (grid = tiles x segments), so the 2*T*H
// independent sorts of one attention call fill the machine together.
//
// Stability: inside a tile, keys are ranked warp by warp, 32 at a time in index order, with a ballot-built
// mask of the lanes holding the same digit giving each lane its rank among equals; tiles and warps are
// combined by exclusive prefix sums in index order.  Equal keys therefore keep ascending original index,
// the tie-break the oracle pins (torch.argsort(stable=True)).  The tile is reordered in shared memory
// first so that the global writes of a digit run are contiguous.
//
// Segments that fit the shared memory of one thread-block cluster take the cluster-resident path below instead: one
// global read of the keys, four passes through distributed shared memory, one global write of the positions.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace hept {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef HEPT_SORT_ITEMS
#define HEPT_SORT_ITEMS 16
#endif
#ifndef HEPT_SORT_BLOCKS
#define HEPT_SORT_BLOCKS 5
#endif
constexpr int kItemsPerThread = HEPT_SORT_ITEMS;
constexpr int kTile = kSortThreads * kItemsPerThread;  // 4096 keys per CTA

// lanes of the warp whose (valid, 8-bit digit) equals mine, from nine ballots.  __match_any_sync gives the same mask
// but its cost grows with the number of distinct values in the warp (measured: the low-byte passes, where all 32
// digits differ, ran 2x slower than the exponent-byte pass); ballots cost the same whatever the data.
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t digit, bool ok) {
  uint32_t m = __ballot_sync(0xffffffffu, ok);
  m = ok ? m : ~m;
#pragma unroll
  for (int b = 0; b < kRadixBits; ++b) {
    const bool bit = (digit >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

__device__ __forceinline__ uint32_t load_key(const void* keys, size_t i, bool as_float) {
  return as_float ? ordered_bits(reinterpret_cast<const float*>(keys)[i])
                  : reinterpret_cast<const uint32_t*>(keys)[i];
}

// tile_hist (segments, tiles, 256)
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const void* __restrict__ keys, bool as_float, int n,
                                                                   int tiles, int shift,
                                                                   uint32_t* __restrict__ tile_hist) {
  __shared__ uint32_t hist[kRadix];
  const int tile = blockIdx.x, seg = blockIdx.y;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)seg * n;
  const int start = tile * kTile;
  // every load of the thread in flight before the first shared-memory atomic
  uint32_t kreg[kItemsPerThread];
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = start + it * kSortThreads + threadIdx.x;
    kreg[it] = i < n ? load_key(keys, base + i, as_float) : 0u;
  }
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = start + it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(kreg[it] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  tile_hist[((size_t)seg * tiles + tile) * kRadix + threadIdx.x] = hist[threadIdx.x];
}

// keys_in/idx_in -> keys_out/idx_out.  idx_in == nullptr means the identity (first pass);
// keys_out == nullptr means the keys are no longer needed (last pass).
__global__ void __launch_bounds__(kSortThreads, HEPT_SORT_BLOCKS) radix_scatter_kernel(const void* __restrict__ keys_in, bool as_float,
                                                                      const int32_t* __restrict__ idx_in, int n,
                                                                      int tiles, int shift,
                                                                      const uint32_t* __restrict__ tile_hist,
                                                                      uint32_t* __restrict__ keys_out,
                                                                      int32_t* __restrict__ idx_out) {
  __shared__ uint32_t warp_off[kSortWarps][kRadix];   // per-warp digit counts, then tile-local start offsets
  __shared__ uint32_t scan_tmp[kRadix];
  __shared__ uint32_t local_start[kRadix];            // first tile-local slot of each digit
  __shared__ uint32_t global_base[kRadix];            // where this tile's run of each digit starts in the segment
  __shared__ uint32_t s_key[kTile];                   // the tile, reordered by digit (stable)
  __shared__ int32_t s_idx[kTile];
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)seg * n;

  // (0) every global load of the CTA is issued before anything depends on one: keys, incoming indices, histograms
  uint32_t key[kItemsPerThread];
  const int wstart = tile * kTile + warp * (32 * kItemsPerThread);
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = wstart + it * 32 + lane;
    key[it] = i < n ? load_key(keys_in, base + i, as_float) : 0u;
  }
  // the incoming indices too: loaded inside the ranking loop their latency was exposed once per item (-18 us per sort)
  int32_t idxr[kItemsPerThread];
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = wstart + it * 32 + lane;
    idxr[it] = (idx_in && i < n) ? idx_in[base + i] : i;
  }
  // (1) where does digit `tid` of this tile start in the segment?
  uint32_t before = 0, total = 0, mine = 0;
  {
    const uint32_t* th = tile_hist + (size_t)seg * tiles * kRadix + tid;
#pragma unroll 8
    for (int t = 0; t < tiles; ++t) {
      const uint32_t c = th[(size_t)t * kRadix];
      total += c;
      before += t < tile ? c : 0u;
      mine = t == tile ? c : mine;
    }
  }
  // exclusive scans over the 256 digits (segment-wide totals, this tile's own counts): shuffle scan inside each
  // warp, then the 8 warp totals
  uint32_t inc_t = total, inc_m = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, inc_t, off), c = __shfl_up_sync(0xffffffffu, inc_m, off);
    if (lane >= off) { inc_t += a; inc_m += c; }
  }
  if (lane == 31) { scan_tmp[warp] = inc_t; scan_tmp[kSortWarps + warp] = inc_m; }
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) warp_off[w][tid] = 0;
  __syncthreads();
  uint32_t pre_t = 0, pre_m = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    pre_t += w < warp ? scan_tmp[w] : 0u;
    pre_m += w < warp ? scan_tmp[kSortWarps + w] : 0u;
  }
  global_base[tid] = pre_t + inc_t - total + before;
  const uint32_t my_local = pre_m + inc_m - mine;
  local_start[tid] = my_local;

  // (2) per-warp digit counts; each warp owns a contiguous run of 512 keys, kept in registers.  Counts only: shared-memory
  // atomics (their order does not matter), so that the ballot masks need not stay in registers until step (4) -- 16 fewer
  // registers per thread are what lets FIVE CTAs share an SM, and the 720 CTAs of an attention call's 48 segments run as one wave
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = wstart + it * 32 + lane;
    if (i < n) atomicAdd(&warp_off[warp][(key[it] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  // (3) exclusive prefix over warps, seeded with the digit's tile-local start
  {
    uint32_t run = my_local;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = warp_off[w][tid];
      warp_off[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
  // (4) rank in index order (ballot masks of the lanes holding the same digit) and park every key at its tile-local slot
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = wstart + it * 32 + lane;
    const bool ok = i < n;
    const uint32_t dg = (key[it] >> shift) & (kRadix - 1);
    const uint32_t peers = same_digit_lanes(dg, ok);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = warp_off[warp][dg] + rank;
    __syncwarp();
    if (ok && lane == __ffs(peers) - 1) warp_off[warp][dg] += __popc(peers);
    __syncwarp();
    if (ok) {
      s_key[pos] = key[it];
      s_idx[pos] = idxr[it];
    }
  }
  __syncthreads();
  // (5) write the tile out slot by slot: consecutive threads hit consecutive addresses inside each digit run
  const int count = min(kTile, n - tile * kTile);
  for (int e = tid; e < count; e += kSortThreads) {
    const uint32_t kk = s_key[e];
    const uint32_t dg = (kk >> shift) & (kRadix - 1);
    const uint32_t pos = global_base[dg] + ((uint32_t)e - local_start[dg]);
    if (keys_out) keys_out[base + pos] = kk;
    idx_out[base + pos] = s_idx[e];
  }
}

// ---- cluster-resident path --------------------------------------------------------------------------------------
// One cluster of `cs` CTAs per segment; CTA r keeps slots [r * cap, (r + 1) * cap) of the segment's current order (keys and
// original indices) in its shared memory.  A pass counts digits per warp (ballots, as above), publishes the CTA's digit
// totals, reads the totals of the whole cluster through DSMEM after a cluster barrier, and scatters every (key, index)
// straight into the slot it belongs to in the shared memory of the CTA that owns that slot.  Order inside a CTA is
// warp run, then iteration, then lane; CTAs are combined in rank order: the same stable order as the global path.
#ifndef HEPT_CS_THREADS
#define HEPT_CS_THREADS 512
#endif
constexpr int kCsThreads = HEPT_CS_THREADS;
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kCsMaxCap = 12288;   // slots per CTA: 4 buffers x 4 B x cap + 17.5 KB of counters <= 227 KB

static size_t cluster_sort_smem(int cap) { return (size_t)cap * 16 + (size_t)(kCsWarps * kRadix + kRadix + 32) * 4; }

__global__ void __launch_bounds__(kCsThreads, 1) cluster_sort_kernel(const float* __restrict__ keys, int n, int cap,
                                                                      int32_t* __restrict__ positions) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  const int seg = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(16) uint8_t smem[];
  // two (keys, indices) orders: order b lives at kbuf + b * cap, ibuf + b * cap
  uint32_t* kbuf = reinterpret_cast<uint32_t*>(smem);
  int32_t* ibuf = reinterpret_cast<int32_t*>(kbuf + 2 * cap);
  uint32_t (*warp_off)[kRadix] = reinterpret_cast<uint32_t (*)[kRadix]>(ibuf + 2 * cap);
  uint32_t* hist_pub = &warp_off[kCsWarps - 1][kRadix - 1] + 1;   // digit totals of this CTA, read by the whole cluster
  uint32_t* scan_tmp = hist_pub + kRadix;
  const size_t base = (size_t)seg * n;
  const int lo = r * cap;
  const int cnt = max(0, min(cap, n - lo));
  for (int e = tid; e < cnt; e += kCsThreads) {
    kbuf[e] = ordered_bits(keys[base + lo + e]);
    ibuf[e] = lo + e;
  }
  // contiguous run of every warp, a multiple of 32 long
  const int per = (((cnt + kCsWarps - 1) / kCsWarps) + 31) & ~31;
  const int wlo = min(cnt, warp * per), whi = min(cnt, wlo + per);
  int cur = 0;
#pragma unroll 1
  for (int shift = 0; shift < 32; shift += kRadixBits) {
    const uint32_t* ck = kbuf + cur * cap;
    const int32_t* ci = ibuf + cur * cap;
    for (int d = tid; d < kCsWarps * kRadix; d += kCsThreads) (&warp_off[0][0])[d] = 0;
    __syncthreads();
    // (1) digit counts of every warp's run
#pragma unroll 1
    for (int e0 = wlo; e0 < whi; e0 += 32) {
      const int e = e0 + lane;
      const bool ok = e < whi;
      const uint32_t dg = ok ? (ck[e] >> shift) & (kRadix - 1) : 0u;
      const uint32_t peers = same_digit_lanes(dg, ok);
      if (ok && lane == __ffs(peers) - 1) warp_off[warp][dg] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    // (2) exclusive prefix over the warps, CTA totals published for the cluster
    if (tid < kRadix) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kCsWarps; ++w) {
        const uint32_t c = warp_off[w][tid];
        warp_off[w][tid] = run;
        run += c;
      }
      hist_pub[tid] = run;
    }
    cluster.sync();
    // (3) digit `tid`: count in the whole segment and in the CTAs before this one; exclusive scan over the digits
    uint32_t total = 0, before = 0;
    if (tid < kRadix) {
      for (int q = 0; q < cs; ++q) {
        const uint32_t c = *cluster.map_shared_rank(hist_pub + tid, q);
        total += c;
        before += q < r ? c : 0u;
      }
    }
    uint32_t inc = total;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t a = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += a;
    }
    if (lane == 31) scan_tmp[warp] = inc;
    __syncthreads();
    if (tid < kRadix) {
      uint32_t pre = 0;
#pragma unroll
      for (int w = 0; w < kRadix / 32; ++w) pre += w < warp ? scan_tmp[w] : 0u;
      const uint32_t first = pre + inc - total + before;     // slot of this CTA's first key with digit `tid`
#pragma unroll
      for (int w = 0; w < kCsWarps; ++w) warp_off[w][tid] += first;
    }
    __syncthreads();
    // (4) rank in the same order; every (key, index) goes to the CTA that owns its slot
    uint32_t* nk = kbuf + (cur ^ 1) * cap;
    int32_t* ni = ibuf + (cur ^ 1) * cap;
#pragma unroll 1
    for (int e0 = wlo; e0 < whi; e0 += 32) {
      const int e = e0 + lane;
      const bool ok = e < whi;
      const uint32_t key = ok ? ck[e] : 0u;
      const int32_t idx = ok ? ci[e] : 0;
      const uint32_t dg = (key >> shift) & (kRadix - 1);
      const uint32_t peers = same_digit_lanes(ok ? dg : 0u, ok);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t pos = 0;
      if (ok) pos = warp_off[warp][dg] + rank;
      __syncwarp();
      if (ok && lane == __ffs(peers) - 1) warp_off[warp][dg] += __popc(peers);
      __syncwarp();
      if (ok) {
        const uint32_t q = pos / (uint32_t)cap, off = pos - q * (uint32_t)cap;
        *cluster.map_shared_rank(nk + off, q) = key;
        *cluster.map_shared_rank(ni + off, q) = idx;
      }
    }
    cluster.sync();     // every slot of the next order is written; nobody reads the old order any more
    cur ^= 1;
  }
  for (int e = tid; e < cnt; e += kCsThreads) positions[base + lo + e] = ibuf[cur * cap + e];
}

// process-wide: 0 = cluster-resident path when the segment fits (default), 1 = global passes always
static std::atomic<int> g_sort_variant{0};

// cluster size for `segments` sorts of n keys, or 0 when the global passes take them.  Measured on B200 with 48 segments
// (us, cluster / global): n = 1 300: 21 / 51, 6 100: 32 / 54, 12 000: 47 / 58, 20 000: 71 / 69, 60 000: 200 / 128 -- both
// paths spend ~3 SM cycles per key and pass on ranking; the cluster path saves the seven extra launches and the
// histogram reads, which is what short segments are made of, and loses once its clusters no longer fit one wave.
// HEPT_SORT_CLUSTER=cs forces a cluster size (experiments), HEPT_SORT_CLUSTER=-1 the global passes.
static int cluster_size_for(int segments, int n) {
  static const int forced = [] { const char* e = getenv("HEPT_SORT_CLUSTER"); return e ? atoi(e) : 0; }();
  if (g_sort_variant.load(std::memory_order_relaxed) == 1 || forced < 0) return 0;
  const int need = (n + kCsMaxCap - 1) / kCsMaxCap;
  if (forced > 0) return need > 8 ? 0 : max(min(forced, 8), need);
  if (n > 16384) return 0;    // measured again with 2 segments of 60 000 keys (prepare.cu): 99 us in 5-CTA clusters (the remote
                              // stores of the ranking serialise) against ~50 us for the eight small launches of the global passes
  const int cs = n <= 2048 ? 1 : 2;
  return (long long)segments * cs <= 2 * 148 ? cs : 0;
}

struct SortPlan {
  int tiles;
  size_t hist_bytes, keys_bytes, idx_bytes, total;
};

int segmented_argsort_launch(const void* keys, int32_t num_segments, int32_t n, int32_t* positions, void* workspace,
                             size_t workspace_bytes, cudaStream_t st, int key_bits = 0);

static SortPlan plan_sort(int segments, int n) {
  SortPlan p;
  p.tiles = (n + kTile - 1) / kTile;
  p.hist_bytes = align_up(sizeof(uint32_t) * (size_t)segments * p.tiles * kRadix, 256);
  p.keys_bytes = align_up(sizeof(uint32_t) * (size_t)segments * n, 256);
  p.idx_bytes = align_up(sizeof(int32_t) * (size_t)segments * n, 256);
  p.total = p.hist_bytes + 2 * p.keys_bytes + 2 * p.idx_bytes;
  return p;
}

}  // namespace hept

using namespace hept;

extern "C" void hept_set_sort_variant(int variant) { g_sort_variant.store(variant == 1 ? 1 : 0, std::memory_order_relaxed); }
extern "C" int hept_get_sort_variant(void) { return g_sort_variant.load(std::memory_order_relaxed); }

extern "C" size_t hept_argsort_workspace_bytes(int32_t num_segments, int32_t n) {
  if (num_segments <= 0 || n <= 0) return 0;
  return plan_sort(num_segments, n).total;
}

extern "C" int hept_segmented_argsort(const float* keys, int32_t num_segments, int32_t n, int32_t* positions,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  return segmented_argsort_launch(keys, num_segments, n, positions, workspace, workspace_bytes, (cudaStream_t)stream);
}

// key_bits == 0: float32 keys (four 8-bit passes over their order-preserving bits).  key_bits > 0: the keys are uint32 whose
// bits above key_bits are zero (the packed region codes of prepare.cu): ceil(key_bits / 8) passes, global path only.
int hept::segmented_argsort_launch(const void* keys_v, int32_t num_segments, int32_t n, int32_t* positions, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream, int key_bits) {
  const float* keys = (const float*)keys_v;
  HEPT_REQUIRE(keys && positions && workspace && num_segments > 0 && n > 0, HEPT_EINVAL,
               "segmented_argsort: bad argument");
  SortPlan p = plan_sort(num_segments, n);
  HEPT_REQUIRE(workspace_bytes >= p.total, HEPT_EWORKSPACE, "segmented_argsort: workspace needs %zu bytes, got %zu",
               p.total, workspace_bytes);
  HEPT_REQUIRE(num_segments <= 65535, HEPT_EINVAL, "segmented_argsort: too many segments (%d)", num_segments);
  cudaStream_t st = (cudaStream_t)stream;
  if (const int cs = key_bits > 0 ? 0 : cluster_size_for(num_segments, n)) {
    const int cap = (n + cs - 1) / cs;
    const size_t smem = cluster_sort_smem(cap);
    static DeviceOnce opted;
    if (opted.needed()) {
      HEPT_REQUIRE(cudaFuncSetAttribute(cluster_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_sort_smem(kCsMaxCap)) == cudaSuccess,
                   HEPT_ECUDA, "segmented_argsort: cannot opt in to %zu bytes of shared memory", cluster_sort_smem(kCsMaxCap));
      opted.mark();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)num_segments * cs);
    cfg.blockDim = dim3(kCsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_sort_kernel, keys, (int)n, cap, positions);
    if (e == cudaSuccess) {
      count_launch(1);
      return HEPT_OK;
    }
    (void)cudaGetLastError();     // a cluster shape this device cannot place: the global passes take it
  }
  char* w = (char*)workspace;
  uint32_t* hist = (uint32_t*)w;            w += p.hist_bytes;
  uint32_t* kbuf[2] = {(uint32_t*)w, (uint32_t*)(w + p.keys_bytes)};  w += 2 * p.keys_bytes;
  int32_t* ibuf[2] = {(int32_t*)w, (int32_t*)(w + p.idx_bytes)};
  dim3 grid(p.tiles, num_segments);
  static DeviceOnce configured;     // five CTAs of 43 KB per SM need the largest shared-memory carve-out
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "segmented_argsort: %s", cudaGetErrorString(e));
    configured.mark();
  }
  const int passes = key_bits > 0 ? (key_bits + kRadixBits - 1) / kRadixBits : 32 / kRadixBits;
  const void* kin = keys;
  const int32_t* iin = nullptr;
  for (int pass = 0; pass < passes; ++pass) {
    const bool first = pass == 0, last = pass == passes - 1;
    const bool as_float = first && key_bits == 0;
    radix_hist_kernel<<<grid, kSortThreads, 0, st>>>(kin, as_float, n, p.tiles, pass * kRadixBits, hist);
    HEPT_CHECK_LAUNCH("radix_hist");
    uint32_t* kout = last ? nullptr : kbuf[pass & 1];
    int32_t* iout = last ? positions : ibuf[pass & 1];
    radix_scatter_kernel<<<grid, kSortThreads, 0, st>>>(kin, as_float, iin, n, p.tiles, pass * kRadixBits, hist, kout,
                                                        iout);
    HEPT_CHECK_LAUNCH("radix_scatter");
    kin = kout;
    iin = iout;
  }
  return HEPT_OK;
}
