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
#include "common.cuh"

namespace hept {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef HEPT_SORT_ITEMS
#define HEPT_SORT_ITEMS 16
#endif
#ifndef HEPT_SORT_BLOCKS
#define HEPT_SORT_BLOCKS 3
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
  uint32_t key[kItemsPerThread], peers[kItemsPerThread];
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

  // (2) per-warp digit counts; each warp owns a contiguous run of 512 keys, kept in registers
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int i = wstart + it * 32 + lane;
    const bool ok = i < n;
    const uint32_t dg = (key[it] >> shift) & (kRadix - 1);
    peers[it] = same_digit_lanes(dg, ok);
    if (ok && lane == __ffs(peers[it]) - 1) warp_off[warp][dg] += __popc(peers[it]);
    __syncwarp();
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
  // (4) rank in the same order and park every key at its tile-local slot
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    int i = wstart + it * 32 + lane;
    bool ok = i < n;
    uint32_t dg = (key[it] >> shift) & (kRadix - 1);
    uint32_t rank = __popc(peers[it] & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = warp_off[warp][dg] + rank;
    __syncwarp();
    if (ok && lane == __ffs(peers[it]) - 1) warp_off[warp][dg] += __popc(peers[it]);
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

struct SortPlan {
  int tiles;
  size_t hist_bytes, keys_bytes, idx_bytes, total;
};

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

extern "C" size_t hept_argsort_workspace_bytes(int32_t num_segments, int32_t n) {
  if (num_segments <= 0 || n <= 0) return 0;
  return plan_sort(num_segments, n).total;
}

extern "C" int hept_segmented_argsort(const float* keys, int32_t num_segments, int32_t n, int32_t* positions,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(keys && positions && workspace && num_segments > 0 && n > 0, HEPT_EINVAL,
               "segmented_argsort: bad argument");
  SortPlan p = plan_sort(num_segments, n);
  HEPT_REQUIRE(workspace_bytes >= p.total, HEPT_EWORKSPACE, "segmented_argsort: workspace needs %zu bytes, got %zu",
               p.total, workspace_bytes);
  HEPT_REQUIRE(num_segments <= 65535, HEPT_EINVAL, "segmented_argsort: too many segments (%d)", num_segments);
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  uint32_t* hist = (uint32_t*)w;            w += p.hist_bytes;
  uint32_t* kbuf[2] = {(uint32_t*)w, (uint32_t*)(w + p.keys_bytes)};  w += 2 * p.keys_bytes;
  int32_t* ibuf[2] = {(int32_t*)w, (int32_t*)(w + p.idx_bytes)};
  dim3 grid(p.tiles, num_segments);
  const int passes = 32 / kRadixBits;
  const void* kin = keys;
  const int32_t* iin = nullptr;
  for (int pass = 0; pass < passes; ++pass) {
    const bool first = pass == 0, last = pass == passes - 1;
    radix_hist_kernel<<<grid, kSortThreads, 0, st>>>(kin, first, n, p.tiles, pass * kRadixBits, hist);
    HEPT_CHECK_LAUNCH("radix_hist");
    uint32_t* kout = last ? nullptr : kbuf[pass & 1];
    int32_t* iout = last ? positions : ibuf[pass & 1];
    radix_scatter_kernel<<<grid, kSortThreads, 0, st>>>(kin, first, iin, n, p.tiles, pass * kRadixBits, hist, kout,
                                                        iout);
    HEPT_CHECK_LAUNCH("radix_scatter");
    kin = kout;
    iin = iout;
  }
  return HEPT_OK;
}
