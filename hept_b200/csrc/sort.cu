// a7: segmented stable argsort of float32 keys (example/hept.py:67-68), as an LSD radix sort over
// order-preserving uint32 keys: 4 passes of 8 bits, each pass = per-tile digit histogram + stable
// scatter.  All segments are processed by the same launches (grid = tiles x segments), so the 2*T*H
// independent sorts of one attention call fill the machine together.
//
// Stability: inside a tile, keys are ranked warp by warp, 32 at a time in index order, with
// __match_any_sync giving each lane its rank among equal digits; tiles and warps are combined by
// exclusive prefix sums in index order.  Equal keys therefore keep ascending original index, the
// tie-break the oracle pins (torch.argsort(stable=True)).
#include "common.cuh"

namespace hept {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerThread = 16;
constexpr int kTile = kSortThreads * kItemsPerThread;  // 4096 keys per CTA

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
#pragma unroll 4
  for (int it = 0; it < kItemsPerThread; ++it) {
    int i = start + it * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(load_key(keys, base + i, as_float) >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  tile_hist[((size_t)seg * tiles + tile) * kRadix + threadIdx.x] = hist[threadIdx.x];
}

// keys_in/idx_in -> keys_out/idx_out.  idx_in == nullptr means the identity (first pass);
// keys_out == nullptr means the keys are no longer needed (last pass).
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const void* __restrict__ keys_in, bool as_float,
                                                                      const int32_t* __restrict__ idx_in, int n,
                                                                      int tiles, int shift,
                                                                      const uint32_t* __restrict__ tile_hist,
                                                                      uint32_t* __restrict__ keys_out,
                                                                      int32_t* __restrict__ idx_out) {
  __shared__ uint32_t warp_off[kSortWarps][kRadix];
  __shared__ uint32_t scan_tmp[kRadix];
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)seg * n;

  // (1) where does digit `tid` of this tile start in the segment?
  uint32_t before = 0, total = 0;
  {
    const uint32_t* th = tile_hist + (size_t)seg * tiles * kRadix + tid;
    for (int t = 0; t < tiles; ++t) {
      uint32_t c = th[(size_t)t * kRadix];
      total += c;
      if (t < tile) before += c;
    }
  }
  // exclusive scan of `total` over the 256 digits (Hillis-Steele in shared memory)
  scan_tmp[tid] = total;
  __syncthreads();
  for (int off = 1; off < kRadix; off <<= 1) {
    uint32_t v = tid >= off ? scan_tmp[tid - off] : 0u;
    __syncthreads();
    scan_tmp[tid] += v;
    __syncthreads();
  }
  const uint32_t digit_start = scan_tmp[tid] - total + before;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) warp_off[w][tid] = 0;
  __syncthreads();

  // (2) per-warp digit counts; each warp owns a contiguous run of 512 keys, kept in registers
  uint32_t key[kItemsPerThread];
  const int wstart = tile * kTile + warp * (32 * kItemsPerThread);
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    int i = wstart + it * 32 + lane;
    bool ok = i < n;
    key[it] = ok ? load_key(keys_in, base + i, as_float) : 0u;
    uint32_t dg = ok ? ((key[it] >> shift) & (kRadix - 1)) : 0xffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, dg);
    if (ok && lane == __ffs(peers) - 1) warp_off[warp][dg] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // (3) exclusive prefix over warps, seeded with the digit's start
  {
    uint32_t run = digit_start;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = warp_off[w][tid];
      warp_off[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
  // (4) rank and scatter in the same order
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    int i = wstart + it * 32 + lane;
    bool ok = i < n;
    uint32_t dg = ok ? ((key[it] >> shift) & (kRadix - 1)) : 0xffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, dg);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = warp_off[warp][dg] + rank;
    __syncwarp();
    if (ok && lane == __ffs(peers) - 1) warp_off[warp][dg] += __popc(peers);
    __syncwarp();
    if (ok) {
      if (keys_out) keys_out[base + pos] = key[it];
      idx_out[base + pos] = idx_in ? idx_in[base + i] : i;
    }
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
