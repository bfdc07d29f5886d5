// SURVEY.md 8(f)-4, first half: the InfoNCE loss of the tracking task (src/utils/losses.py:8-74, InfoNCELoss with
// dist_metric l2_rbf / l2_inverse / cosine) forward and backward, on the device, deterministic.
//
//   pos[p]  = cid[a] == cid[b]  and  recons[a], recons[b] != 0  and  pts[a], pts[b] > pt_thres      (losses.py:15-18, metrics.py:8-16)
//   s[p]    = sim(x[a], x[b]) / tau;   m = max_p s[p];   e[p] = exp(s[p] - m)                      (losses.py:20-30, 41-42)
//   D[i]    = sum over NEGATIVE pairs with first point i of e[p]                                     (losses.py:47-49)
//   l[p]    = -log(e[p] / (e[p] + D[a]))  for POSITIVE pairs                                        (losses.py:51-52)
//   loss    = mean over labels g (cluster ids that own a positive pair) of mean over g's positive pairs of l[p]   (losses.py:33-37)
//
// The reference groups pairs with torch_scatter after an argsort ("deterministic_scatter").  Here the pairs of a point are
// found through a CSR index built by counting (integer atomics: the counts and offsets are order-free), whose segments are
// then sorted by pair number, so every floating-point sum runs over a fixed order: same bits every run, no floating-point
// atomics.  Points are grouped by cluster id with two chained 32-bit radix sorts (the ids are 64-bit particle ids).
//
// One deliberate difference: the reference indexes the COMPACTED result of its segment sum (one entry per point that owns a
// negative pair) with raw point numbers (losses.py:48-51), which is only meaningful when every point up to the largest
// first index owns a negative pair (true for the radius-graph pairs of the dataset).  D[] here is indexed by point number.
#include "common.cuh"

namespace hept {

int segmented_argsort_launch(const void* keys, int32_t num_segments, int32_t n, int32_t* positions, void* workspace,
                             size_t workspace_bytes, cudaStream_t st, int key_bits = 0);

constexpr int kLossThreads = 256;
constexpr int kLossDim = 16;          // embedding width <= 16 (tracking: h_dim / 2 = 12)
constexpr int kSegSortMax = 1024;     // pairs of one point sorted in shared memory; longer segments: slow path

struct NcePlan {
  size_t s_bytes, flag_bytes, rowptr_bytes, csr_bytes, d_bytes, nlab_bytes, scal_bytes, total;   // the SAVED state
};
static NcePlan plan_nce(int N, long long P) {
  NcePlan p;
  p.s_bytes = align_up(sizeof(float) * (size_t)P, 256);
  p.flag_bytes = align_up((size_t)P, 256);
  p.rowptr_bytes = align_up(sizeof(int32_t) * ((size_t)N + 1), 256);
  p.csr_bytes = align_up(sizeof(int32_t) * (size_t)P, 256);
  p.d_bytes = align_up(sizeof(float) * (size_t)N, 256);
  p.nlab_bytes = align_up(sizeof(int32_t) * (size_t)N, 256);
  p.scal_bytes = 256;                                         // [0] m, [1] number of labels G, [2] loss
  p.total = p.s_bytes + p.flag_bytes + p.rowptr_bytes + p.csr_bytes + p.d_bytes + p.nlab_bytes + p.scal_bytes;
  return p;
}

// similarity of two rows and what the backward needs: returns sim; dist = |xa - xb| (l2 metrics)
__device__ __forceinline__ float pair_sim(const float* __restrict__ x, int d, int a, int b, int metric, float& dist) {
  float dd = 0.f, na = 0.f, nb = 0.f, ab = 0.f;
  for (int j = 0; j < d; ++j) {
    const float xa = __ldg(x + (size_t)a * d + j), xb = __ldg(x + (size_t)b * d + j);
    const float df = xa - xb;
    dd = fmaf(df, df, dd);
    na = fmaf(xa, xa, na);
    nb = fmaf(xb, xb, nb);
    ab = fmaf(xa, xb, ab);
  }
  dist = sqrtf(dd);
  if (metric == 0) return expf(-dist / (2.f * 0.75f * 0.75f));           // l2_rbf, sigma = 0.75 (losses.py:25-26)
  if (metric == 1) return 1.f / (dist + 1.f);                            // l2_inverse (losses.py:28-29)
  return ab / (fmaxf(sqrtf(na), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f));       // cosine, eps = 1e-8 like F.cosine_similarity
}

__global__ void __launch_bounds__(kLossThreads) nce_pair_fwd_kernel(const float* __restrict__ x, int d, const int64_t* __restrict__ pairs,
                                                                    int P, const int64_t* __restrict__ cid, const float* __restrict__ recons,
                                                                    const float* __restrict__ pts, float pt_thres, int metric, float tau,
                                                                    float* __restrict__ s, uint8_t* __restrict__ flag,
                                                                    int32_t* __restrict__ count, uint32_t* __restrict__ block_max) {
  __shared__ uint32_t red[kLossThreads / 32];
  const int p = blockIdx.x * kLossThreads + threadIdx.x;
  uint32_t mine = 0u;                                     // ordered_bits of anything real is > 0
  if (p < P) {
    const int a = (int)__ldg(pairs + p), b = (int)__ldg(pairs + (size_t)P + p);
    const bool pos = __ldg(cid + a) == __ldg(cid + b) && __ldg(recons + a) != 0.f && __ldg(recons + b) != 0.f &&
                     __ldg(pts + a) > pt_thres && __ldg(pts + b) > pt_thres;
    float dist;
    const float sv = __fdiv_rn(pair_sim(x, d, a, b, metric, dist), tau);
    s[p] = sv;
    flag[p] = pos ? 1 : 0;
    atomicAdd(count + a, 1);
    mine = ordered_bits(sv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = red[0];
#pragma unroll
    for (int w = 1; w < kLossThreads / 32; ++w) m = max(m, red[w]);
    block_max[blockIdx.x] = m;
  }
}

// count (N) -> exclusive prefix row_ptr (N + 1); also the maximum over the blocks' maxima -> scal[0].  One CTA.
__global__ void __launch_bounds__(1024) nce_scan_kernel(const int32_t* __restrict__ count, int N, int32_t* __restrict__ row_ptr,
                                                        const uint32_t* __restrict__ block_max, int blocks, float* __restrict__ scal) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  __shared__ uint32_t smax[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + tid;
    const int v = i < N ? count[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int pre = 0;
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    const int c = carry;
    if (i < N) row_ptr[i] = c + pre + inc - v;
    __syncthreads();
    if (tid == 1023) carry = c + pre + inc;
    __syncthreads();
  }
  if (tid == 0) row_ptr[N] = carry;
  if (block_max) {
    uint32_t m = 0u;
    for (int b = tid; b < blocks; b += 1024) m = max(m, block_max[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    if (tid == 0) {
      uint32_t t = smax[0];
      for (int w = 1; w < 32; ++w) t = max(t, smax[w]);
      scal[0] = from_ordered_bits(t);
    }
  }
}

// csr[row_ptr[key] + slot] = pair number; which == 0: keyed by the first point of the pair, 1: by the second
__global__ void __launch_bounds__(kLossThreads) nce_fill_kernel(const int64_t* __restrict__ pairs, int P, int which,
                                                                const int32_t* __restrict__ row_ptr, int32_t* __restrict__ cursor,
                                                                int32_t* __restrict__ csr) {
  const int p = blockIdx.x * kLossThreads + threadIdx.x;
  if (p >= P) return;
  const int key = (int)__ldg(pairs + (size_t)which * P + p);
  csr[__ldg(row_ptr + key) + atomicAdd(cursor + key, 1)] = p;
}
__global__ void __launch_bounds__(kLossThreads) nce_count_kernel(const int64_t* __restrict__ pairs, int P, int which,
                                                                 int32_t* __restrict__ count) {
  const int p = blockIdx.x * kLossThreads + threadIdx.x;
  if (p < P) atomicAdd(count + (int)__ldg(pairs + (size_t)which * P + p), 1);
}

// every segment ascending by pair number (the fill order above is whatever the atomics gave): one warp per point
__global__ void __launch_bounds__(128) nce_segsort_kernel(const int32_t* __restrict__ row_ptr, int32_t* __restrict__ csr, int N) {
  __shared__ int32_t buf[4][kSegSortMax];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= N) return;
  const int lo = row_ptr[i], L = row_ptr[i + 1] - lo;
  if (L <= 1) return;
  if (L > kSegSortMax) {                     // not expected (the dataset caps a point's pairs at 256): plain insertion sort
    if (lane == 0)
      for (int u = 1; u < L; ++u) {
        const int v = csr[lo + u];
        int w = u - 1;
        while (w >= 0 && csr[lo + w] > v) { csr[lo + w + 1] = csr[lo + w]; --w; }
        csr[lo + w + 1] = v;
      }
    return;
  }
  int n = 32;
  while (n < L) n <<= 1;
  int32_t* b = buf[warp];
  for (int u = lane; u < n; u += 32) b[u] = u < L ? csr[lo + u] : 0x7fffffff;
  __syncwarp();
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int u = lane; u < n; u += 32) {
        const int v = u ^ j;
        if (v > u) {
          const int32_t x0 = b[u], x1 = b[v];
          const bool up = (u & k) == 0;
          if ((x0 > x1) == up) { b[u] = x1; b[v] = x0; }
        }
      }
      __syncwarp();
    }
  for (int u = lane; u < L; u += 32) csr[lo + u] = b[u];
}

// per point: D = sum of e over its negative pairs, then L = sum of l over its positive pairs and their count
__global__ void __launch_bounds__(kLossThreads) nce_point_fwd_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ csr,
                                                                     const float* __restrict__ s, const uint8_t* __restrict__ flag,
                                                                     const float* __restrict__ scal, int N, float* __restrict__ D,
                                                                     float* __restrict__ lsum, int32_t* __restrict__ npos) {
  const int i = blockIdx.x * kLossThreads + threadIdx.x;
  if (i >= N) return;
  const float m = scal[0];
  const int lo = row_ptr[i], hi = row_ptr[i + 1];
  float den = 0.f;
  for (int r = lo; r < hi; ++r) {
    const int p = csr[r];
    if (!flag[p]) den += expf(s[p] - m);
  }
  den = fmaxf(den, 0.f);
  float l = 0.f;
  int c = 0;
  for (int r = lo; r < hi; ++r) {
    const int p = csr[r];
    if (flag[p]) {
      const float e = expf(s[p] - m);
      l += -logf(e / (e + den));
      ++c;
    }
  }
  D[i] = den;
  lsum[i] = l;
  npos[i] = c;
}

__global__ void __launch_bounds__(kLossThreads) nce_cid_words_kernel(const int64_t* __restrict__ cid, const int32_t* __restrict__ order,
                                                                     int N, int word, uint32_t* __restrict__ keys) {
  const int r = blockIdx.x * kLossThreads + threadIdx.x;
  if (r >= N) return;
  const uint64_t c = (uint64_t)cid[order ? order[r] : r] ^ 0x8000000000000000ull;      // signed order -> unsigned order
  keys[r] = word ? (uint32_t)(c >> 32) : (uint32_t)c;
}
__global__ void __launch_bounds__(kLossThreads) nce_compose_kernel(const int32_t* __restrict__ first, const int32_t* __restrict__ second,
                                                                   int N, int32_t* __restrict__ out) {
  const int r = blockIdx.x * kLossThreads + threadIdx.x;
  if (r < N) out[r] = first[second[r]];
}

// points in cluster-id order: the first point of every label sums its label's losses and counts; labels with a positive pair
// contribute their mean.  Per-block (sum of means, number of labels) in a fixed order; n_label[point] = pairs of its label.
__global__ void __launch_bounds__(kLossThreads) nce_label_kernel(const int32_t* __restrict__ order, const int64_t* __restrict__ cid,
                                                                 const float* __restrict__ lsum, const int32_t* __restrict__ npos, int N,
                                                                 int32_t* __restrict__ n_label, float* __restrict__ part_sum,
                                                                 int32_t* __restrict__ part_cnt) {
  __shared__ float ssum[kLossThreads];
  __shared__ int scnt[kLossThreads];
  const int r = blockIdx.x * kLossThreads + threadIdx.x;
  float mean = 0.f;
  int has = 0;
  if (r < N) {
    const int64_t c = cid[order[r]];
    if (r == 0 || cid[order[r - 1]] != c) {
      float S = 0.f;
      int n = 0, e = r;
      for (; e < N && cid[order[e]] == c; ++e) { S += lsum[order[e]]; n += npos[order[e]]; }
      for (int u = r; u < e; ++u) n_label[order[u]] = n;
      if (n > 0) { mean = S / (float)n; has = 1; }
    }
  }
  ssum[threadIdx.x] = mean;
  scnt[threadIdx.x] = has;
  __syncthreads();
  for (int o = kLossThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_sum[blockIdx.x] = ssum[0]; part_cnt[blockIdx.x] = scnt[0]; }
}
__global__ void __launch_bounds__(kLossThreads) nce_final_kernel(const float* __restrict__ part_sum, const int32_t* __restrict__ part_cnt,
                                                                 int blocks, float* __restrict__ scal, float* __restrict__ loss) {
  __shared__ float ssum[kLossThreads];
  __shared__ int scnt[kLossThreads];
  float S = 0.f;
  int c = 0;
  for (int b = threadIdx.x; b < blocks; b += kLossThreads) { S += part_sum[b]; c += part_cnt[b]; }
  ssum[threadIdx.x] = S;
  scnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = kLossThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = ssum[0] / (float)scnt[0];      // no positive pair at all: nan, like torch.mean of an empty tensor
    scal[1] = (float)scnt[0];
    scal[2] = l;
    *loss = l;
  }
}

// ---- backward -----------------------------------------------------------------------------------------------------
// d loss / d s[p]: positive pair p of point a: w (e / (e + D) - 1), w = gout / (G n_label[a]); negative pair q of point a:
// e[q] * R[a], R[a] = sum over a's positive pairs of w / (e + D).
__global__ void __launch_bounds__(kLossThreads) nce_point_bwd_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ csr,
                                                                     const float* __restrict__ s, const uint8_t* __restrict__ flag,
                                                                     const float* __restrict__ scal, const float* __restrict__ D,
                                                                     const int32_t* __restrict__ n_label, const float* __restrict__ gout,
                                                                     int N, float* __restrict__ gs) {
  const int i = blockIdx.x * kLossThreads + threadIdx.x;
  if (i >= N) return;
  const float m = scal[0], G = scal[1], den = D[i];
  const int lo = row_ptr[i], hi = row_ptr[i + 1];
  const int nl = n_label[i];
  const float w = nl > 0 ? __ldg(gout) / (G * (float)nl) : 0.f;
  float R = 0.f;
  for (int r = lo; r < hi; ++r) {
    const int p = csr[r];
    if (flag[p]) {
      const float e = expf(s[p] - m), t = 1.f / (e + den);
      R = fmaf(w, t, R);
      gs[p] = w * (e * t - 1.f);
    }
  }
  for (int r = lo; r < hi; ++r) {
    const int p = csr[r];
    if (!flag[p]) gs[p] = expf(s[p] - m) * R;
  }
}

// d loss / d x[i] = sum over the pairs with i first of gs[p] ds/dxa + sum over the pairs with i second of gs[p] ds/dxb
__global__ void __launch_bounds__(kLossThreads) nce_grad_x_kernel(const float* __restrict__ x, int d, const int64_t* __restrict__ pairs,
                                                                  int P, const int32_t* __restrict__ rp_a, const int32_t* __restrict__ csr_a,
                                                                  const int32_t* __restrict__ rp_b, const int32_t* __restrict__ csr_b,
                                                                  const float* __restrict__ gs, int metric, float tau, int N,
                                                                  float* __restrict__ dx) {
  const int i = blockIdx.x * kLossThreads + threadIdx.x;
  if (i >= N) return;
  float xi[kLossDim], acc[kLossDim];
  float ni = 0.f;
#pragma unroll
  for (int j = 0; j < kLossDim; ++j) {
    xi[j] = j < d ? __ldg(x + (size_t)i * d + j) : 0.f;
    acc[j] = 0.f;
    ni = fmaf(xi[j], xi[j], ni);
  }
  ni = fmaxf(sqrtf(ni), 1e-8f);
  for (int side = 0; side < 2; ++side) {
    const int32_t* rp = side ? rp_b : rp_a;
    const int32_t* csr = side ? csr_b : csr_a;
    for (int r = rp[i]; r < rp[i + 1]; ++r) {
      const int p = csr[r];
      const int o = (int)__ldg(pairs + (size_t)(side ? 0 : P) + p);       // the other point of the pair
      float xo[kLossDim], dd = 0.f, no = 0.f, io = 0.f;
#pragma unroll
      for (int j = 0; j < kLossDim; ++j) {
        xo[j] = j < d ? __ldg(x + (size_t)o * d + j) : 0.f;
        const float df = xi[j] - xo[j];
        dd = fmaf(df, df, dd);
        no = fmaf(xo[j], xo[j], no);
        io = fmaf(xi[j], xo[j], io);
      }
      const float g = gs[p] / tau;                                         // d loss / d sim
      if (metric == 2) {
        no = fmaxf(sqrtf(no), 1e-8f);
        const float c = io / (ni * no);
#pragma unroll
        for (int j = 0; j < kLossDim; ++j) acc[j] = fmaf(g, xo[j] / (ni * no) - c * xi[j] / (ni * ni), acc[j]);
      } else {
        const float dist = sqrtf(dd);
        float k;                                                           // d sim / d dist
        if (metric == 0) k = -expf(-dist / (2.f * 0.75f * 0.75f)) / (2.f * 0.75f * 0.75f);
        else k = -1.f / ((dist + 1.f) * (dist + 1.f));
        const float f = dist > 0.f ? g * k / dist : 0.f;                   // d dist / d x_i = (x_i - x_o) / dist for either side
#pragma unroll
        for (int j = 0; j < kLossDim; ++j) acc[j] = fmaf(f, xi[j] - xo[j], acc[j]);
      }
    }
  }
  for (int j = 0; j < d; ++j) dx[(size_t)i * d + j] = acc[j];
}

// points in ascending cluster-id order (64-bit ids): stable radix sort by the low word, then by the high word.
// scratch: keys, pos1, pos2 (N words each) + sort workspace.  Shared with metrics.cu.
int order_by_cluster_id(const int64_t* cid, int N, uint32_t* keys, int32_t* pos1, int32_t* pos2, int32_t* order, void* sort_ws,
                        size_t sort_bytes, cudaStream_t st) {
  const unsigned ng = (unsigned)((N + kLossThreads - 1) / kLossThreads);
  nce_cid_words_kernel<<<ng, kLossThreads, 0, st>>>(cid, nullptr, N, 0, keys);
  HEPT_CHECK_LAUNCH("cid_words");
  if (int rc = segmented_argsort_launch(keys, 1, N, pos1, sort_ws, sort_bytes, st, 32)) return rc;
  nce_cid_words_kernel<<<ng, kLossThreads, 0, st>>>(cid, pos1, N, 1, keys);
  HEPT_CHECK_LAUNCH("cid_words");
  if (int rc = segmented_argsort_launch(keys, 1, N, pos2, sort_ws, sort_bytes, st, 32)) return rc;
  nce_compose_kernel<<<ng, kLossThreads, 0, st>>>(pos1, pos2, N, order);
  HEPT_CHECK_LAUNCH("cid_compose");
  return HEPT_OK;
}

static size_t nce_scratch_bytes(int N, long long P, bool backward) {
  const size_t blocks = (size_t)((P + kLossThreads - 1) / kLossThreads);
  size_t b = align_up(sizeof(int32_t) * (size_t)N, 256) * 2 + align_up(sizeof(uint32_t) * blocks, 256);   // count, cursor, block maxima
  b += align_up(sizeof(float) * (size_t)N, 256) + align_up(sizeof(int32_t) * (size_t)N, 256);             // lsum, npos
  b += 5 * align_up(sizeof(int32_t) * (size_t)N, 256) + align_up(hept_argsort_workspace_bytes(1, N), 256);  // cid sort
  b += 2 * align_up(sizeof(float) * (((size_t)N + kLossThreads - 1) / kLossThreads), 256);                // label partials
  if (backward) b += align_up(sizeof(float) * (size_t)P, 256) + align_up(sizeof(int32_t) * ((size_t)N + 1), 256) + align_up(sizeof(int32_t) * (size_t)P, 256);
  return b;
}

// CSR index of the pairs by their first (which = 0) or second (1) point, segments ascending by pair number
static int build_csr(const int64_t* pairs, int P, int N, int which, int32_t* count, int32_t* cursor, int32_t* row_ptr, int32_t* csr,
                     bool counted, cudaStream_t st) {
  const unsigned pg = (unsigned)((P + kLossThreads - 1) / kLossThreads);
  if (!counted) {
    HEPT_REQUIRE(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)N, st) == cudaSuccess, HEPT_ECUDA, "infonce: memset failed");
    nce_count_kernel<<<pg, kLossThreads, 0, st>>>(pairs, P, which, count);
    HEPT_CHECK_LAUNCH("nce_count");
    nce_scan_kernel<<<1, 1024, 0, st>>>(count, N, row_ptr, nullptr, 0, nullptr);
    HEPT_CHECK_LAUNCH("nce_scan");
  }
  HEPT_REQUIRE(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)N, st) == cudaSuccess, HEPT_ECUDA, "infonce: memset failed");
  nce_fill_kernel<<<pg, kLossThreads, 0, st>>>(pairs, P, which, row_ptr, cursor, csr);
  HEPT_CHECK_LAUNCH("nce_fill");
  nce_segsort_kernel<<<(N + 3) / 4, 128, 0, st>>>(row_ptr, csr, N);
  HEPT_CHECK_LAUNCH("nce_segsort");
  return HEPT_OK;
}

}  // namespace hept

using namespace hept;

extern "C" size_t hept_infonce_saved_bytes(int32_t N, int64_t P) { return N > 0 && P > 0 ? plan_nce(N, P).total : 0; }
extern "C" size_t hept_infonce_workspace_bytes(int32_t N, int64_t P, int32_t backward) {
  return N > 0 && P > 0 ? nce_scratch_bytes(N, P, backward != 0) : 0;
}

extern "C" int hept_infonce_fwd(const float* x, int32_t N, int32_t d, const int64_t* point_pairs, int64_t P, const int64_t* cluster_ids,
                                const float* recons, const float* pts, float pt_thres, int32_t metric, float tau, float* loss,
                                void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(x && point_pairs && cluster_ids && recons && pts && loss && saved && workspace, HEPT_EINVAL, "infonce_fwd: null pointer");
  HEPT_REQUIRE(N > 0 && P > 0 && P < (1ll << 31) && d > 0 && d <= kLossDim && metric >= 0 && metric <= 2 && tau > 0.f, HEPT_EINVAL,
               "infonce_fwd: bad argument (N=%d P=%lld d=%d metric=%d)", N, (long long)P, d, metric);
  NcePlan pl = plan_nce(N, P);
  HEPT_REQUIRE(saved_bytes >= pl.total && workspace_bytes >= nce_scratch_bytes(N, P, false), HEPT_EWORKSPACE, "infonce_fwd: buffers too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* sv = (char*)saved;
  float* s = (float*)sv;                 sv += pl.s_bytes;
  uint8_t* flag = (uint8_t*)sv;          sv += pl.flag_bytes;
  int32_t* row_ptr = (int32_t*)sv;       sv += pl.rowptr_bytes;
  int32_t* csr = (int32_t*)sv;           sv += pl.csr_bytes;
  float* D = (float*)sv;                 sv += pl.d_bytes;
  int32_t* n_label = (int32_t*)sv;       sv += pl.nlab_bytes;
  float* scal = (float*)sv;
  const int Pi = (int)P;
  const unsigned pg = (unsigned)((Pi + kLossThreads - 1) / kLossThreads), ng = (unsigned)((N + kLossThreads - 1) / kLossThreads);
  const size_t nb = align_up(sizeof(int32_t) * (size_t)N, 256);
  char* w = (char*)workspace;
  int32_t* count = (int32_t*)w;          w += nb;
  int32_t* cursor = (int32_t*)w;         w += nb;
  uint32_t* block_max = (uint32_t*)w;    w += align_up(sizeof(uint32_t) * pg, 256);
  float* lsum = (float*)w;               w += align_up(sizeof(float) * (size_t)N, 256);
  int32_t* npos = (int32_t*)w;           w += nb;
  uint32_t* keys = (uint32_t*)w;         w += nb;
  int32_t* pos1 = (int32_t*)w;           w += nb;
  int32_t* pos2 = (int32_t*)w;           w += nb;
  int32_t* order = (int32_t*)w;          w += nb;
  w += nb;                                                   // (spare: keeps the layout of nce_scratch_bytes)
  void* sort_ws = w;                     const size_t sort_bytes = align_up(hept_argsort_workspace_bytes(1, N), 256);
  w += sort_bytes;
  float* part_sum = (float*)w;           w += align_up(sizeof(float) * ng, 256);
  int32_t* part_cnt = (int32_t*)w;
  HEPT_REQUIRE(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)N, st) == cudaSuccess, HEPT_ECUDA, "infonce_fwd: memset failed");
  nce_pair_fwd_kernel<<<pg, kLossThreads, 0, st>>>(x, d, point_pairs, Pi, cluster_ids, recons, pts, pt_thres, metric, tau, s, flag, count,
                                                   block_max);
  HEPT_CHECK_LAUNCH("nce_pair_fwd");
  nce_scan_kernel<<<1, 1024, 0, st>>>(count, N, row_ptr, block_max, (int)pg, scal);
  HEPT_CHECK_LAUNCH("nce_scan");
  if (int rc = build_csr(point_pairs, Pi, N, 0, count, cursor, row_ptr, csr, true, st)) return rc;
  nce_point_fwd_kernel<<<ng, kLossThreads, 0, st>>>(row_ptr, csr, s, flag, scal, N, D, lsum, npos);
  HEPT_CHECK_LAUNCH("nce_point_fwd");
  if (int rc = order_by_cluster_id(cluster_ids, N, keys, pos1, pos2, order, sort_ws, sort_bytes, st)) return rc;
  nce_label_kernel<<<ng, kLossThreads, 0, st>>>(order, cluster_ids, lsum, npos, N, n_label, part_sum, part_cnt);
  HEPT_CHECK_LAUNCH("nce_label");
  nce_final_kernel<<<1, kLossThreads, 0, st>>>(part_sum, part_cnt, (int)ng, scal, loss);
  HEPT_CHECK_LAUNCH("nce_final");
  return HEPT_OK;
}

extern "C" int hept_infonce_bwd(const float* x, int32_t N, int32_t d, const int64_t* point_pairs, int64_t P, int32_t metric, float tau,
                                const float* grad_loss, const void* saved, size_t saved_bytes, float* dx, void* workspace,
                                size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(x && point_pairs && grad_loss && saved && dx && workspace, HEPT_EINVAL, "infonce_bwd: null pointer");
  HEPT_REQUIRE(N > 0 && P > 0 && P < (1ll << 31) && d > 0 && d <= kLossDim && metric >= 0 && metric <= 2, HEPT_EINVAL, "infonce_bwd: bad argument");
  NcePlan pl = plan_nce(N, P);
  HEPT_REQUIRE(saved_bytes >= pl.total && workspace_bytes >= nce_scratch_bytes(N, P, true), HEPT_EWORKSPACE, "infonce_bwd: buffers too small");
  cudaStream_t st = (cudaStream_t)stream;
  const char* sv = (const char*)saved;
  const float* s = (const float*)sv;               sv += pl.s_bytes;
  const uint8_t* flag = (const uint8_t*)sv;        sv += pl.flag_bytes;
  const int32_t* row_ptr = (const int32_t*)sv;     sv += pl.rowptr_bytes;
  const int32_t* csr = (const int32_t*)sv;         sv += pl.csr_bytes;
  const float* D = (const float*)sv;               sv += pl.d_bytes;
  const int32_t* n_label = (const int32_t*)sv;     sv += pl.nlab_bytes;
  const float* scal = (const float*)sv;
  const int Pi = (int)P;
  const unsigned ng = (unsigned)((N + kLossThreads - 1) / kLossThreads);
  const size_t nb = align_up(sizeof(int32_t) * (size_t)N, 256);
  char* w = (char*)workspace;
  int32_t* count = (int32_t*)w;          w += nb;
  int32_t* cursor = (int32_t*)w;         w += nb;
  float* gs = (float*)w;                 w += align_up(sizeof(float) * (size_t)P, 256);
  int32_t* rp_b = (int32_t*)w;           w += align_up(sizeof(int32_t) * ((size_t)N + 1), 256);
  int32_t* csr_b = (int32_t*)w;
  nce_point_bwd_kernel<<<ng, kLossThreads, 0, st>>>(row_ptr, csr, s, flag, scal, D, n_label, grad_loss, N, gs);
  HEPT_CHECK_LAUNCH("nce_point_bwd");
  if (int rc = build_csr(point_pairs, Pi, N, 1, count, cursor, rp_b, csr_b, false, st)) return rc;
  nce_grad_x_kernel<<<ng, kLossThreads, 0, st>>>(x, d, point_pairs, Pi, row_ptr, csr, rp_b, csr_b, gs, metric, tau, N, dx);
  HEPT_CHECK_LAUNCH("nce_grad_x");
  return HEPT_OK;
}
