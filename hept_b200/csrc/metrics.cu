// SURVEY.md 8(f)-4, second half: the kNN metrics of the tracking task, `acc_and_pr_at_k` + `calc_scores`
// (src/utils/metrics.py:23-93), on the device.  The reference builds the dense (queries x N) distance matrix with cdist,
// takes topk(K + 1) per row, copies the indices to the host and scores them in a numba loop; at 60 000 hits that matrix is
// 14 GB.  Here a CTA of 128 queries streams one slice of the candidates through shared memory in tiles; every thread keeps the
// K + 1 nearest of its query in that slice (ties: lower index first), a second kernel merges the slices and scores:
//   neighbours = the K nearest after dropping the first (the query itself)               metrics.py:76
//   k = cluster size - 1 (points with k == 0 are skipped)                               metrics.py:51,72-73
//   accuracy = matches among the first k / k;  precision = matches / K;  recall = matches / k      metrics.py:84-86
// and the means over the scored queries come from a fixed-order reduction (deterministic).
#include "common.cuh"

namespace hept {

int order_by_cluster_id(const int64_t* cid, int N, uint32_t* keys, int32_t* pos1, int32_t* pos2, int32_t* order, void* sort_ws,
                        size_t sort_bytes, cudaStream_t st);

constexpr int kKnnThreads = 128, kKnnTile = 256, kKnnDim = 16, kKnnMaxK = 32;

// size_of[point] = number of points with the same cluster id (points in cluster-id order)
__global__ void __launch_bounds__(256) cluster_size_kernel(const int32_t* __restrict__ order, const int64_t* __restrict__ cid, int N,
                                                           int32_t* __restrict__ size_of) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= N) return;
  const int64_t c = cid[order[r]];
  if (r == 0 || cid[order[r - 1]] != c) {
    int e = r;
    while (e < N && cid[order[e]] == c) ++e;
    for (int u = r; u < e; ++u) size_of[order[u]] = e - r;
  }
}

__global__ void __launch_bounds__(256) inv_norm_kernel(const float* __restrict__ x, int d, int N, float* __restrict__ inv) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int j = 0; j < d; ++j) { const float v = x[(size_t)i * d + j]; s = fmaf(v, v, s); }
  inv[i] = 1.f / fmaxf(sqrtf(s), 1e-8f);
}

// Scan: a CTA = 128 queries (one per thread) x one of kKnnSlices slices of the candidate range (blockIdx.y); the candidates
// stream through shared memory in tiles shared by the 128 queries.  One slice per query would leave the GPU with eight warps
// per SM: the slices multiply the CTAs, not the tile loads.  Every (query, slice) leaves its K + 1 nearest, sorted by
// (distance, index), in `lists`; knn_merge_kernel merges the slices and scores.
constexpr int kKnnSlices = 8;
__global__ void __launch_bounds__(kKnnThreads, 6) knn_scan_kernel(const float* __restrict__ x, int d, int N, const int64_t* __restrict__ queries,
                                                                  int M, const float* __restrict__ inv_norm, int cosine, int K,
                                                                  float* __restrict__ list_d, int32_t* __restrict__ list_i) {
  __shared__ __align__(16) float tile[kKnnTile * kKnnDim];
  __shared__ float tile_inv[kKnnTile];
  const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
  const bool live = qi < M;
  const int q = live ? (int)queries[qi] : 0;
  const int tiles = (N + kKnnTile - 1) / kKnnTile, per = (tiles + kKnnSlices - 1) / kKnnSlices;
  const int lo = blockIdx.y * per * kKnnTile, hi = min(N, lo + per * kKnnTile);
  float xq[kKnnDim];
#pragma unroll
  for (int j = 0; j < kKnnDim; ++j) xq[j] = (live && j < d) ? __ldg(x + (size_t)q * d + j) : 0.f;
  const float inv_q = cosine ? __ldg(inv_norm + q) : 0.f;
  // The K + 1 nearest so far live in a per-thread buffer: a sorted prefix plus the candidates accepted since the last
  // compaction (anything below the threshold tau = the current (K + 1)-th distance).  Accepting is one predicated store; the
  // buffer is sorted and cut back to K + 1 at WARP-UNIFORM points (when any lane runs short of room), so the lanes never
  // diverge over their individual insertions — a sorted insert per accepted candidate made every lane wait for every other
  // lane's inserts (ncu: 1.7 G local-memory instructions, 38 ms; the scan itself is ~3 ms of work).
  constexpr int BUF = 2 * kKnnMaxK;
  float bd[BUF];
  int bi[BUF];
  const int keep = K + 1;
  int cnt = 0;
  float tau = live ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);     // dead lanes accept nothing
  auto offer = [&](float dist, int idx) {
    if (dist < tau) {                                 // strict: among equal distances the lower index stays in front
      bd[cnt] = dist;
      bi[cnt] = idx;
      ++cnt;
    }
  };
  auto compact = [&]() {                               // every lane of the warp, together: insertion sort by (distance, index)
    const int top = __reduce_max_sync(0xffffffffu, cnt);
    for (int u = 1; u < top; ++u) {
      if (u < cnt) {
        const float dcur = bd[u];
        const int icur = bi[u];
        int w = u - 1;
        while (w >= 0 && (bd[w] > dcur || (bd[w] == dcur && bi[w] > icur))) { bd[w + 1] = bd[w]; bi[w + 1] = bi[w]; --w; }
        bd[w + 1] = dcur;
        bi[w + 1] = icur;
      }
    }
    if (cnt >= keep) { cnt = keep; tau = bd[keep - 1]; }
  };
  for (int c0 = lo; c0 < hi; c0 += kKnnTile) {
    const int ncand = min(kKnnTile, hi - c0);
    __syncthreads();
    for (int u = threadIdx.x; u < kKnnTile * kKnnDim; u += kKnnThreads) {       // columns >= d and rows >= cnt are zero
      const int c = u / kKnnDim, j = u - c * kKnnDim;
      tile[u] = (c < ncand && j < d) ? __ldg(x + (size_t)(c0 + c) * d + j) : 0.f;
    }
    if (cosine) for (int u = threadIdx.x; u < kKnnTile; u += kKnnThreads) tile_inv[u] = u < ncand ? __ldg(inv_norm + c0 + u) : 0.f;
    __syncthreads();
    // four candidates at a time: independent accumulators, 16-byte broadcast reads of the tile
    for (int c = 0; c < ncand; c += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j4 = 0; j4 < kKnnDim / 4; ++j4) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 v = *reinterpret_cast<const float4*>(tile + (c + t) * kKnnDim + 4 * j4);
          if (cosine) {
            acc[t] = fmaf(xq[4 * j4], v.x, acc[t]); acc[t] = fmaf(xq[4 * j4 + 1], v.y, acc[t]);
            acc[t] = fmaf(xq[4 * j4 + 2], v.z, acc[t]); acc[t] = fmaf(xq[4 * j4 + 3], v.w, acc[t]);
          } else {
            float df = xq[4 * j4] - v.x;      acc[t] = fmaf(df, df, acc[t]);
            df = xq[4 * j4 + 1] - v.y;        acc[t] = fmaf(df, df, acc[t]);
            df = xq[4 * j4 + 2] - v.z;        acc[t] = fmaf(df, df, acc[t]);
            df = xq[4 * j4 + 3] - v.w;        acc[t] = fmaf(df, df, acc[t]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (c + t < ncand) offer(cosine ? 1.f - acc[t] * inv_q * tile_inv[c + t] : acc[t], c0 + c + t);
      if (__any_sync(0xffffffffu, cnt > BUF - 4)) compact();          // room for the next four candidates in every lane
    }
  }
  compact();
  if (!live) return;
  const size_t base = ((size_t)qi * kKnnSlices + blockIdx.y) * kKnnMaxK;
  for (int u = 0; u < kKnnMaxK; ++u) {
    const bool have = u < cnt && u < keep;
    list_d[base + u] = have ? bd[u] : __int_as_float(0x7f800000);
    list_i[base + u] = have ? bi[u] : 0x7fffffff;
  }
}

// res (M, 4): accuracy, precision, recall, scored (1 / 0) per query; res_k (M): k of the query (for the reference's K check)
__global__ void __launch_bounds__(kKnnThreads) knn_merge_kernel(const float* __restrict__ list_d, const int32_t* __restrict__ list_i,
                                                                const int64_t* __restrict__ queries, int M, const int64_t* __restrict__ cid,
                                                                const int32_t* __restrict__ size_of, int K, float* __restrict__ res,
                                                                int32_t* __restrict__ res_k) {
  const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
  if (qi >= M) return;
  const int q = (int)queries[qi];
  const int keep = K + 1;
  int bi[kKnnMaxK];
  int head[kKnnSlices];
#pragma unroll
  for (int s2 = 0; s2 < kKnnSlices; ++s2) head[s2] = 0;
  const size_t base = (size_t)qi * kKnnSlices * kKnnMaxK;
  for (int u = 0; u < keep; ++u) {                     // slices cover ascending index ranges: ties resolve to the lower slice
    int best = 0;
    float bdist = list_d[base + head[0]];
    int bidx = list_i[base + head[0]];
#pragma unroll
    for (int s2 = 1; s2 < kKnnSlices; ++s2) {
      const float dd = list_d[base + s2 * kKnnMaxK + head[s2]];
      const int ii = list_i[base + s2 * kKnnMaxK + head[s2]];
      if (dd < bdist || (dd == bdist && ii < bidx)) { best = s2; bdist = dd; bidx = ii; }
    }
#pragma unroll
    for (int s2 = 0; s2 < kKnnSlices; ++s2) head[s2] += (s2 == best);
    bi[u] = bidx;
  }
  const int k = __ldg(size_of + q) - 1;
  res_k[qi] = k;
  float a = 0.f, p = 0.f, r = 0.f, scored = 0.f;
  if (k > 0) {
    const int64_t mine = __ldg(cid + q);
    int m_all = 0, m_k = 0;
    for (int u = 1; u <= K; ++u) {
      const bool match = bi[u] != 0x7fffffff && __ldg(cid + bi[u]) == mine;
      m_all += match;
      if (u <= k) m_k += match;
    }
    a = (float)m_k / (float)k;
    p = (float)m_all / (float)K;
    r = (float)m_all / (float)k;
    scored = 1.f;
  }
  res[(size_t)qi * 4 + 0] = a;
  res[(size_t)qi * 4 + 1] = p;
  res[(size_t)qi * 4 + 2] = r;
  res[(size_t)qi * 4 + 3] = scored;
}

// out[0..2] = means of accuracy / precision / recall over the scored queries, out[3] = their number, out[4] = max k
__global__ void __launch_bounds__(256) knn_reduce_kernel(const float* __restrict__ res, const int32_t* __restrict__ res_k, int M,
                                                         float* __restrict__ out) {
  __shared__ double red[4][256];
  __shared__ int redk[256];
  double s[4] = {0, 0, 0, 0};
  int mk = 0;
  for (int i = threadIdx.x; i < M; i += 256) {
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] += (double)res[(size_t)i * 4 + c];
    mk = max(mk, res_k[i]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) red[c][threadIdx.x] = s[c];
  redk[threadIdx.x] = mk;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int c = 0; c < 4; ++c) red[c][threadIdx.x] += red[c][threadIdx.x + o];
      redk[threadIdx.x] = max(redk[threadIdx.x], redk[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = red[3][0];
    out[0] = (float)(red[0][0] / n);
    out[1] = (float)(red[1][0] / n);
    out[2] = (float)(red[2][0] / n);
    out[3] = (float)n;
    out[4] = (float)redk[0];
  }
}

}  // namespace hept

using namespace hept;

extern "C" size_t hept_knn_metrics_workspace_bytes(int32_t N, int32_t M) {
  if (N <= 0 || M <= 0) return 0;
  const size_t nb = align_up(sizeof(int32_t) * (size_t)N, 256);
  return 6 * nb + align_up(hept_argsort_workspace_bytes(1, N), 256) + align_up(sizeof(float) * 4 * (size_t)M, 256) +
         align_up(sizeof(int32_t) * (size_t)M, 256) + 2 * align_up(sizeof(float) * (size_t)M * kKnnSlices * kKnnMaxK, 256);
}

extern "C" int hept_knn_metrics(const float* x, int32_t N, int32_t d, const int64_t* cluster_ids, const int64_t* queries, int32_t M,
                                int32_t cosine, int32_t K, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(x && cluster_ids && queries && out && workspace, HEPT_EINVAL, "knn_metrics: null pointer");
  HEPT_REQUIRE(N > 0 && M > 0 && d > 0 && d <= kKnnDim && K > 0 && K + 1 <= kKnnMaxK && K + 1 <= N, HEPT_EINVAL,
               "knn_metrics: bad argument (N=%d M=%d d=%d K=%d)", N, M, d, K);
  HEPT_REQUIRE(workspace_bytes >= hept_knn_metrics_workspace_bytes(N, M), HEPT_EWORKSPACE, "knn_metrics: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = align_up(sizeof(int32_t) * (size_t)N, 256);
  char* w = (char*)workspace;
  uint32_t* keys = (uint32_t*)w;     w += nb;
  int32_t* pos1 = (int32_t*)w;       w += nb;
  int32_t* pos2 = (int32_t*)w;       w += nb;
  int32_t* order = (int32_t*)w;      w += nb;
  int32_t* size_of = (int32_t*)w;    w += nb;
  float* inv = (float*)w;            w += nb;
  void* sort_ws = w;                 const size_t sort_bytes = align_up(hept_argsort_workspace_bytes(1, N), 256);
  w += sort_bytes;
  float* res = (float*)w;            w += align_up(sizeof(float) * 4 * (size_t)M, 256);
  int32_t* res_k = (int32_t*)w;      w += align_up(sizeof(int32_t) * (size_t)M, 256);
  float* list_d = (float*)w;         w += align_up(sizeof(float) * (size_t)M * kKnnSlices * kKnnMaxK, 256);
  int32_t* list_i = (int32_t*)w;
  if (int rc = order_by_cluster_id(cluster_ids, N, keys, pos1, pos2, order, sort_ws, sort_bytes, st)) return rc;
  cluster_size_kernel<<<(N + 255) / 256, 256, 0, st>>>(order, cluster_ids, N, size_of);
  HEPT_CHECK_LAUNCH("cluster_size");
  if (cosine) {
    inv_norm_kernel<<<(N + 255) / 256, 256, 0, st>>>(x, d, N, inv);
    HEPT_CHECK_LAUNCH("inv_norm");
  }
  const unsigned qg = (unsigned)((M + kKnnThreads - 1) / kKnnThreads);
  knn_scan_kernel<<<dim3(qg, kKnnSlices), kKnnThreads, 0, st>>>(x, d, N, queries, M, inv, cosine, K, list_d, list_i);
  HEPT_CHECK_LAUNCH("knn_scan");
  knn_merge_kernel<<<qg, kKnnThreads, 0, st>>>(list_d, list_i, queries, M, cluster_ids, size_of, K, res, res_k);
  HEPT_CHECK_LAUNCH("knn_merge");
  knn_reduce_kernel<<<1, 256, 0, st>>>(res, res_k, M, out);
  HEPT_CHECK_LAUNCH("knn_reduce");
  return HEPT_OK;
}
