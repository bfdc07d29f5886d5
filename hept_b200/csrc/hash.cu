// a3-a6: coordinate scale, E2LSH projection with fused min/max, AND-construction of the sort keys.
// HBM-bound streaming kernels (reference: example/hept.py:21-28,61-65; example/hept_utils.py:45-47,64-71;
// src/models/attention/hept.py:46-56,93-101).
#include "common.cuh"

namespace hept {

// ---------------------------------------------------------------------------------------------------
// coordinate scale: one small CTA; H*R*K sums of D terms.
// ---------------------------------------------------------------------------------------------------
// thread <-> (h, r, k): wbar = sum_d w[h*D+d, r*K+k] (D independent loads), e = exp(min(wbar, 50)); then thread <-> (h, r)
// adds the K terms in order.  One CTA of H*R*K threads (400 for the tracking model).
__global__ void coord_scale_fwd_kernel(const float* __restrict__ w, int H, int D, int R, int K,
                                       float* __restrict__ scale) {
  extern __shared__ float s_e[];                   // (H*R, K)
  const int n = H * R * K;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int h = idx / (R * K), rk = idx - h * (R * K);
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += __ldg(w + (size_t)(h * D + d) * (R * K) + rk);
    s_e[idx] = expf(fminf(s, 50.f));
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < H * R; idx += blockDim.x) {
    const int h = idx / R, r = idx % R;
    float qw = 0.f;
    for (int kk = 0; kk < K; ++kk) qw += s_e[idx * K + kk];
    const float sc = sqrtf(2.f * qw);
    scale[h * (R + 1) + r + 1] = sc;
    if (r == 0) scale[h * (R + 1)] = sc;  // eta and phi share weight 0 (example/hept.py:23)
  }
}

// thread <-> (h, r, k), any number of CTAs
__global__ void coord_scale_bwd_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                       const float* __restrict__ dscale, int H, int D, int R, int K,
                                       float* __restrict__ dw) {
  // d scale / d qw = 1 / scale ; d qw / d wbar = exp(wbar) * [wbar <= 50] ; d wbar / d w[h,d,r,k] = 1
  const int C = R + 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * R * K) return;
  const int h = idx / (R * K), r = (idx / K) % R, kk = idx % K;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += __ldg(w + (size_t)(h * D + d) * (R * K) + r * K + kk);
  float dqw = dscale[h * C + r + 1] / scale[h * C + r + 1];
  if (r == 0) dqw += dscale[h * C] / scale[h * C];
  const float g = (s <= 50.f) ? dqw * expf(s) : 0.f;
  for (int d = 0; d < D; ++d) dw[(size_t)(h * D + d) * (R * K) + r * K + kk] = g;
}

// ---------------------------------------------------------------------------------------------------
// projection + min/max.  One warp = 32 consecutive hits of one head; a CTA covers all heads of its
// hits, so every 32-byte sector of the q/k rows it touches is consumed inside the CTA.
// proj layout (2, T, H, N); extrema kept as order-preserving uint32 (atomicMin / atomicMax).
// ---------------------------------------------------------------------------------------------------
__global__ void init_extrema_kernel(uint32_t* __restrict__ ext, int th) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < th) {
    ext[2 * i + 0] = 0xffffffffu;  // running min
    ext[2 * i + 1] = 0u;           // running max
  }
}

// One CTA = 32 consecutive hits x all heads.  The q and k rows of those hits are 32 x H*D contiguous floats each: they
// are copied into shared memory with fully coalesced 16-byte loads (every sector fetched is used, adjacent lanes read
// adjacent addresses), then one thread per (hit, head) reads its own D-float slices back (row stride padded by 4 floats:
// the 32 lanes of a warp — 32 hits of one head — hit distinct bank groups) and runs the sequential FMA chains.
template <int D, int C>
__global__ void __launch_bounds__(256) hash_project_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                           const float* __restrict__ coords,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ alpha, int N, int H, int T,
                                                           int raw_size, float* __restrict__ proj,
                                                           uint32_t* __restrict__ ext) {
  constexpr int E = D + C;
  extern __shared__ float s_dyn[];
  const int HD = H * D, stride = HD + 4;
  float* s_alpha = s_dyn;                         // (H, E, T)
  float* s_scale = s_alpha + H * E * T;           // (H, C)
  float* s_q = s_scale + ((H * C + 3) & ~3);      // 32 rows x stride
  float* s_k = s_q + 32 * stride;
  for (int i = threadIdx.x; i < H * E * T; i += blockDim.x) s_alpha[i] = alpha[i];
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) s_scale[i] = scale[i];
  const int n0 = blockIdx.x * 32;
  const int rows = min(32, N - n0);
  const int vec_per_row = HD / 4;
  for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c4 = i - r * vec_per_row;
    const size_t g = ((size_t)(n0 + r) * HD) / 4 + c4;
    *reinterpret_cast<float4*>(s_q + r * stride + 4 * c4) = __ldg(reinterpret_cast<const float4*>(q) + g);
    *reinterpret_cast<float4*>(s_k + r * stride + 4 * c4) = __ldg(reinterpret_cast<const float4*>(k) + g);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int n = n0 + lane;
  const bool live = n < N;
  const bool real = live && n < raw_size;
  float cc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) cc[c] = 0.f;
  if (real) load_row<C>(coords + (size_t)n * C, cc);
  for (int h = threadIdx.x >> 5; h < H; h += warps) {
    float qa[E], ka[E];
#pragma unroll
    for (int c4 = 0; c4 < D / 4; ++c4) {
      const float4 a = real ? *reinterpret_cast<const float4*>(s_q + lane * stride + h * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b = real ? *reinterpret_cast<const float4*>(s_k + lane * stride + h * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      qa[4 * c4] = a.x; qa[4 * c4 + 1] = a.y; qa[4 * c4 + 2] = a.z; qa[4 * c4 + 3] = a.w;
      ka[4 * c4] = b.x; ka[4 * c4 + 1] = b.y; ka[4 * c4 + 2] = b.z; ka[4 * c4 + 3] = b.w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = real ? __fmul_rn(s_scale[h * C + c], cc[c]) : 0.f;
      qa[D + c] = v;
      ka[D + c] = v;
    }
    for (int t = 0; t < T; ++t) {
      float pq = 0.f, pk = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float a = s_alpha[(h * E + e) * T + t];
        pq = fmaf(qa[e], a, pq);
        pk = fmaf(ka[e], a, pk);
      }
      const size_t th_n = (size_t)(t * H + h) * N;
      if (live) {
        proj[th_n + n] = pq;
        proj[(size_t)T * H * N + th_n + n] = pk;
      }
      uint32_t lo = live ? min(ordered_bits(pq), ordered_bits(pk)) : 0xffffffffu;
      uint32_t hi = live ? max(ordered_bits(pq), ordered_bits(pk)) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == 0) {
        atomicMin(&ext[2 * (t * H + h) + 0], lo);
        atomicMax(&ext[2 * (t * H + h) + 1], hi);
      }
    }
  }
}

// hat_coords[n,h,0:8] = scale[h,c] * coords[n,c] (c < C), zero padded: the coordinate part of q_hat / k_hat, which is
// the same for every table and for the query and the key side, materialised once so the tile gathers read two
// 16-byte chunks instead of C scalars and C multiplies per row.
__global__ void hat_coords_kernel(const float* __restrict__ coords, const float* __restrict__ scale, int N, int H, int C,
                                  int raw_size, float* __restrict__ hat) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (n, h, c8)
  if (i >= (size_t)N * H * 8) return;
  const int c = (int)(i & 7);
  const int h = (int)((i >> 3) % H);
  const size_t n = (i >> 3) / H;
  hat[i] = (c < C && (int)n < raw_size) ? __fmul_rn(scale[h * C + c], coords[n * C + c]) : 0.f;
}


// ---------------------------------------------------------------------------------------------------
// keys.  Explicit round-to-nearest multiply and add: an FMA contraction would not match eager torch.
// ---------------------------------------------------------------------------------------------------
template <typename ShiftT>     // int64 as the reference's prepare_input delivers them, or the same values as int32
__global__ void keys_packed_kernel(const float* __restrict__ proj, const float* __restrict__ span,
                                   const ShiftT* __restrict__ shifts, int N, size_t thn, float* __restrict__ keys) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= thn) return;
  float sp = span[i / N];
  float sh = __fmul_rn((float)shifts[i], sp);  // integer -> f32 (rne), then one rounding for the product
  keys[i] = __fadd_rn(proj[i], sh);
  keys[thn + i] = __fadd_rn(proj[thn + i], sh);
}

__global__ void keys_regions_kernel(const float* __restrict__ proj, const float* __restrict__ span,
                                    const float* __restrict__ eta, const float* __restrict__ phi,
                                    const float* __restrict__ regions_h, int N, int raw_size, size_t thn,
                                    float* __restrict__ keys) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= thn) return;
  const int th = (int)(i / N), n = (int)(i % N);
  const float sp = span[th];
  // (phi * span) * (ceil(regions_h[0]) + 1) + eta * span      src/models/attention/hept.py:50-55
  float mult = __fadd_rn(ceilf(regions_h[th]), 1.f);
  float shift = __fadd_rn(__fmul_rn(__fmul_rn(phi[i], sp), mult), __fmul_rn(eta[i], sp));
  const float inf = __int_as_float(0x7f800000);
  float pq = n < raw_size ? proj[i] : inf;
  float pk = n < raw_size ? proj[thn + i] : inf;
  keys[i] = __fadd_rn(pq, shift);
  keys[thn + i] = __fadd_rn(pk, shift);
}

// Second generation, T known at compile time (T <= 4).
//  * persistent grid: a CTA walks groups of 32 hits and keeps the running min / max of its warps in registers; it
//    writes ONE partial per (table, head) at the end and finish_span reduces the partials.  The first generation issued
//    2 T atomics per warp on 2 T H addresses that share four cache lines: 90 k serialised L2 atomics were the whole
//    75 us of the kernel (a rewrite of the arithmetic alone did not move it).  No atomics: the span is deterministic.
//  * the q / k rows go to shared memory with cp.async (no registers, no issue slots while they fly) and a thread walks
//    its row four elements at a time with every table's accumulator live: a step is 2 + T LDS.128 for 8 T FMAs.
//    alpha is re-laid as (H, EP = 32, T) with zero rows past E, so the T 16-byte chunks of a step are aligned.
// The per-table FMA chain runs over e = 0 .. E-1 in order, exactly like the first generation: same bits.
constexpr int kHashMaxCtas = HEPT_HASH_MAX_CTAS;

template <int D, int C, int T, int HPW>
__global__ void __launch_bounds__(256, HPW == 1 ? 4 : 2) hash_project_v2_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ coords,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ alpha, int N, int H, int raw_size,
                                                              float* __restrict__ proj, uint32_t* __restrict__ partial,
                                                              float* __restrict__ hat) {
  constexpr int E = D + C, EP = 32;
  static_assert(E <= EP && D % 4 == 0 && T <= 4 && C <= 8, "row shapes");
  extern __shared__ __align__(16) float s_dyn[];
  const int HD = H * D, stride = HD + 4;
  float* s_alpha = s_dyn;                         // (H, EP, T), 16-byte aligned chunks of 4 e x T
  float* s_scale = s_alpha + H * EP * T;          // (H, C)
  float* s_q = s_scale + ((H * C + 3) & ~3);      // 32 rows x stride
  float* s_k = s_q + 32 * stride;
  for (int i = threadIdx.x; i < H * EP * T; i += blockDim.x) {
    const int t = i % T, e = (i / T) % EP, h = i / (T * EP);
    s_alpha[i] = e < E ? alpha[(h * E + e) * T + t] : 0.f;
  }
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) s_scale[i] = scale[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const int vec_per_row = HD / 4;
  const int groups = (N + 31) / 32;
  // running extrema of the heads this warp owns (h = warp, warp + warps, ...): lane-private until the end
  // HPW = heads per warp (H <= warps * HPW): 1 for the shipped 8-head shapes, which keeps the kernel at 4 CTAs per SM
  uint32_t lo[HPW][T], hi[HPW][T];
#pragma unroll
  for (int j = 0; j < HPW; ++j)
#pragma unroll
    for (int t = 0; t < T; ++t) { lo[j][t] = 0xffffffffu; hi[j][t] = 0u; }

  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int n0 = g * 32;
    const int rows = min(32, N - n0);
    __syncthreads();                              // the previous group's rows have been consumed
    for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
      const int r = i / vec_per_row, c4 = i - r * vec_per_row;
      const size_t gi = ((size_t)(n0 + r) * HD) / 4 + c4;
      cp_async16_ca(s_q + r * stride + 4 * c4, reinterpret_cast<const float4*>(q) + gi);
      cp_async16_ca(s_k + r * stride + 4 * c4, reinterpret_cast<const float4*>(k) + gi);
    }
    const int n = n0 + lane;
    const bool live = n < N;
    const bool real = live && n < raw_size;
    float cc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cc[c] = 0.f;
    if (real) {
#pragma unroll
      for (int c = 0; c < C; ++c) cc[c] = __ldg(coords + (size_t)n * C + c);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

#pragma unroll
    for (int j = 0; j < HPW; ++j) {
      const int h = warp + j * warps;
      if (h >= H) break;
      float pq[T], pk[T];
#pragma unroll
      for (int t = 0; t < T; ++t) { pq[t] = 0.f; pk[t] = 0.f; }
      const float* al = s_alpha + h * EP * T;
#pragma unroll
      for (int c4 = 0; c4 < EP / 4; ++c4) {
        if (4 * c4 >= E) break;
        float xq[4], xk[4];
        if (4 * c4 < D) {
          const float4 a = real ? *reinterpret_cast<const float4*>(s_q + lane * stride + h * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 b = real ? *reinterpret_cast<const float4*>(s_k + lane * stride + h * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          xq[0] = a.x; xq[1] = a.y; xq[2] = a.z; xq[3] = a.w;
          xk[0] = b.x; xk[1] = b.y; xk[2] = b.z; xk[3] = b.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = 4 * c4 + u - D;
            const float v = (c < C && real) ? __fmul_rn(s_scale[h * C + (c < C ? c : 0)], cc[c < C ? c : 0]) : 0.f;
            xq[u] = v; xk[u] = v;
          }
          // by-product: the scaled coordinates are the hat_coords rows the tile kernels gather (saves hat_coords_kernel)
          if (hat != nullptr && live && 4 * c4 - D < 8)
            *reinterpret_cast<float4*>(hat + ((size_t)n * H + h) * 8 + (4 * c4 - D)) = make_float4(xq[0], xq[1], xq[2], xq[3]);
        }
        float av[4 * T];                                  // alpha[e = 4 c4 + u][t] at av[u * T + t]
#pragma unroll
        for (int jj = 0; jj < T; ++jj) {
          const float4 a4 = *reinterpret_cast<const float4*>(al + 4 * c4 * T + 4 * jj);
          av[4 * jj] = a4.x; av[4 * jj + 1] = a4.y; av[4 * jj + 2] = a4.z; av[4 * jj + 3] = a4.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (4 * c4 + u < E) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
              pq[t] = fmaf(xq[u], av[u * T + t], pq[t]);
              pk[t] = fmaf(xk[u], av[u * T + t], pk[t]);
            }
          }
        }
      }
      if constexpr (C <= 4) {                        // the second hat chunk holds no coordinate: zero it
        if (hat != nullptr && live) *reinterpret_cast<float4*>(hat + ((size_t)n * H + h) * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const size_t th_n = (size_t)(t * H + h) * N;
        if (live) {
          proj[th_n + n] = pq[t];
          proj[(size_t)T * H * N + th_n + n] = pk[t];
          lo[j][t] = min(lo[j][t], min(ordered_bits(pq[t]), ordered_bits(pk[t])));
          hi[j][t] = max(hi[j][t], max(ordered_bits(pq[t]), ordered_bits(pk[t])));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < HPW; ++j) {
    const int h = warp + j * warps;
    if (h >= H) break;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint32_t l = __reduce_min_sync(0xffffffffu, lo[j][t]), m = __reduce_max_sync(0xffffffffu, hi[j][t]);
      if (lane == 0) {
        partial[((size_t)blockIdx.x * T * H + t * H + h) * 2 + 0] = l;
        partial[((size_t)blockIdx.x * T * H + t * H + h) * 2 + 1] = m;
      }
    }
  }
}

// span[th] = max - min over the per-CTA partials (first generation: ctas == 1, the atomically maintained pair).
// One CTA of 128 threads per (table, head).
__global__ void __launch_bounds__(128) finish_span_kernel(const uint32_t* __restrict__ partial, int ctas, int th,
                                                          float* __restrict__ span) {
  __shared__ uint32_t s_lo[4], s_hi[4];
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int b = threadIdx.x; b < ctas; b += 128) {
    lo = min(lo, partial[((size_t)b * th + i) * 2 + 0]);
    hi = max(hi, partial[((size_t)b * th + i) * 2 + 1]);
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    lo = min(min(s_lo[0], s_lo[1]), min(s_lo[2], s_lo[3]));
    hi = max(max(s_hi[0], s_hi[1]), max(s_hi[2], s_hi[3]));
    span[i] = __fsub_rn(from_ordered_bits(hi), from_ordered_bits(lo));
  }
}

template <int D, int C, int T>
static int launch_project_v2(const hept_shape* s, const float* q, const float* k, const float* coords, const float* scale,
                             const float* alpha, float* proj, uint32_t* partial, int* ctas, float* hat, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)s->H * 32 * T + (((size_t)s->H * C + 3) & ~(size_t)3) +
                                       2 * 32 * ((size_t)s->H * D + 4));
  static DeviceOnce configured;
  int per_sm = 0;
  const bool wide = s->H > 8;                     // more than one head per warp
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(hash_project_v2_kernel<D, C, T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(hash_project_v2_kernel<D, C, T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "hash_project: %s", cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "hash_project: cannot read the SM count");
  HEPT_REQUIRE(smem <= 100 * 1024, HEPT_EUNSUPPORTED, "hash_project: H*D=%d too wide for the staging buffer", s->H * D);
  HEPT_REQUIRE(s->H <= 8 * 4, HEPT_EUNSUPPORTED, "hash_project: more than 32 heads");
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hash_project_v2_kernel<D, C, T, 4>, 256, smem)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hash_project_v2_kernel<D, C, T, 1>, 256, smem);
  HEPT_REQUIRE(e == cudaSuccess && per_sm > 0, HEPT_ECUDA, "hash_project: occupancy query failed");
  const int groups = (s->N + 31) / 32;
  int grid = sms * per_sm;                       // one resident wave
  if (grid > groups) grid = groups;
  if (grid > kHashMaxCtas) grid = kHashMaxCtas;
  *ctas = grid;
  if (wide) hash_project_v2_kernel<D, C, T, 4><<<grid, 256, smem, st>>>(q, k, coords, scale, alpha, s->N, s->H, s->raw_size, proj, partial, hat);
  else hash_project_v2_kernel<D, C, T, 1><<<grid, 256, smem, st>>>(q, k, coords, scale, alpha, s->N, s->H, s->raw_size, proj, partial, hat);
  HEPT_CHECK_LAUNCH("hash_project");
  return HEPT_OK;
}

template <int D, int C>
static int launch_project(const hept_shape* s, const float* q, const float* k, const float* coords, const float* scale,
                          const float* alpha, float* proj, uint32_t* ext, cudaStream_t st) {
  const int E = D + C;
  const size_t smem = sizeof(float) * ((size_t)s->H * E * s->T + (((size_t)s->H * C + 3) & ~(size_t)3) +
                                       2 * 32 * ((size_t)s->H * D + 4));
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(hash_project_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "hash_project: %s", cudaGetErrorString(e));
    configured.mark();
  }
  HEPT_REQUIRE(smem <= 100 * 1024, HEPT_EUNSUPPORTED, "hash_project: H*D=%d too wide for the staging buffer", s->H * D);
  dim3 grid((s->N + 31) / 32);
  hash_project_kernel<D, C><<<grid, 256, smem, st>>>(q, k, coords, scale, alpha, s->N, s->H, s->T, s->raw_size, proj, ext);
  HEPT_CHECK_LAUNCH("hash_project");
  return HEPT_OK;
}

}  // namespace hept

using namespace hept;

extern "C" int hept_coord_scale_fwd(const float* w, int32_t H, int32_t D, int32_t R, int32_t K, float* scale,
                                    void* stream) {
  HEPT_REQUIRE(w && scale && H > 0 && D > 0 && R > 0 && K > 0, HEPT_EINVAL, "coord_scale_fwd: bad argument");
  const int items = H * R * K;
  HEPT_REQUIRE(items * sizeof(float) <= 48 * 1024, HEPT_EUNSUPPORTED, "coord_scale_fwd: H*R*K=%d too large", items);
  const int threads = items < 1024 ? (items + 31) / 32 * 32 : 1024;
  coord_scale_fwd_kernel<<<1, threads, items * sizeof(float), (cudaStream_t)stream>>>(w, H, D, R, K, scale);
  HEPT_CHECK_LAUNCH("coord_scale_fwd");
  return HEPT_OK;
}

extern "C" int hept_coord_scale_bwd(const float* w, const float* scale, const float* dscale, int32_t H, int32_t D,
                                    int32_t R, int32_t K, float* dw, void* stream) {
  HEPT_REQUIRE(w && scale && dscale && dw && H > 0 && D > 0 && R > 0 && K > 0, HEPT_EINVAL,
               "coord_scale_bwd: bad argument");
  coord_scale_bwd_kernel<<<(H * R * K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w, scale, dscale, H, D, R, K, dw);
  HEPT_CHECK_LAUNCH("coord_scale_bwd");
  return HEPT_OK;
}

extern "C" size_t hept_hash_workspace_bytes(const hept_shape* s) {
  if (!s || s->T <= 0 || s->H <= 0) return 0;
  return sizeof(uint32_t) * 2 * (size_t)s->T * s->H * kHashMaxCtas;
}

namespace hept {
// hat != nullptr: also emit hat_coords (N, H, 8) when the fused kernel is used; *hat_done tells whether it was
int hash_project_impl(const hept_shape* s, const float* q, const float* k, const float* coords, const float* scale,
                      const float* alpha, float* proj, float* span, void* workspace, size_t workspace_bytes, float* hat,
                      bool* hat_done, void* stream) {
  if (hat_done) *hat_done = false;
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(q && k && coords && scale && alpha && proj && span && workspace, HEPT_EINVAL,
               "hash_project: null pointer");
  const int th = s->T * s->H;
  HEPT_REQUIRE(workspace_bytes >= hept_hash_workspace_bytes(s), HEPT_EWORKSPACE,
               "hash_project: workspace needs %zu bytes", hept_hash_workspace_bytes(s));
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ext = (uint32_t*)workspace;
  int rc, ctas = 1;
  bool fused = true;
  if (s->D == 24 && s->C == 6 && s->T == 3) rc = launch_project_v2<24, 6, 3>(s, q, k, coords, scale, alpha, proj, ext, &ctas, hat, st);
  else if (s->D == 24 && s->C == 4 && s->T == 3) rc = launch_project_v2<24, 4, 3>(s, q, k, coords, scale, alpha, proj, ext, &ctas, hat, st);
  else if (s->D == 8 && s->C == 6 && s->T == 2) rc = launch_project_v2<8, 6, 2>(s, q, k, coords, scale, alpha, proj, ext, &ctas, hat, st);
  else {
    fused = false;
    // any other table count: the first-generation kernel with one atomically maintained (min, max) pair per (table, head)
    init_extrema_kernel<<<(th + 127) / 128, 128, 0, st>>>(ext, th);
    HEPT_CHECK_LAUNCH("init_extrema");
    if (s->D == 24 && s->C == 6) rc = launch_project<24, 6>(s, q, k, coords, scale, alpha, proj, ext, st);
    else if (s->D == 24 && s->C == 4) rc = launch_project<24, 4>(s, q, k, coords, scale, alpha, proj, ext, st);
    else if (s->D == 8 && s->C == 6) rc = launch_project<8, 6>(s, q, k, coords, scale, alpha, proj, ext, st);
    else {
      set_error("hash_project: (D=%d, C=%d) not compiled in", s->D, s->C);
      return HEPT_EUNSUPPORTED;
    }
  }
  if (rc) return rc;
  finish_span_kernel<<<th, 128, 0, st>>>(ext, ctas, th, span);
  HEPT_CHECK_LAUNCH("finish_span");
  if (hat_done) *hat_done = fused && hat != nullptr;
  return HEPT_OK;
}
}  // namespace hept

extern "C" int hept_hash_project(const hept_shape* s, const float* q, const float* k, const float* coords,
                                 const float* scale, const float* alpha, float* proj, float* span, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  return hept::hash_project_impl(s, q, k, coords, scale, alpha, proj, span, workspace, workspace_bytes, nullptr, nullptr, stream);
}

extern "C" int hept_hat_coords(const hept_shape* s, const float* coords, const float* scale, float* hat_coords,
                               void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(coords && scale && hat_coords && s->C <= 8, HEPT_EINVAL, "hat_coords: bad argument");
  const size_t total = (size_t)s->N * s->H * 8;
  hat_coords_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(coords, scale, s->N, s->H, s->C,
                                                                                     s->raw_size, hat_coords);
  HEPT_CHECK_LAUNCH("hat_coords");
  return HEPT_OK;
}

extern "C" int hept_keys_from_packed_shifts(const hept_shape* s, const float* proj, const float* span,
                                            const int64_t* combined_shifts, float* keys, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(proj && span && combined_shifts && keys, HEPT_EINVAL, "keys_from_packed_shifts: null pointer");
  size_t thn = (size_t)s->T * s->H * s->N;
  keys_packed_kernel<int64_t><<<(unsigned)((thn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proj, span, combined_shifts,
                                                                                              s->N, thn, keys);
  HEPT_CHECK_LAUNCH("keys_packed");
  return HEPT_OK;
}

extern "C" int hept_keys_from_packed_shifts32(const hept_shape* s, const float* proj, const float* span,
                                              const int32_t* combined_shifts32, float* keys, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(proj && span && combined_shifts32 && keys, HEPT_EINVAL, "keys_from_packed_shifts32: null pointer");
  size_t thn = (size_t)s->T * s->H * s->N;
  keys_packed_kernel<int32_t><<<(unsigned)((thn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proj, span, combined_shifts32,
                                                                                              s->N, thn, keys);
  HEPT_CHECK_LAUNCH("keys_packed");
  return HEPT_OK;
}

extern "C" int hept_keys_from_region_indices(const hept_shape* s, const float* proj, const float* span,
                                             const float* region_eta, const float* region_phi,
                                             const float* regions_h, float* keys, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(proj && span && region_eta && region_phi && regions_h && keys, HEPT_EINVAL,
               "keys_from_region_indices: null pointer");
  size_t thn = (size_t)s->T * s->H * s->N;
  keys_regions_kernel<<<(unsigned)((thn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      proj, span, region_eta, region_phi, regions_h, s->N, s->raw_size, thn, keys);
  HEPT_CHECK_LAUNCH("keys_regions");
  return HEPT_OK;
}
