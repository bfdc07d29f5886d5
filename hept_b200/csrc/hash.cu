// a3-a6: coordinate scale, E2LSH projection with fused min/max, AND-construction of the sort keys.
// HBM-bound streaming kernels (reference: example/hept.py:21-28,61-65; example/hept_utils.py:45-47,64-71;
// src/models/attention/hept.py:46-56,93-101).
#include "common.cuh"

namespace hept {

// ---------------------------------------------------------------------------------------------------
// coordinate scale: one small CTA; H*R*K sums of D terms.
// ---------------------------------------------------------------------------------------------------
__global__ void coord_scale_fwd_kernel(const float* __restrict__ w, int H, int D, int R, int K,
                                       float* __restrict__ scale) {
  // thread <-> (h, r); qw = sum_k exp(min(sum_d w[h*D+d, r*K+k], 50))
  for (int idx = threadIdx.x; idx < H * R; idx += blockDim.x) {
    int h = idx / R, r = idx % R;
    float qw = 0.f;
    for (int kk = 0; kk < K; ++kk) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += w[(size_t)(h * D + d) * (R * K) + r * K + kk];
      qw += expf(fminf(s, 50.f));
    }
    float sc = sqrtf(2.f * qw);
    scale[h * (R + 1) + r + 1] = sc;
    if (r == 0) scale[h * (R + 1)] = sc;  // eta and phi share weight 0 (example/hept.py:23)
  }
}

__global__ void coord_scale_bwd_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                       const float* __restrict__ dscale, int H, int D, int R, int K,
                                       float* __restrict__ dw) {
  // d scale / d qw = 1 / scale ; d qw / d wbar = exp(wbar) * [wbar <= 50] ; d wbar / d w[h,d,r,k] = 1
  const int C = R + 1;
  for (int idx = threadIdx.x; idx < H * R * K; idx += blockDim.x) {
    int h = idx / (R * K), r = (idx / K) % R, kk = idx % K;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += w[(size_t)(h * D + d) * (R * K) + r * K + kk];
    float dqw = dscale[h * C + r + 1] / scale[h * C + r + 1];
    if (r == 0) dqw += dscale[h * C] / scale[h * C];
    float g = (s <= 50.f) ? dqw * expf(s) : 0.f;
    for (int d = 0; d < D; ++d) dw[(size_t)(h * D + d) * (R * K) + r * K + kk] = g;
  }
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

__global__ void finish_span_kernel(const uint32_t* __restrict__ ext, int th, float* __restrict__ span) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < th) span[i] = __fsub_rn(from_ordered_bits(ext[2 * i + 1]), from_ordered_bits(ext[2 * i + 0]));
}

// ---------------------------------------------------------------------------------------------------
// keys.  Explicit round-to-nearest multiply and add: an FMA contraction would not match eager torch.
// ---------------------------------------------------------------------------------------------------
__global__ void keys_packed_kernel(const float* __restrict__ proj, const float* __restrict__ span,
                                   const int64_t* __restrict__ shifts, int N, size_t thn, float* __restrict__ keys) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= thn) return;
  float sp = span[i / N];
  float sh = __fmul_rn((float)shifts[i], sp);  // int64 -> f32 (rne), then one rounding for the product
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

template <int D, int C>
static int launch_project(const hept_shape* s, const float* q, const float* k, const float* coords, const float* scale,
                          const float* alpha, float* proj, uint32_t* ext, cudaStream_t st) {
  const int E = D + C;
  const size_t smem = sizeof(float) * ((size_t)s->H * E * s->T + (((size_t)s->H * C + 3) & ~(size_t)3) +
                                       2 * 32 * ((size_t)s->H * D + 4));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hash_project_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "hash_project: %s", cudaGetErrorString(e));
    configured = true;
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
  coord_scale_fwd_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(w, H, D, R, K, scale);
  HEPT_CHECK_LAUNCH("coord_scale_fwd");
  return HEPT_OK;
}

extern "C" int hept_coord_scale_bwd(const float* w, const float* scale, const float* dscale, int32_t H, int32_t D,
                                    int32_t R, int32_t K, float* dw, void* stream) {
  HEPT_REQUIRE(w && scale && dscale && dw && H > 0 && D > 0 && R > 0 && K > 0, HEPT_EINVAL,
               "coord_scale_bwd: bad argument");
  coord_scale_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w, scale, dscale, H, D, R, K, dw);
  HEPT_CHECK_LAUNCH("coord_scale_bwd");
  return HEPT_OK;
}

extern "C" int hept_hash_project(const hept_shape* s, const float* q, const float* k, const float* coords,
                                 const float* scale, const float* alpha, float* proj, float* span, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(q && k && coords && scale && alpha && proj && span && workspace, HEPT_EINVAL,
               "hash_project: null pointer");
  const int th = s->T * s->H;
  HEPT_REQUIRE(workspace_bytes >= sizeof(uint32_t) * 2 * (size_t)th, HEPT_EWORKSPACE,
               "hash_project: workspace needs %zu bytes", sizeof(uint32_t) * 2 * (size_t)th);
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ext = (uint32_t*)workspace;
  init_extrema_kernel<<<(th + 127) / 128, 128, 0, st>>>(ext, th);
  HEPT_CHECK_LAUNCH("init_extrema");
  int rc;
  if (s->D == 24 && s->C == 6) rc = launch_project<24, 6>(s, q, k, coords, scale, alpha, proj, ext, st);
  else if (s->D == 24 && s->C == 4) rc = launch_project<24, 4>(s, q, k, coords, scale, alpha, proj, ext, st);
  else if (s->D == 8 && s->C == 6) rc = launch_project<8, 6>(s, q, k, coords, scale, alpha, proj, ext, st);
  else {
    set_error("hash_project: (D=%d, C=%d) not compiled in", s->D, s->C);
    return HEPT_EUNSUPPORTED;
  }
  if (rc) return rc;
  finish_span_kernel<<<(th + 127) / 128, 128, 0, st>>>(ext, th, span);
  HEPT_CHECK_LAUNCH("finish_span");
  return HEPT_OK;
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
  keys_packed_kernel<<<(unsigned)((thn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(proj, span, combined_shifts,
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
