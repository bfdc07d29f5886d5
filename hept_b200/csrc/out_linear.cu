// a12, second half: the output projection out = out_pre W^T + b (example/hept.py:80, nn.Linear(H*D, D)) and its
// backward, as streaming fp32 kernels.  The library GEMMs this replaces were 12 % of a tracking-60k step: the
// 24 x 60000 x 192 weight-gradient product alone took 132 us in cuBLAS (one long reduction, poorly split), the
// two 60000 x 192 x 24 products 29 us each.  All three are HBM-sized problems (46 MB of out_pre / d_out_pre):
//   out_linear_fwd        reads out_pre once, W from shared memory                      -> out (N, D)        44 us
//   out_linear_bwd_input  d_out_pre = g W, W from shared memory                         -> d_out_pre (N, H*D) 41 us
// (measured at 60k hits; both are bound by the L1 tag stage: a lane-per-hit LDG.128 / STG.128 touches 32 lines.  Staging
// the rows through shared memory with cp.async is the next step; tile-shape variants HEPT_OL_OUTS / HEPT_OL_HT do not
// move them.)
//   out_linear_bwd_params dW = g^T out_pre, db = sum_n g: per-CTA partials over slabs of hits, then a fixed-order
//                         reduction over CTAs (deterministic, no floating-point atomics)
#include "common.cuh"

namespace hept {

#ifndef HEPT_OL_OUTS
#define HEPT_OL_OUTS 12
#define HEPT_OL_HT 4
#endif
constexpr int kOlRows = 64;        // hits per CTA pass
constexpr int kOlThreads = 256;
constexpr int kOlMaxCtas = HEPT_OUT_LINEAR_MAX_CTAS;

// Both row-wise products put one hit on one lane, so every weight a warp touches is the same address for all 32 lanes:
// a broadcast LDS.128 (one wavefront) feeds four FMAs per lane.  (A first version gave each lane its own weight
// columns: four wavefronts per LDS.128 made it shared-memory bound, 48 / 54 us, slower than the library.)  The rows
// themselves are read / written straight from / to global memory, 16 bytes per lane per instruction; consecutive
// instructions use up the sectors a warp has touched.
constexpr int kOlWarps = 4;

// out[n, :] = b + x[n, :] W^T.  A lane owns HT hits (n, n + 32, ...) and OUTS of the OUT outputs (the outputs of a hit
// are split over OUT / OUTS warps): a step loads HT + OUTS 16-byte values for 4 HT OUTS FMAs.  A (constant-bank
// version, weights through LDCU / uniform registers, was 3x slower: the 18 KB of weights thrash the constant cache.)
template <int OUT, int OUTS, int HT>
__global__ void __launch_bounds__(32 * kOlWarps) out_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                       const float* __restrict__ b, int N, int IN,
                                                                       float* __restrict__ out) {
  static_assert(OUT % OUTS == 0 && OUTS % 4 == 0, "output split");
  constexpr int SPLIT = OUT / OUTS;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                              // (OUT, IN)
  for (int i = threadIdx.x; i < OUT * IN / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x * kOlWarps + warp;   // (group of 32 HT hits, output part)
  const int part = unit % SPLIT;
  const int n0 = (unit / SPLIT) * 32 * HT + lane;
  if (n0 >= N) return;
  const float* xr[HT];
#pragma unroll
  for (int t = 0; t < HT; ++t) xr[t] = x + (size_t)min(n0 + 32 * t, N - 1) * IN;   // rows past N: recomputed, not stored
  const float* wr = s_w + part * OUTS * IN;
  float acc[HT][OUTS];
#pragma unroll
  for (int u = 0; u < OUTS; ++u) {
    const float bu = __ldg(b + part * OUTS + u);
#pragma unroll
    for (int t = 0; t < HT; ++t) acc[t][u] = bu;
  }
  const int vec = IN / 4;
  float4 xn[HT];
#pragma unroll
  for (int t = 0; t < HT; ++t) xn[t] = ldg4(xr[t]);
  for (int c4 = 0; c4 < vec; ++c4) {
    float4 xv[HT];
#pragma unroll
    for (int t = 0; t < HT; ++t) xv[t] = xn[t];
    if (c4 + 1 < vec) {
#pragma unroll
      for (int t = 0; t < HT; ++t) xn[t] = ldg4(xr[t] + 4 * (c4 + 1));   // next step's rows fly under this step's FMAs
    }
#pragma unroll
    for (int u = 0; u < OUTS; ++u) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + u * IN + 4 * c4);
#pragma unroll
      for (int t = 0; t < HT; ++t)
        acc[t][u] = fmaf(xv[t].w, wv.w, fmaf(xv[t].z, wv.z, fmaf(xv[t].y, wv.y, fmaf(xv[t].x, wv.x, acc[t][u]))));
    }
  }
#pragma unroll
  for (int t = 0; t < HT; ++t) {
    if (n0 + 32 * t < N) {
      float* dst = out + (size_t)(n0 + 32 * t) * OUT + part * OUTS;
#pragma unroll
      for (int u4 = 0; u4 < OUTS / 4; ++u4)
        *reinterpret_cast<float4*>(dst + 4 * u4) = make_float4(acc[t][4 * u4], acc[t][4 * u4 + 1], acc[t][4 * u4 + 2], acc[t][4 * u4 + 3]);
    }
  }
}

// dx[n, :] = g[n, :] W.  A lane owns two hits (n, n + 32); the IN columns of a hit are split over SPLIT warps.
template <int OUT, int SPLIT>
__global__ void __launch_bounds__(32 * kOlWarps) out_linear_bwd_input_kernel(const float* __restrict__ g,
                                                                             const float* __restrict__ w, int N, int IN,
                                                                             float* __restrict__ dx) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                              // (OUT, IN)
  for (int i = threadIdx.x; i < OUT * IN / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x * kOlWarps + warp;
  const int part = unit % SPLIT;
  const int n0 = (unit / SPLIT) * 64 + lane;
  if (n0 >= N) return;
  const int n1 = n0 + 32;
  const bool has1 = n1 < N;
  float ga[OUT], gb[OUT];
#pragma unroll
  for (int j4 = 0; j4 < OUT / 4; ++j4) {
    const float4 t = ldg4(g + (size_t)n0 * OUT + 4 * j4);
    const float4 v = has1 ? ldg4(g + (size_t)n1 * OUT + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
    ga[4 * j4] = t.x; ga[4 * j4 + 1] = t.y; ga[4 * j4 + 2] = t.z; ga[4 * j4 + 3] = t.w;
    gb[4 * j4] = v.x; gb[4 * j4 + 1] = v.y; gb[4 * j4 + 2] = v.z; gb[4 * j4 + 3] = v.w;
  }
  const int vec = IN / 4, per = (vec + SPLIT - 1) / SPLIT;
  const int c_end = min(vec, (part + 1) * per);
  float* d0 = dx + (size_t)n0 * IN;
  float* d1 = dx + (size_t)(has1 ? n1 : n0) * IN;
#pragma unroll 2
  for (int c4 = part * per; c4 < c_end; ++c4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(s_w + j * IN + 4 * c4);
      a.x = fmaf(ga[j], wv.x, a.x); a.y = fmaf(ga[j], wv.y, a.y); a.z = fmaf(ga[j], wv.z, a.z); a.w = fmaf(ga[j], wv.w, a.w);
      c.x = fmaf(gb[j], wv.x, c.x); c.y = fmaf(gb[j], wv.y, c.y); c.z = fmaf(gb[j], wv.z, c.z); c.w = fmaf(gb[j], wv.w, c.w);
    }
    *reinterpret_cast<float4*>(d0 + 4 * c4) = a;
    if (has1) *reinterpret_cast<float4*>(d1 + 4 * c4) = c;
  }
}

// Stage 1 of dW / db: CTA b sums the slabs b, b + grid, ... of kOlRows hits.  Threads [0, IN) are two row groups of
// IN / 2 threads: group q takes the hits q, q + 2, ... of a slab and a thread owns columns c and c + IN / 2 of dW (all OUT
// rows of both in registers), so one broadcast LDS.128 of the gradient row feeds 8 FMAs.  Threads [IN, IN + OUT) own db.
// The two groups are added through shared memory in a fixed order.  partial (grid, OUT + 1, IN): row OUT holds db in its
// first OUT entries.
template <int OUT>
__global__ void __launch_bounds__(kOlThreads) out_linear_bwd_params_kernel(const float* __restrict__ g,
                                                                           const float* __restrict__ x, int N, int IN,
                                                                           float* __restrict__ partial) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_g = s_dyn;                              // (kOlRows, OUT)
  float* s_acc = s_g + kOlRows * OUT;              // (OUT, IN): group 1's sums
  const int half = IN / 2;
  const int t = threadIdx.x;
  const int q = t / half, c = t - q * half;        // q >= 2: not a dW thread
  const int slabs = (N + kOlRows - 1) / kOlRows;
  float acc0[OUT], acc1[OUT];
#pragma unroll
  for (int j = 0; j < OUT; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
  float bsum = 0.f;
  for (int slab = blockIdx.x; slab < slabs; slab += gridDim.x) {
    const int n0 = slab * kOlRows;
    const int rows = min(kOlRows, N - n0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * OUT / 4; i += kOlThreads)
      reinterpret_cast<float4*>(s_g)[i] = __ldg(reinterpret_cast<const float4*>(g + (size_t)n0 * OUT) + i);
    __syncthreads();
    if (q < 2) {
      const float* xp = x + (size_t)n0 * IN + c;
      auto row_fma = [&](int r, float xa, float xb) {
#pragma unroll
        for (int j4 = 0; j4 < OUT / 4; ++j4) {
          const float4 gv = *reinterpret_cast<const float4*>(s_g + r * OUT + 4 * j4);
          acc0[4 * j4] = fmaf(gv.x, xa, acc0[4 * j4]);         acc1[4 * j4] = fmaf(gv.x, xb, acc1[4 * j4]);
          acc0[4 * j4 + 1] = fmaf(gv.y, xa, acc0[4 * j4 + 1]); acc1[4 * j4 + 1] = fmaf(gv.y, xb, acc1[4 * j4 + 1]);
          acc0[4 * j4 + 2] = fmaf(gv.z, xa, acc0[4 * j4 + 2]); acc1[4 * j4 + 2] = fmaf(gv.z, xb, acc1[4 * j4 + 2]);
          acc0[4 * j4 + 3] = fmaf(gv.w, xa, acc0[4 * j4 + 3]); acc1[4 * j4 + 3] = fmaf(gv.w, xb, acc1[4 * j4 + 3]);
        }
      };
      int r = q;
      for (; r + 6 < rows; r += 8) {               // four hits (eight loads) in flight per thread
        float xa[4], xb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xa[u] = __ldg(xp + (size_t)(r + 2 * u) * IN);
          xb[u] = __ldg(xp + (size_t)(r + 2 * u) * IN + half);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) row_fma(r + 2 * u, xa[u], xb[u]);
      }
      for (; r < rows; r += 2) row_fma(r, __ldg(xp + (size_t)r * IN), __ldg(xp + (size_t)r * IN + half));
    } else if (t < IN + OUT) {
      for (int r = 0; r < rows; ++r) bsum += s_g[r * OUT + (t - IN)];
    }
  }
  __syncthreads();
  if (q == 1) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) { s_acc[j * IN + c] = acc0[j]; s_acc[j * IN + c + half] = acc1[j]; }
  }
  __syncthreads();
  float* p = partial + (size_t)blockIdx.x * (OUT + 1) * IN;
  if (q == 0) {
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
      p[(size_t)j * IN + c] = acc0[j] + s_acc[j * IN + c];
      p[(size_t)j * IN + c + half] = acc1[j] + s_acc[j * IN + c + half];
    }
  } else if (q >= 2 && t < IN + OUT) {
    p[(size_t)OUT * IN + (t - IN)] = bsum;
  }
}

// Stage 2: fixed-order sum over CTAs.  A CTA of 256 threads = 32 entries of (OUT + 1, IN) x 8 parts; part p sums the
// partials p, p + 8, ... in order, then the eight sums are added in order.  Entries [OUT][OUT..IN) are unused.
__global__ void __launch_bounds__(256) out_linear_reduce_kernel(const float* __restrict__ partial, int ctas, int OUT, int IN,
                                                                float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[8][32];
  const int e = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  const int entries = (OUT + 1) * IN;
  float s = 0.f;
  if (i < entries)
    for (int b = part; b < ctas; b += 8) s += partial[(size_t)b * entries + i];
  red[part][e] = s;
  __syncthreads();
  if (part == 0 && i < entries) {
    float t = red[0][e];
#pragma unroll
    for (int p = 1; p < 8; ++p) t += red[p][e];
    const int j = i / IN, c = i - j * IN;
    if (j < OUT) dw[i] = t;
    else if (c < OUT) db[c] = t;
  }
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) sms = n;
  }
  return sms;
}

template <int OUT, int OUTS, int HT>
static int launch_ol_fwd(const hept_shape* s, const float* x, const float* w, const float* b, float* out, cudaStream_t st) {
  const int IN = s->H * s->D;
  const size_t smem = sizeof(float) * (size_t)OUT * IN;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(out_linear_fwd_kernel<OUT, OUTS, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_fwd: %s", cudaGetErrorString(e));
    configured = true;
  }
  HEPT_REQUIRE(smem <= 96 * 1024 && IN % 4 == 0, HEPT_EUNSUPPORTED, "out_linear_fwd: H*D=%d not supported", IN);
  const int units = (s->N + 32 * HT - 1) / (32 * HT) * (OUT / OUTS);
  out_linear_fwd_kernel<OUT, OUTS, HT><<<(units + kOlWarps - 1) / kOlWarps, 32 * kOlWarps, smem, st>>>(x, w, b, s->N, IN, out);
  HEPT_CHECK_LAUNCH("out_linear_fwd");
  return HEPT_OK;
}

template <int OUT, int SPLIT>
static int launch_ol_bwd(const hept_shape* s, const float* g, const float* w, const float* x, float* dx, float* dw,
                         float* db, float* partial, cudaStream_t st) {
  const int IN = s->H * s->D;
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "out_linear_bwd: cannot read the SM count");
  HEPT_REQUIRE(IN + OUT <= kOlThreads && IN % 4 == 0 && (size_t)(kOlRows + IN) * OUT * 4 <= 48 * 1024, HEPT_EUNSUPPORTED,
               "out_linear_bwd: H*D=%d too wide", IN);
  if (dx) {
    const size_t smem = sizeof(float) * (size_t)OUT * IN;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(out_linear_bwd_input_kernel<OUT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_bwd: %s", cudaGetErrorString(e));
      configured = true;
    }
    HEPT_REQUIRE(smem <= 96 * 1024, HEPT_EUNSUPPORTED, "out_linear_bwd: H*D=%d too wide", IN);
    const int units = (s->N + 63) / 64 * SPLIT;
    const unsigned grid = (unsigned)((units + kOlWarps - 1) / kOlWarps);
    out_linear_bwd_input_kernel<OUT, SPLIT><<<grid, 32 * kOlWarps, smem, st>>>(g, w, s->N, IN, dx);
    HEPT_CHECK_LAUNCH("out_linear_bwd_input");
  }
  const int slabs = (s->N + kOlRows - 1) / kOlRows;
  int ctas = sms * 4;
  if (ctas > slabs) ctas = slabs;
  if (ctas > kOlMaxCtas) ctas = kOlMaxCtas;
  const size_t psmem = sizeof(float) * ((size_t)kOlRows * OUT + (size_t)OUT * IN);
  out_linear_bwd_params_kernel<OUT><<<ctas, kOlThreads, psmem, st>>>(g, x, s->N, IN, partial);
  HEPT_CHECK_LAUNCH("out_linear_bwd_params");
  const int entries = (OUT + 1) * IN;
  out_linear_reduce_kernel<<<(entries + 31) / 32, 256, 0, st>>>(partial, ctas, OUT, IN, dw, db);
  HEPT_CHECK_LAUNCH("out_linear_reduce");
  return HEPT_OK;
}

}  // namespace hept

using namespace hept;

extern "C" int hept_out_linear_fwd(const hept_shape* s, const float* out_pre, const float* weight, const float* bias,
                                   float* out, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(out_pre && weight && bias && out, HEPT_EINVAL, "out_linear_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->D == 24) return launch_ol_fwd<24, HEPT_OL_OUTS, HEPT_OL_HT>(s, out_pre, weight, bias, out, st);
  if (s->D == 8) return launch_ol_fwd<8, 8, 2>(s, out_pre, weight, bias, out, st);
  set_error("out_linear_fwd: D=%d not compiled in", s->D);
  return HEPT_EUNSUPPORTED;
}

extern "C" size_t hept_out_linear_bwd_workspace_bytes(const hept_shape* s) {
  if (!s || s->H <= 0 || s->D <= 0) return 0;
  return sizeof(float) * (size_t)kOlMaxCtas * (s->D + 1) * s->H * s->D;
}

extern "C" int hept_out_linear_bwd(const hept_shape* s, const float* d_out, const float* weight, const float* out_pre,
                                   float* d_out_pre, float* d_weight, float* d_bias, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(d_out && weight && out_pre && d_weight && d_bias && workspace, HEPT_EINVAL, "out_linear_bwd: null pointer");
  HEPT_REQUIRE(workspace_bytes >= hept_out_linear_bwd_workspace_bytes(s), HEPT_EWORKSPACE,
               "out_linear_bwd: workspace needs %zu bytes", hept_out_linear_bwd_workspace_bytes(s));
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (s->D == 24) return launch_ol_bwd<24, 4>(s, d_out, weight, out_pre, d_out_pre, d_weight, d_bias, partial, st);
  if (s->D == 8) return launch_ol_bwd<8, 1>(s, d_out, weight, out_pre, d_out_pre, d_weight, d_bias, partial, st);
  set_error("out_linear_bwd: D=%d not compiled in", s->D);
  return HEPT_EUNSUPPORTED;
}
