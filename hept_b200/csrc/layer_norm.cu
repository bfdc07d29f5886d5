// The other LayerNorms of the caller: the Attn block's norm2 (example/transformer.py:163,
// src/models/baselines/transformer.py:216) and the four 256-wide norms of the model head (torch_geometric MLP,
// example/transformer.py:84), forward and backward.  Same definition as torch.nn.functional.layer_norm: biased variance over
// the last dimension, eps inside the square root, y = (x - mean) rstd gamma + beta.
//
// Why not the library's: on (60 000, 24) and (60 000, 256) fp32 rows torch's backward spends 0.25-0.35 ms per call in
// GammaBetaBackward alone -- 3.0 ms of a 15 ms training step of the tracking model (torch profiler, tools/prof_model.py),
// next to 3.2 ms for all four layers' backward attention tiles.  Here a row is shared by LPR lanes (8 for D = 24, 32 for
// D = 256) holding one or two 16-byte vectors each, row statistics go over the lanes by shuffles, and d gamma / d beta are
// per-thread sums over the thread's rows, added over the CTA's row groups in a fixed order and over CTAs by
// ln_params_reduce (attn_block.cu): deterministic, one streaming pass over x and dy.
#include "common.cuh"

namespace hept {

void ln_params_reduce_launch(const float* partial, int ctas, int DM, float* dgamma, float* dbeta, cudaStream_t st);

constexpr int kLnThreads = 256, kLnMaxCtas = 1024;

// lane `sub` of a row's LPR lanes holds the vectors sub, sub + LPR (VPL of them) of the row's D / 4
template <int LPR>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int LPR, int VPL>
__global__ void __launch_bounds__(kLnThreads) layer_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, int N, int D, float eps,
                                                                     float* __restrict__ y, float* __restrict__ mean_rstd) {
  constexpr int RPC = kLnThreads / LPR;                      // rows per CTA pass
  const int sub = threadIdx.x % LPR, rin = threadIdx.x / LPR;
  const int nv = D / 4;
  float4 g[VPL], b[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    const int c = sub + u * LPR;
    g[u] = c < nv ? ldg4(gamma + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    b[u] = c < nv ? ldg4(beta + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv = 1.f / (float)D;
  for (int r0 = blockIdx.x * RPC; r0 < N; r0 += gridDim.x * RPC) {
    const int r = r0 + rin;
    const bool live = r < N;                                 // dead rows run the shuffles with zeros
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = sub + u * LPR;
      v[u] = (live && c < nv) ? ldg4(x + (size_t)r * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    const float mean = group_sum<LPR>(s) * inv;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      if (sub + u * LPR < nv) {
        const float dx = v[u].x - mean, dy = v[u].y - mean, dz = v[u].z - mean, dw = v[u].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = 1.f / sqrtf(group_sum<LPR>(q) * inv + eps);
    if (live) {
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        const int c = sub + u * LPR;
        if (c < nv) {
          float4 o;
          o.x = fmaf((v[u].x - mean) * rstd, g[u].x, b[u].x);
          o.y = fmaf((v[u].y - mean) * rstd, g[u].y, b[u].y);
          o.z = fmaf((v[u].z - mean) * rstd, g[u].z, b[u].z);
          o.w = fmaf((v[u].w - mean) * rstd, g[u].w, b[u].w);
          *reinterpret_cast<float4*>(y + (size_t)r * D + 4 * c) = o;
        }
      }
      if (sub == 0) *reinterpret_cast<float2*>(mean_rstd + 2 * (size_t)r) = make_float2(mean, rstd);
    }
  }
}

// dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy gamma;  d gamma += dy xhat;  d beta += dy.
// partial (ctas, 2 D): the CTA's sums, row groups added in order
template <int LPR, int VPL>
__global__ void __launch_bounds__(kLnThreads) layer_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mean_rstd,
                                                                     const float* __restrict__ gamma, const float* __restrict__ dy,
                                                                     int N, int D, float* __restrict__ dx, float* __restrict__ partial) {
  constexpr int RPC = kLnThreads / LPR;
  __shared__ float s_red[RPC][2 * 4 * LPR * VPL + 1];       // + 1: rows of different groups start in different banks
  const int sub = threadIdx.x % LPR, rin = threadIdx.x / LPR;
  const int nv = D / 4;
  float4 g[VPL], dg[VPL], db[VPL];
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    const int c = sub + u * LPR;
    g[u] = c < nv ? ldg4(gamma + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv = 1.f / (float)D;
  for (int r0 = blockIdx.x * RPC; r0 < N; r0 += gridDim.x * RPC) {
    const int r = r0 + rin;
    const bool live = r < N;
    const float2 ms = live ? ldg2(mean_rstd + 2 * (size_t)r) : make_float2(0.f, 0.f);
    float4 xh[VPL], gy[VPL];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int u = 0; u < VPL; ++u) {
      const int c = sub + u * LPR;
      const bool on = live && c < nv;
      const float4 xv = on ? ldg4(x + (size_t)r * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 d = on ? ldg4(dy + (size_t)r * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      xh[u] = on ? make_float4((xv.x - ms.x) * ms.y, (xv.y - ms.x) * ms.y, (xv.z - ms.x) * ms.y, (xv.w - ms.x) * ms.y)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
      gy[u] = make_float4(d.x * g[u].x, d.y * g[u].y, d.z * g[u].z, d.w * g[u].w);
      m1 += (gy[u].x + gy[u].y) + (gy[u].z + gy[u].w);
      m2 += (gy[u].x * xh[u].x + gy[u].y * xh[u].y) + (gy[u].z * xh[u].z + gy[u].w * xh[u].w);
      dg[u].x = fmaf(d.x, xh[u].x, dg[u].x); dg[u].y = fmaf(d.y, xh[u].y, dg[u].y);
      dg[u].z = fmaf(d.z, xh[u].z, dg[u].z); dg[u].w = fmaf(d.w, xh[u].w, dg[u].w);
      db[u].x += d.x; db[u].y += d.y; db[u].z += d.z; db[u].w += d.w;
    }
    m1 = group_sum<LPR>(m1) * inv;
    m2 = group_sum<LPR>(m2) * inv;
    if (live) {
#pragma unroll
      for (int u = 0; u < VPL; ++u) {
        const int c = sub + u * LPR;
        if (c < nv) {
          float4 o;
          o.x = ms.y * (gy[u].x - m1 - xh[u].x * m2);
          o.y = ms.y * (gy[u].y - m1 - xh[u].y * m2);
          o.z = ms.y * (gy[u].z - m1 - xh[u].z * m2);
          o.w = ms.y * (gy[u].w - m1 - xh[u].w * m2);
          *reinterpret_cast<float4*>(dx + (size_t)r * D + 4 * c) = o;
        }
      }
    }
  }
  // the CTA's sums: row groups in order 0 .. RPC - 1 (fixed order)
#pragma unroll
  for (int u = 0; u < VPL; ++u) {
    float* d = s_red[rin] + 4 * (sub + u * LPR);
    d[0] = dg[u].x; d[1] = dg[u].y; d[2] = dg[u].z; d[3] = dg[u].w;
    float* e = d + 4 * LPR * VPL;
    e[0] = db[u].x; e[1] = db[u].y; e[2] = db[u].z; e[3] = db[u].w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += kLnThreads) {
    const int col = i < D ? i : i - D;                       // vector col / 4, element col % 4 -> slot 4 (vector) + element
    const int slot = (i < D ? 0 : 4 * LPR * VPL) + col;
    float s = 0.f;
#pragma unroll 8
    for (int rg = 0; rg < RPC; ++rg) s += s_red[rg][slot];
    partial[(size_t)blockIdx.x * 2 * D + i] = s;
  }
}

template <int LPR, int VPL>
static int launch_ln(bool fwd, const float* x, const float* mr_in, const float* gamma, const float* beta_or_dy, int N, int D, float eps,
                     float* out, float* mr_out, float* dgamma, float* dbeta, float* ws, cudaStream_t st) {
  constexpr int RPC = kLnThreads / LPR;
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "layer_norm: cannot read the SM count");
  const int passes = (N + RPC - 1) / RPC;
  int ctas = 4 * sms < passes ? 4 * sms : passes;            // four CTAs of 256 threads per SM, each walks its share of the rows
  if (ctas > kLnMaxCtas) ctas = kLnMaxCtas;
  if (fwd) {
    layer_norm_fwd_kernel<LPR, VPL><<<ctas, kLnThreads, 0, st>>>(x, gamma, beta_or_dy, N, D, eps, out, mr_out);
    HEPT_CHECK_LAUNCH("layer_norm_fwd");
  } else {
    layer_norm_bwd_kernel<LPR, VPL><<<ctas, kLnThreads, 0, st>>>(x, mr_in, gamma, beta_or_dy, N, D, out, ws);
    HEPT_CHECK_LAUNCH("layer_norm_bwd");
    ln_params_reduce_launch(ws, ctas, D, dgamma, dbeta, st);
    HEPT_CHECK_LAUNCH("ln_params_reduce");
  }
  return HEPT_OK;
}

static int dispatch_ln(bool fwd, const float* x, const float* mr_in, const float* gamma, const float* beta_or_dy, int N, int D,
                       float eps, float* out, float* mr_out, float* dgamma, float* dbeta, float* ws, cudaStream_t st) {
  const int nv = D / 4;
#define HEPT_LN_CASE(LPR, VPL) \
  return launch_ln<LPR, VPL>(fwd, x, mr_in, gamma, beta_or_dy, N, D, eps, out, mr_out, dgamma, dbeta, ws, st)
  if (nv <= 4) HEPT_LN_CASE(4, 1);
  if (nv <= 8) HEPT_LN_CASE(8, 1);
  if (nv <= 16) HEPT_LN_CASE(16, 1);
  if (nv <= 32) HEPT_LN_CASE(32, 1);
  HEPT_LN_CASE(32, 2);
#undef HEPT_LN_CASE
}

}  // namespace hept

using namespace hept;

extern "C" int hept_layer_norm_supported(int32_t D) { return D >= 4 && D <= 256 && D % 4 == 0; }

extern "C" int hept_layer_norm_fwd(const float* x, const float* weight, const float* bias, int32_t N, int32_t D, float eps, float* y,
                                   float* mean_rstd, void* stream) {
  HEPT_REQUIRE(x && weight && bias && y && mean_rstd && N > 0, HEPT_EINVAL, "layer_norm_fwd: bad argument");
  HEPT_REQUIRE(hept_layer_norm_supported(D), HEPT_EUNSUPPORTED, "layer_norm_fwd: D=%d (4 <= D <= 256, D %% 4 == 0)", D);
  HEPT_REQUIRE(aligned16({x, weight, bias, y}) && !(reinterpret_cast<uintptr_t>(mean_rstd) & 7u), HEPT_EINVAL,
               "layer_norm_fwd: array pointers must be 16-byte aligned");
  return dispatch_ln(true, x, nullptr, weight, bias, N, D, eps, y, mean_rstd, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" size_t hept_layer_norm_bwd_workspace_bytes(int32_t N, int32_t D) {
  if (N <= 0 || D <= 0) return 0;
  return sizeof(float) * (size_t)kLnMaxCtas * 2 * D;
}

extern "C" int hept_layer_norm_bwd(const float* x, const float* mean_rstd, const float* weight, const float* dy, int32_t N, int32_t D,
                                   float* dx, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(x && mean_rstd && weight && dy && dx && d_weight && d_bias && workspace && N > 0, HEPT_EINVAL,
               "layer_norm_bwd: bad argument");
  HEPT_REQUIRE(hept_layer_norm_supported(D), HEPT_EUNSUPPORTED, "layer_norm_bwd: D=%d (4 <= D <= 256, D %% 4 == 0)", D);
  HEPT_REQUIRE(workspace_bytes >= hept_layer_norm_bwd_workspace_bytes(N, D), HEPT_EWORKSPACE, "layer_norm_bwd: workspace needs %zu bytes",
               hept_layer_norm_bwd_workspace_bytes(N, D));
  HEPT_REQUIRE(aligned16({x, weight, dy, dx, workspace}) && !(reinterpret_cast<uintptr_t>(mean_rstd) & 7u), HEPT_EINVAL,
               "layer_norm_bwd: array pointers must be 16-byte aligned");
  return dispatch_ln(false, x, mean_rstd, weight, dy, N, D, 0.f, dx, nullptr, d_weight, d_bias, (float*)workspace, (cudaStream_t)stream);
}
