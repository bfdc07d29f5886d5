// Lane-pair tiles: two adjacent lanes share R resident rows, each holding one half of every row (half of the
// 32-float hat row, half of the D-wide value / gradient row), and combine their partial dot products with one
// shuffle.  Compared with one lane per row this doubles the FMAs done per broadcast LDS.128 at the same register
// budget — the fp32 tiles are bounded by shared-memory load bandwidth (each LDS.128 costs ~3 LSU cycles however
// many lanes share the address), not by the FMA pipe.  All multiply-adds are packed FFMA2 (fma.rn.f32x2, new on
// sm_100): two fp32 FMAs per issue slot.
//
// Canonical arithmetic of the backward tiles (dq and dk kernels must see bit-identical dS, see tile.cuh):
//   half_dot(h) = (A.x + A.y) + (B.x + B.y), A / B = FFMA2 accumulators from (0,0) over the half's even / odd
//                 16-byte chunks in order, two elements at a time
//   dot = half_dot(0) + half_dot(1)          t = (dot + nq) + nk          P = ex2(min(t log2e, 0))
//   dP  = (half_dp(0) + half_dp(1)) - gy     with half_dp built the same way from (gd, v)
#pragma once

#include "tile.cuh"

namespace hept {

template <int D_, int C_, int B_, int G_, int R_>
struct PairLayout {
  static constexpr int D = D_, C = C_, B = B_, G = G_, R = R_;
  static constexpr int E = D + C;
  static_assert(D % 8 == 0, "value rows split in two halves of whole 16-byte chunks");
  static_assert(E + 2 <= 32, "hash_dim + 2 side slots must fit one 32-float row");
  static_assert(B % R == 0, "block size must be a multiple of the rows per lane pair");
  static constexpr int ROW_CHUNKS = 8;
  static constexpr int USED_CHUNKS = (E + 2 + 3) / 4;
  static_assert(USED_CHUNKS % 2 == 0, "hat rows split in two halves of whole chunks");
  static constexpr int HCH = USED_CHUNKS / 2;          // hat chunks per half
  static constexpr int HE = 4 * HCH;                   // hat elements per half
  static constexpr int VCH = D / 4;
  static constexpr int VH = VCH / 2;                   // value chunks per half
  static constexpr int RG = B / R;                     // row groups (= lane pairs) per block
  static constexpr int LANES = G * RG * 2;
  static constexpr int THREADS = (LANES + 31) / 32 * 32;
  static constexpr size_t SMEM_BYTES = (size_t)G * B * (ROW_CHUNKS + VCH) * sizeof(float4);
  // the gather helpers are written against TileLayout; same geometry, thread count from here
  using Gather = TileLayout<D, C, B, G, 1>;
};

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }

// acc2 += a2 * b2 over one 16-byte chunk (two FFMA2)
__device__ __forceinline__ void chunk_fma2(float2& acc, const float2* a, float4 b) {
  acc = __ffma2_rn(a[0], f2(b.x, b.y), acc);
  acc = __ffma2_rn(a[1], f2(b.z, b.w), acc);
}

// One half of a resident hat row: x' = x_hat[n] - centre for this lane's HCH chunks (zero at and beyond E, so the
// side slots of streamed rows drop out of dot products) and the half's share of |x'|^2 (pairwise tree over chunks).
template <class P>
__device__ __forceinline__ void load_half_row(const float* __restrict__ x, const float* __restrict__ kx,
                                              const float* __restrict__ coords, const float* __restrict__ scale_h,
                                              int n, int n0, int h, int H, int raw_size, int hf, float2* a2,
                                              float& half_sq) {
  float part[P::HCH];
#pragma unroll
  for (int cc = 0; cc < P::HCH; ++cc) {
    const int c = hf * P::HCH + cc;
    float4 d = load_hat_chunk<P::D, P::C>(x, coords, scale_h, n, h, H, c, n < raw_size);
    const float4 ctr = load_hat_chunk<P::D, P::C>(kx, coords, scale_h, n0, h, H, c, n0 < raw_size);
    d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
    float t[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (4 * c + u >= P::E) t[u] = 0.f;
    part[cc] = chunk_sq<P::E>(d, c);
    a2[2 * cc + 0] = f2(t[0], t[1]);
    a2[2 * cc + 1] = f2(t[2], t[3]);
  }
  if constexpr (P::HCH == 4) half_sq = (part[0] + part[1]) + (part[2] + part[3]);
  else if constexpr (P::HCH == 2) half_sq = part[0] + part[1];
  else half_sq = part[0];
}

// One half of gd = g / den of hit n and the half's share of gd . y (chunk partials summed in order).
template <class P>
__device__ __forceinline__ void load_half_grad(const float* __restrict__ g, const float* __restrict__ y,
                                               const float* __restrict__ den, int n, int h, int H, int hf, float2* gd2,
                                               float& half_gy) {
  const float inv_den = 1.f / __ldg(den + (size_t)n * H + h);
  half_gy = 0.f;
#pragma unroll
  for (int cc = 0; cc < P::VH; ++cc) {
    const int c = hf * P::VH + cc;
    const float4 gg = ldg4(g + ((size_t)n * H + h) * P::D + 4 * c);
    const float4 yy = ldg4(y + ((size_t)n * H + h) * P::D + 4 * c);
    const float gx = gg.x * inv_den, gy_ = gg.y * inv_den, gz = gg.z * inv_den, gw = gg.w * inv_den;
    const float part = fmaf(gw, yy.w, fmaf(gz, yy.z, fmaf(gy_, yy.y, fmaf(gx, yy.x, 0.f))));
    half_gy = cc == 0 ? part : half_gy + part;
    gd2[2 * cc + 0] = f2(gx, gy_);
    gd2[2 * cc + 1] = f2(gz, gw);
  }
}

}  // namespace hept
