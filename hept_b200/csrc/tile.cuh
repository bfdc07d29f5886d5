// Shared pieces of the block-local kernel-attention tiles (forward and both backward passes).
//
// Work decomposition (all three kernels): one CTA owns G consecutive blocks of one (table, head).
// "Streamed" rows of a block (keys in the forward / dq pass, queries in the dkv pass) are gathered
// through the sort permutation into shared memory once; every lane owns R "resident" rows of one
// block in registers and walks the block's streamed rows.  Lanes of a warp that belong to the same
// block read the same shared-memory address (broadcast), so one LDS.128 wavefront feeds 32 lanes x 4
// values; that is what keeps the fp32 FMA pipe, not the LSU, the limiter.
//
// Numerics: scores are evaluated on rows re-centred on the block's first key, q' = q_hat - c,
// k' = k_hat - c (S depends only on q_hat - k_hat), which removes the catastrophic cancellation of
// q.k - |q|^2/2 - |k|^2/2 at trained-weight magnitudes.  exp is evaluated as
// ex2(S * log2 e) with log2 e folded into the resident rows.
#pragma once

#include "common.cuh"

namespace hept {

template <int D_, int C_, int B_, int G_, int R_>
struct TileLayout {
  static constexpr int D = D_, C = C_, B = B_, G = G_, R = R_;
  static constexpr int E = D + C;
  static_assert(D % 4 == 0, "dims per head must be a multiple of 4 (float4 rows)");
  static_assert(E + 2 <= 32, "hash_dim + 2 side slots must fit one 32-float row");
  // 32-float (128-byte) rows: E values, then two side slots (E, E+1).  Eight 16-byte chunks per row,
  // stored with chunk ^= (row & 7) so that eight consecutive rows written by eight lanes hit eight
  // different bank groups (conflict-free STS.128); readers use the same XOR (uniform per row).
  static constexpr int ROW_CHUNKS = 8;
  static constexpr int USED_CHUNKS = (E + 2 + 3) / 4;  // chunks that carry data
  static constexpr int VCH = D / 4;                    // chunks of a value / gradient row
  static constexpr int LPB = (B + R - 1) / R;          // lanes per block
  static constexpr int LANES = G * LPB;
  static constexpr int THREADS = (LANES + 31) / 32 * 32;
  static constexpr size_t SMEM_BYTES = (size_t)G * B * (ROW_CHUNKS + VCH) * sizeof(float4);
  __device__ static __forceinline__ int hat_off(int row, int chunk) { return row * ROW_CHUNKS + (chunk ^ (row & 7)); }
};

// q_hat / k_hat row of hit n, head h: [x[n,h,:] | scale[h,:] * coords[n,:]]; zero for src/ padding rows.
template <int D, int C>
__device__ __forceinline__ void load_hat_row(const float* __restrict__ x, const float* __restrict__ coords,
                                             const float* sc, int n, int h, int H, bool real, float* out) {
  if (real) {
    load_row<D>(x + ((size_t)n * H + h) * D, out);
    float cc[C];
    load_row<C>(coords + (size_t)n * C, cc);
#pragma unroll
    for (int c = 0; c < C; ++c) out[D + c] = __fmul_rn(sc[c], cc[c]);
  } else {
#pragma unroll
    for (int e = 0; e < D + C; ++e) out[e] = 0.f;
  }
}

// Store a 32-float row (values[0..E) then side0, side1, zero padding) into swizzled shared memory.
template <class L>
__device__ __forceinline__ void store_hat_row(float4* base, int row, const float* vals, float side0, float side1) {
#pragma unroll
  for (int c = 0; c < L::USED_CHUNKS; ++c) {
    float t[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int e = 4 * c + x;
      t[x] = e < L::E ? vals[e] : (e == L::E ? side0 : (e == L::E + 1 ? side1 : 0.f));
    }
    base[L::hat_off(row, c)] = make_float4(t[0], t[1], t[2], t[3]);
  }
}

// dot of a resident row with a streamed row; also hands back the two side slots of the streamed row.
// Two partial sums (even / odd chunks) halve the dependent-FMA chain.
template <class L>
__device__ __forceinline__ void dot_rows(const float4* __restrict__ base, int row, const float (&a)[L::R][L::E],
                                         float (&s)[L::R], float& side0, float& side1, float4 (&keep)[L::USED_CHUNKS]) {
  float s1[L::R];
#pragma unroll
  for (int r = 0; r < L::R; ++r) s1[r] = 0.f;
#pragma unroll
  for (int c = 0; c < L::USED_CHUNKS; ++c) {
    const float4 kk = base[L::hat_off(row, c)];
    keep[c] = kk;
    const float t[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int e = 4 * c + x;
      if (e < L::E) {
#pragma unroll
        for (int r = 0; r < L::R; ++r) {
          if (c & 1) s1[r] = fmaf(a[r][e], t[x], s1[r]);
          else s[r] = fmaf(a[r][e], t[x], s[r]);
        }
      } else if (e == L::E) {
        side0 = t[x];
      } else if (e == L::E + 1) {
        side1 = t[x];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < L::R; ++r) s[r] += s1[r];
}

// Gather the key side of the G blocks a CTA owns: k' rows (+ nk2 = -|k'|^2/2 * log2 e in side slot 0)
// into `ks`, value rows into `vs`.  One thread per row; rows of blocks past the end are left untouched.
template <class L>
__device__ __forceinline__ void gather_key_rows(const float* __restrict__ k, const float* __restrict__ v,
                                                const float* __restrict__ coords, const float* sc,
                                                const int32_t* __restrict__ kpos, int blk0, int nb, int h, int H,
                                                int raw_size, float4* ks, float4* vs) {
  constexpr int D = L::D, C = L::C, B = L::B, E = L::E;
  for (int rr = threadIdx.x; rr < L::G * B; rr += L::THREADS) {
    const int g = rr / B, blk = blk0 + g;
    if (blk >= nb) continue;
    const int n = __ldg(kpos + (size_t)blk * B + (rr - g * B));
    const int n0 = __ldg(kpos + (size_t)blk * B + (B - 1));  // centre = last key of the block
    float kr[E], ctr[E];
    load_hat_row<D, C>(k, coords, sc, n, h, H, n < raw_size, kr);
    load_hat_row<D, C>(k, coords, sc, n0, h, H, n0 < raw_size, ctr);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      kr[e] -= ctr[e];
      sq = fmaf(kr[e], kr[e], sq);
    }
    store_hat_row<L>(ks, rr, kr, -0.5f * kLog2e * sq, 0.f);
    const bool real = n < raw_size;
#pragma unroll
    for (int c = 0; c < L::VCH; ++c)
      vs[rr * L::VCH + c] = real ? ldg4(v + ((size_t)n * H + h) * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Resident row of a lane: centred, log2(e)-scaled copy of x_hat[n] and its -|x'|^2/2 * log2 e.
template <class L>
__device__ __forceinline__ void load_resident_row(const float* __restrict__ x, const float* __restrict__ coords,
                                                  const float* sc, const float* ctr, int n, int h, int H,
                                                  int raw_size, float* a, float& half_sq) {
  float xr[L::E];
  load_hat_row<L::D, L::C>(x, coords, sc, n < 0 ? 0 : n, h, H, n >= 0 && n < raw_size, xr);
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < L::E; ++e) {
    const float d = xr[e] - ctr[e];
    sq = fmaf(d, d, sq);
    a[e] = d * kLog2e;
  }
  half_sq = -0.5f * kLog2e * sq;
}

}  // namespace hept
