// Shared pieces of the block-local kernel-attention tiles (forward and both backward passes).
//
// Work decomposition (all three kernels): one CTA owns G consecutive blocks of one (table, head).
// "Streamed" rows of a block (keys in the forward / dq pass, queries in the dkv pass) are gathered
// through the sort permutation into shared memory once; every lane owns R "resident" rows of one
// block in registers and walks the block's streamed rows.  Lanes of a warp that belong to the same
// block read the same shared-memory address (broadcast), so one LDS.128 feeds 32 lanes x 4 values;
// that is what keeps the fp32 FMA pipe, not the LSU, the limiter.
//
// Gather: eight lanes per row, one 16-byte chunk each — the 96-byte q/k/v row of a hit is read by six
// adjacent lanes (coalesced) and a warp writes 512 contiguous bytes of shared memory (conflict-free
// without swizzling, so readers address rows with immediates).
//
// Numerics: scores are evaluated on rows re-centred on the block's last key, q' = q_hat - c,
// k' = k_hat - c (S depends only on q_hat - k_hat), which tames the cancellation of
// q.k - |q|^2/2 - |k|^2/2 when blocks are spatially tight (trained weights).
//
// Canonical arithmetic: all three tile kernels evaluate  t = (dot(q', k') + nq) + nk,
// P = ex2(min(t * log2 e, 0)),  dP = chain(gd . v) - gy  with the SAME operation order, and every
// row-wise reduction (|x'|^2, gd . y) is "4-element FMA chain per 16-byte chunk, then a fixed
// pairwise tree over the 8 chunks" whether it is computed by one lane or by eight.  The forward
// pass and the two backward passes therefore see bit-identical P and dS.  That matters: d scale is a
// sum over all hits of coords * (dq^ + dk^) that cancels by ~|q^|^2 / |q^ - k^|^2, and any mismatch
// between the dS used for dq^ and the dS used for dk^ is amplified by that factor.
#pragma once

#include "common.cuh"

namespace hept {

// tile -> (head, table, block) for tiles ordered (head, table, block) without integer division on the device: the
// quotients come from one IMAD.HI each with multipliers computed on the host (a runtime `/` costs ~40 instructions, and
// the persistent tile kernels decode several tiles per iteration in every warp role).
// magic(d) = floor(2^32 / d) + 1 gives floor(x / d) = umulhi(x, magic) for x * d < 2^32; d == 1 is flagged by magic 0.
// Two orders of the (head, table) groups of nb tiles each:
//   plain   (h0,t0) (h0,t1) ... (h0,tT-1) (h1,t0) ...                 one head's rows stay in L2 for its T tables;
//   grouped, G heads at a time: (h0,t0) (h1,t0) (h2,t0) (h0,t1) (h1,t1) (h2,t1) ... then the next G heads; the last
//           H % G heads form a smaller group of their own.  G heads' rows stay in L2, and the tiles of (h, t + 1) start
//           G - 1 whole groups after the last tile of (h, t) -- what the backward needs to add a table's rows onto the
//           previous table's without waiting (attn_bwd_tc.cu).
struct TileDecoder {
  uint32_t nb, T, magic_nb, magic_T;
  uint32_t G, tail_base, tail_h0, tail_G, magic_GT, magic_G, magic_tG;   // G == 0: plain order
  static uint32_t magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / d) + 1u; }
  static TileDecoder make(int nb, int T, int H = 0, int G = 0) {
    TileDecoder d{(uint32_t)nb, (uint32_t)T, magic((uint32_t)nb), magic((uint32_t)T), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (G > 1 && H > 0) {
      d.G = (uint32_t)G;
      d.tail_G = (uint32_t)(H % G);
      d.tail_base = (uint32_t)((H / G) * G * T);
      d.tail_h0 = (uint32_t)((H / G) * G);
      d.magic_GT = magic((uint32_t)(G * T));
      d.magic_G = magic((uint32_t)G);
      d.magic_tG = magic(d.tail_G);
    }
    return d;
  }
  // largest tile count the multipliers are exact for
  static bool exact_for(long long tiles, int nb, int T) { return tiles * (long long)nb < (1ll << 32) && tiles * 8ll * T < (1ll << 32); }
  __device__ __forceinline__ void operator()(int tile, int& h, int& t, int& blk) const {
    const uint32_t hl = magic_nb ? __umulhi((uint32_t)tile, magic_nb) : (uint32_t)tile;
    blk = (int)((uint32_t)tile - hl * nb);
    if (G) {
      uint32_t r, h0, gs, mg;                     // position in the head group, its first head, its size
      if (hl < tail_base) {
        const uint32_t hp = __umulhi(hl, magic_GT);
        r = hl - hp * G * T; h0 = hp * G; gs = G; mg = magic_G;
      } else {
        r = hl - tail_base; h0 = tail_h0; gs = tail_G; mg = magic_tG;
      }
      const uint32_t tt = mg ? __umulhi(r, mg) : r;
      t = (int)tt;
      h = (int)(h0 + r - tt * gs);
    } else {
      const uint32_t hh = magic_T ? __umulhi(hl, magic_T) : hl;
      h = (int)hh;
      t = (int)(hl - hh * T);
    }
  }
};


template <int D_, int C_, int B_, int G_, int R_>
struct TileLayout {
  static constexpr int D = D_, C = C_, B = B_, G = G_, R = R_;
  static constexpr int E = D + C;
  static_assert(D % 4 == 0, "dims per head must be a multiple of 4 (float4 rows)");
  static_assert(E + 2 <= 32, "hash_dim + 2 side slots must fit one 32-float row");
  static constexpr int ROW_CHUNKS = 8;                 // 32-float (128-byte) rows: E values, side slots E, E+1
  static constexpr int USED_CHUNKS = (E + 2 + 3) / 4;  // chunks that carry data
  static constexpr int VCH = D / 4;                    // chunks of a value / gradient row
  static_assert(VCH <= 8 && VCH % 2 == 0, "value rows are gathered by the same 8 lanes and split in two halves");
  static constexpr int LPB = (B + R - 1) / R;          // lanes per block
  static constexpr int LANES = G * LPB;
  static constexpr int THREADS = (LANES + 31) / 32 * 32;
  static constexpr size_t SMEM_BYTES = (size_t)G * B * (ROW_CHUNKS + VCH) * sizeof(float4);
};

// Canonical row reduction of the D-wide rows (gd . y): the row's VCH chunk partials are summed in order inside each
// half of the row, then the two halves are added — the order a lane pair that splits the row in halves produces.
template <int VCH>
__device__ __forceinline__ float halves_sum(const float* p) {
  constexpr int VH = VCH / 2;
  float a = p[0], b = p[VH];
#pragma unroll
  for (int c = 1; c < VH; ++c) { a += p[c]; b += p[VH + c]; }
  return a + b;
}
// the same reduction across the 8 lanes of a gather row group (lane c holds the partial of chunk c)
template <int VCH>
__device__ __forceinline__ float halves_sum_lanes(float part) {
  float v[VCH];
  const int base = (threadIdx.x & 31) & ~7;
#pragma unroll
  for (int c = 0; c < VCH; ++c) v[c] = __shfl_sync(0xffffffffu, part, base + c);
  return halves_sum<VCH>(v);
}

// fixed pairwise tree over eight partial sums: ((p0+p1)+(p2+p3)) + ((p4+p5)+(p6+p7))
__device__ __forceinline__ float tree8(const float* p) {
  return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
}
// the same tree evaluated across the 8 lanes of a row group (xor butterfly: every lane gets the total)
__device__ __forceinline__ float tree8_lanes(float p) {
  p += __shfl_xor_sync(0xffffffffu, p, 1);
  p += __shfl_xor_sync(0xffffffffu, p, 2);
  p += __shfl_xor_sync(0xffffffffu, p, 4);
  return p;
}

// Raw 16-byte chunk `chunk` of the hat row of hit n, head h, BEFORE the coordinate scale is applied:
// x[n,h,e] for e < D, coords[n,e-D] for D <= e < E, else 0.  Loads only — no arithmetic — so that a gather can
// put many of these in flight before the first use (the chunk index is a lane-dependent branch).
template <int D, int C>
__device__ __forceinline__ float4 load_raw_chunk(const float* __restrict__ x, const float* __restrict__ coords, int n,
                                                 int h, int H, int chunk, bool real) {
  if (!real) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (4 * chunk + 3 < D) return ldg4(x + ((size_t)n * H + h) * D + 4 * chunk);
  float t[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = 4 * chunk + u;
    if (e < D) t[u] = __ldg(x + ((size_t)n * H + h) * D + e);
    else if (e < D + C) t[u] = __ldg(coords + (size_t)n * C + (e - D));
    else t[u] = 0.f;
  }
  return make_float4(t[0], t[1], t[2], t[3]);
}
// per-chunk multipliers: 1 for feature columns, scale[h, e-D] for coordinate columns, 0 for padding
template <int D, int C>
__device__ __forceinline__ float4 chunk_multiplier(const float* __restrict__ scale_h, int chunk) {
  float t[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = 4 * chunk + u;
    t[u] = e < D ? 1.f : (e < D + C ? __ldg(scale_h + (e - D)) : 0.f);
  }
  return make_float4(t[0], t[1], t[2], t[3]);
}
// scale * coords with one rounding (x * 1.0f is exact for the feature columns), like the reference's fp32 product
__device__ __forceinline__ float4 apply_multiplier(float4 raw, float4 m) {
  return make_float4(__fmul_rn(m.x, raw.x), __fmul_rn(m.y, raw.y), __fmul_rn(m.z, raw.z), __fmul_rn(m.w, raw.w));
}
// chunk `chunk` of the hat row of hit n, head h: x[n,h,e] for e < D, scale[h,e-D] * coords[n,e-D] up to E, else 0
template <int D, int C>
__device__ __forceinline__ float4 load_hat_chunk(const float* __restrict__ x, const float* __restrict__ coords,
                                                 const float* __restrict__ scale_h, int n, int h, int H, int chunk,
                                                 bool real) {
  return apply_multiplier(load_raw_chunk<D, C>(x, coords, n, h, H, chunk, real), chunk_multiplier<D, C>(scale_h, chunk));
}

// sum of squares of the (at most 4) real elements of a chunk, FMA chain from 0 in element order
template <int E>
__device__ __forceinline__ float chunk_sq(float4 d, int chunk) {
  const float t[4] = {d.x, d.y, d.z, d.w};
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (4 * chunk + u < E) s = fmaf(t[u], t[u], s);
  return s;
}

// Gather the streamed side of the G blocks a CTA owns, eight lanes per row.
//   hat rows  -> `hs` [G*B][8] float4: x' = x_hat - centre in [0,E), side0 = -|x'|^2/2 at slot E, side1 at E+1
//   aux rows  -> `as` [G*B][VCH] float4: value rows (GRAD == false) or gd = g / den rows (GRAD == true, in which
//               case side1 = gy = gd . y; otherwise side1 = 0)
// `spos` are the sorted positions of the streamed side, `kpos` those of the keys (the centre is the block's last key).
template <class L, bool GRAD>
__device__ __forceinline__ void gather_streamed_rows(const float* __restrict__ x, const float* __restrict__ kx,
                                                     const float* __restrict__ aux, const float* __restrict__ y,
                                                     const float* __restrict__ den, const float* __restrict__ coords,
                                                     const float* __restrict__ scale_h, const int32_t* __restrict__ spos,
                                                     const int32_t* __restrict__ kpos, int blk0, int nb, int h, int H,
                                                     int raw_size, float4* hs, float4* as) {
  constexpr int D = L::D, C = L::C, B = L::B, E = L::E;
  constexpr int ROWS_PER_PASS = L::THREADS / 8;
  constexpr int PASSES = (L::G * B + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
  constexpr int BATCH = 4;  // rows in flight per lane: the gather is latency-bound, not bandwidth-bound
  const int sub = threadIdx.x >> 3, c = threadIdx.x & 7;
  const float4 mult = chunk_multiplier<D, C>(scale_h, c);
  // all permutation indices first (one round trip), then the row data BATCH passes at a time
  int nidx[PASSES], n0idx[PASSES];
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int rr = ps * ROWS_PER_PASS + sub;
    const int g = rr / B, blk = blk0 + g;
    const bool valid = rr < L::G * B && blk < nb;
    nidx[ps] = valid ? __ldg(spos + (size_t)blk * B + (rr - g * B)) : -1;
    n0idx[ps] = valid ? __ldg(kpos + (size_t)blk * B + (B - 1)) : 0;
  }
#pragma unroll
  for (int p0 = 0; p0 < PASSES; p0 += BATCH) {
    float4 d[BATCH], ctr[BATCH], av[BATCH], yy[BATCH];
    float inv_den[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int ps = p0 + u;
      if (ps >= PASSES) continue;
      const int n = nidx[ps], n0 = n0idx[ps];
      const bool valid = n >= 0, real = valid && n < raw_size;
      d[u] = load_raw_chunk<D, C>(x, coords, valid ? n : 0, h, H, c, real);
      ctr[u] = load_raw_chunk<D, C>(kx, coords, n0, h, H, c, valid && n0 < raw_size);
      av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      yy[u] = av[u];
      inv_den[u] = 0.f;
      if (GRAD) {
        if (c < L::VCH && valid) {
          inv_den[u] = 1.f / __ldg(den + (size_t)n * H + h);
          av[u] = ldg4(aux + ((size_t)n * H + h) * D + 4 * c);
          yy[u] = ldg4(y + ((size_t)n * H + h) * D + 4 * c);
        }
      } else if (c < L::VCH && real) {
        av[u] = ldg4(aux + ((size_t)n * H + h) * D + 4 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int ps = p0 + u;
      if (ps >= PASSES) continue;
      const int rr = ps * ROWS_PER_PASS + sub;
      const bool valid = nidx[ps] >= 0;
      float4 dd = apply_multiplier(d[u], mult);
      const float4 cc = apply_multiplier(ctr[u], mult);
      dd.x -= cc.x; dd.y -= cc.y; dd.z -= cc.z; dd.w -= cc.w;
      const float half_sq = -0.5f * tree8_lanes(chunk_sq<E>(dd, c));
      float side1 = 0.f;
      float4 a4 = av[u];
      if (GRAD) {
        a4 = make_float4(a4.x * inv_den[u], a4.y * inv_den[u], a4.z * inv_den[u], a4.w * inv_den[u]);
        const float part = fmaf(a4.w, yy[u].w, fmaf(a4.z, yy[u].z, fmaf(a4.y, yy[u].y, fmaf(a4.x, yy[u].x, 0.f))));
        side1 = halves_sum_lanes<L::VCH>(part);
      }
      if (valid) {
        float t[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int e = 4 * c + w;
          if (e == E) t[w] = half_sq;
          else if (e == E + 1) t[w] = side1;
          else if (e > E + 1) t[w] = 0.f;
        }
        if (c < L::USED_CHUNKS) hs[rr * L::ROW_CHUNKS + c] = make_float4(t[0], t[1], t[2], t[3]);
        if (c < L::VCH) as[rr * L::VCH + c] = a4;
      }
    }
  }
}

// Resident row of a lane: x' = x_hat[n] - centre (E values) and -|x'|^2/2, canonical reduction order.
template <class L>
__device__ __forceinline__ void load_resident_row(const float* __restrict__ x, const float* __restrict__ kx,
                                                  const float* __restrict__ coords, const float* __restrict__ scale_h,
                                                  int n, int n0, int h, int H, int raw_size, float* a, float& half_sq) {
  float part[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    part[c] = 0.f;
    if (c < L::USED_CHUNKS && 4 * c < L::E) {
      float4 d = load_hat_chunk<L::D, L::C>(x, coords, scale_h, n < 0 ? 0 : n, h, H, c, n >= 0 && n < raw_size);
      const float4 ctr = load_hat_chunk<L::D, L::C>(kx, coords, scale_h, n0, h, H, c, n0 < raw_size);
      d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
      part[c] = chunk_sq<L::E>(d, c);
      const float t[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (4 * c + u < L::E) a[4 * c + u] = t[u];
    }
  }
  half_sq = -0.5f * tree8(part);
}

// gd = g / den (D values) and gy = gd . y of hit n, canonical reduction order.
template <class L>
__device__ __forceinline__ void load_resident_grad(const float* __restrict__ g, const float* __restrict__ y,
                                                   const float* __restrict__ den, int n, int h, int H, float* gd,
                                                   float& gy) {
  const float inv_den = 1.f / __ldg(den + (size_t)n * H + h);
  float part[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    part[c] = 0.f;
    if (c < L::VCH) {
      const float4 gg = ldg4(g + ((size_t)n * H + h) * L::D + 4 * c);
      const float4 yy = ldg4(y + ((size_t)n * H + h) * L::D + 4 * c);
      gd[4 * c + 0] = gg.x * inv_den; gd[4 * c + 1] = gg.y * inv_den;
      gd[4 * c + 2] = gg.z * inv_den; gd[4 * c + 3] = gg.w * inv_den;
      part[c] = fmaf(gd[4 * c + 3], yy.w, fmaf(gd[4 * c + 2], yy.z, fmaf(gd[4 * c + 1], yy.y, fmaf(gd[4 * c], yy.x, 0.f))));
    }
  }
  gy = halves_sum<L::VCH>(part);
}

// dot of R resident rows with one streamed row; also hands back the two side slots of the streamed row.
// Two partial sums (even / odd chunks, both from 0) halve the dependent-FMA chain.
template <class L>
__device__ __forceinline__ void dot_rows(const float4* __restrict__ row, const float (&a)[L::R][L::E], float (&s)[L::R],
                                         float& side0, float& side1, float4 (&keep)[L::USED_CHUNKS]) {
  float s1[L::R];
#pragma unroll
  for (int r = 0; r < L::R; ++r) { s[r] = 0.f; s1[r] = 0.f; }
#pragma unroll
  for (int c = 0; c < L::USED_CHUNKS; ++c) {
    const float4 kk = row[c];
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

// dP = (gd . v) - gy with a fixed order: two interleaved FMA chains, the first seeded with -gy.
// (gd, v) may come from registers or shared memory in either role: fmaf is commutative in its factors.
__device__ __forceinline__ void dp_step(float4 gg, float4 vv, float& dp0, float& dp1) {
  dp0 = fmaf(gg.x, vv.x, dp0);
  dp1 = fmaf(gg.y, vv.y, dp1);
  dp0 = fmaf(gg.z, vv.z, dp0);
  dp1 = fmaf(gg.w, vv.w, dp1);
}

}  // namespace hept
