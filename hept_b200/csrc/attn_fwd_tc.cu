// a8-a11 on the 5th-generation tensor cores: the same fused gather -> block attention -> scatter as
// attn_fwd.cu, with both contractions of a tile issued as tcgen05.mma (kind::tf32, M = 128) and the
// accumulators in TMEM.
//
// fp32 fidelity from tf32 tensor cores: every operand x is split as x = hi + lo with hi = rn_tf32(x) and
// lo = x - hi (exact), and each product is evaluated as hi*hi + hi*lo + lo*hi ("3xTF32", fp32 accumulate);
// the dropped lo*lo term is 2^-22 relative.  The split only works because rows are re-centred on the
// block's last key first (tile.cuh): |q'|, |k'| are block-sized, not detector-sized.
//
// One tile = one block of B <= 100 sorted hits of one (table, head):
//   S2[i,j] = log2e * (q'_i . k'_j + nk_j)         SS MMAs, K = 32: slots [0,E) carry q' log2e / k',
//                                                   slots 30, 31 carry 1 / (nk hi, nk mid) so the key-side
//                                                   norm rides along in the contraction (3-way split, exact)
//   P[i,j]  = ex2(min(S2 + nq2_i, 0))              one thread per TMEM lane (= query row), written back to
//                                                   TMEM as (hi, lo)
//   O[i,:]  = P V                                   TS MMAs (A = P from TMEM, B = V MN-major), K = 112
// Padded key rows get nk = -1e30 (P = 0), padded value rows are zero, padded query rows are never stored.
#include "tile.cuh"
#include "umma.cuh"

namespace hept {

constexpr int kTcM = 128;    // UMMA M: query rows, padded
constexpr int kTcN = 112;    // key rows, padded to a multiple of 16
constexpr int kTcVN = 32;    // value columns, padded

template <int D, int C, int B>
struct TcFwdSmem {
  static constexpr int A_BYTES = kTcM * 128;   // one 128-row K-major tile
  static constexpr int K_BYTES = kTcN * 128;
  static constexpr int OFF_AH = 0, OFF_AL = A_BYTES, OFF_KH = 2 * A_BYTES, OFF_KL = OFF_KH + K_BYTES,
                       OFF_VH = OFF_KL + K_BYTES, OFF_VL = OFF_VH + K_BYTES, OFF_NQ = OFF_VL + K_BYTES,
                       TOTAL = OFF_NQ + kTcM * 4;
};

constexpr int kTcThreads = 256;   // warps 0-3 and 4-7 both map onto TMEM lanes 0-127; they split the columns

template <int D, int C, int B, int TILES>
__global__ void __launch_bounds__(kTcThreads, 2)
    block_attn_fwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ hatc, const int32_t* __restrict__ positions, int N, int H, int T,
                             int raw_size, float* __restrict__ stage) {
  constexpr int E = D + C;
  static_assert(E + 2 <= 32 && D % 4 == 0 && D <= kTcVN && B <= kTcN && B <= kTcM, "tile shape");
  using SM = TcFwdSmem<D, C, B>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  float* s_nq = reinterpret_cast<float*>(smem + SM::OFF_NQ);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_l[kTcM];               // row sums of the upper column half

  const int h = blockIdx.y / T, t = blockIdx.y % T, th = t * H + h;  // tables of one head run back to back: its q/k/v slices stay in L2
  const int nb = N / B;
  const int32_t* qpos = positions + (size_t)th * N;
  const int32_t* kpos = positions + ((size_t)T * H + th) * N;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int sub = tid >> 3, c = tid & 7;                 // gather role: 8 lanes per row, 32 rows per pass
  const int row = (warp & 3) * 32 + (tid & 31);          // epilogue role: TMEM lane = query row
  const int half = warp >> 2;                            //                column half of that row

  if (tid == 0) umma::mbar_init(&mbar, 1);
  if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tPl = tmem + kTcN, tO = tmem + 2 * kTcN;   // columns [0,112) [112,224) [224,256)
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t sbase = umma::smem_u32(smem);
  uint32_t phase = 0;

  // chunk c of a hat row: feature columns come from q / k, the coordinate columns from hat_coords (already scaled)
  constexpr int XCH = D / 4;
  auto hat_chunk = [&](const float* __restrict__ x, int n) -> float4 {
    const float* src = c < XCH ? x + ((size_t)n * H + h) * D + 4 * c : hatc + ((size_t)n * H + h) * 8 + 4 * (c - XCH);
    return (c < XCH + 2 && n < raw_size) ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
  };

  // rows that never hold data are written once: query rows [B,128) zero, key rows [B,112) zero with nk = -1e30
  // (P = ex2(-1e30) = 0), value rows [B,112) zero
  for (int rr = B + sub; rr < kTcM; rr += kTcThreads / 8) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + SM::OFF_AH + umma::sw128_offset(rr, c)) = z;
    *reinterpret_cast<float4*>(smem + SM::OFF_AL + umma::sw128_offset(rr, c)) = z;
    if (rr < kTcN) {
      *reinterpret_cast<float4*>(smem + SM::OFF_KH + umma::sw128_offset(rr, c)) =
          c == 7 ? make_float4(0.f, 0.f, umma::tf32_hi(-1e30f), 0.f) : z;
      *reinterpret_cast<float4*>(smem + SM::OFF_KL + umma::sw128_offset(rr, c)) = z;
      *reinterpret_cast<float4*>(smem + SM::OFF_VH + umma::sw128b32_offset(rr, c)) = z;
      *reinterpret_cast<float4*>(smem + SM::OFF_VL + umma::sw128b32_offset(rr, c)) = z;
    }
    if (c == 0) s_nq[rr] = 0.f;
  }

#pragma unroll 1
  for (int it = 0; it < TILES; ++it) {
    const int blk = blockIdx.x * TILES + it;
    if (blk >= nb) break;

    // ---- gather: all indices, then all rows (every load of the tile is in flight before the first use) ------
    constexpr int PASSES = (B + 31) / 32;                 // 32 rows per pass; the last pass is partial
    int nk_idx[PASSES], nq_idx[PASSES];
    const int n0 = __ldg(kpos + (size_t)blk * B + (B - 1));
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ps * 32 + sub;
      nk_idx[ps] = r < B ? __ldg(kpos + (size_t)blk * B + r) : -1;
      nq_idx[ps] = r < B ? __ldg(qpos + (size_t)blk * B + r) : -1;
    }
    const float4 ctr = hat_chunk(k, n0);
    float4 dk[PASSES], vk[PASSES], dq[PASSES];
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int nkk = nk_idx[ps], nqq = nq_idx[ps];
      dk[ps] = nkk >= 0 ? hat_chunk(k, nkk) : make_float4(0.f, 0.f, 0.f, 0.f);
      dq[ps] = nqq >= 0 ? hat_chunk(q, nqq) : make_float4(0.f, 0.f, 0.f, 0.f);
      vk[ps] = (nkk >= 0 && c < XCH && nkk < raw_size) ? ldg4(v + ((size_t)nkk * H + h) * D + 4 * c)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ps * 32 + sub;
      // the last pass covers rows [32*(PASSES-1), B): warps whose 4 rows are all past B skip it
      if (ps == PASSES - 1 && (ps * 32 + (warp << 2)) >= B) continue;
      const bool in = r < B;
      // ---- key + value row r -> Kh, Kl, Vh, Vl
      {
        float4 d = dk[ps];
        d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
        const float sq = tree8_lanes(chunk_sq<E>(d, c));
        const float nk2 = kLog2e * (-0.5f * sq);
        float hi[4], lo[4];
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float x = (4 * c + u < E) ? dv[u] : 0.f;
          hi[u] = umma::tf32_hi(x);
          lo[u] = x - hi[u];
        }
        if (c == 7) {  // K slots 30, 31 carry the key-side norm, split three ways: nk2 = h0 + h1 + l0
          const float h0 = umma::tf32_hi(nk2);
          const float r1 = nk2 - h0;
          const float h1 = umma::tf32_hi(r1);
          hi[2] = h0; hi[3] = h1;
          lo[2] = r1 - h1; lo[3] = 0.f;
        }
        if (in) {
          const uint32_t off = umma::sw128_offset(r, c);
          *reinterpret_cast<float4*>(smem + SM::OFF_KH + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(smem + SM::OFF_KL + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          const float4 vv = vk[ps];
          const float vh[4] = {umma::tf32_hi(vv.x), umma::tf32_hi(vv.y), umma::tf32_hi(vv.z), umma::tf32_hi(vv.w)};
          const uint32_t voff = umma::sw128b32_offset(r, c);
          *reinterpret_cast<float4*>(smem + SM::OFF_VH + voff) = make_float4(vh[0], vh[1], vh[2], vh[3]);
          *reinterpret_cast<float4*>(smem + SM::OFF_VL + voff) = make_float4(vv.x - vh[0], vv.y - vh[1], vv.z - vh[2], vv.w - vh[3]);
        }
      }
      // ---- query row r -> Ah, Al, nq2
      {
        float4 d = dq[ps];
        d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
        const float nq2 = kLog2e * (-0.5f * tree8_lanes(chunk_sq<E>(d, c)));
        float hi[4], lo[4];
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = 4 * c + u;
          const float x = e < E ? dv[u] * kLog2e : 0.f;
          hi[u] = umma::tf32_hi(x);
          lo[u] = x - hi[u];
          if (e == 30 || e == 31) { hi[u] = 1.f; lo[u] = 0.f; }
        }
        if (in) {
          if (c == 0) s_nq[r] = nq2;
          const uint32_t off = umma::sw128_offset(r, c);
          *reinterpret_cast<float4*>(smem + SM::OFF_AH + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(smem + SM::OFF_AL + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    // ---- S2 = A K^T, 3xTF32 -----------------------------------------------------------------------------
    if (tid == 0) {
      constexpr uint32_t idesc = umma::idesc_tf32(kTcM, kTcN, false, false);
      // The accumulator is truncated (not rounded) after every MMA and q'.k' is far larger than the score it
      // cancels to, so: small cross terms first, and the four hi*hi steps split over two accumulators
      // (tS: cross terms + k-steps 0,1; tPl: k-steps 2,3) that are added in fp32 in the epilogue.
      bool acc = false;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const uint32_t a_off = part == 1 ? SM::OFF_AL : SM::OFF_AH;
        const uint32_t b_off = part == 0 ? SM::OFF_KL : SM::OFF_KH;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma::mma_ss(tS, umma::smem_desc_sw128(sbase + a_off + 32 * kk, 1024, 16),
                       umma::smem_desc_sw128(sbase + b_off + 32 * kk, 1024, 16), idesc, acc);
          acc = true;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma::mma_ss(kk < 2 ? tS : tPl, umma::smem_desc_sw128(sbase + SM::OFF_AH + 32 * kk, 1024, 16),
                     umma::smem_desc_sw128(sbase + SM::OFF_KH + 32 * kk, 1024, 16), idesc, kk != 2);
      umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, phase);
    phase ^= 1;
    umma::fence_after_sync();

    // ---- P = ex2(min(S2 + nq2, 0)), row sums, split back into TMEM; each warp-half owns 56 columns --------
    const float nq2 = s_nq[row];
    float l = 0.f;
    {
      constexpr int HC = kTcN / 2;   // 56 columns per half, 7 chunks of 8
      const uint32_t col0 = half * HC;
#pragma unroll
      for (int cc = 0; cc < HC / 8; ++cc) {
        uint32_t ra[8], rb[8];
        umma::tmem_ld8_nowait(tS + lane_base + col0 + 8 * cc, ra);
        umma::tmem_ld8_nowait(tPl + lane_base + col0 + 8 * cc, rb);
        umma::tmem_wait_ld();
        float ph[8], pl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float p = exp2_fast(fminf((__uint_as_float(ra[u]) + __uint_as_float(rb[u])) + nq2, 0.f));
          l += p;
          ph[u] = umma::tf32_hi(p);
          pl[u] = p - ph[u];
        }
        umma::tmem_st8(tS + lane_base + col0 + 8 * cc, ph);
        umma::tmem_st8(tPl + lane_base + col0 + 8 * cc, pl);
      }
    }
    if (half == 1) s_l[row] = l;
    umma::tmem_wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    // ---- O = P V, 3xTF32 ----------------------------------------------------------------------------------
    if (tid == 0) {
      constexpr uint32_t idesc = umma::idesc_tf32(kTcM, kTcVN, false, true);
      bool acc = false;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t a_t = part == 1 ? tPl : tS;
        const uint32_t b_off = part == 0 ? SM::OFF_VL : SM::OFF_VH;
#pragma unroll
        for (int kk = 0; kk < kTcN / 8; ++kk) {
          umma::mma_ts(tO, a_t + 8 * kk, umma::smem_desc(sbase + b_off + 1024 * kk, 512, 1024, umma::kLayoutSw128Base32),
                       idesc, acc);
          acc = true;
        }
      }
      umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, phase);
    phase ^= 1;
    umma::fence_after_sync();

    // ---- scatter numerator and normaliser back to original order: half 0 writes columns [0,16), half 1 the rest ---
    {
      float ov[16];
      umma::tmem_ld16(tO + lane_base + 16 * half, ov);
      if (row < B) {
        const int n = __ldg(qpos + (size_t)blk * B + row);
        float4* dst = reinterpret_cast<float4*>(stage + (((size_t)h * N + n) * T + t) * kStageRow);
        constexpr int VCH = D / 4;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int chunk = 4 * half + cc;
          if (chunk < VCH) dst[chunk] = make_float4(ov[4 * cc], ov[4 * cc + 1], ov[4 * cc + 2], ov[4 * cc + 3]);
        }
        if (half == 0) {
          dst[VCH] = make_float4((l + s_l[row]) + 1e-20f, 0.f, 0.f, 0.f);   // denom = rowsum + 1e-20
          if constexpr ((VCH + 1) % 2 == 1) dst[VCH + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();   // TMEM and shared memory are free for the next tile
    umma::fence_after_sync();
  }

  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

template <int D, int C, int B, int TILES>
static int launch_fwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* hatc,
                         const int32_t* positions, float* stage, cudaStream_t st) {
  using SM = TcFwdSmem<D, C, B>;
  auto kern = block_attn_fwd_tc_kernel<D, C, B, TILES>;
  const size_t smem = SM::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "block_attn_fwd_tc: cannot reserve %zu B of shared memory: %s", smem,
                 cudaGetErrorString(e));
    configured = true;
  }
  const int nb = s->N / s->B;
  dim3 grid((nb + TILES - 1) / TILES, s->T * s->H);
  kern<<<grid, kTcThreads, smem, st>>>(q, k, v, hatc, positions, s->N, s->H, s->T, s->raw_size, stage);
  HEPT_CHECK_LAUNCH("block_attn_fwd_tc");
  return HEPT_OK;
}

int block_attention_fwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* hatc,
                           const int32_t* positions, float* stage, cudaStream_t st) {
  if (s->D == 24 && s->C == 6 && s->B == 100) return launch_fwd_tc<24, 6, 100, 4>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 100) return launch_fwd_tc<24, 4, 100, 4>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 8 && s->C == 6 && s->B == 10) return launch_fwd_tc<8, 6, 10, 4>(s, q, k, v, hatc, positions, stage, st);
  set_error("block_attention_fwd (tensor-core engine): (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}

}  // namespace hept
