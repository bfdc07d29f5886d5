// a8-a11 on the 5th-generation tensor cores: the fused gather -> block attention -> scatter of attn_fwd.cu with both
// contractions of a tile issued as tcgen05.mma (kind::tf32, M = 128) and the accumulators in TMEM, as a persistent,
// warp-specialised pipeline: one CTA per SM walks the tiles, two tiles are in flight at any time.
//
// fp32 fidelity from tf32 tensor cores: every operand x is split as x = hi + lo with hi = rn_tf32(x) and
// lo = x - hi (exact), and each product is evaluated as hi*hi + hi*lo + lo*hi ("3xTF32", fp32 accumulate);
// the dropped lo*lo term is 2^-22 relative.  The split only works because rows are re-centred on the
// block's last key first: |q'|, |k'| are block-sized, not detector-sized.
//
// One tile = one block of B <= 104 sorted hits of one (table, head):
//   S[i,j]  = q'_i . k'_j + nk_j                    SS MMAs: K slots [0,E) carry q' / k', slots E, E+1 carry 1 /
//                                                   (nk hi, nk mid) so the key-side norm nk = -|k'|^2 / 2 rides along in
//                                                   the contraction (3-way split, exact)
//   P[i,j]  = ex2(min(log2e S + nq2_i, 0))          one thread per TMEM lane (= query row), written back to TMEM as (hi, lo)
//   O[i,:]  = P [V | 1]                             TS MMAs (A = P from TMEM, B = V MN-major); column D is the row sum
// Padded key rows get nk = -1e30 (P = 0), padded value rows are zero, padded query rows are never stored.
//
// The tensor-core accumulator truncates after every MMA and q'.k' is far larger than the score it cancels to, so the
// small cross terms are accumulated first and the hi*hi steps are split over two accumulators that the epilogue adds
// in fp32 (S0: cross terms + first half of the k-steps, S1: the rest).  (One accumulator -- half the epilogue's TMEM loads --
// passes the parity suite too but measured 1 % SLOWER: twelve score MMAs in one dependency chain.)
//
// Warp roles (896 threads = 7 warpgroups, register budgets 104 / 56 / 56 rebalanced with setmaxnreg):
//   warps 0-7   epilogue: TMEM lane = (warp & 3) * 32 + lane, columns split in two parts (warp >> 2); software
//               pipelined: P of tile n+1 is produced before the output rows of tile n are read and scattered, so the
//               P V MMAs of tile n run under SIMT work
//   warps 8-23  producer: gather q^ / k^ rows through the sort permutation into registers (one tile ahead), centre,
//               split, write the K-major operand tiles of stage n & 1 as soon as the score MMAs of tile n-2 are done
//   warp 24     one elected lane issues every tcgen05.mma: S(n+1), then P V(n)
//   warps 25-27 value operand: raw v rows travel global -> shared with cp.async four tiles ahead; as soon as P V of tile
//               n is done (its value tiles are free) each thread splits its own chunks into the MN-major (hi, lo) value
//               tiles of tile n+2.  (Until round 2 the epilogue warps did this between two tiles: with 16 producer warps
//               the epilogue was the kernel's serial chain -- 2 000 of its 6 200 traced cycles per tile -- and these
//               three warps idle: 288 -> 271 us.)
// TMEM: two tile slots of 256 columns (S0 | S1 | O).  Hand-offs are mbarriers that complete once per use of a slot /
// stage, so the wait parity is bit 1 of the tile counter.
#include "tile.cuh"
#include "trace.cuh"
#include "umma.cuh"

namespace hept {

using umma::split4;
using umma::split_tf32;
using umma::trunc_tf32;

constexpr int kFtEpiThreads = 256, kFtProdThreads = 512;
constexpr int kFtThreads = kFtEpiThreads + kFtProdThreads + 128;
constexpr int kFtParts = kFtEpiThreads / 128;
// Launch budget 65536 / 896 -> 72 registers per thread = 64512 per CTA; setmaxnreg moves registers inside that pool.
// The gather / centre / split work of the producer is latency bound (ncu: every producer warp busy or stalled on its
// own loads the whole time, the epilogue waiting for scores), so it gets 16 warps: two 64-row passes per tile.
constexpr int kFtLaunchRegs = 72;
constexpr int kFtRegsEpi = 104, kFtRegsProd = 56, kFtRegsMma = 56;
constexpr int kFtVWarps = 3, kFtVThreads = 32 * kFtVWarps;   // the MMA warp's three siblings prepare the value operand
static_assert(kFtEpiThreads * kFtRegsEpi + kFtProdThreads * kFtRegsProd + 128 * kFtRegsMma <= kFtThreads * kFtLaunchRegs, "register pool");

template <int D, int C, int B>
struct TcFwd {
  static constexpr int E = D + C;
  static constexpr int NP = (B + 15) / 16 * 16;   // N of the score MMAs (multiple of 16 at M = 128)
  static constexpr int KC = (B + 7) / 8 * 8;      // contraction length of the TS MMAs
  static constexpr int KSTEPS = KC / 8;
  static constexpr int SK = (E + 2 + 7) / 8;      // k-steps of Q^ K^^T (two side slots at E, E + 1)
  static constexpr int SK0 = (SK + 1) / 2;        // hi*hi k-steps that go to the first accumulator
  static constexpr int VCH = D / 4;
  static constexpr int SLOT_CH = E / 4, SLOT_U = E % 4;
  // Operand tiles hold KC rows of 128 B.  The MMAs read M = 128 (A) or NP (B) rows: what lies past row KC belongs to
  // the next tile and only reaches TMEM lanes / columns >= KC, which nothing reads.
  static constexpr int TILE = KC * 128;
  static constexpr int QH = 0, QL = 1, KH = 2, KL = 3;    // K-major tiles of a stage
  static constexpr int OFF_V = 2 * 4 * TILE;              // then per stage VH, VL (MN-major)
  static constexpr int OFF_VSTG = OFF_V + 2 * 2 * TILE;    // raw value rows, two sets of B x (D / 4) 16-byte chunks
  static constexpr int VSTG = (B * (D / 4) * 16 + 127) / 128 * 128;
  static constexpr int OFF_AUX = OFF_VSTG + 2 * VSTG;
  static constexpr int AUX_BYTES = (4 * 128 + 8 * 128) * 4;   // nq2[4][128], qidx[8][128]
  static constexpr int TOTAL = OFF_AUX + AUX_BYTES;
  static constexpr int RPP = kFtProdThreads / 8;
  static constexpr int PASSES = (B + RPP - 1) / RPP;
  static constexpr int SLOT_COLS = 2 * NP + 32;
  static constexpr int TMEM_COLS = SLOT_COLS <= 64 ? 128 : (SLOT_COLS <= 128 ? 256 : 512);   // two slots, power of two
  static constexpr int SLOT_STRIDE = TMEM_COLS / 2;
  static_assert(E % 2 == 0 && E + 2 <= 32 && D + 1 <= 32 && D % 4 == 0, "row shapes");
  static_assert(B <= 128 && NP <= 128 && SLOT_COLS <= 256, "tile shape");
  static_assert(7 * TILE + 128 * 128 <= TOTAL, "the M = 128 overrun of the last K-major tile stays inside the allocation");
  static_assert(TOTAL + 1024 <= 227 * 1024, "shared memory");
};

// timeline probe events (trace.cuh)
enum FtEv { EV_E_SREADY, EV_E_PREADY, EV_E_ODONE, EV_E_OUT, EV_P_QKFREE, EV_P_QKFULL, EV_P_VFREE, EV_P_VFULL, EV_P_ISSUED,
            EV_M_S_GO, EV_M_S_ISSUED, EV_M_PV_GO, EV_M_PV_ISSUED };
HEPT_TRACE_SETTER(hept_debug_trace_fwd)

// barriers, two of each (stage / slot = tile & 1)
enum FtBar { QKFULL = 0, VFULL = 2, QKFREE = 4, VFREE = 6, SREADY = 8, PREADY = 10, ODONE = 12, FT_NBAR = 14 };

template <int D, int C, int B>
__global__ void __launch_bounds__(kFtThreads, 1)
    block_attn_fwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ hatc, const int32_t* __restrict__ positions, int N, int H, int T,
                             int raw_size, int total_tiles, TileDecoder dec, float* __restrict__ stage) {
  using CF = TcFwd<D, C, B>;
  constexpr int E = CF::E, NP = CF::NP, KSTEPS = CF::KSTEPS, VCH = CF::VCH, PASSES = CF::PASSES, RPP = CF::RPP;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  float* s_nq2 = reinterpret_cast<float*>(smem + CF::OFF_AUX);   // [4][128]  log2e * -|q'|^2 / 2, by tile & 3
  int* s_qidx = reinterpret_cast<int*>(s_nq2 + 512);             // [8][128]  original hit index of query row r, by tile & 7
  __shared__ uint64_t mbar[FT_NBAR];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int EW = kFtEpiThreads / 32, PW = kFtProdThreads / 32;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&mbar[QKFULL + s], PW);
      umma::mbar_init(&mbar[VFULL + s], kFtVWarps);    // the value warps own the value operand
      umma::mbar_init(&mbar[QKFREE + s], 1);
      umma::mbar_init(&mbar[VFREE + s], 1);
      umma::mbar_init(&mbar[SREADY + s], 1);
      umma::mbar_init(&mbar[PREADY + s], EW);
      umma::mbar_init(&mbar[ODONE + s], 1);
    }
  }
  if (warp == EW + PW) umma::tmem_alloc<CF::TMEM_COLS>(&tmem_slot);
  if (warp >= EW && warp < EW + PW) {
    // rows that never hold data are written once: zero operands, key norm -1e30 (P = ex2(-1e30) = 0)
    const int ptid = tid - kFtEpiThreads, sub = ptid >> 3, c = ptid & 7;
    for (int rr = B + sub; rr < CF::KC; rr += RPP) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 kz = z;
      if (c == CF::SLOT_CH) umma::elem<CF::SLOT_U>(kz) = umma::tf32_hi(-1e30f);
#pragma unroll
      for (int st = 0; st < 2; ++st) {
#pragma unroll
        for (int tl = 0; tl < 4; ++tl)
          *reinterpret_cast<float4*>(smem + (st * 4 + tl) * CF::TILE + umma::sw128_offset(rr, c)) = tl == CF::KH ? kz : z;
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
          *reinterpret_cast<float4*>(smem + CF::OFF_V + (st * 2 + tl) * CF::TILE + umma::sw128b32_offset(rr, c)) = z;
      }
    }
    // value tiles: only chunks [0, D / 4) of a row change from tile to tile; chunk D / 4 holds the ones column
    // (O[:, D] = row sum of P) and the rest is zero
    for (int rr = sub; rr < B; rr += kFtProdThreads / 8) {
      if (c >= VCH) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(c == VCH ? 1.f : 0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          *reinterpret_cast<float4*>(smem + CF::OFF_V + (st * 2 + 0) * CF::TILE + umma::sw128b32_offset(rr, c)) = one;
          *reinterpret_cast<float4*>(smem + CF::OFF_V + (st * 2 + 1) * CF::TILE + umma::sw128b32_offset(rr, c)) = z;
        }
      }
    }
    for (int i = ptid; i < 4 * 128; i += kFtProdThreads) s_nq2[i] = 0.f;
    for (int i = ptid; i < 8 * 128; i += kFtProdThreads) s_qidx[i] = -1;
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = umma::smem_u32(smem);

  // tiles are ordered (head, table, block): the CTAs of a wave work on one head's rows, which stay in L2
  const TileDecoder decode = dec;

  if (warp < EW) {
    // =========================================== epilogue warps =================================================
    umma::setmaxnreg_inc<kFtRegsEpi>();
    const int row = (warp & 3) * 32 + lane;                // TMEM lane
    const int part = warp >> 2;                            // column part of that lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    constexpr int MAXCH = (KSTEPS + kFtParts - 1) / kFtParts;   // 8-column chunks per thread

    // P of tile number `it` (slot it & 1): S0 + S1 -> (P hi, P lo) in place
    auto make_p = [&](int it) {
      const int sl = it & 1;
      const uint32_t tS0 = tmem + sl * CF::SLOT_STRIDE, tS1 = tS0 + NP;
      umma::mbar_wait_parked(&mbar[SREADY + sl], (it >> 1) & 1);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(EV_E_SREADY, it);
      const float nq2 = s_nq2[(it & 3) * 128 + row];
      // (issuing the loads of chunk ci + 1 before chunk ci is turned into P measured 2 % SLOWER: the TMEM port is shared
      // with the P V MMAs of the other slot, spreading the loads does not help)
#pragma unroll
      for (int ci = 0; ci < MAXCH; ++ci) {
        const int ch = part + ci * kFtParts;
        if (ch < KSTEPS) {
          uint32_t xa[8], xb[8];
          umma::tmem_ld8_nowait(tS0 + lane_base + 8 * ch, xa);
          umma::tmem_ld8_nowait(tS1 + lane_base + 8 * ch, xb);
          umma::tmem_wait_ld(xa, xb);
          float ph[8], pl[8];
          // two elements per issue slot where the ISA has packed fp32 (add, fma: sm_100); same IEEE results per element
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(xa[u]), __uint_as_float(xa[u + 1])),
                                         make_float2(__uint_as_float(xb[u]), __uint_as_float(xb[u + 1])));
            const float2 x2 = __ffma2_rn(s2, make_float2(kLog2e, kLog2e), make_float2(nq2, nq2));
            const float p0 = exp2_fast(fminf(x2.x, 0.f)), p1 = exp2_fast(fminf(x2.y, 0.f));   // exp(min(S, 0)), example/hept.py:12
            ph[u] = __uint_as_float(__float_as_uint(p0) & 0xffffe000u);
            ph[u + 1] = __uint_as_float(__float_as_uint(p1) & 0xffffe000u);
            const float2 l2 = __ffma2_rn(make_float2(ph[u], ph[u + 1]), make_float2(-1.f, -1.f), make_float2(p0, p1));   // p - hi, exact
            pl[u] = l2.x;
            pl[u + 1] = l2.y;
          }
          umma::tmem_st8(tS0 + lane_base + 8 * ch, ph);
          umma::tmem_st8(tS1 + lane_base + 8 * ch, pl);
        }
      }
      umma::tmem_wait_st();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[PREADY + sl]);
      if (warp == 0) HEPT_TRACE_EVENT(EV_E_PREADY, it);
    };

    int it = 0;
    // (tried: the read-out below on the producer warps, which wait for the score MMAs most of a tile -- their wait for
    // P V then delays the next tile's operands, and three parties on the TMEM port stretch the P V MMAs: 283 us against 271)
    if ((int)blockIdx.x < total_tiles) make_p(0);
#pragma unroll 1
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if (tile + (int)gridDim.x < total_tiles) make_p(it + 1);   // the other slot: runs while P V of this tile executes
      const int sl = it & 1;
      int h, t, blk;
      decode(tile, h, t, blk);
      // ---- numerator and normaliser back to original hit order: part 0 writes columns [0,16), part 1 the rest ------
      umma::mbar_wait_parked(&mbar[ODONE + sl], (it >> 1) & 1);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(EV_E_ODONE, it);
      float ov[16];
      umma::tmem_ld16(tmem + sl * CF::SLOT_STRIDE + 2 * NP + lane_base + 16 * part, ov);
      umma::fence_before_sync();                       // these loads precede the P V MMAs of tile it + 2 into the same columns
      const int n = row < B ? s_qidx[(it & 7) * 128 + row] : -1;
      if (n >= 0) {
        if (part == D / 16) ov[D % 16] += 1e-20f;      // column D: denom = rowsum + 1e-20 (example/hept.py:14)
        float4* dst = reinterpret_cast<float4*>(stage + (((size_t)h * N + n) * T + t) * kStageRow);
        constexpr int USED = (D + 1 + 3) / 4;           // 16-byte chunks that carry data
        constexpr int WR = (USED + 1) / 2 * 2;          // keep 32-byte sectors whole
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          if (4 * part + cc < WR) dst[4 * part + cc] = make_float4(ov[4 * cc], ov[4 * cc + 1], ov[4 * cc + 2], ov[4 * cc + 3]);
      }
      if (warp == 0) HEPT_TRACE_EVENT(EV_E_OUT, it);
    }
  } else if (warp < EW + PW) {
    // =========================================== producer warps =================================================
    umma::setmaxnreg_dec<kFtRegsProd>();
    const int ptid = tid - kFtEpiThreads, sub = ptid >> 3, c = ptid & 7;   // 8 lanes per row, RPP rows per pass
    int nk_idx[PASSES], nq_idx[PASSES], n0 = 0;
    float4 xq[PASSES], xk[PASSES], ctr;

    auto load_indices = [&](int tile) {
      int h, t, blk;
      decode(tile, h, t, blk);
      const int th = t * H + h;
      const int32_t* qpos = positions + (size_t)th * N + (size_t)blk * B;
      const int32_t* kpos = positions + ((size_t)T * H + th) * N + (size_t)blk * B;
      n0 = __ldg(kpos + (B - 1));
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int r = ps * RPP + sub;
        nk_idx[ps] = r < B ? __ldg(kpos + r) : -1;
        nq_idx[ps] = r < B ? __ldg(qpos + r) : -1;
      }
    };
    // issue every row load of a tile (registers); chunk c of a hat row: feature columns from q / k, coordinate
    // columns from hat_coords (already scaled)
    // The q^ / k^ loads of the next tile are issued as soon as this tile's have been consumed (before the value phase and
    // its wait), the value loads after the value phase: every load gets as much flight time as its registers allow.
    auto issue_qk = [&](int tile, int it) {
      int h, t, blk;
      decode(tile, h, t, blk);
      // chunk c of a hat row: feature columns from q / k, coordinate columns from hat_coords.  The per-lane base and
      // row stride are fixed for the tile, so a load is one multiply-add and one predicated LDG (no divergent paths).
      const size_t hstride = c < VCH ? (size_t)H * D : (size_t)H * 8;
      const size_t hoff = c < VCH ? (size_t)h * D + 4 * c : (size_t)h * 8 + 4 * (c - VCH);
      auto hat_chunk = [&](const float* __restrict__ x, int n) -> float4 {
        const float* src = (c < VCH ? x : hatc) + hoff + (size_t)(n < 0 ? 0 : n) * hstride;
        return (c < VCH + 2 && n >= 0 && n < raw_size) ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      ctr = hat_chunk(k, n0);
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int nkk = nk_idx[ps], nqq = nq_idx[ps];
        xk[ps] = hat_chunk(k, nkk);
        xq[ps] = hat_chunk(q, nqq);
        const int r = ps * RPP + sub;
        if (c == 0 && r < B) s_qidx[(it & 7) * 128 + r] = nqq;
      }
    };
    int tile = blockIdx.x;
    if (tile < total_tiles) {
      load_indices(tile);
      issue_qk(tile, 0);
      if (tile + (int)gridDim.x < total_tiles) load_indices(tile + gridDim.x);
    }
    int it = 0;
#pragma unroll 1
    for (; tile < total_tiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      uint8_t* km = smem + st * 4 * CF::TILE;
      // ---- q^ / k^ tiles of this stage: free once the score MMAs of tile it - 2 are done ---------------------------
      if (it >= 2) umma::mbar_wait_parked(&mbar[QKFREE + st], ph ^ 1);
      if (warp == EW) HEPT_TRACE_EVENT(EV_P_QKFREE, it);
      float* nq2s = s_nq2 + (it & 3) * 128;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int r = ps * RPP + sub;
        if (ps == PASSES - 1 && (ps * RPP + ((warp - EW) << 2)) >= B) continue;   // warps whose 4 rows are all past B
        const bool in = r < B;
        const uint32_t okm = umma::sw128_offset(r, c);
        float4 hi, lo;
        {  // key row: k' = k^ - centre; side slots carry nk = -|k'|^2 / 2 split three ways (exact)
          float4 d = xk[ps];
          d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
          const float nk = -0.5f * tree8_lanes(chunk_sq<E>(d, c));
          split4(d, hi, lo);
          if (c == CF::SLOT_CH) {
            const float h0 = umma::tf32_hi(nk), r1 = nk - h0, h1 = umma::tf32_hi(r1);
            umma::elem<CF::SLOT_U>(hi) = h0; umma::elem<CF::SLOT_U + 1>(hi) = h1;
            umma::elem<CF::SLOT_U>(lo) = r1 - h1; umma::elem<CF::SLOT_U + 1>(lo) = 0.f;
          }
          if (in) {
            *reinterpret_cast<float4*>(km + CF::KH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(km + CF::KL * CF::TILE + okm) = lo;
          }
        }
        {  // query row: q' = q^ - centre; side slots carry 1
          float4 d = xq[ps];
          d.x -= ctr.x; d.y -= ctr.y; d.z -= ctr.z; d.w -= ctr.w;
          const float nq2 = kLog2e * (-0.5f * tree8_lanes(chunk_sq<E>(d, c)));
          split4(d, hi, lo);
          if (c == CF::SLOT_CH) {
            umma::elem<CF::SLOT_U>(hi) = 1.f; umma::elem<CF::SLOT_U + 1>(hi) = 1.f;
            umma::elem<CF::SLOT_U>(lo) = 0.f; umma::elem<CF::SLOT_U + 1>(lo) = 0.f;
          }
          if (in) {
            *reinterpret_cast<float4*>(km + CF::QH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(km + CF::QL * CF::TILE + okm) = lo;
            if (c == 0) nq2s[r] = nq2;
          }
        }
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[QKFULL + st]);
      if (warp == EW) HEPT_TRACE_EVENT(EV_P_QKFULL, it);
      const int next = tile + gridDim.x;
      if (next < total_tiles) issue_qk(next, it + 1);

      // ---- registers are free: fetch the indices of the tile after the next (its rows are already in flight) ------
      if (next < total_tiles && next + (int)gridDim.x < total_tiles) load_indices(next + gridDim.x);
      if (warp == EW) HEPT_TRACE_EVENT(EV_P_ISSUED, it);
    }
  } else {
    umma::setmaxnreg_dec<kFtRegsMma>();
  }
  if (warp > EW + PW) {
    // =========================================== value warps ===================================================
    const int vt = tid - (kFtEpiThreads + kFtProdThreads + 32);
    // ---- the value operand: raw v rows travel global -> shared with cp.async (no registers) four tiles before they are
    // used; thread <-> items (row r, chunk c < D / 4) idx = vt, vt + 96, ...; a thread splits its own items into (hi, lo)
    // and writes the MN-major tiles of stage it & 1 as soon as P V of tile it - 2 is done (VFREE).  Round 1 gave this to the
    // epilogue warps, when the producers bounded the kernel; with 16 producer warps the epilogue became the serial chain
    // (pipeline trace: 2 000 of its 6 200 cycles per tile went here while S of the next tile sat ready in TMEM), and the MMA
    // warp's three siblings were idle.
    constexpr int ITEMS = B * VCH, VPER = (ITEMS + kFtVThreads - 1) / kFtVThreads;
    const int g = (int)gridDim.x;
    int hit[VPER];                                     // key-side hit index of this thread's items, loaded one step before `v_issue`
    auto v_hits = [&](int tile) {
      int h, t, blk;
      decode(tile, h, t, blk);
      const int32_t* kpos = positions + ((size_t)T * H + t * H + h) * N + (size_t)blk * B;
#pragma unroll
      for (int u = 0; u < VPER; ++u) {
        const int idx = vt + u * kFtVThreads;
        hit[u] = idx < ITEMS ? __ldg(kpos + idx / VCH) : -1;
      }
    };
    auto v_issue = [&](int tile, int set) {            // uses hit[] loaded for `tile`; always commits one group
      if (tile < total_tiles) {
        int h, t, blk;
        decode(tile, h, t, blk);
        float4* stg = reinterpret_cast<float4*>(smem + CF::OFF_VSTG + set * CF::VSTG);
        const float* vh = v + (size_t)h * D;
#pragma unroll
        for (int u = 0; u < VPER; ++u) {
          const int idx = vt + u * kFtVThreads;
          if (idx < ITEMS) {
            const int n = hit[u];
            const bool ok = n < raw_size;
            umma::cp_async16(stg + idx, ok ? vh + (size_t)n * H * D + 4 * (idx % VCH) : v, ok ? 16 : 0);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto v_write = [&](int vit) {                      // value tiles of tile number `vit` (stage vit & 1) from staging set vit & 1
      const int st = vit & 1;
      uint8_t* vm = smem + CF::OFF_V + st * 2 * CF::TILE;
      const float4* stg = reinterpret_cast<const float4*>(smem + CF::OFF_VSTG + st * CF::VSTG);
#pragma unroll
      for (int u = 0; u < VPER; ++u) {
        const int idx = vt + u * kFtVThreads;
        if (idx < ITEMS) {
          float4 hi, lo;
          split4(stg[idx], hi, lo);
          const uint32_t omn = umma::sw128b32_offset(idx / VCH, idx % VCH);
          *reinterpret_cast<float4*>(vm + omn) = hi;
          *reinterpret_cast<float4*>(vm + CF::TILE + omn) = lo;
        }
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[VFULL + st]);
      if (warp == EW + PW + 1) HEPT_TRACE_EVENT(EV_P_VFULL, vit);
    };
    // prologue: rows of tiles 0, 1 (written before the loop), 2, 3 (in flight); groups are committed in tile order
    const int tile0 = blockIdx.x;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (tile0 + j * g < total_tiles) v_hits(tile0 + j * g);
      if (j == 2) {                                    // sets are reused: tiles 0 and 1 must be out of them first
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tile0 < total_tiles) v_write(0);
        if (tile0 + g < total_tiles) v_write(1);
      }
      v_issue(tile0 + j * g, j & 1);
    }
    if (tile0 + 4 * g < total_tiles) v_hits(tile0 + 4 * g);

    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if (tile + 2 * g < total_tiles) {
        umma::mbar_wait_parked(&mbar[VFREE + (it & 1)], (it >> 1) & 1);   // P V of this tile is done: its value tiles are free
        umma::fence_after_sync();
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // rows of tile it + 2 have landed (it + 3 may be in flight)
        v_write(it + 2);
      }
      v_issue(tile + 4 * g, it & 1);                             // hit[] holds tile it + 4
      if (tile + 5 * g < total_tiles) v_hits(tile + 5 * g);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (warp == EW + PW) {
    // =========================================== MMA issuer ====================================================
    // the warp runs the schedule uniformly; the lane chosen by elect.sync issues (always the same lane, so its
    // tcgen05.commit covers every MMA issued before)
    constexpr uint32_t idesc_s = umma::idesc_tf32(128, NP, false, false);
    constexpr uint32_t idesc_o = umma::idesc_tf32(128, 32, false, true);
    const uint64_t kdesc0 = umma::smem_desc_sw128(sbase, 1024, 16);
    const uint64_t vdesc0 = umma::smem_desc(sbase + CF::OFF_V, 512, 1024, umma::kLayoutSw128Base32);
    auto wait = [&](int b, uint32_t parity) {
      umma::mbar_wait_parked(&mbar[b], parity);
      umma::fence_after_sync();
    };
    // S = Q^ K^^T of tile number `it`: cross terms and the first SK0 hi*hi k-steps into S0, the rest into S1
    auto scores = [&](int it) {
      const int st = it & 1;
      const uint32_t tS0 = tmem + st * CF::SLOT_STRIDE, tS1 = tS0 + NP;
      if (umma::elect_one()) {
        uint64_t base = kdesc0 + (uint64_t)((st * 4 * CF::TILE) >> 4);
        asm volatile("" : "+l"(base));   // opaque: descriptors are re-derived here (one add each), not hoisted and spilled
        const uint64_t qh = base + ((CF::QH * CF::TILE) >> 4), ql = base + ((CF::QL * CF::TILE) >> 4),
                       kh = base + ((CF::KH * CF::TILE) >> 4), kl = base + ((CF::KL * CF::TILE) >> 4);
#pragma unroll
        for (int kk = 0; kk < CF::SK; ++kk) umma::mma_ss(tS0, qh + 2 * kk, kl + 2 * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < CF::SK; ++kk) umma::mma_ss(tS0, ql + 2 * kk, kh + 2 * kk, idesc_s, true);
#pragma unroll
        for (int kk = 0; kk < CF::SK; ++kk)
          umma::mma_ss(kk < CF::SK0 ? tS0 : tS1, qh + 2 * kk, kh + 2 * kk, idesc_s, kk != CF::SK0);
        umma::commit(&mbar[SREADY + st]);
        umma::commit(&mbar[QKFREE + st]);
      }
      __syncwarp();
    };
    int it = 0;
    int tile = blockIdx.x;
    HEPT_TRACE_CTA(0);
    HEPT_TRACE_CTA(2);
    if (tile < total_tiles) {
      wait(QKFULL + 0, 0);
      scores(0);
    }
#pragma unroll 1
    for (; tile < total_tiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      if (tile + (int)gridDim.x < total_tiles) {
        // S of the next tile goes into the other slot: its P (tile it - 1) must have been consumed by P V(it - 1)
        wait(QKFULL + (st ^ 1), ((it + 1) >> 1) & 1);
        if (it >= 1) wait(ODONE + (st ^ 1), ((it - 1) >> 1) & 1);
        HEPT_TRACE_EVENT(EV_M_S_GO, it + 1);
        scores(it + 1);
        HEPT_TRACE_EVENT(EV_M_S_ISSUED, it + 1);
      }
      wait(PREADY + st, ph);
      wait(VFULL + st, ph);
      HEPT_TRACE_EVENT(EV_M_PV_GO, it);
      if (umma::elect_one()) {
        const uint32_t tS0 = tmem + st * CF::SLOT_STRIDE, tS1 = tS0 + NP, tO = tS0 + 2 * NP;
        uint64_t base = vdesc0 + (uint64_t)((st * 2 * CF::TILE) >> 4);
        asm volatile("" : "+l"(base));
#pragma unroll
        for (int p3 = 0; p3 < 3; ++p3) {                 // P_hi V_lo, P_lo V_hi, P_hi V_hi
          const uint32_t a = p3 == 1 ? tS1 : tS0;
          const uint64_t db = base + (p3 == 0 ? (uint64_t)(CF::TILE >> 4) : 0);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) umma::mma_ts(tO, a + 8 * kk, db + 64 * kk, idesc_o, (p3 | kk) != 0);
        }
        umma::commit(&mbar[ODONE + st]);
        umma::commit(&mbar[VFREE + st]);
      }
      __syncwarp();
      HEPT_TRACE_EVENT(EV_M_PV_ISSUED, it);
    }
    HEPT_TRACE_CTA(1);
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == EW + PW) umma::tmem_dealloc<CF::TMEM_COLS>(tmem);
}

template <int D, int C, int B>
static int launch_fwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* hatc,
                         const int32_t* positions, float* stage, cudaStream_t st) {
  using CF = TcFwd<D, C, B>;
  auto kern = block_attn_fwd_tc_kernel<D, C, B>;
  const size_t smem = CF::TOTAL + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "block_attn_fwd_tc: cannot reserve %zu B of shared memory: %s", smem,
                 cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "block_attn_fwd_tc: cannot read the SM count");
  const int tiles = s->T * s->H * (s->N / s->B);
  const int grid = tiles < sms ? tiles : sms;   // one CTA per SM (two tiles in flight use all 512 TMEM columns)
  HEPT_REQUIRE(TileDecoder::exact_for(tiles, s->N / s->B, s->T), HEPT_EUNSUPPORTED, "block_attn_fwd_tc: too many tiles (%d)", tiles);
  kern<<<grid, kFtThreads, smem, st>>>(q, k, v, hatc, positions, s->N, s->H, s->T, s->raw_size, tiles,
                                       TileDecoder::make(s->N / s->B, s->T), stage);
  HEPT_CHECK_LAUNCH("block_attn_fwd_tc");
  return HEPT_OK;
}

int block_attention_fwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* hatc,
                           const int32_t* positions, float* stage, cudaStream_t st) {
  if (s->D == 24 && s->C == 6 && s->B == 100) return launch_fwd_tc<24, 6, 100>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 100) return launch_fwd_tc<24, 4, 100>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 24 && s->C == 6 && s->B == 64) return launch_fwd_tc<24, 6, 64>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 64) return launch_fwd_tc<24, 4, 64>(s, q, k, v, hatc, positions, stage, st);
  if (s->D == 8 && s->C == 6 && s->B == 10) return launch_fwd_tc<8, 6, 10>(s, q, k, v, hatc, positions, stage, st);
  set_error("block_attention_fwd (tensor-core engine): (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}

}  // namespace hept
