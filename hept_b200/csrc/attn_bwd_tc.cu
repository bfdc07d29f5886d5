// a18 on the 5th-generation tensor cores: backward of the block-local attention tiles (autograd of
// example/hept.py:7-18,70-79) with every contraction of a tile issued as tcgen05.mma (kind::tf32, M = 128,
// 3xTF32 operand splits, accumulators in TMEM).  One persistent, warp-specialised CTA per SM walks the tiles; a tile
// is one block of B <= 104 sorted hits of one (table, head) and is processed from BOTH sides, because the tensor
// core wants the reduced-over index on the TMEM columns:
//
//   query side (TMEM lane = query i)                     key side (TMEM lane = key j)
//   S   = Q^ K^^T            SS, N = NP                   S^T  = K^ Q^^T            SS (operands exchanged)
//   dP  = G' V'^T            SS                           dP^T = V' G'^T            SS
//   dS  = P dP  -> TMEM (hi, lo)                          P^T -> TMEM (hi, lo), dS^T kept in registers
//   dQ  = dS [K^ | 1]        TS, B = K^ MN-major          dV   = P^T G'             TS, B = G' MN-major
//   dq^_i = dQ_i - rs_i q'_i                              dK   = dS^T [Q^ | 1]      TS (dS^T written where dS was, once dQ is done)
//                                                         dk^_j = dK_j - cs_j k'_j
// The TS products are N = 64 MMAs over [hi | lo] operand tiles: per k-step A_hi meets [B_hi | B_lo] once (columns [0,32) =
// hi*hi, [32,64) = hi*lo) and A_lo meets B_hi in an N = 32 MMA into the upper half; the epilogue adds the halves.
// Order on the tensor pipe: dV (accumulator tO), dQ (accumulator over the P^T columns dV has consumed), dK (A = dS^T in
// the dS columns, accumulator tO again once dv has been read out), then the next tile's scores.
// The row / column sums rs_i = sum_j dS_ij, cs_j = sum_i dS_ij come out of the same MMAs (a ones column in the
// MN-major operand), so they are the sums of exactly the (hi, lo) values that produced dQ / dK.
// The reference's clamp(max=0) passes gradient only where S <= 0 (example/hept.py:12, ClampBackward): dS = [x <= 0] P dP with
// the x this kernel computes (centred rows: x > 0 happens only by rounding, for coincident points), on both sides alike.
// with x = log2e (q'.k' - |k'|^2/2) - log2e |q'|^2/2 (the key norm rides in two spare K slots of the contraction,
// split three ways so it is exact), P = ex2(min(x, 0)), G'_i = [g_i / den_i, -(g_i . y_i) / den_i], V'_j = [v_j, 1]
// (so dP = gd . v - gy comes out of one contraction).  The tf32 MMA is bitwise symmetric under an exchange of its
// operands (tests/test_gpu_umma.py) and both sides apply the same fmaf, so they see bit-identical P and dS.
//
// Warp roles (640 threads = 5 warpgroups, register budgets rebalanced with setmaxnreg):
//   warps 0-7   epilogue: TMEM lane = (warp & 3) * 32 + lane, columns split in two parts (warp >> 2)
//   warps 8-15  producer: gather rows through the sort permutation (q^, k^ into registers, v and G' with cp.async into
//               a staging buffer, all issued one tile ahead), centre, split, write the K-major operand tiles once the
//               previous tile's dV is done (earlier would be legal, but the conversion and the TS MMAs contend for
//               shared memory), copy them to the MN-major tiles once the previous tile's last MMA is done
//   warp 16     one thread issues every tcgen05.mma (warps 17-19 only give their registers away); the next tile's
//               score MMAs are issued as soon as their TMEM columns are free
// Hand-offs are mbarriers (tcgen05.commit on the MMA side, one arrive per warp on the others); each completes once
// per tile, so the wait parity is the tile counter's low bit.
//
// d scale[h,c] = sum_n coords[n,c] (dq^ + dk^)[n,h,D+c] cancels by ~|x|^2 / |x_i - x_j|^2 when summed over hits in
// detector coordinates.  Per tile sum_i dq^_i + sum_j dk^_j = 0, so the same sum may be taken with the tile's
// CENTRED rows, sum_i q'_ic dq^_ic + sum_j k'_jc dk^_jc = scale_c * (the tile's share of d scale_c): no large
// terms, nothing to cancel.  Per-tile partials are reduced in a fixed order (deterministic).
#include "tile.cuh"
#include "trace.cuh"
#include "umma.cuh"

namespace hept {

constexpr int kBtEpiThreads = 256, kBtProdThreads = 256;
constexpr int kBtThreads = kBtEpiThreads + kBtProdThreads + 128;
// Launch budget 65536 / 640 -> 96 registers per thread = 61440 per CTA; setmaxnreg moves registers inside that pool:
// 256 x 128 (epilogue) + 256 x 88 (producer) + 128 x 40 (MMA warpgroup) = 60416.
constexpr int kBtRegsEpi = 128, kBtRegsProd = 88, kBtRegsMma = 40;
static_assert(kBtEpiThreads * kBtRegsEpi + kBtProdThreads * kBtRegsProd + 128 * kBtRegsMma <= kBtThreads * 96, "register pool");
constexpr int kBtParts = kBtEpiThreads / 128;   // epilogue warps w and w+4 share TMEM lane quarter w & 3

constexpr int pow2_cols(int need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

template <int D, int C, int B>
struct TcBwd {
  static constexpr int E = D + C;
  static constexpr int NP = (B + 15) / 16 * 16;   // N of the score MMAs (multiple of 16 at M = 128)
  static constexpr int KC = (B + 7) / 8 * 8;      // contraction length of the TS MMAs
  static constexpr int KSTEPS = KC / 8;
  static constexpr int SK = (E + 2 + 7) / 8;      // k-steps of Q^ K^^T (two side slots at E, E + 1)
  static constexpr int GK = (D + 1 + 7) / 8;      // k-steps of G' V'^T
  static constexpr int VCH = D / 4;
  static constexpr int SLOT_CH = E / 4, SLOT_U = E % 4;   // chunk / element of side slot E
  // Operand tiles hold KC rows of 128 B.  The MMAs read M = 128 (A) or NP (B) rows: what lies past row KC belongs to
  // the next tile and only reaches TMEM lanes / columns >= KC, which nothing reads.
  static constexpr int TILE = KC * 128;
  static constexpr int QH = 0, QL = 1, KH = 2, KL = 3, GH = 4, GL = 5, VH = 6, VL = 7;   // K-major tiles
  static constexpr int MKH = 0, MKL = 1, MGH = 2, MGL = 3, MQH = 4, MQL = 5;             // MN-major tiles
  static constexpr int RPP = kBtProdThreads / 8;   // producer: rows per pass (8 lanes per row)
  static constexpr int PASSES = (B + RPP - 1) / RPP;
  static constexpr int OFF_MN = 8 * TILE;
  static constexpr int OFF_STG_V = OFF_MN + 6 * TILE;
  static constexpr int OFF_STG_G = OFF_STG_V + PASSES * kBtProdThreads * 16;
  static constexpr int OFF_AUX = OFF_STG_G + PASSES * kBtProdThreads * 16;
  // aux words: nq2[2][128], qidx[3][128], kidx[3][128], dsc[8][8]
  static constexpr int AUX_BYTES = (8 * 128 + 64) * 4;
  static constexpr int TOTAL = OFF_AUX + AUX_BYTES;
  // dQ accumulates over 64 columns from the start of the P^T region (2 NP): the dV / dK accumulator must start past it, or
  // dQ lands on dV before the epilogue has read it (NP < 32, i.e. blocks of at most 16 hits)
  static constexpr int O_COL = 4 * NP > 2 * NP + 64 ? 4 * NP : 2 * NP + 64;
  static constexpr int TMEM_NEED = O_COL + 64;
  static constexpr int TMEM_COLS = pow2_cols(TMEM_NEED);
  static_assert(E % 2 == 0 && E + 2 <= 32 && D + 1 <= 32 && D % 4 == 0, "row shapes");
  static_assert(B <= 128 && NP <= 128 && TMEM_NEED <= 512, "tile shape");
  static_assert(7 * TILE + 128 * 128 <= TOTAL, "the M = 128 overrun of the last K-major tile stays inside the allocation");
  static_assert(TOTAL + 1024 <= 227 * 1024, "shared memory");
};

// timeline probe events (trace.cuh)
enum BtEv { BE_QREADY, BE_DSRDY, BE_KREADY, BE_PTRDY, BE_DQDONE, BE_DQOUT, BE_DVDONE, BE_DSTRDY, BE_DVOUT, BE_DKDONE, BE_END,
            BP_KFREE, BP_KFULL, BP_ISSUED, BP_MFREE, BP_MFULL, BM_DQ_GO, BM_DV_GO, BM_DK_GO, BM_SQ_GO, BM_SK_GO, BM_END,
            BE_DSTISS, BE_DVWAIT, BE_DQACC, BE_DKACC };
HEPT_TRACE_SETTER(hept_debug_trace_bwd)

#ifndef HEPT_BWD_PROD_WAIT
#define HEPT_BWD_PROD_WAIT DVDONE
#endif
#ifndef HEPT_BWD_SPLIT
#define HEPT_BWD_SPLIT 0
#endif
// HEPT_BWD_SPLIT=1 (experiment, `make VARIANT=split EXTRA=-DHEPT_BWD_SPLIT=1`, tools/ab_split.sh): the three TS products start
// on the FIRST HALF of their A operand (k-steps [0, HK)) while the epilogue still writes the second half, and dS^T is written
// back in two halves behind dQ's two halves.  Correct (same bits: the parity suite passes), but SLOWER: 834 us against 805 at
// 60k hits.  The MMAs that now run under the epilogue's TMEM traffic are stretched by more than the overlap hides — the
// thread-side TMEM port, not the dependency chain, is what bounds the tile.
enum BtBar { KFULL, MFULL, MFREE, QREADY, KREADY, DSRDY, DQDONE, PTRDY, DVDONE, DSTRDY, DKDONE, STFREE, PTHALF, DSHALF, DQHALF, DSTHALF, BT_NBAR };

using umma::split4;
using umma::split_tf32;
using umma::trunc_tf32;

// ---------------------------------------------------------------------------------------------------------------
// G' rows (N, H, 32): [g / den (D), -(g . y) / den, 0 ...] — the gradient arriving at numerator and normaliser of
// every table (d so_t = g / den, d denom_t = -(g . y) / den, attn_bwd.cu).  Eight lanes per row.
// ---------------------------------------------------------------------------------------------------------------
// The same launch also writes hat_coords (N, H, 8) = scale[h,c] * coords[n,c] (hash.cu, hat_coords_kernel): lanes 0 and 1
// of a row write its two 16-byte chunks, which saves a 15 us streaming launch per backward.
template <int D>
__global__ void __launch_bounds__(256) grad_rows_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                        const float* __restrict__ den, size_t rows,
                                                        float* __restrict__ out, const float* __restrict__ coords,
                                                        const float* __restrict__ scale, int H, int C, int raw_size,
                                                        float* __restrict__ hat, uint32_t* __restrict__ zero, int zero_words) {
  constexpr int VCH = D / 4;
  constexpr int RPT = 2;                                 // rows per thread: both rows' loads are in flight before the first use
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the tile kernel's per-CTA d scale partials and (table, head) tile counters start at zero (was a memset node)
  for (size_t z = idx; z < (size_t)zero_words; z += (size_t)gridDim.x * blockDim.x) zero[z] = 0u;
  const size_t half = (rows + RPT - 1) / RPT;
  const int c = (int)(idx & 7);
  size_t r[RPT];
  bool live[RPT];
  float dn[RPT];
  float4 gg[RPT], yy[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    r[u] = (idx >> 3) + u * half;
    live[u] = (idx >> 3) < half && r[u] < rows;
    dn[u] = 1.f;
    gg[u] = yy[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live[u] && c < VCH) {
      dn[u] = __ldg(den + r[u]);
      gg[u] = ldg4(g + r[u] * D + 4 * c);
      yy[u] = ldg4(y + r[u] * D + 4 * c);
    }
  }
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    if (live[u] && c < 2) {
      const size_t n = r[u] / H;
      const int h = (int)(r[u] - n * H);
      float hv[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int cc = 4 * c + w;
        hv[w] = (cc < C && (int)n < raw_size) ? __fmul_rn(__ldg(scale + h * C + cc), __ldg(coords + n * C + cc)) : 0.f;
      }
      *reinterpret_cast<float4*>(hat + r[u] * 8 + 4 * c) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    float4 gd = make_float4(0.f, 0.f, 0.f, 0.f);
    float part = 0.f;
    if (live[u] && c < VCH) {
      const float inv = 1.f / dn[u];
      gd = make_float4(gg[u].x * inv, gg[u].y * inv, gg[u].z * inv, gg[u].w * inv);
      part = fmaf(gd.w, yy[u].w, fmaf(gd.z, yy[u].z, fmaf(gd.y, yy[u].y, gd.x * yy[u].x)));
    }
    const float gy = tree8_lanes(part);
    if (c == VCH) gd.x = -gy;
    if (live[u]) *reinterpret_cast<float4*>(out + r[u] * 32 + 4 * c) = gd;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------------------
template <int D, int C, int B, bool DIRECT>
__global__ void __launch_bounds__(kBtThreads, 1)
    block_attn_bwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ hatc, const float* __restrict__ grows,
                             const int32_t* __restrict__ positions, int N, int H, int T, int raw_size, int total_tiles,
                             TileDecoder dec, float* __restrict__ stage_dq, float* __restrict__ stage_dk, float* __restrict__ stage_dv,
                             float* __restrict__ ds_partial, int* __restrict__ done) {
  using CF = TcBwd<D, C, B>;
  constexpr int E = CF::E, NP = CF::NP, KSTEPS = CF::KSTEPS, VCH = CF::VCH, PASSES = CF::PASSES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  uint8_t* mn = smem + CF::OFF_MN;
  float* s_nq2 = reinterpret_cast<float*>(smem + CF::OFF_AUX);   // [2][128]  log2e * -|q'|^2 / 2, by tile parity
  int* s_qidx = reinterpret_cast<int*>(s_nq2 + 256);             // [3][128]  original hit index of query row r, tile % 3
  int* s_kidx = s_qidx + 384;                                    // [3][128]
  float* s_dsc = reinterpret_cast<float*>(s_kidx + 384);         // [8][8]    d scale partials per epilogue warp
  __shared__ uint64_t mbar[BT_NBAR];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_dep[2];                                       // DIRECT: (group, tile count seen) fetched at the last group change

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int EW = kBtEpiThreads / 32, PW = kBtProdThreads / 32;

  if (tid == 0) {
    s_dep[0] = -1;
    s_dep[1] = 0;
    umma::mbar_init(&mbar[KFULL], PW);
    umma::mbar_init(&mbar[MFULL], PW);
    umma::mbar_init(&mbar[MFREE], 1 + EW);
    umma::mbar_init(&mbar[QREADY], 1);
    umma::mbar_init(&mbar[KREADY], 1);
    umma::mbar_init(&mbar[DSRDY], EW);
    umma::mbar_init(&mbar[DQDONE], 1);
    umma::mbar_init(&mbar[PTRDY], EW);
    umma::mbar_init(&mbar[DVDONE], 1);
    umma::mbar_init(&mbar[DSTRDY], EW);
    umma::mbar_init(&mbar[DKDONE], 1);
    umma::mbar_init(&mbar[STFREE], EW);
    umma::mbar_init(&mbar[PTHALF], EW);
    umma::mbar_init(&mbar[DSHALF], EW);
    umma::mbar_init(&mbar[DQHALF], 1);
    umma::mbar_init(&mbar[DSTHALF], EW);
  }
  if (warp == EW + PW) umma::tmem_alloc<CF::TMEM_COLS>(&tmem_slot);
  if (warp >= EW && warp < EW + PW) {
    // rows that never hold data are written once: zero operands, key norm -1e30 (P = ex2(-1e30) = 0)
    const int ptid = tid - kBtEpiThreads, sub = ptid >> 3, c = ptid & 7;
    for (int rr = B + sub; rr < CF::KC; rr += kBtProdThreads / 8) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 kz = z;
      if (c == CF::SLOT_CH) umma::elem<CF::SLOT_U>(kz) = umma::tf32_hi(-1e30f);
#pragma unroll
      for (int tl = 0; tl < 8; ++tl)
        *reinterpret_cast<float4*>(smem + tl * CF::TILE + umma::sw128_offset(rr, c)) = tl == CF::KH ? kz : z;
#pragma unroll
      for (int tl = 0; tl < 6; ++tl)
        *reinterpret_cast<float4*>(mn + tl * CF::TILE + umma::sw128b32_offset(rr, c)) = z;
    }
    for (int rr = B + ptid; rr < 128; rr += kBtProdThreads) {
      s_nq2[rr] = -1e30f; s_nq2[128 + rr] = -1e30f;
#pragma unroll
      for (int u = 0; u < 3; ++u) { s_qidx[u * 128 + rr] = -1; s_kidx[u * 128 + rr] = -1; }
    }
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + NP, tST = tmem + 2 * NP, tDPT = tmem + 3 * NP, tO = tmem + CF::O_COL;
  const uint32_t sbase = umma::smem_u32(smem), mbase = sbase + CF::OFF_MN;

  // tiles are ordered (head, table, block): the CTAs of a wave work on one head's rows, which stay in L2
  const TileDecoder decode = dec;

  if (warp < EW) {
    // =========================================== epilogue warps =================================================
    umma::setmaxnreg_inc<kBtRegsEpi>();
    const int row = (warp & 3) * 32 + lane;                // TMEM lane
    const int srow = row < CF::KC ? row : 0;               // lanes past the tile read row 0 of the operand tiles (unused)
    const int part = warp >> 2;                            // column part of that lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    constexpr int MAXCH = (KSTEPS + kBtParts - 1) / kBtParts;   // 8-column chunks per thread
    constexpr int HK = (KSTEPS + 1) / 2;                        // k-steps of the first half of a TS product
    constexpr int HCI = (HK - 1) / kBtParts;                    // after chunk iteration HCI every chunk < HK of this thread is done
    static_assert(HCI < MAXCH, "half split");
    // warp-level arrive: every lane's TMEM stores are complete and fenced before lane 0 signals
    auto arrive_tmem = [&](BtBar b) {
      umma::tmem_wait_st();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[b]);
    };
    // d scale: each thread keeps its share of sum_i q'_ic dq^_ic + sum_j k'_jc dk^_jc over the consecutive tiles of one
    // head; when the head changes the CTA reduces them in a fixed order (warp tree, then warps in order) into
    // ds_partial[cta][head][c].  A CTA visits its tiles in a fixed order, so the result is deterministic.
    float dsc[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) dsc[cc] = 0.f;
    int dsc_head = -1;
    int ready = -1, cur_grp = -1;                        // DIRECT: see below
    auto flush_dscale = [&](int dep = -1) {   // dep: the group whose count the CTA's next tiles will need (DIRECT), or -1
      if (dsc_head < 0) return;
      int dep_cnt = 0;                         // asked for here, used after the barrier: the L2 round trip hides behind the sums
      if (DIRECT && tid == 32 && dep >= 0) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(dep_cnt) : "l"(done + dep) : "memory");
#pragma unroll
      for (int cc = 0; cc < C; ++cc) {
        float x = dsc[cc];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_dsc[warp * 8 + cc] = x;
        dsc[cc] = 0.f;
      }
      umma::bar_sync(1, kBtEpiThreads);
      if (DIRECT && tid == 32) { s_dep[0] = dep; s_dep[1] = dep_cnt; }   // published by the barrier below
      if (tid < C) {
        float x = 0.f;
#pragma unroll
        for (int w = 0; w < EW; ++w) x += s_dsc[w * 8 + tid];
        float* dst = ds_partial + ((size_t)blockIdx.x * H + dsc_head) * 8 + tid;
        *dst = DIRECT ? *dst + x : x;                    // paired order: a head comes back T times (the buffer starts at zero)
      }
      umma::bar_sync(1, kBtEpiThreads);                  // s_dsc may be rewritten
    };
    // DIRECT: rows go straight into dq / dk / dv (n, h, D): table 0 stores, later tables add (16-byte vector atomics), so
    // the per-table staging rows and the kernel that sums them are gone.  Order of the adds = table order, enforced, so
    // the sums are the same bits every run and the same as the staged sum: the CTA's tiles of a (table, head) group are
    // counted into done[t * H + h] with release semantics once their rows are out (by a producer thread, see there), and
    // a tile of table t >= 1 writes its first row once done[(t - 1) * H + h] has reached the group's nb tiles.  With the
    // grouped tile order that group ended at least a whole group ago, so with groups of two waves of tiles or more nobody
    // waits; the count is fetched when the CTA changes group, at the barrier flush_dscale needs anyway, so a tile normally
    // starts its rows without a global load.  (A fence per warp and tile in the epilogue costs 10 % of the kernel, a
    // __threadfence per lane 40 %.)
    auto put_chunk = [&](float4* dst, const float4 x, bool live, int t_) {
      if (!DIRECT || t_ == 0) *dst = live ? x : make_float4(0.f, 0.f, 0.f, 0.f);
      else if (live) atomicAdd(dst, x);
    };
    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      int h, t, blk;
      decode(tile, h, t, blk);
      const float* nq2s = s_nq2 + ph * 128;
      const int* qidx = s_qidx + (it % 3) * 128;
      const int* kidx = s_kidx + (it % 3) * 128;

      // ---- key side first: P^T -> TMEM (hi over S^T, lo over dP^T), dS^T kept in registers ---------------------------
      float dsr[MAXCH * 8];
      umma::mbar_wait(&mbar[KREADY], ph);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(BE_KREADY, it);
      {
#pragma unroll
        for (int ci = 0; ci < MAXCH; ++ci) {
          const int ch = part + ci * kBtParts;
          if (ch < KSTEPS) {
            uint32_t ra[8], rb[8];
            umma::tmem_ld8_nowait(tST + lane_base + 8 * ch, ra);
            umma::tmem_ld8_nowait(tDPT + lane_base + 8 * ch, rb);
            const float4 n0v = *reinterpret_cast<const float4*>(nq2s + 8 * ch), n1v = *reinterpret_cast<const float4*>(nq2s + 8 * ch + 4);
            const float nq2[8] = {n0v.x, n0v.y, n0v.z, n0v.w, n1v.x, n1v.y, n1v.z, n1v.w};
            umma::tmem_wait_ld(ra, rb);
            float phv[8], plv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float x = fmaf(__uint_as_float(ra[u]), kLog2e, nq2[u]);   // the same fmaf as the query side
              const float p = exp2_fast(fminf(x, 0.f));
              dsr[ci * 8 + u] = x <= 0.f ? p * __uint_as_float(rb[u]) : 0.f;   // clamp(max=0) backward: [S <= 0]
              trunc_tf32(p, phv[u], plv[u]);
            }
            umma::tmem_st8(tST + lane_base + 8 * ch, phv);
            umma::tmem_st8(tDPT + lane_base + 8 * ch, plv);
          }
          if (HEPT_BWD_SPLIT && ci == HCI) arrive_tmem(PTHALF);     // this warp's chunks of k-steps [0, HK) are in TMEM
        }
      }
      arrive_tmem(PTRDY);
      if (warp == 0) HEPT_TRACE_EVENT(BE_PTRDY, it);

      // ---- query side: dS -> TMEM (hi over S, lo over dP) -----------------------------------------------------------
      umma::mbar_wait(&mbar[QREADY], ph);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(BE_QREADY, it);
      {
        const float nq2 = nq2s[row];
#pragma unroll 1                                   // nothing is carried between chunks: rolled, the kernel's code has to stay
        for (int ci = 0; ci < MAXCH; ++ci) {      // inside the instruction cache (see the note on code size in DESIGN.md)
          const int ch = part + ci * kBtParts;
          if (ch < KSTEPS) {
            uint32_t ra[8], rb[8];
            umma::tmem_ld8_nowait(tS + lane_base + 8 * ch, ra);
            umma::tmem_ld8_nowait(tDP + lane_base + 8 * ch, rb);
            umma::tmem_wait_ld(ra, rb);
            float dh[8], dl[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float x = fmaf(__uint_as_float(ra[u]), kLog2e, nq2);
              const float p = exp2_fast(fminf(x, 0.f));
              trunc_tf32(x <= 0.f ? p * __uint_as_float(rb[u]) : 0.f, dh[u], dl[u]);   // the same mask, the same bits
            }
            umma::tmem_st8(tS + lane_base + 8 * ch, dh);
            umma::tmem_st8(tDP + lane_base + 8 * ch, dl);
          }
          if (HEPT_BWD_SPLIT && ci == HCI) arrive_tmem(DSHALF);
        }
      }
      arrive_tmem(DSRDY);
      if (warp == 0) HEPT_TRACE_EVENT(BE_DSRDY, it);

      // first tile of another head (DIRECT: of another (table, head) group -- a CTA that skips groups can come back to the
      // same head): hand in the finished group's sums and, DIRECT, its tile count
      if (DIRECT ? t * H + h != cur_grp : h != dsc_head) {
        flush_dscale();
        dsc_head = h;
        cur_grp = t * H + h;
      }
      // x'_e (e in this thread's 16 columns) of row `row` of an MN-major (hi, lo) tile pair: hi + lo is exact
      auto centred_row = [&](int th_, int tl_, float (&xr)[16]) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const uint32_t off = umma::sw128b32_offset(srow, 4 * part + cc);
          const float4 a = *reinterpret_cast<const float4*>(mn + th_ * CF::TILE + off);
          const float4 b = *reinterpret_cast<const float4*>(mn + tl_ * CF::TILE + off);
          xr[4 * cc] = a.x + b.x; xr[4 * cc + 1] = a.y + b.y; xr[4 * cc + 2] = a.z + b.z; xr[4 * cc + 3] = a.w + b.w;
        }
      };
      // out_e = acc_e - sum * x'_e; rows of D floats per (head, hit, table); coordinate columns feed d scale
      auto finish_rows = [&](const float (&acc)[16], const float (&xr)[16], float sum, int n, float* __restrict__ stage) {
        float o[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) o[u] = fmaf(-sum, xr[u], acc[u]);
        if (n >= 0) {
          float4* dst = reinterpret_cast<float4*>(stage + (DIRECT ? ((size_t)n * H + h) * D : (((size_t)h * N + n) * T + t) * D));
          const bool live = !DIRECT || n < raw_size;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (4 * (4 * part + cc) < D)
              put_chunk(dst + 4 * part + cc, make_float4(o[4 * cc], o[4 * cc + 1], o[4 * cc + 2], o[4 * cc + 3]), live, t);
#pragma unroll
          for (int cc = 0; cc < C; ++cc)     // coordinate column D + cc lives in part (D + cc) / 16 (part is warp-uniform)
            if (part == (D + cc) / 16) dsc[cc] = fmaf(xr[(D + cc) % 16], o[(D + cc) % 16], dsc[cc]);
        }
      };
      // a 64-column accumulator = [hi*hi | hi*lo + lo*hi]: this thread's 16 columns of the sum of the two halves, and (SUM)
      // column E of both halves -- every load is issued before the one wait (four waits cost four TMEM round trips)
      auto ld_acc = [&](uint32_t taddr, float (&acc)[16], float* sum) {
        uint32_t a[16], l[16], s0 = 0, s1 = 0;
        umma::tmem_ld16_nowait(taddr + lane_base + 16 * part, a);
        umma::tmem_ld16_nowait(taddr + 32 + lane_base + 16 * part, l);
        if (sum) {
          umma::tmem_ld1_nowait(taddr + lane_base + E, s0);
          umma::tmem_ld1_nowait(taddr + 32 + lane_base + E, s1);
        }
        umma::tmem_wait_ld();
        umma::tmem_tie(a);
        umma::tmem_tie(l);
        if (sum) {
          umma::tmem_tie(s0, s1);
          *sum = __uint_as_float(s0) + __uint_as_float(s1);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = __uint_as_float(a[u]) + __uint_as_float(l[u]);
      };

      // ---- dv rows (dV was the first product into tO) ------------------------------------------------------------------
      umma::mbar_wait(&mbar[DVDONE], ph);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(BE_DVDONE, it);
      if (DIRECT) {
        if (t > 0 && ready != (t - 1) * H + h) {         // the rows of table t - 1 of this head are all in place?
          ready = (t - 1) * H + h;
          const int want = (int)decode.nb;
          // usually the count was already complete when the CTA left its previous group (flush_dscale asked for it)
          if (!(s_dep[0] == ready && s_dep[1] >= want) && lane == 0) {
            // poll relaxed (an acquire load invalidates the SM's L1 every time), acquire once the count is there
            auto count = [&]() {
              int c;
              asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(c) : "l"(done + ready) : "memory");
              return c;
            };
            int spins = 0;
            while (count() < want && ++spins < (1 << 24)) __nanosleep(32);
            int c;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(c) : "l"(done + ready) : "memory");
            if (c < want) __trap();                      // a second of waiting: the launch is broken, do not add out of order
          }
          __syncwarp();                                  // the other lanes' adds are ordered after lane 0's acquire
        }
      }
      if (warp == 0) HEPT_TRACE_EVENT(BE_DVWAIT, it);
      {
        float acc[16];
        ld_acc(tO, acc, nullptr);
        const int n = row < B ? kidx[row] : -1;
        if (n >= 0) {
          float4* dst = reinterpret_cast<float4*>(stage_dv + (DIRECT ? ((size_t)n * H + h) * D : (((size_t)h * N + n) * T + t) * D));
          const bool live = !DIRECT || n < raw_size;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (4 * (4 * part + cc) < D)
              put_chunk(dst + 4 * part + cc, make_float4(acc[4 * cc], acc[4 * cc + 1], acc[4 * cc + 2], acc[4 * cc + 3]), live, t);
        }
      }
      if (warp == 0) HEPT_TRACE_EVENT(BE_DVOUT, it);

      // ---- dS^T goes where dS was, as soon as dQ has consumed dS; dK then accumulates into tO (dv rows are out) --------
      umma::mbar_wait(&mbar[HEPT_BWD_SPLIT ? DQHALF : DQDONE], ph);      // dQ has consumed k-steps [0, HK) of dS (or all of it)
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(BE_DQDONE, it);
      bool second_half = false;
      auto open_second_half = [&]() {                    // the chunks of k-steps >= HK: dQ must be done with ALL of dS
        arrive_tmem(DSTHALF);                            // (also: this warp has read dV out of tO)
        umma::mbar_wait(&mbar[DQDONE], ph);
        umma::fence_after_sync();
        second_half = true;
      };
#pragma unroll
      for (int ci = 0; ci < MAXCH; ++ci) {
        const int ch = part + ci * kBtParts;
        if (ch < KSTEPS) {
          if (HEPT_BWD_SPLIT && ch >= HK && !second_half) open_second_half();
          float dh[8], dl[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) trunc_tf32(dsr[ci * 8 + u], dh[u], dl[u]);
          umma::tmem_st8(tS + lane_base + 8 * ch, dh);
          umma::tmem_st8(tDP + lane_base + 8 * ch, dl);
        }
      }
      if (HEPT_BWD_SPLIT && !second_half) open_second_half();
      if (warp == 0) HEPT_TRACE_EVENT(BE_DSTISS, it);
      arrive_tmem(DSTRDY);                               // also: this warp has read dV out of tO
      if (warp == 0) HEPT_TRACE_EVENT(BE_DSTRDY, it);

      // ---- dq^ rows: dQ was accumulated over the first columns of the (consumed) P^T region ---------------------------
      {
        float acc[16], xr[16], rs;
        ld_acc(tST, acc, &rs);                                     // rs = column E of dQ: sum_j dS_ij
        if (warp == 0) HEPT_TRACE_EVENT(BE_DQACC, it);
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&mbar[STFREE]);           // the next tile's key-side scores may overwrite tST
        centred_row(CF::MQH, CF::MQL, xr);
        finish_rows(acc, xr, rs, row < B ? qidx[row] : -1, stage_dq);
      }
      if (warp == 0) HEPT_TRACE_EVENT(BE_DQOUT, it);

      // ---- dk^ rows ---------------------------------------------------------------------------------------------------
      umma::mbar_wait(&mbar[DKDONE], ph);
      umma::fence_after_sync();
      if (warp == 0) HEPT_TRACE_EVENT(BE_DKDONE, it);
      {
        float acc[16], xr[16], cs;
        ld_acc(tO, acc, &cs);                                      // cs = column E of dK: sum_i dS_ij
        if (warp == 0) HEPT_TRACE_EVENT(BE_DKACC, it);
        centred_row(CF::MKH, CF::MKL, xr);
        umma::fence_before_sync();                       // the TMEM loads above precede the next tile's MMAs into tO
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&mbar[MFREE]);  // this warp is done with the MN-major tiles (its rows are in registers)
        finish_rows(acc, xr, cs, row < B ? kidx[row] : -1, stage_dk);
      }
      if (DIRECT) {
        // last tile of this CTA in the group: sums handed in, and the count the next group depends on asked for, now
        int hn = -1, tn = 0, bn;
        if (tile + (int)gridDim.x < total_tiles) decode(tile + (int)gridDim.x, hn, tn, bn);
        if (hn < 0 || tn * H + hn != cur_grp) {
          flush_dscale(hn >= 0 && tn > 0 ? (tn - 1) * H + hn : -1);
          dsc_head = hn;
          cur_grp = tn * H + hn;
        }
      }
      if (warp == 0) HEPT_TRACE_EVENT(BE_END, it);

    }
    flush_dscale();
  } else if (warp < EW + PW) {
    // =========================================== producer warps =================================================
    umma::setmaxnreg_dec<kBtRegsProd>();
    const int ptid = tid - kBtEpiThreads, sub = ptid >> 3, c = ptid & 7;   // 8 lanes per row, RPP rows per pass
    constexpr int RPP = CF::RPP;
    float4* stg_v = reinterpret_cast<float4*>(smem + CF::OFF_STG_V);
    float4* stg_g = reinterpret_cast<float4*>(smem + CF::OFF_STG_G);
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
    // issue every row load of a tile: q^ / k^ chunks into registers, v / G' chunks into the staging buffer
    auto issue_rows = [&](int tile, int it) {
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
        const bool vok = nkk >= 0 && c < VCH && nkk < raw_size;
        umma::cp_async16(stg_v + ps * kBtProdThreads + ptid, vok ? v + ((size_t)nkk * H + h) * D + 4 * c : v, vok ? 16 : 0);
        umma::cp_async16(stg_g + ps * kBtProdThreads + ptid, nqq >= 0 ? grows + ((size_t)nqq * H + h) * 32 + 4 * c : grows,
                         nqq >= 0 ? 16 : 0);
        const int r = ps * RPP + sub;
        if (c == 0 && r < B) { s_qidx[(it % 3) * 128 + r] = nqq; s_kidx[(it % 3) * 128 + r] = nkk; }
      }
    };

    int tile = blockIdx.x;
    if (tile < total_tiles) {
      load_indices(tile);
      issue_rows(tile, 0);
      if (tile + (int)gridDim.x < total_tiles) load_indices(tile + gridDim.x);
    }
    // DIRECT: this CTA's finished tiles of a (table, head) group are counted into done[] here, not by the epilogue warps:
    // once the producer has seen MFREE of tile it - 1, every epilogue warp has issued that tile's rows (they arrive on
    // MFREE after their last row), so one producer thread can publish them with a release -- and its MEMBAR stalls a
    // warp that is about to wait for the MMAs anyway instead of the epilogue, which bounds the kernel.
    int rel_grp = -1, rel_cnt = 0;                        // group of tile it - 1 (-1: last table, nobody reads the count)
    auto hand_in = [&]() {
      if (rel_grp >= 0 && rel_cnt > 0 && ptid == 0)
        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(done + rel_grp), "r"(rel_cnt) : "memory");
      rel_cnt = 0;
    };
    int it = 0;
#pragma unroll 1
    for (; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      // ---- K-major operand tiles: free once the previous tile's score MMAs (both sides) are done ------------------
      if (it > 0) umma::mbar_wait(&mbar[HEPT_BWD_PROD_WAIT], ph ^ 1);   // dV done (hence both sides' score MMAs): see DESIGN.md on contention
      if (warp == EW) HEPT_TRACE_EVENT(BP_KFREE, it);
      umma::cp_async_wait_all();
      float* nq2s = s_nq2 + ph * 128;
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
            *reinterpret_cast<float4*>(smem + CF::KH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(smem + CF::KL * CF::TILE + okm) = lo;
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
            *reinterpret_cast<float4*>(smem + CF::QH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(smem + CF::QL * CF::TILE + okm) = lo;
            if (c == 0) nq2s[r] = nq2;
          }
        }
        {  // value row with a 1 in slot D
          float4 d = stg_v[ps * kBtProdThreads + ptid];
          if (c == VCH) d.x = 1.f;
          split4(d, hi, lo);
          if (in) {
            *reinterpret_cast<float4*>(smem + CF::VH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(smem + CF::VL * CF::TILE + okm) = lo;
          }
        }
        {  // gradient row (gd, -gy)
          split4(stg_g[ps * kBtProdThreads + ptid], hi, lo);
          if (in) {
            *reinterpret_cast<float4*>(smem + CF::GH * CF::TILE + okm) = hi;
            *reinterpret_cast<float4*>(smem + CF::GL * CF::TILE + okm) = lo;
          }
        }
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[KFULL]);
      if (warp == EW) HEPT_TRACE_EVENT(BP_KFULL, it);

      // ---- registers and staging are free: put the next tile's loads in flight, fetch the indices of the one after -
      const int next = tile + gridDim.x;
      if (next < total_tiles) {
        issue_rows(next, it + 1);
        if (next + (int)gridDim.x < total_tiles) load_indices(next + gridDim.x);
      }

      // ---- MN-major copies: free once the previous tile's last MMA is done and its epilogue has read them ---------
      if (warp == EW) HEPT_TRACE_EVENT(BP_ISSUED, it);
      if (it > 0) umma::mbar_wait(&mbar[MFREE], ph ^ 1);
      if (warp == EW) HEPT_TRACE_EVENT(BP_MFREE, it);
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int r = ps * RPP + sub;
        if (r < B) {
          const uint32_t okm = umma::sw128_offset(r, c), omn = umma::sw128b32_offset(r, c);
          constexpr int src[6] = {CF::KH, CF::KL, CF::GH, CF::GL, CF::QH, CF::QL};   // -> MKH, MKL, MGH, MGL, MQH, MQL
#pragma unroll
          for (int tl = 0; tl < 6; ++tl) {
            float4 x = *reinterpret_cast<const float4*>(smem + src[tl] * CF::TILE + okm);
            if (tl < 2 && c == CF::SLOT_CH) {   // the MN-major K^ carries (1, 0) in the side slots: column E of dQ = sum_j dS_ij
              umma::elem<CF::SLOT_U>(x) = tl == 0 ? 1.f : 0.f;
              umma::elem<CF::SLOT_U + 1>(x) = 0.f;
            }
            *reinterpret_cast<float4*>(mn + tl * CF::TILE + omn) = x;
          }
        }
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&mbar[MFULL]);
      if (warp == EW) HEPT_TRACE_EVENT(BP_MFULL, it);
      if (DIRECT && warp == EW) {
        // MFREE of tile it - 1 was seen above: count it, and hand the count in when this tile belongs to another group
        int h, t, blk;
        decode(tile, h, t, blk);
        const int g = t + 1 < T ? t * H + h : -1;
        if (it > 0) {
          ++rel_cnt;
          if (g != rel_grp || t + 1 >= T) hand_in();      // (last-table groups all read -1: nothing to hand in, counter reset)
        }
        rel_grp = g;
      }
    }
    if (DIRECT && warp == EW && it > 0) {                 // the CTA's last tile
      umma::mbar_wait(&mbar[MFREE], (it - 1) & 1);
      ++rel_cnt;
      hand_in();
    }
  } else {
    umma::setmaxnreg_dec<kBtRegsMma>();
  }
  if (warp == EW + PW) {
    // =========================================== MMA issuer ====================================================
    // the warp runs the schedule uniformly; the lane chosen by elect.sync issues (always the same lane, so its
    // tcgen05.commit covers every MMA issued before)
    constexpr uint32_t idesc_s = umma::idesc_tf32(128, NP, false, false);
    constexpr uint32_t idesc_o = umma::idesc_tf32(128, 32, false, true);
    constexpr uint32_t idesc_o64 = umma::idesc_tf32(128, 64, false, true);
    // descriptors differ only in their start-address field (bits [0,14), 16-byte units): base + offset / 16
    const uint64_t kdesc0 = umma::smem_desc_sw128(sbase, 1024, 16);
    // MN-major B: LBO = the distance between the 32-column tiles of an N = 64 operand, i.e. from a hi tile to its lo tile
    const uint64_t mdesc0 = umma::smem_desc(mbase, 512, CF::TILE, umma::kLayoutSw128Base32);
    // D = A B^T, both K-major, 3xTF32 (small cross terms first).  `swap` exchanges the operand roles while keeping the
    // product order, so D^T comes out bit-identical.  Called by the elected lane only.
    auto ss_product = [&](uint32_t d, int xh, int xl, int yh, int yl, int ksteps, bool swap) {
      uint64_t base = kdesc0;
      asm volatile("" : "+l"(base));   // opaque: descriptors are re-derived here (one add each), not hoisted and spilled
#pragma unroll
      for (int p3 = 0; p3 < 3; ++p3) {
        const int x = p3 == 1 ? xl : xh, y = p3 == 0 ? yl : yh;
        const uint64_t dx = base + (uint64_t)((x * CF::TILE) >> 4), dy = base + (uint64_t)((y * CF::TILE) >> 4);
#pragma unroll 1
        for (int kk = 0; kk < ksteps; ++kk) {
          if (swap) umma::mma_ss(d, dy + 2 * kk, dx + 2 * kk, idesc_s, (p3 | kk) != 0);
          else umma::mma_ss(d, dx + 2 * kk, dy + 2 * kk, idesc_s, (p3 | kk) != 0);
        }
      }
    };
    // D[64 columns] = A[tmem hi/lo] * B[MN-major hi/lo], 3xTF32 with two MMAs per k-step instead of three: A_hi meets
    // [B_hi | B_lo] in ONE N = 64 MMA (columns [0,32) = hi*hi, [32,64) = hi*lo) and A_lo meets B_hi in an N = 32 MMA that
    // accumulates into the upper half; the epilogue adds the halves in fp32.  A third fewer MMAs and a third fewer reads
    // of A from TMEM, whose read bandwidth bounds this kernel.  `bh` = the hi tile; its lo tile follows it.
    // Called by the elected lane only.
    [[maybe_unused]] constexpr int HK = (KSTEPS + 1) / 2;
    auto ts_product = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, int bh, int k0, int k1) {
      uint64_t base = mdesc0;
      asm volatile("" : "+l"(base));
      const uint64_t db = base + (uint64_t)((bh * CF::TILE) >> 4);
#pragma unroll 1
      for (int kk = k0; kk < k1; ++kk) {
        umma::mma_ts(d, a_hi + 8 * kk, db + 64 * kk, idesc_o64, kk != 0);
        umma::mma_ts(d + 32, a_lo + 8 * kk, db + 64 * kk, idesc_o, true);
      }
    };
    auto wait = [&](BtBar b, uint32_t parity) {
      umma::mbar_wait(&mbar[b], parity);
      umma::fence_after_sync();
    };
    auto scores_query_side = [&]() {
      if (umma::elect_one()) {
        ss_product(tS, CF::QH, CF::QL, CF::KH, CF::KL, CF::SK, false);
        ss_product(tDP, CF::GH, CF::GL, CF::VH, CF::VL, CF::GK, false);
        umma::commit(&mbar[QREADY]);
      }
      __syncwarp();
    };
    auto scores_key_side = [&]() {
      if (umma::elect_one()) {
        ss_product(tST, CF::QH, CF::QL, CF::KH, CF::KL, CF::SK, true);
        ss_product(tDPT, CF::GH, CF::GL, CF::VH, CF::VL, CF::GK, true);
        umma::commit(&mbar[KREADY]);
      }
      __syncwarp();
    };
    int it = 0;
    int tile = blockIdx.x;
    HEPT_TRACE_CTA(0);
    HEPT_TRACE_CTA(2);
    if (tile < total_tiles) {
      wait(KFULL, 0);
      scores_key_side();
      scores_query_side();
    }
#pragma unroll 1
    for (; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const bool more = tile + (int)gridDim.x < total_tiles;
#if HEPT_BWD_SPLIT
      wait(PTHALF, ph);
      wait(MFULL, ph);
      HEPT_TRACE_EVENT(BM_DV_GO, it);
      if (umma::elect_one()) ts_product(tO, tST, tDPT, CF::MGH, 0, HK);          // dV = P^T G', first half of P^T
      __syncwarp();
      wait(PTRDY, ph);
      if (umma::elect_one()) {
        ts_product(tO, tST, tDPT, CF::MGH, HK, KSTEPS);
        umma::commit(&mbar[DVDONE]);
      }
      __syncwarp();
      wait(DSHALF, ph);
      HEPT_TRACE_EVENT(BM_DQ_GO, it);
      if (umma::elect_one()) {
        ts_product(tST, tS, tDP, CF::MKH, 0, HK);       // dQ = dS K^: accumulator over the P^T columns dV has consumed
        umma::commit(&mbar[DQHALF]);                    // k-steps [0, HK) of dS are consumed: dS^T may go there
      }
      __syncwarp();
      wait(DSRDY, ph);
      if (umma::elect_one()) {
        ts_product(tST, tS, tDP, CF::MKH, HK, KSTEPS);
        umma::commit(&mbar[DQDONE]);
      }
      __syncwarp();
      wait(DSTHALF, ph);
      HEPT_TRACE_EVENT(BM_DK_GO, it);
      if (umma::elect_one()) ts_product(tO, tS, tDP, CF::MQH, 0, HK);            // dK = dS^T Q^: A where dS was, accumulator where dV was
      __syncwarp();
      wait(DSTRDY, ph);
      if (umma::elect_one()) {
        ts_product(tO, tS, tDP, CF::MQH, HK, KSTEPS);
        umma::commit(&mbar[DKDONE]);
        umma::commit(&mbar[MFREE]);
      }
      __syncwarp();
#else
      wait(PTRDY, ph);
      wait(MFULL, ph);
      HEPT_TRACE_EVENT(BM_DV_GO, it);
      if (umma::elect_one()) {
        ts_product(tO, tST, tDPT, CF::MGH, 0, KSTEPS);  // dV = P^T G'
        umma::commit(&mbar[DVDONE]);
      }
      __syncwarp();
      wait(DSRDY, ph);
      HEPT_TRACE_EVENT(BM_DQ_GO, it);
      if (umma::elect_one()) {
        ts_product(tST, tS, tDP, CF::MKH, 0, KSTEPS);   // dQ = dS K^: accumulator over the P^T columns dV has consumed
        umma::commit(&mbar[DQDONE]);
      }
      __syncwarp();
      wait(DSTRDY, ph);
      HEPT_TRACE_EVENT(BM_DK_GO, it);
      if (umma::elect_one()) {
        ts_product(tO, tS, tDP, CF::MQH, 0, KSTEPS);    // dK = dS^T Q^: A where dS was, accumulator where dV was (read out)
        umma::commit(&mbar[DKDONE]);
        umma::commit(&mbar[MFREE]);
      }
      __syncwarp();
#endif
      // The next tile's key-side scores (the epilogue starts with them) overwrite tST / tDPT, where dQ accumulated: they
      // wait for the epilogue to read dq out; the query-side scores follow (tS / tDP: behind dK in pipe order).
      if (more) {
        wait(KFULL, ph ^ 1);
        wait(STFREE, ph);
        HEPT_TRACE_EVENT(BM_SK_GO, it + 1);
        scores_key_side();
        HEPT_TRACE_EVENT(BM_SQ_GO, it + 1);
        scores_query_side();
      }
      HEPT_TRACE_EVENT(BM_END, it);
    }
    HEPT_TRACE_CTA(1);
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == EW + PW) umma::tmem_dealloc<CF::TMEM_COLS>(tmem);
}

// Sum the T per-table staging rows of every (hit, head) into dq, dk, dv.  thread <-> (n, h, 16-byte chunk).
// Rows >= raw_size are the src/ flavour's padding: the reference overwrites their q^, k^ and v with zeros
// (src/models/attention/hept.py:89-91), so no gradient reaches the caller's rows.
template <int D>
__global__ void __launch_bounds__(256) bwd_table_sum_kernel(const float* __restrict__ sq, const float* __restrict__ sk,
                                                            const float* __restrict__ sv, int N, int H, int T,
                                                            int raw_size, float* __restrict__ dq, float* __restrict__ dk,
                                                            float* __restrict__ dv) {
  constexpr int VCH = D / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * H * VCH) return;
  const int c = (int)(idx % VCH);
  const size_t nh = idx / VCH;
  const int h = (int)(nh % H);
  const size_t n = nh / H;
  const size_t src = (((size_t)h * N + n) * T) * D + 4 * c;
  const size_t dst = (n * H + h) * D + 4 * c;
  auto sum_tables = [&](const float* __restrict__ in, float* __restrict__ out) {
    if (n >= (size_t)raw_size) {
      *reinterpret_cast<float4*>(out + dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      return;
    }
    float4 s = ldg4(in + src);
    if (T <= 4) {                   // every load in flight before the first add (T is a run-time value)
      float4 x[3];
#pragma unroll
      for (int t = 1; t < 4; ++t) x[t - 1] = t < T ? ldg4(in + src + (size_t)t * D) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 1; t < 4; ++t)   // table order, like the reference's sum over dim 0
        if (t < T) { s.x += x[t - 1].x; s.y += x[t - 1].y; s.z += x[t - 1].z; s.w += x[t - 1].w; }
    } else {
      for (int t = 1; t < T; ++t) {
        const float4 x = ldg4(in + src + (size_t)t * D);
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
    }
    *reinterpret_cast<float4*>(out + dst) = s;
  };
  sum_tables(sq, dq);
  sum_tables(sk, dk);
  sum_tables(sv, dv);
}

// dscale[h,c] = (sum over CTAs of partial[cta][h][c]) / scale[h,c]; fixed-order tree -> deterministic.
__global__ void __launch_bounds__(256) dscale_tiles_kernel(const float* __restrict__ partial, const float* __restrict__ scale,
                                                           int ctas, int H, int C, float* __restrict__ dscale) {
  __shared__ float red[256];
  const int h = blockIdx.x / C, c = blockIdx.x % C;
  float x = 0.f;
  for (int b = threadIdx.x; b < ctas; b += 256) x += partial[((size_t)b * H + h) * 8 + c];
  red[threadIdx.x] = x;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float s = scale[h * C + c];
    dscale[h * C + c] = s != 0.f ? red[0] / s : 0.f;
  }
}

constexpr int kBtMaxCtas = 1024;
struct BwdTcPlan {
  size_t hat_bytes, grow_bytes, stage_bytes, partial_bytes, total;
  int tiles;
};
static BwdTcPlan plan_bwd_tc(const hept_shape* s) {
  BwdTcPlan p;
  p.tiles = s->T * s->H * (s->N / s->B);
  p.hat_bytes = align_up(sizeof(float) * (size_t)s->N * s->H * 8, 256);
  p.grow_bytes = align_up(sizeof(float) * (size_t)s->N * s->H * 32, 256);
  p.stage_bytes = align_up(sizeof(float) * (size_t)s->H * s->N * s->T * s->D, 256);
  p.partial_bytes = align_up(sizeof(float) * (size_t)kBtMaxCtas * s->H * 8 + sizeof(int) * (size_t)s->T * s->H, 256);
  p.total = p.hat_bytes + p.grow_bytes + 3 * p.stage_bytes + p.partial_bytes;
  return p;
}
size_t bwd_tc_workspace_bytes(const hept_shape* s) { return plan_bwd_tc(s).total; }

template <int D, int C, int B>
static int launch_bwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                         const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                         const float* d_out_pre, float* dq, float* dk, float* dv, float* dscale, char* ws,
                         cudaStream_t st) {
  using CF = TcBwd<D, C, B>;
  auto kern = block_attn_bwd_tc_kernel<D, C, B, false>;
  auto kern_direct = block_attn_bwd_tc_kernel<D, C, B, true>;
  const size_t smem = CF::TOTAL + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "block_attn_bwd_tc: cannot reserve %zu B of shared memory: %s", smem,
                 cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "block_attn_bwd_tc: cannot read the SM count");
  BwdTcPlan p = plan_bwd_tc(s);
  float* hat = (float*)ws;
  float* grows = (float*)(ws + p.hat_bytes);
  float* sq = (float*)(ws + p.hat_bytes + p.grow_bytes);
  float* sk = (float*)((char*)sq + p.stage_bytes);
  float* sv = (float*)((char*)sk + p.stage_bytes);
  float* partial = (float*)((char*)sv + p.stage_bytes);
  HEPT_REQUIRE(s->C <= 8, HEPT_EUNSUPPORTED, "block_attn_bwd_tc: more than 8 coordinates");
  const size_t rows = (size_t)s->N * s->H;
  const int mask = bwd_stage_mask();
  int grid = p.tiles < sms ? p.tiles : sms;             // one CTA per SM (the tile uses all 512 TMEM columns)
  if (grid > kBtMaxCtas) grid = kBtMaxCtas;
  // d scale partials of heads a CTA never visits and the per-(table, head) tile counts of the direct form start at zero
  const int zero_words = (mask & 3) ? grid * s->H * 8 + s->T * s->H : 0;
  grad_rows_kernel<D><<<(unsigned)((((rows + 1) / 2) * 8 + 255) / 256), 256, 0, st>>>(d_out_pre, out_pre, den_sum, rows, grows, coords, scale,
                                                                         s->H, s->C, s->raw_size, hat, (uint32_t*)partial, zero_words);
  HEPT_CHECK_LAUNCH("grad_rows");
  // Rows of the T tables: added straight into dq / dk / dv in table order by the tile kernel (needs the paired tile order,
  // i.e. an even number of heads; variant 3 does it when a (head, table) group is at least two waves of tiles, so that
  // no tile ever waits for the previous table; 4 = whenever possible, 5 = never), or staged per table and summed by
  // bwd_table_sum.  Same bits either way.
  const int nb = s->N / s->B;
  const int variant = bwd_variant();
  // heads per group: two when the number of heads is even (measured at 60k hits, 8 heads: 798 us with pairs, 806 with
  // threes, ~826 with fours -- the rows of more heads no longer stay in L2), else the smallest group size that does not
  // leave a last group of one head (whose tables would follow each other directly).  HEPT_BWD_GROUP overrides (experiments).
  static const int forced_g = [] { const char* e = getenv("HEPT_BWD_GROUP"); return e ? atoi(e) : 0; }();
  int hg = 0;
  for (int g = 2; g <= 4 && !hg; ++g)
    if (s->H >= g && s->H % g != 1) hg = g;
  if (forced_g > 1 && s->H >= forced_g && s->H % forced_g != 1) hg = forced_g;
  bool direct = hg > 0 && variant != 5 && (variant == 4 || nb >= 2 * grid) && (mask & 3) == 3;
  if (mask & 3) {
    int* done = (int*)(partial + (size_t)grid * s->H * 8);
    HEPT_REQUIRE(TileDecoder::exact_for(p.tiles, nb, s->T), HEPT_EUNSUPPORTED, "block_attn_bwd_tc: too many tiles (%d)", p.tiles);
    bool launched = false;
    if (direct) {
      // The direct form's CTAs wait for each other's tile counts, so every CTA of the grid must be resident at once: a
      // COOPERATIVE launch makes the driver guarantee that (it places the whole grid or nothing, whatever other streams hold)
      // instead of this code assuming it.  If the device cannot place `grid` CTAs together the staged form takes over.
      const TileDecoder dec = TileDecoder::make(nb, s->T, s->H, hg);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(kBtThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kern_direct, q, k, v, (const float*)hat, (const float*)grows, positions, (int)s->N,
                                         (int)s->H, (int)s->T, (int)s->raw_size, (int)p.tiles, dec, dq, dk, dv, partial, done);
      if (e == cudaSuccess) launched = true;
      else (void)cudaGetLastError();
    }
    direct = launched;
    if (!launched)
      kern<<<grid, kBtThreads, smem, st>>>(q, k, v, hat, grows, positions, s->N, s->H, s->T, s->raw_size, p.tiles,
                                           TileDecoder::make(nb, s->T), sq, sk, sv, partial, done);
    HEPT_CHECK_LAUNCH("block_attn_bwd_tc");
  }
  if (!(mask & 4)) return HEPT_OK;
  if (!direct) {
    const size_t total = rows * (D / 4);
    bwd_table_sum_kernel<D><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sq, sk, sv, s->N, s->H, s->T, s->raw_size, dq, dk, dv);
    HEPT_CHECK_LAUNCH("bwd_table_sum");
  }
  dscale_tiles_kernel<<<s->H * C, 256, 0, st>>>(partial, scale, grid, s->H, C, dscale);
  HEPT_CHECK_LAUNCH("dscale_tiles");
  return HEPT_OK;
}

int block_attention_bwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                           const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                           const float* d_out_pre, float* dq, float* dk, float* dv, float* dscale, char* ws,
                           cudaStream_t st) {
  if (s->D == 24 && s->C == 6 && s->B == 100)
    return launch_bwd_tc<24, 6, 100>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 4 && s->B == 100)
    return launch_bwd_tc<24, 4, 100>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 6 && s->B == 64)
    return launch_bwd_tc<24, 6, 64>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 4 && s->B == 64)
    return launch_bwd_tc<24, 4, 64>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 8 && s->C == 6 && s->B == 10)
    return launch_bwd_tc<8, 6, 10>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  set_error("block_attention_bwd (tensor-core engine): (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}

}  // namespace hept
