// a18, second-generation fp32 tiles: lane pairs + packed FFMA2 (pair.cuh).  Same mathematics and the same staging
// layout as attn_bwd.cu; dq and dk see bit-identical dS (canonical arithmetic in pair.cuh).
//   dq kernel : lane pairs own R query rows, keys/values streamed            (84 FMA / pair of hits)
//   dk kernel : lane pairs own R key rows, queries/gd streamed               (84 FMA / pair)
//   dv kernel : lane pairs own R key rows, queries/gd streamed, P only       (54 FMA / pair)
// Splitting dk and dv costs one extra evaluation of S but lets every kernel keep two rows per lane pair inside
// 128 registers; the fused dk/dv kernel of attn_bwd.cu needs 168 registers for ONE row per lane.
#include "pair.cuh"

namespace hept {

enum class BwdRole { DQ, DK, DV };

// gather geometry (tile.cuh helpers) with the pair kernel's thread count
template <class P>
struct PairGather : P::Gather {
  static constexpr int THREADS = P::THREADS;
};

// ROLE == DQ: resident rows are queries (x = q), streamed hat rows are keys with values as the aux rows.
// ROLE == DK / DV: resident rows are keys (x = k) with their value rows, streamed hat rows are queries (+nq, gy)
//                  with gd = g / den as the aux rows.
template <int D, int C, int B, int G, int R, int MINB, BwdRole ROLE>
__global__ void __launch_bounds__((PairLayout<D, C, B, G, R>::THREADS), MINB)
    block_attn_bwd_pair_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                               const float* __restrict__ coords, const float* __restrict__ scale,
                               const int32_t* __restrict__ positions, const float* __restrict__ out_pre,
                               const float* __restrict__ den_sum, const float* __restrict__ d_out_pre, int N, int H,
                               int T, int raw_size, float* __restrict__ stage_out) {
  using P = PairLayout<D, C, B, G, R>;
  using GL = PairGather<P>;
  constexpr int E = P::E, HCH = P::HCH, VH = P::VH;
  extern __shared__ float4 smem[];
  float4* hs = smem;                                   // [G*B][8]   streamed hat rows
  float4* as = smem + G * B * P::ROW_CHUNKS;           // [G*B][D/4] streamed aux rows (v or gd)

  const int h = blockIdx.y / T, t = blockIdx.y % T, th = t * H + h;  // tables of one head run back to back: its q/k/v slices stay in L2
  const int nb = N / B;
  const int blk0 = blockIdx.x * G;
  const int32_t* qpos = positions + (size_t)th * N;
  const int32_t* kpos = positions + ((size_t)T * H + th) * N;
  const float* scale_h = scale + h * C;
  const int tid = threadIdx.x;

  if constexpr (ROLE == BwdRole::DQ)
    gather_streamed_rows<GL, false>(k, k, v, nullptr, nullptr, coords, scale_h, kpos, kpos, blk0, nb, h, H, raw_size, hs, as);
  else
    gather_streamed_rows<GL, true>(q, k, d_out_pre, out_pre, den_sum, coords, scale_h, qpos, kpos, blk0, nb, h, H, raw_size, hs, as);
  __syncthreads();

  if (tid >= P::LANES) return;
  const int pair = tid >> 1, hf = tid & 1;
  const int g = pair / P::RG, pp = pair - g * P::RG;
  const int blk = blk0 + g;
  if (blk >= nb) return;
  const int n0 = __ldg(kpos + (size_t)blk * B + (B - 1));
  const int32_t* rpos = (ROLE == BwdRole::DQ ? qpos : kpos) + (size_t)blk * B;

  int nrow[R];
  float2 a2[R][2 * HCH];        // resident x' half
  float2 w2[R][2 * VH];         // DQ: gd half; DK / DV: v half
  float2 acc2[R][2 * HCH];      // DQ / DK: accumulated dS * streamed hat half
  float2 dv2[R][2 * VH];        // DV: accumulated P * gd half
  float nres[R], gyr[R], sds[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    nrow[r] = __ldg(rpos + pp + r * P::RG);
    float hsq;
    load_half_row<P>(ROLE == BwdRole::DQ ? q : k, k, coords, scale_h, nrow[r], n0, h, H, raw_size, hf, a2[r], hsq);
    nres[r] = -0.5f * (hsq + __shfl_xor_sync(0xffffffffu, hsq, 1));
    gyr[r] = 0.f;
    if constexpr (ROLE == BwdRole::DQ) {
      float hgy;
      load_half_grad<P>(d_out_pre, out_pre, den_sum, nrow[r], h, H, hf, w2[r], hgy);
      gyr[r] = hgy + __shfl_xor_sync(0xffffffffu, hgy, 1);
    } else {
#pragma unroll
      for (int cc = 0; cc < VH; ++cc) {
        const float4 vv = nrow[r] < raw_size ? ldg4(v + ((size_t)nrow[r] * H + h) * D + 4 * (hf * VH + cc))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        w2[r][2 * cc] = f2(vv.x, vv.y);
        w2[r][2 * cc + 1] = f2(vv.z, vv.w);
      }
    }
    sds[r] = 0.f;
#pragma unroll
    for (int x = 0; x < 2 * HCH; ++x) acc2[r][x] = f2(0.f, 0.f);
#pragma unroll
    for (int x = 0; x < 2 * VH; ++x) dv2[r][x] = f2(0.f, 0.f);
  }

  // side slots of a streamed row live in the second half: slot E (nk / nq) and E+1 (gy)
  constexpr int SIDE_CHUNK = E / 4 - HCH, SIDE_POS = E % 4;            // within the half-1 lane's chunks
  constexpr int SIDE1_CHUNK = (E + 1) / 4 - HCH, SIDE1_POS = (E + 1) % 4;
  const float4* hrow = hs + (size_t)g * B * P::ROW_CHUNKS + hf * HCH;
  const float4* arow = as + (size_t)g * B * P::VCH + hf * VH;
#pragma unroll 1
  for (int j = 0; j < B; ++j, hrow += P::ROW_CHUNKS, arow += P::VCH) {
    // all shared-memory loads of the iteration first, then the multiply-adds with the R rows interleaved and every
    // dot product split over two accumulators (even / odd chunks): four independent FFMA2 chains instead of one
    float4 kk[HCH], gv[VH];
#pragma unroll
    for (int cc = 0; cc < HCH; ++cc) kk[cc] = hrow[cc];
#pragma unroll
    for (int cc = 0; cc < VH; ++cc) gv[cc] = arow[cc];
    float2 sA[R], sB[R], dA[R], dB[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { sA[r] = f2(0.f, 0.f); sB[r] = sA[r]; dA[r] = sA[r]; dB[r] = sA[r]; }
#pragma unroll
    for (int cc = 0; cc < HCH; ++cc) {
#pragma unroll
      for (int r = 0; r < R; ++r) chunk_fma2((cc & 1) ? sB[r] : sA[r], &a2[r][2 * cc], kk[cc]);
    }
    if constexpr (ROLE != BwdRole::DV) {
#pragma unroll
      for (int cc = 0; cc < VH; ++cc) {
#pragma unroll
        for (int r = 0; r < R; ++r) chunk_fma2((cc & 1) ? dB[r] : dA[r], &w2[r][2 * cc], gv[cc]);
      }
    }
    float hd[R], hdp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      hd[r] = (sA[r].x + sA[r].y) + (sB[r].x + sB[r].y);
      hdp[r] = (dA[r].x + dA[r].y) + (dB[r].x + dB[r].y);
    }
    // side slots of the streamed row, from the lane that holds the second half
    float side0 = 0.f, side1 = 0.f;
    {
      const float4 sc4 = kk[SIDE_CHUNK < 0 ? 0 : SIDE_CHUNK];
      const float tsc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
      const float4 sd4 = kk[SIDE1_CHUNK < 0 ? 0 : SIDE1_CHUNK];
      const float tsd[4] = {sd4.x, sd4.y, sd4.z, sd4.w};
      side0 = __shfl_sync(0xffffffffu, tsc[SIDE_POS], (tid & 31) | 1);
      if constexpr (ROLE != BwdRole::DQ) side1 = __shfl_sync(0xffffffffu, tsd[SIDE1_POS], (tid & 31) | 1);
    }
    float dotv[R], dpv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      dotv[r] = __shfl_xor_sync(0xffffffffu, hd[r], 1);
      dpv[r] = ROLE != BwdRole::DV ? __shfl_xor_sync(0xffffffffu, hdp[r], 1) : 0.f;
    }
    float tt[R], pr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float dot = hd[r] + dotv[r];
      // canonical order (dot + nq) + nk: nq is the query-side norm, nk the key-side one
      tt[r] = ROLE == BwdRole::DQ ? (dot + nres[r]) + side0 : (dot + side0) + nres[r];
      pr[r] = exp2_fast(fminf(tt[r] * kLog2e, 0.f));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float p = pr[r];
      if constexpr (ROLE == BwdRole::DV) {
        const float2 p2 = f2(p, p);
#pragma unroll
        for (int cc = 0; cc < VH; ++cc) {
          dv2[r][2 * cc] = __ffma2_rn(p2, f2(gv[cc].x, gv[cc].y), dv2[r][2 * cc]);
          dv2[r][2 * cc + 1] = __ffma2_rn(p2, f2(gv[cc].z, gv[cc].w), dv2[r][2 * cc + 1]);
        }
      } else {
        const float gy = ROLE == BwdRole::DQ ? gyr[r] : side1;
        const float dp = (hdp[r] + dpv[r]) - gy;
        const float ds = tt[r] <= 0.f ? p * dp : 0.f;      // clamp(max=0) passes gradient where S <= 0
        sds[r] += ds;
        const float2 ds2 = f2(ds, ds);
#pragma unroll
        for (int cc = 0; cc < HCH; ++cc) {
          acc2[r][2 * cc] = __ffma2_rn(ds2, f2(kk[cc].x, kk[cc].y), acc2[r][2 * cc]);
          acc2[r][2 * cc + 1] = __ffma2_rn(ds2, f2(kk[cc].z, kk[cc].w), acc2[r][2 * cc + 1]);
        }
      }
    }
  }

  // ---- write this lane's half of every resident row to the per-table staging rows ---------------------------------
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const size_t srow = ((size_t)h * N + nrow[r]) * T + t;
    if constexpr (ROLE == BwdRole::DV) {
      float4* dst = reinterpret_cast<float4*>(stage_out + srow * D) + hf * VH;
#pragma unroll
      for (int cc = 0; cc < VH; ++cc)
        dst[cc] = make_float4(dv2[r][2 * cc].x, dv2[r][2 * cc].y, dv2[r][2 * cc + 1].x, dv2[r][2 * cc + 1].y);
    } else {
      // d x^_row = sum dS * streamed' - (sum dS) * x'_row
      float4* dst = reinterpret_cast<float4*>(stage_out + srow * kStageRow) + hf * HCH;
#pragma unroll
      for (int cc = 0; cc < HCH; ++cc) {
        const float o[4] = {fmaf(-sds[r], a2[r][2 * cc].x, acc2[r][2 * cc].x), fmaf(-sds[r], a2[r][2 * cc].y, acc2[r][2 * cc].y),
                            fmaf(-sds[r], a2[r][2 * cc + 1].x, acc2[r][2 * cc + 1].x),
                            fmaf(-sds[r], a2[r][2 * cc + 1].y, acc2[r][2 * cc + 1].y)};
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = (4 * (hf * HCH + cc) + u < E) ? o[u] : 0.f;
        dst[cc] = make_float4(w[0], w[1], w[2], w[3]);
      }
      if constexpr (P::USED_CHUNKS < P::ROW_CHUNKS) {      // short hat rows: keep the unused tail of the 128-byte row zero
        if (hf == 1) {
          float4* tail = reinterpret_cast<float4*>(stage_out + srow * kStageRow) + P::USED_CHUNKS;
#pragma unroll
          for (int cc = 0; cc < P::ROW_CHUNKS - P::USED_CHUNKS; ++cc) tail[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
}

template <int D, int C, int B, int G, int R, int MINB, BwdRole ROLE>
static int launch_pair(const char* name, const hept_shape* s, const float* q, const float* k, const float* v,
                       const float* coords, const float* scale, const int32_t* positions, const float* out_pre,
                       const float* den_sum, const float* d_out_pre, float* stage_out, cudaStream_t st) {
  using P = PairLayout<D, C, B, G, R>;
  auto kern = block_attn_bwd_pair_kernel<D, C, B, G, R, MINB, ROLE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM_BYTES);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "%s: cannot reserve %zu B of shared memory: %s", name, (size_t)P::SMEM_BYTES,
                 cudaGetErrorString(e));
    configured = true;
  }
  const int nb = s->N / s->B;
  kern<<<dim3((nb + G - 1) / G, s->T * s->H), P::THREADS, P::SMEM_BYTES, st>>>(
      q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, s->N, s->H, s->T, s->raw_size, stage_out);
  HEPT_CHECK_LAUNCH(name);
  return HEPT_OK;
}

template <int D, int C, int B, int G, int R, int MINB>
static int launch_all(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                      const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                      const float* d_out_pre, float* stage_dq, float* stage_dk, float* stage_dv, int mask, cudaStream_t st) {
  int rc = HEPT_OK;
  if ((mask & 1) && (rc = launch_pair<D, C, B, G, R, MINB, BwdRole::DQ>("block_attn_bwd_dq2", s, q, k, v, coords, scale, positions,
                                                                         out_pre, den_sum, d_out_pre, stage_dq, st)))
    return rc;
  if (mask & 2) {
    if ((rc = launch_pair<D, C, B, G, R, MINB, BwdRole::DK>("block_attn_bwd_dk2", s, q, k, v, coords, scale, positions, out_pre,
                                                            den_sum, d_out_pre, stage_dk, st)))
      return rc;
    if ((rc = launch_pair<D, C, B, G, R, MINB, BwdRole::DV>("block_attn_bwd_dv2", s, q, k, v, coords, scale, positions, out_pre,
                                                            den_sum, d_out_pre, stage_dv, st)))
      return rc;
  }
  return rc;
}

// second-generation tile kernels of hept_block_attention_bwd: bit 0 of `mask` = dq, bit 1 = dk and dv
int block_attention_bwd_tiles2(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                               const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                               const float* d_out_pre, float* stage_dq, float* stage_dk, float* stage_dv, int mask,
                               cudaStream_t st) {
  if (s->D == 24 && s->C == 6 && s->B == 100)
    return launch_all<24, 6, 100, 5, 2, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, stage_dq, stage_dk, stage_dv, mask, st);
  if (s->D == 24 && s->C == 4 && s->B == 100)
    return launch_all<24, 4, 100, 5, 2, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, stage_dq, stage_dk, stage_dv, mask, st);
  if (s->D == 8 && s->C == 6 && s->B == 10)
    return launch_all<8, 6, 10, 4, 2, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, stage_dq, stage_dk, stage_dv, mask, st);
  set_error("block_attention_bwd (pair tiles): (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}

}  // namespace hept
