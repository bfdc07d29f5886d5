// a18: backward of a8-a12 (autograd of example/hept.py:7-18,70-79 and of the coordinate part of
// prep_qk, example/hept.py:25-27).  The reference saves the (T,H,nb,B,B) score tensors; here P is
// recomputed per tile from the gathered rows (only the two sort permutations, out_pre and the summed
// normaliser are saved).
//
// With g = d out_pre[n,h,:], y = out_pre[n,h,:], den = sum_t denom_t:
//   d so_t = g / den =: gd          d denom_t = -(g . y) / den =: -gy       (same for every table t)
//   dP_ij = gd_i . v_j - gy_i       dS_ij = [S_ij <= 0] P_ij dP_ij          (clamp(max=0) passes S <= 0)
//   dv_j  = sum_i P_ij gd_i
//   dq^_i = sum_j dS_ij (k^_j - q^_i)          dk^_j = sum_i dS_ij (q^_i - k^_j)
// Two tile kernels with the forward's structure: "dq" (lanes own query rows, keys/values streamed) and
// "dkv" (lanes own key rows, queries/gd streamed).  Per-table results go to staging rows in original
// hit order and a streaming kernel sums the T tables (deterministic: no floating-point atomics).
#include "tile.cuh"

namespace hept {

// ---------------------------------------------------------------------------------------------------
// dq: resident query row i (a = q', gd, gy); streamed k'_j (+nk), v_j.
// ---------------------------------------------------------------------------------------------------
template <int D, int C, int B, int G, int MINB>
__global__ void __launch_bounds__((TileLayout<D, C, B, G, 1>::THREADS), MINB)
    block_attn_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ coords, const float* __restrict__ scale,
                             const int32_t* __restrict__ positions, const float* __restrict__ out_pre,
                             const float* __restrict__ den_sum, const float* __restrict__ d_out_pre, int N, int H,
                             int T, int raw_size, float* __restrict__ stage_dq) {
  using L = TileLayout<D, C, B, G, 1>;
  constexpr int E = L::E;
  extern __shared__ float4 smem[];
  float4* ks = smem;
  float4* vs = smem + G * B * L::ROW_CHUNKS;

  const int h = blockIdx.y / T, t = blockIdx.y % T, th = t * H + h;  // tables of one head run back to back: its q/k/v slices stay in L2
  const int nb = N / B;
  const int blk0 = blockIdx.x * G;
  const int32_t* qpos = positions + (size_t)th * N;
  const int32_t* kpos = positions + ((size_t)T * H + th) * N;
  const float* scale_h = scale + h * C;
  const int tid = threadIdx.x;

  gather_streamed_rows<L, false>(k, k, v, nullptr, nullptr, coords, scale_h, kpos, kpos, blk0, nb, h, H, raw_size, ks, vs);
  __syncthreads();

  if (tid >= L::LANES) return;
  const int g = tid / L::LPB, i = tid - g * L::LPB;
  const int blk = blk0 + g;
  if (blk >= nb) return;

  const int n = __ldg(qpos + (size_t)blk * B + i);
  float a[1][E], nq, gd[D], gy;
  load_resident_row<L>(q, k, coords, scale_h, n, __ldg(kpos + (size_t)blk * B + (B - 1)), h, H, raw_size, a[0], nq);
  load_resident_grad<L>(d_out_pre, out_pre, den_sum, n, h, H, gd, gy);
  float acc[E], sds = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;

  const float4* krow = ks + (size_t)g * B * L::ROW_CHUNKS;
  const float4* vrow = vs + (size_t)g * B * L::VCH;
#pragma unroll 1
  for (int j = 0; j < B; ++j, krow += L::ROW_CHUNKS, vrow += L::VCH) {
    float s[1], nk = 0.f, unused = 0.f;
    float4 keep[L::USED_CHUNKS];
    dot_rows<L>(krow, a, s, nk, unused, keep);
    const float s2 = (s[0] + nq) + nk;  // canonical order (dot + nq) + nk, see tile.cuh
    const float p = exp2_fast(fminf(s2 * kLog2e, 0.f));
    float dp0 = -gy, dp1 = 0.f;
#pragma unroll
    for (int c = 0; c < L::VCH; ++c)
      dp_step(make_float4(gd[4 * c], gd[4 * c + 1], gd[4 * c + 2], gd[4 * c + 3]), vrow[c], dp0, dp1);
    const float ds = s2 <= 0.f ? p * (dp0 + dp1) : 0.f;  // clamp(max=0) passes gradient where S <= 0
    sds += ds;
#pragma unroll
    for (int c = 0; c < L::USED_CHUNKS; ++c) {
      const float tk[4] = {keep[c].x, keep[c].y, keep[c].z, keep[c].w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
        if (4 * c + x < E) acc[4 * c + x] = fmaf(ds, tk[x], acc[4 * c + x]);
    }
  }

  // dq^_i = sum_j dS_ij k'_j - (sum_j dS_ij) q'_i   (a[0] holds q' exactly)
  float4* dst = reinterpret_cast<float4*>(stage_dq + (((size_t)h * N + n) * T + t) * kStageRow);
#pragma unroll
  for (int c = 0; c < L::ROW_CHUNKS; ++c) {
    float o4[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int e = 4 * c + x;
      o4[x] = e < E ? fmaf(-sds, a[0][e < E ? e : 0], acc[e < E ? e : 0]) : 0.f;
    }
    dst[c] = make_float4(o4[0], o4[1], o4[2], o4[3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// dkv: resident key row j (a = k', v_j); streamed q'_i (+nq, gy_i) and gd_i.
// ---------------------------------------------------------------------------------------------------
template <int D, int C, int B, int G, int MINB>
__global__ void __launch_bounds__((TileLayout<D, C, B, G, 1>::THREADS), MINB)
    block_attn_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                              const float* __restrict__ coords, const float* __restrict__ scale,
                              const int32_t* __restrict__ positions, const float* __restrict__ out_pre,
                              const float* __restrict__ den_sum, const float* __restrict__ d_out_pre, int N, int H,
                              int T, int raw_size, float* __restrict__ stage_dk, float* __restrict__ stage_dv) {
  using L = TileLayout<D, C, B, G, 1>;
  constexpr int E = L::E;
  extern __shared__ float4 smem[];
  float4* qs = smem;                           // [G*B][8]   q' rows: q'[0..E), nq, gy
  float4* gs = smem + G * B * L::ROW_CHUNKS;   // [G*B][D/4] gd rows

  const int h = blockIdx.y / T, t = blockIdx.y % T, th = t * H + h;  // tables of one head run back to back: its q/k/v slices stay in L2
  const int nb = N / B;
  const int blk0 = blockIdx.x * G;
  const int32_t* qpos = positions + (size_t)th * N;
  const int32_t* kpos = positions + ((size_t)T * H + th) * N;
  const float* scale_h = scale + h * C;
  const int tid = threadIdx.x;

  gather_streamed_rows<L, true>(q, k, d_out_pre, out_pre, den_sum, coords, scale_h, qpos, kpos, blk0, nb, h, H,
                                raw_size, qs, gs);
  __syncthreads();

  if (tid >= L::LANES) return;
  const int g = tid / L::LPB, j = tid - g * L::LPB;
  const int blk = blk0 + g;
  if (blk >= nb) return;

  const int n = __ldg(kpos + (size_t)blk * B + j);
  float a[1][E], nk, vj[D];
  load_resident_row<L>(k, k, coords, scale_h, n, __ldg(kpos + (size_t)blk * B + (B - 1)), h, H, raw_size, a[0], nk);
  if (n < raw_size) load_row<D>(v + ((size_t)n * H + h) * D, vj);
  else {
#pragma unroll
    for (int d = 0; d < D; ++d) vj[d] = 0.f;
  }
  float dk[E], dv[D], sds = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) dk[e] = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) dv[d] = 0.f;

  const float4* qrow = qs + (size_t)g * B * L::ROW_CHUNKS;
  const float4* grow = gs + (size_t)g * B * L::VCH;
#pragma unroll 1
  for (int i = 0; i < B; ++i, qrow += L::ROW_CHUNKS, grow += L::VCH) {
    float s[1], nq = 0.f, gy = 0.f;
    float4 keep[L::USED_CHUNKS];
    dot_rows<L>(qrow, a, s, nq, gy, keep);
    const float s2 = (s[0] + nq) + nk;  // canonical order (dot + nq) + nk: bit-identical to the dq pass
    const float p = exp2_fast(fminf(s2 * kLog2e, 0.f));
    float dp0 = -gy, dp1 = 0.f;
#pragma unroll
    for (int c = 0; c < L::VCH; ++c) {
      const float4 gg = grow[c];
      dp_step(gg, make_float4(vj[4 * c], vj[4 * c + 1], vj[4 * c + 2], vj[4 * c + 3]), dp0, dp1);
      dv[4 * c + 0] = fmaf(p, gg.x, dv[4 * c + 0]);
      dv[4 * c + 1] = fmaf(p, gg.y, dv[4 * c + 1]);
      dv[4 * c + 2] = fmaf(p, gg.z, dv[4 * c + 2]);
      dv[4 * c + 3] = fmaf(p, gg.w, dv[4 * c + 3]);
    }
    const float ds = s2 <= 0.f ? p * (dp0 + dp1) : 0.f;
    sds += ds;
#pragma unroll
    for (int c = 0; c < L::USED_CHUNKS; ++c) {
      const float tq[4] = {keep[c].x, keep[c].y, keep[c].z, keep[c].w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
        if (4 * c + x < E) dk[4 * c + x] = fmaf(ds, tq[x], dk[4 * c + x]);
    }
  }

  // dk^_j = sum_i dS_ij q'_i - (sum_i dS_ij) k'_j   (a[0] holds k' exactly)
  const size_t srow = ((size_t)h * N + n) * T + t;
  float4* dstk = reinterpret_cast<float4*>(stage_dk + srow * kStageRow);
#pragma unroll
  for (int c = 0; c < L::ROW_CHUNKS; ++c) {
    float o4[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int e = 4 * c + x;
      o4[x] = e < E ? fmaf(-sds, a[0][e < E ? e : 0], dk[e < E ? e : 0]) : 0.f;
    }
    dstk[c] = make_float4(o4[0], o4[1], o4[2], o4[3]);
  }
  float4* dstv = reinterpret_cast<float4*>(stage_dv + srow * D);
#pragma unroll
  for (int c = 0; c < L::VCH; ++c) dstv[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
}

// ---------------------------------------------------------------------------------------------------
// Sum the T per-table staging rows of every (hit, head) into dq, dk, dv and fold the coordinate
// columns into per-CTA partial sums of d scale[h,c] = sum_n coords[n,c] (dq^ + dk^)[n,h,D+c].
// One warp = 32 consecutive hits of one head.
// ---------------------------------------------------------------------------------------------------
template <int D, int C>
__global__ void __launch_bounds__(256) bwd_reduce_kernel(const float* __restrict__ stage_dq,
                                                         const float* __restrict__ stage_dk,
                                                         const float* __restrict__ stage_dv,
                                                         const float* __restrict__ coords, int N, int H, int T,
                                                         int raw_size, float* __restrict__ dq, float* __restrict__ dk,
                                                         float* __restrict__ dv, float* __restrict__ partial) {
  constexpr int E = D + C;
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool live = n < N, real = live && n < raw_size;
  float cc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) cc[c] = 0.f;
  if (real) load_row<C>(coords + (size_t)n * C, cc);
  for (int h = threadIdx.x >> 5; h < H; h += warps) {
    float sq[E], sk[E], sv[D];
#pragma unroll
    for (int e = 0; e < E; ++e) { sq[e] = 0.f; sk[e] = 0.f; }
#pragma unroll
    for (int d = 0; d < D; ++d) sv[d] = 0.f;
    if (real) {
      const size_t srow = ((size_t)h * N + n) * T;
      for (int t = 0; t < T; ++t) {
        const float* rq = stage_dq + (srow + t) * kStageRow;
        const float* rk = stage_dk + (srow + t) * kStageRow;
        const float* rv = stage_dv + (srow + t) * D;
#pragma unroll
        for (int c4 = 0; c4 < (E + 3) / 4; ++c4) {
          const float4 x = ldg4(rq + 4 * c4), y = ldg4(rk + 4 * c4);
          const float tx[4] = {x.x, x.y, x.z, x.w}, ty[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (4 * c4 + u < E) { sq[4 * c4 + u] += tx[u]; sk[4 * c4 + u] += ty[u]; }
        }
#pragma unroll
        for (int c4 = 0; c4 < D / 4; ++c4) {
          const float4 x = ldg4(rv + 4 * c4);
          sv[4 * c4] += x.x; sv[4 * c4 + 1] += x.y; sv[4 * c4 + 2] += x.z; sv[4 * c4 + 3] += x.w;
        }
      }
    }
    if (live) {
      float4* oq = reinterpret_cast<float4*>(dq + ((size_t)n * H + h) * D);
      float4* ok = reinterpret_cast<float4*>(dk + ((size_t)n * H + h) * D);
      float4* ov = reinterpret_cast<float4*>(dv + ((size_t)n * H + h) * D);
#pragma unroll
      for (int c4 = 0; c4 < D / 4; ++c4) {
        oq[c4] = make_float4(sq[4 * c4], sq[4 * c4 + 1], sq[4 * c4 + 2], sq[4 * c4 + 3]);
        ok[c4] = make_float4(sk[4 * c4], sk[4 * c4 + 1], sk[4 * c4 + 2], sk[4 * c4 + 3]);
        ov[c4] = make_float4(sv[4 * c4], sv[4 * c4 + 1], sv[4 * c4 + 2], sv[4 * c4 + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float x = cc[c] * (sq[D + c] + sk[D + c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) partial[((size_t)blockIdx.x * H + h) * C + c] = x;
    }
  }
}

// dscale[hc] = sum over CTAs of partial[cta][hc], fixed-order tree -> deterministic.
__global__ void __launch_bounds__(256) dscale_finish_kernel(const float* __restrict__ partial, int nblk, int hc,
                                                            float* __restrict__ dscale) {
  __shared__ float red[256];
  const int col = blockIdx.x;
  float x = 0.f;
  for (int b = threadIdx.x; b < nblk; b += 256) x += partial[(size_t)b * hc + col];
  red[threadIdx.x] = x;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dscale[col] = red[0];
}

int block_attention_bwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                           const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                           const float* d_out_pre, float* dq, float* dk, float* dv, float* dscale, char* ws,
                           cudaStream_t st);
size_t bwd_tc_workspace_bytes(const hept_shape* s);

struct BwdPlan {
  size_t dq_bytes, dv_bytes, partial_bytes, total;
  int nblk;
};
static BwdPlan plan_bwd(const hept_shape* s) {
  BwdPlan p;
  p.nblk = (s->N + 31) / 32;
  p.dq_bytes = align_up(sizeof(float) * (size_t)s->H * s->N * s->T * kStageRow, 256);
  p.dv_bytes = align_up(sizeof(float) * (size_t)s->H * s->N * s->T * s->D, 256);
  p.partial_bytes = align_up(sizeof(float) * (size_t)p.nblk * s->H * s->C, 256);
  p.total = 2 * p.dq_bytes + p.dv_bytes + p.partial_bytes;
  return p;
}

template <int D, int C, int B, int GQ, int MINQ, int GK, int MINK>
static int launch_bwd(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                      const float* scale, const int32_t* positions, const float* out_pre, const float* den_sum,
                      const float* d_out_pre, float* dq, float* dk, float* dv, float* dscale, char* ws,
                      cudaStream_t st) {
  using LQ = TileLayout<D, C, B, GQ, 1>;
  using LK = TileLayout<D, C, B, GK, 1>;
  auto kq = block_attn_bwd_dq_kernel<D, C, B, GQ, MINQ>;
  auto kk = block_attn_bwd_dkv_kernel<D, C, B, GK, MINK>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LQ::SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LK::SMEM_BYTES);
    HEPT_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, HEPT_ECUDA, "block_attn_bwd: cannot reserve shared memory");
    configured.mark();
  }
  BwdPlan p = plan_bwd(s);
  float* stage_dq = (float*)ws;
  float* stage_dk = (float*)(ws + p.dq_bytes);
  float* stage_dv = (float*)(ws + 2 * p.dq_bytes);
  float* partial = (float*)(ws + 2 * p.dq_bytes + p.dv_bytes);
  const int nb = s->N / s->B;
  const int mask = bwd_stage_mask();  // profiling aid: all stages unless hept_set_bwd_stage_mask() says otherwise
  if (mask & 1) {
    kq<<<dim3((nb + GQ - 1) / GQ, s->T * s->H), LQ::THREADS, LQ::SMEM_BYTES, st>>>(
        q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, s->N, s->H, s->T, s->raw_size, stage_dq);
    HEPT_CHECK_LAUNCH("block_attn_bwd_dq");
  }
  if (mask & 2) {
    kk<<<dim3((nb + GK - 1) / GK, s->T * s->H), LK::THREADS, LK::SMEM_BYTES, st>>>(
        q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, s->N, s->H, s->T, s->raw_size, stage_dk, stage_dv);
    HEPT_CHECK_LAUNCH("block_attn_bwd_dkv");
  }
  if (!(mask & 4)) return HEPT_OK;
  bwd_reduce_kernel<D, C><<<p.nblk, 256, 0, st>>>(stage_dq, stage_dk, stage_dv, coords, s->N, s->H, s->T, s->raw_size,
                                                  dq, dk, dv, partial);
  HEPT_CHECK_LAUNCH("bwd_reduce");
  dscale_finish_kernel<<<s->H * s->C, 256, 0, st>>>(partial, p.nblk, s->H * s->C, dscale);
  HEPT_CHECK_LAUNCH("dscale_finish");
  return HEPT_OK;
}

}  // namespace hept

using namespace hept;

extern "C" size_t hept_attention_bwd_workspace_bytes(const hept_shape* s) {
  if (!s || s->N <= 0) return 0;
  const size_t a = plan_bwd(s).total, b = bwd_tc_workspace_bytes(s);
  return a > b ? a : b;   // whichever backward engine is selected later fits
}

extern "C" int hept_block_attention_bwd(const hept_shape* s, const float* q, const float* k, const float* v,
                                        const float* coords, const float* scale, const int32_t* positions,
                                        const float* out_pre, const float* den_sum, const float* d_out_pre, float* dq,
                                        float* dk, float* dv, float* dscale, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(q && k && v && coords && scale && positions && out_pre && den_sum && d_out_pre && dq && dk && dv &&
                   dscale && workspace,
               HEPT_EINVAL, "block_attention_bwd: null pointer");
  HEPT_REQUIRE(workspace_bytes >= hept_attention_bwd_workspace_bytes(s), HEPT_EWORKSPACE,
               "block_attention_bwd: workspace needs %zu bytes", hept_attention_bwd_workspace_bytes(s));
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  if (bwd_variant() >= 3 && tc_tiles_supported(s->D, s->C, s->B))
    return block_attention_bwd_tc(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 6 && s->B == 100)
    return launch_bwd<24, 6, 100, 5, 1, 3, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 4 && s->B == 100)
    return launch_bwd<24, 4, 100, 5, 1, 3, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 6 && s->B == 64)
    return launch_bwd<24, 6, 64, 8, 1, 5, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 4 && s->B == 64)
    return launch_bwd<24, 4, 64, 8, 1, 5, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 6 && s->B == 128)
    return launch_bwd<24, 6, 128, 4, 1, 2, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 24 && s->C == 4 && s->B == 128)
    return launch_bwd<24, 4, 128, 4, 1, 2, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  if (s->D == 8 && s->C == 6 && s->B == 10)
    return launch_bwd<8, 6, 10, 4, 1, 4, 1>(s, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre, dq, dk, dv, dscale, ws, st);
  set_error("block_attention_bwd: (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}
