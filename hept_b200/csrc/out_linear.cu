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
#include "mma_tf32.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// Register-tiled versions for IN = 8 OUT (the shipped H = 8 shapes).  The row-per-lane kernels above are bound by the
// L1 tag stage (a warp's 16-byte accesses touch 32 lines) and by the 128 B/clk shared-memory return path (one
// LDS.128 — broadcast or not — delivers 512 B per warp for 4 FMAs per lane).  Here the rows travel global -> shared
// with fully coalesced cp.async (double-buffered), and a thread owns a 4-hit x 6-output (forward), 4-hit x 8-column
// (input gradient) or 8 x 8 (weight gradient) register tile, so a 16-byte shared-memory load feeds 16-32 FMAs.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTlHits = 128;                       // hits per CTA tile (forward, input gradient)
constexpr int kTlThreads = 128;

// forward: thread = (q = tid / 4, og = tid % 4): hits q, q+32, q+64, q+96 x outputs [6 og, 6 og + 6)
template <int OUT, int IN>
__global__ void __launch_bounds__(kTlThreads) out_linear_fwd_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                          const float* __restrict__ b, int N, float* __restrict__ out) {
  constexpr int KC = 48, NCH = IN / KC, XS = KC + 4, WS = IN + 4, PER = OUT / 4;
  static_assert(IN % KC == 0 && OUT % 4 == 0 && PER % 2 == 0, "tile shape");
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                              // (OUT, WS)
  float* s_x = s_w + OUT * WS;                     // (kTlHits, XS): one buffer; 45 KB per CTA -> 5 CTAs per SM, so the 469 tiles
                                                   // of a 60k-hit event are ONE wave and other CTAs cover a CTA's load latency
  const int tid = threadIdx.x, q = tid >> 2, og = tid & 3;
  const int n0 = blockIdx.x * kTlHits;
  const int rows = min(kTlHits, N - n0);
  auto load_chunk = [&](int kc) {
    float* dst = s_x;
    for (int i = tid; i < kTlHits * (KC / 4); i += kTlThreads) {
      const int r = i / (KC / 4), c4 = i - r * (KC / 4);
      if (r < rows) cp_async16_cg(dst + r * XS + 4 * c4, x + (size_t)(n0 + r) * IN + kc * KC + 4 * c4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_chunk(0);
  for (int i = tid; i < OUT * (IN / 4); i += kTlThreads) {
    const int j = i / (IN / 4), c4 = i - j * (IN / 4);
    *reinterpret_cast<float4*>(s_w + j * WS + 4 * c4) = ldg4(w + (size_t)j * IN + 4 * c4);
  }
  // packed fp32 FMAs (fma.rn.f32x2, sm_100): an accumulator pair holds the sums over the even and the odd k of a
  // 16-byte chunk pair; two FMAs per issue slot, the halves are added at the end
  float2 acc[4][PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const float bu = __ldg(b + og * PER + u);
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][u] = make_float2(bu, 0.f);
  }
#pragma unroll 1
  for (int kc = 0; kc < NCH; ++kc) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float* xs = s_x + q * XS;
    const float* ws = s_w + og * PER * WS + kc * KC;
#pragma unroll 4
    for (int c4 = 0; c4 < KC / 4; ++c4) {
      float4 xv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) xv[t] = *reinterpret_cast<const float4*>(xs + 32 * t * XS + 4 * c4);
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + u * WS + 4 * c4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc[t][u] = __ffma2_rn(make_float2(xv[t].x, xv[t].y), make_float2(wv.x, wv.y), acc[t][u]);
          acc[t][u] = __ffma2_rn(make_float2(xv[t].z, xv[t].w), make_float2(wv.z, wv.w), acc[t][u]);
        }
      }
    }
    __syncthreads();
    if (kc + 1 < NCH) load_chunk(kc + 1);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int r = q + 32 * t;
    if (r < rows) {
      float2* dst = reinterpret_cast<float2*>(out + (size_t)(n0 + r) * OUT + og * PER);
#pragma unroll
      for (int u2 = 0; u2 < PER / 2; ++u2)
        dst[u2] = make_float2(acc[t][2 * u2].x + acc[t][2 * u2].y, acc[t][2 * u2 + 1].x + acc[t][2 * u2 + 1].y);
    }
  }
}

// input gradient: thread = (q = tid / 4, cg = tid % 4 + 4 i): hits q + 32 u x columns [8 cg, 8 cg + 8), i < IN / 32
template <int OUT, int IN>
__global__ void __launch_bounds__(kTlThreads, 4) out_linear_bwd_input_tiled_kernel(const float* __restrict__ g,
                                                                                const float* __restrict__ w, int N,
                                                                                float* __restrict__ dx) {
  constexpr int GS = OUT + 4;
  static_assert(OUT % 4 == 0 && IN % 32 == 0, "tile shape");
  extern __shared__ __align__(16) float s_dyn[];
  float* s_w = s_dyn;                              // (OUT, IN)
  float* s_g = s_w + OUT * IN;                     // (kTlHits, GS)
  const int tid = threadIdx.x, q = tid >> 2, c0 = tid & 3;
  const int n0 = blockIdx.x * kTlHits;
  const int rows = min(kTlHits, N - n0);
  for (int i = tid; i < OUT * IN / 4; i += kTlThreads) reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  for (int i = tid; i < kTlHits * (OUT / 4); i += kTlThreads) {
    const int r = i / (OUT / 4), j4 = i - r * (OUT / 4);
    *reinterpret_cast<float4*>(s_g + r * GS + 4 * j4) =
        r < rows ? ldg4(g + (size_t)(n0 + r) * OUT + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
#pragma unroll 1
  for (int i = 0; i < IN / 32; ++i) {
    asm volatile("" ::: "memory");                 // keep the gradient rows in shared memory: hoisting their loads out of
                                                   // this loop costs 96 registers (spills, one CTA less per SM)
    const int cg = c0 + 4 * i;
    float4 a0[4], a1[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { a0[t] = make_float4(0.f, 0.f, 0.f, 0.f); a1[t] = a0[t]; }
#pragma unroll
    for (int j4 = 0; j4 < OUT / 4; ++j4) {
      float4 gv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) gv[t] = *reinterpret_cast<const float4*>(s_g + (q + 32 * t) * GS + 4 * j4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (4 * j4 + jj) * IN + 8 * cg);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + (4 * j4 + jj) * IN + 8 * cg + 4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float gj = jj == 0 ? gv[t].x : (jj == 1 ? gv[t].y : (jj == 2 ? gv[t].z : gv[t].w));
          const float2 g2 = make_float2(gj, gj);     // packed fp32 FMAs: two columns per issue slot, same sums per element
          float2 r;
          r = __ffma2_rn(g2, make_float2(w0.x, w0.y), make_float2(a0[t].x, a0[t].y)); a0[t].x = r.x; a0[t].y = r.y;
          r = __ffma2_rn(g2, make_float2(w0.z, w0.w), make_float2(a0[t].z, a0[t].w)); a0[t].z = r.x; a0[t].w = r.y;
          r = __ffma2_rn(g2, make_float2(w1.x, w1.y), make_float2(a1[t].x, a1[t].y)); a1[t].x = r.x; a1[t].y = r.y;
          r = __ffma2_rn(g2, make_float2(w1.z, w1.w), make_float2(a1[t].z, a1[t].w)); a1[t].z = r.x; a1[t].w = r.y;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int r = q + 32 * t;
      if (r < rows) {
        st_global_v8(dx + (size_t)(n0 + r) * IN + 8 * cg, a0[t], a1[t]);
      }
    }
  }
}

// weight / bias gradient, stage 1: a CTA = kPgGroups hit groups x (OUT / 8) x (IN / 8) threads; a thread owns an 8 x 8 tile
// of dW; slabs of kPgRows hits are staged in shared memory (double-buffered cp.async); group p takes the hits p, p + G, ...
// of a slab.  The groups' tiles are added through shared memory in group order (deterministic), db by the ct == 0 threads.
constexpr int kPgGroups = 4, kPgRows = 32, kPgStages = 2;   // (three stages measured: no gain, the loop is barrier-bound)
template <int OUT, int IN>
__global__ void __launch_bounds__(kPgGroups * (OUT / 8) * (IN / 8), 2) out_linear_bwd_params_tiled_kernel(
    const float* __restrict__ g, const float* __restrict__ x, int N, float* __restrict__ partial) {
  constexpr int JT = OUT / 8, CT = IN / 8, PER = JT * CT, THREADS = kPgGroups * PER;
  constexpr int XS = IN + 4, GS = OUT + 4, SLAB = kPgRows * (XS + GS);
  static_assert(2 * SLAB >= (OUT + 1) * IN, "the staging buffers double as the reduction buffer");
  extern __shared__ __align__(16) float s_dyn[];   // kPgStages x [ (kPgRows, XS) | (kPgRows, GS) ]: the loads of two slabs are in
                                                   // flight while one is consumed (a slab's compute is shorter than its load latency)
  const int tid = threadIdx.x, p = tid / PER, u = tid % PER;
  const int jt = u / CT, ct = u % CT;
  float2 acc[8][4];                                // 8 x 8 tile as column pairs: packed fp32 FMAs, two columns per issue slot
  float bsum[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    bsum[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = make_float2(0.f, 0.f);
  }
  const int slabs = (N + kPgRows - 1) / kPgRows;
  auto load_slab = [&](int slab, int buf) {
    float* xs = s_dyn + buf * SLAB;
    float* gs = xs + kPgRows * XS;
    const int n0 = slab * kPgRows, rows = min(kPgRows, N - n0);
    for (int i = tid; i < kPgRows * (IN / 4); i += THREADS) {
      const int r = i / (IN / 4), c4 = i - r * (IN / 4);
      if (r < rows) cp_async16_cg(xs + r * XS + 4 * c4, x + (size_t)(n0 + r) * IN + 4 * c4);
    }
    for (int i = tid; i < kPgRows * (OUT / 4); i += THREADS) {
      const int r = i / (OUT / 4), j4 = i - r * (OUT / 4);
      if (r < rows) cp_async16_cg(gs + r * GS + 4 * j4, g + (size_t)(n0 + r) * OUT + 4 * j4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int slab = blockIdx.x, it = 0;
#pragma unroll
  for (int pre = 0; pre < kPgStages - 1; ++pre) {      // one commit group per stage, empty ones included: the waits count groups
    if (slab + pre * (int)gridDim.x < slabs) load_slab(slab + pre * gridDim.x, pre);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll 1
  for (; slab < slabs; slab += gridDim.x, ++it) {
    const int next = slab + (kPgStages - 1) * (int)gridDim.x;
    if (next < slabs) load_slab(next, (it + kPgStages - 1) % kPgStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" :: "n"(kPgStages - 1) : "memory");
    __syncthreads();
    const float* xs = s_dyn + (it % kPgStages) * SLAB;
    const float* gs = xs + kPgRows * XS;
    const int rows = min(kPgRows, N - slab * kPgRows);
#pragma unroll 2
    for (int r = p; r < rows; r += kPgGroups) {
      const float4 g0 = *reinterpret_cast<const float4*>(gs + r * GS + 8 * jt), g1 = *reinterpret_cast<const float4*>(gs + r * GS + 8 * jt + 4);
      // the thread's eight columns are 4 ct .. 4 ct + 3 and IN / 2 + 4 ct ..: a warp's 128-bit loads cover consecutive words
      const float4 x0 = *reinterpret_cast<const float4*>(xs + r * XS + 4 * ct), x1 = *reinterpret_cast<const float4*>(xs + r * XS + IN / 2 + 4 * ct);
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float2 xv[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y), make_float2(x1.z, x1.w)};
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float2 g2 = make_float2(gv[a], gv[a]);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = __ffma2_rn(g2, xv[c], acc[a][c]);
        bsum[a] += gv[a];                          // only the ct == 0 threads' sums are used
      }
    }
    __syncthreads();
  }
  // groups add their tiles in order 0, 1, ... through shared memory (the staging buffers are free now)
  float* s_acc = s_dyn;                            // (OUT + 1, IN)
#pragma unroll 1
  for (int turn = 0; turn < kPgGroups; ++turn) {
    if (p == turn) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4* dst = reinterpret_cast<float4*>(s_acc + (8 * jt + a) * IN + h * (IN / 2) + 4 * ct);
          float4 v = make_float4(acc[a][2 * h].x, acc[a][2 * h].y, acc[a][2 * h + 1].x, acc[a][2 * h + 1].y);
          if (turn != 0) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *dst = v;
        }
        if (ct == 0) s_acc[OUT * IN + 8 * jt + a] = turn == 0 ? bsum[a] : s_acc[OUT * IN + 8 * jt + a] + bsum[a];
      }
    }
    __syncthreads();
  }
  float* dstp = partial + (size_t)blockIdx.x * (OUT + 1) * IN;
  for (int i = tid; i < OUT * IN + OUT; i += THREADS) dstp[i] = s_acc[i];
}

// The same product on the legacy tensor path (mma.sync m16n8k8, 3xTF32: lo*hi + hi*lo + hi*hi with fp32 accumulation): the
// CUDA-core kernel above is bound by its shared-memory wavefronts (70 % of the LSU data pipe, profiles/r2_ncu_prof_front_r2.csv),
// an mma needs an eighth of the operand loads per product.  P[c][j] = sum_n x[n][c] g[n][j]: M = the IN columns of x, N = the
// OUT columns of g, K = hits.  Slabs of 32 hits go through a cp.async ring of three slots (two slabs in flight per CTA, two
// CTAs per SM: 120 KB of loads in flight per SM); a warp = (k half hg, column group mg) multiplies 48 columns (three M tiles) x
// all OUT columns (three N tiles) over two of a slab's four k-steps.  k-slot t of k-step s is hit 4 s + t of the slab, k-slot
// t + 4 is hit 16 + 4 s + t (K is the contraction index: any bijection will do) -- with rows padded to 8 mod 32 words every
// fragment load is conflict-free.  blockIdx.y selects one of up to three operand pairs (the q / k / v weight gradients of the
// attention block in ONE launch).  The depth of one accumulation chain is N / (2 gridDim.x) hits; the CTAs' partials are
// added by out_linear_reduce_kernel in fixed order.
constexpr int kPmWarps = 8, kPmThreads = 32 * kPmWarps, kPmHalves = 2, kPmRows = 32, kPmStages = 3;
struct PgOperands {
  const float* g[3];       // (N, OUT)
  const float* x[3];       // (N, IN)
};
template <int OUT, int IN>
constexpr size_t params_mma_smem_bytes() { return sizeof(float) * kPmStages * (size_t)kPmRows * (IN + 8 + OUT + 16); }

template <int OUT, int IN, bool BIAS>
__global__ void __launch_bounds__(kPmThreads, 2) out_linear_bwd_params_mma_kernel(PgOperands ops, int N, float* __restrict__ partial) {
  constexpr int NT = OUT / 8, MG = kPmWarps / kPmHalves, MT = IN / (16 * MG);
  constexpr int XS = IN + 8, GS = OUT + 16, SLAB = kPmRows * (XS + GS);
  static_assert(OUT % 8 == 0 && IN % (16 * MG) == 0 && XS % 32 == 8 && GS % 32 == 8, "tile shape / bank padding");
  static_assert(kPmStages * SLAB >= MG * (MT * NT * 4 + NT) * 32, "the ring doubles as the reduction buffer");
  extern __shared__ __align__(16) float s_dyn[];   // kPmStages x [ (32, XS) | (32, GS) ]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mg = warp % MG, hg = warp / MG;
  const int gq = lane >> 2, t = lane & 3;
  const int pair = blockIdx.y;                   // selects, not a dynamic index: the struct stays in parameter space
  const float* __restrict__ g = pair == 0 ? ops.g[0] : (pair == 1 ? ops.g[1] : ops.g[2]);
  const float* __restrict__ x = pair == 0 ? ops.x[0] : (pair == 1 ? ops.x[1] : ops.x[2]);
  float acc[MT][NT][4];
  float bs[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bs[nt] = 0.f;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  }
  const int slabs = (N + kPmRows - 1) / kPmRows;
  // a thread's share of a slab: x rows tr, tr + 16 x column quads tc, tc + 16, .. (all offsets are constants from here on),
  // and one quad of g for the first 32 (OUT / 4) threads
  static_assert(kPmThreads == 256 && kPmRows == 32 && (IN / 4) % 16 == 0 && kPmRows * (OUT / 4) <= kPmThreads, "slab load mapping");
  const int tr = tid >> 4, tc = tid & 15;
  const int gr = tid / (OUT / 4), gj = tid - gr * (OUT / 4);
  auto load_slab = [&](int slab, int slot) {
    float* xs = s_dyn + slot * SLAB;
    float* gs = xs + kPmRows * XS;
    const int n0 = slab * kPmRows, rows = min(kPmRows, N - n0);
    const float* xsrc = x + (size_t)n0 * IN;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int r = tr + 16 * a;
#pragma unroll
      for (int q = 0; q < IN / 64; ++q) {
        const int c4 = tc + 16 * q;
        if (r < rows) cp_async16_cg(xs + r * XS + 4 * c4, xsrc + (size_t)r * IN + 4 * c4);
        else *reinterpret_cast<float4*>(xs + r * XS + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);   // a short last slab adds zeros
      }
    }
    if (gr < kPmRows) {
      if (gr < rows) cp_async16_cg(gs + gr * GS + 4 * gj, g + (size_t)(n0 + gr) * OUT + 4 * gj);
      else *reinterpret_cast<float4*>(gs + gr * GS + 4 * gj) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
#pragma unroll
  for (int pre = 0; pre < kPmStages - 1; ++pre) {      // one commit group per slab, empty ones included: the waits count groups
    const int slab = blockIdx.x + pre * (int)gridDim.x;
    if (slab < slabs) load_slab(slab, pre);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int it = 0;
#pragma unroll 1
  for (int slab = blockIdx.x; slab < slabs; slab += gridDim.x, ++it) {
    asm volatile("cp.async.wait_group %0;" :: "n"(kPmStages - 2) : "memory");
    __syncthreads();                                   // this slab has landed for every thread; the previous one's slot is free
    const int next = slab + (kPmStages - 1) * (int)gridDim.x;
    if (next < slabs) load_slab(next, (it + kPmStages - 1) % kPmStages);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* xs = s_dyn + (it % kPmStages) * SLAB + 16 * MT * mg + gq;
    const float* gs = s_dyn + (it % kPmStages) * SLAB + kPmRows * XS + gq;
#pragma unroll
    for (int sh = 0; sh < 4 / kPmHalves; ++sh) {
      const int ra = 4 * ((4 / kPmHalves) * hg + sh) + t, rb = ra + 16;
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float b0 = gs[ra * GS + 8 * nt], b1 = gs[rb * GS + 8 * nt];
        split_tf32(b0, bh[nt][0], bl[nt][0]);
        split_tf32(b1, bh[nt][1], bl[nt][1]);
        if (BIAS) bs[nt] += b0 + b1;
      }
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float a[4] = {xs[ra * XS + 16 * mt], xs[ra * XS + 16 * mt + 8], xs[rb * XS + 16 * mt], xs[rb * XS + 16 * mt + 8]};
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(a[e], ah[mt][e], al[mt][e]);
      }
      // the small terms first; product kind outermost, so that consecutive mma write different accumulators
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (BIAS) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {                          // the four k-slot lanes of a column: fixed order
      bs[nt] += __shfl_xor_sync(0xffffffffu, bs[nt], 1);
      bs[nt] += __shfl_xor_sync(0xffffffffu, bs[nt], 2);
    }
  }
  __syncthreads();                                             // the ring is free: the second k half parks its sums in it
  float* s_acc = s_dyn + mg * (MT * NT * 4 + NT) * 32;
  if (hg == 1) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s_acc[((mt * NT + nt) * 4 + e) * 32 + lane] = acc[mt][nt][e];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s_acc[(MT * NT * 4 + nt) * 32 + lane] = bs[nt];
  }
  __syncthreads();
  if (hg == 0) {
    float* dstp = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (OUT + 1) * IN;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 16 * MT * mg + 16 * mt + gq, j = 8 * nt + 2 * t;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = acc[mt][nt][e] + s_acc[((mt * NT + nt) * 4 + e) * 32 + lane];
        dstp[(size_t)j * IN + c] = v[0];
        dstp[(size_t)(j + 1) * IN + c] = v[1];
        dstp[(size_t)j * IN + c + 8] = v[2];
        dstp[(size_t)(j + 1) * IN + c + 8] = v[3];
      }
    if (BIAS && mg == 0 && t == 0) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dstp[(size_t)OUT * IN + 8 * nt + gq] = bs[nt] + s_acc[(MT * NT * 4 + nt) * 32 + lane];
    }
  }
}

// Stage 2: fixed-order sum over CTAs.  A CTA of 512 threads = 32 entries of (OUT + 1, IN) x 16 parts; part p sums the
// partials p, p + 16, ... in order (eight loads in flight ahead of the adds), then the sixteen sums are added in order.
// Entries [OUT][OUT..IN) are unused.
constexpr int kOlParts = 16;
struct PgResults { float* dw[3]; };     // blockIdx.y selects the operand pair; db belongs to pair 0
__global__ void __launch_bounds__(32 * kOlParts) out_linear_reduce_kernel(const float* __restrict__ partial, int ctas, int OUT,
                                                                          int IN, PgResults res, float* __restrict__ db,
                                                                          bool transposed = false) {
  __shared__ float red[kOlParts][32];
  const int e = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  const int entries = (OUT + 1) * IN;
  float s = 0.f;
  float* __restrict__ dw = blockIdx.y == 0 ? res.dw[0] : (blockIdx.y == 1 ? res.dw[1] : res.dw[2]);
  if (i < entries) {
    const float* src = partial + (size_t)blockIdx.y * ctas * entries + i;
    int b = part;
    for (; b + 7 * kOlParts < ctas; b += 8 * kOlParts) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(b + u * kOlParts) * entries];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; b < ctas; b += kOlParts) s += src[(size_t)b * entries];
  }
  red[part][e] = s;
  __syncthreads();
  if (part == 0 && i < entries) {
    float t = red[0][e];
#pragma unroll
    for (int p = 1; p < kOlParts; ++p) t += red[p][e];
    const int j = i / IN, c = i - j * IN;
    if (j < OUT) dw[transposed ? c * OUT + j : i] = t;     // transposed: dw is (IN, OUT)
    else if (c < OUT && db && blockIdx.y == 0) db[c] = t;
  }
}

// HEPT_PG_STAGED=1 selects the staged CUDA-core kernel (A/B against the tensor-path one); read once.
static bool pg_staged() {
  static const bool v = [] { const char* e = getenv("HEPT_PG_STAGED"); return e && e[0] == '1'; }();
  return v;
}
// CTAs per operand pair: two resident CTAs per SM over all pairs
static int pg_ctas(int sms, int slabs, int pairs) {
  int ctas = pg_staged() ? 2 * sms : 2 * sms / pairs;
  if (ctas > slabs) ctas = slabs;
  return ctas > kOlMaxCtas / pairs ? kOlMaxCtas / pairs : ctas;     // the workspace holds kOlMaxCtas partials
}
// partial sums of `pairs` products g[m]^T x[m] -> partial (pairs, ctas, (OUT + 1) IN), then their fixed-order sums -> res
template <int OUT, int TIN>
static int launch_pg(const PgOperands& ops, int pairs, int N, float* partial, int ctas, const PgResults& res, float* db,
                     bool transposed, cudaStream_t st, const char* what) {
  const int entries = (OUT + 1) * TIN;
  if (pg_staged()) {
    constexpr int THREADS = kPgGroups * (OUT / 8) * (TIN / 8);
    const size_t tsmem = sizeof(float) * kPgStages * (size_t)kPgRows * (TIN + 4 + OUT + 4);
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(out_linear_bwd_params_tiled_kernel<OUT, TIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
      HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
      configured.mark();
    }
    for (int m = 0; m < pairs; ++m)
      out_linear_bwd_params_tiled_kernel<OUT, TIN><<<ctas, THREADS, tsmem, st>>>(ops.g[m], ops.x[m], N, partial + (size_t)m * ctas * entries);
  } else {
    const size_t msmem = params_mma_smem_bytes<OUT, TIN>();
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(out_linear_bwd_params_mma_kernel<OUT, TIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(out_linear_bwd_params_mma_kernel<OUT, TIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
      HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
      configured.mark();
    }
    if (db) out_linear_bwd_params_mma_kernel<OUT, TIN, true><<<dim3(ctas, pairs), kPmThreads, msmem, st>>>(ops, N, partial);
    else out_linear_bwd_params_mma_kernel<OUT, TIN, false><<<dim3(ctas, pairs), kPmThreads, msmem, st>>>(ops, N, partial);
  }
  cudaError_t e = cudaGetLastError();
  HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "%s (partial sums): %s", what, cudaGetErrorString(e));
  out_linear_reduce_kernel<<<dim3((entries + 31) / 32, pairs), 32 * kOlParts, 0, st>>>(partial, ctas, OUT, TIN, res, db, transposed);
  e = cudaGetLastError();
  HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "%s (reduce): %s", what, cudaGetErrorString(e));
  return HEPT_OK;
}

template <int OUT, int OUTS, int HT>
static int launch_ol_fwd(const hept_shape* s, const float* x, const float* w, const float* b, float* out, cudaStream_t st) {
  const int IN = s->H * s->D;
  const size_t smem = sizeof(float) * (size_t)OUT * IN;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(out_linear_fwd_kernel<OUT, OUTS, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_fwd: %s", cudaGetErrorString(e));
    configured.mark();
  }
  HEPT_REQUIRE(smem <= 96 * 1024 && IN % 4 == 0, HEPT_EUNSUPPORTED, "out_linear_fwd: H*D=%d not supported", IN);
  if constexpr (OUT == 24) if (IN == OUT * 8) {    // the shipped H = 8, D = 24 shape: staged, register-tiled
    constexpr int TIN = OUT * 8;
    const size_t tsmem = sizeof(float) * ((size_t)OUT * (TIN + 4) + (size_t)kTlHits * (48 + 4));
    static DeviceOnce tconfigured;
    if (tconfigured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(out_linear_fwd_tiled_kernel<OUT, TIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
      HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_fwd: %s", cudaGetErrorString(e));
      tconfigured.mark();
    }
    out_linear_fwd_tiled_kernel<OUT, TIN><<<(s->N + kTlHits - 1) / kTlHits, kTlThreads, tsmem, st>>>(x, w, b, s->N, out);
    HEPT_CHECK_LAUNCH("out_linear_fwd");
    return HEPT_OK;
  }
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
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(out_linear_bwd_input_kernel<OUT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_bwd: %s", cudaGetErrorString(e));
      configured.mark();
    }
    HEPT_REQUIRE(smem <= 96 * 1024, HEPT_EUNSUPPORTED, "out_linear_bwd: H*D=%d too wide", IN);
    bool tiled = false;
    if constexpr (OUT == 24) if (IN == OUT * 8) {
      tiled = true;
      constexpr int TIN = OUT * 8;
      const size_t tsmem = sizeof(float) * ((size_t)OUT * TIN + (size_t)kTlHits * (OUT + 4));
      static DeviceOnce tconfigured;
      if (tconfigured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(out_linear_bwd_input_tiled_kernel<OUT, TIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
        HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_bwd: %s", cudaGetErrorString(e));
        tconfigured.mark();
      }
      out_linear_bwd_input_tiled_kernel<OUT, TIN><<<(s->N + kTlHits - 1) / kTlHits, kTlThreads, tsmem, st>>>(g, w, s->N, dx);
    }
    if (!tiled) {
      const int units = (s->N + 63) / 64 * SPLIT;
      const unsigned grid = (unsigned)((units + kOlWarps - 1) / kOlWarps);
      out_linear_bwd_input_kernel<OUT, SPLIT><<<grid, 32 * kOlWarps, smem, st>>>(g, w, s->N, IN, dx);
    }
    HEPT_CHECK_LAUNCH("out_linear_bwd_input");
  }
  const int slabs = (s->N + kOlRows - 1) / kOlRows;
  int ctas = sms * 4;
  if (ctas > slabs) ctas = slabs;
  if (ctas > kOlMaxCtas) ctas = kOlMaxCtas;
  if constexpr (OUT == 24) if (IN == OUT * 8) {
    constexpr int TIN = OUT * 8;
    PgOperands ops{};
    PgResults res{};
    ops.g[0] = g; ops.x[0] = x; res.dw[0] = dw;
    return launch_pg<OUT, TIN>(ops, 1, s->N, partial, pg_ctas(sms, (s->N + kPgRows - 1) / kPgRows, 1), res, db, false, st, "out_linear_bwd");
  }
  const size_t psmem = sizeof(float) * ((size_t)kOlRows * OUT + (size_t)OUT * IN);
  out_linear_bwd_params_kernel<OUT><<<ctas, kOlThreads, psmem, st>>>(g, x, s->N, IN, partial);
  HEPT_CHECK_LAUNCH("out_linear_bwd_params");
  const int entries = (OUT + 1) * IN;
  PgResults res{};
  res.dw[0] = dw;
  out_linear_reduce_kernel<<<(entries + 31) / 32, 32 * kOlParts, 0, st>>>(partial, ctas, OUT, IN, res, db);
  HEPT_CHECK_LAUNCH("out_linear_reduce");
  return HEPT_OK;
}

// The three weight gradients of the attention block's projections (attn_block.cu): dW_m (H*D, D) = dq_m^T xn is the
// parameter-gradient product above with the operands' roles exchanged (g := xn (N, D), x := dq_m (N, H*D)), written transposed.
size_t qkv_weight_grads_partial_floats(int H, int D) { return (size_t)kOlMaxCtas * (D + 1) * H * D; }

int qkv_weight_grads(const float* xn, const float* dq, const float* dk, const float* dv, int N, int H, int D, float* dwq,
                     float* dwk, float* dwv, float* partial, size_t partial_floats, cudaStream_t st) {
  HEPT_REQUIRE(D == 24 && H == 8, HEPT_EUNSUPPORTED, "qkv_weight_grads: (H=%d, D=%d) not compiled in", H, D);
  HEPT_REQUIRE(partial_floats >= qkv_weight_grads_partial_floats(H, D), HEPT_EWORKSPACE, "qkv_weight_grads: workspace too small");
  constexpr int OUT = 24, TIN = 192;
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "qkv_weight_grads: cannot read the SM count");
  PgOperands ops{};
  PgResults res{};
  const float* src[3] = {dq, dk, dv};
  float* dst[3] = {dwq, dwk, dwv};
  for (int m = 0; m < 3; ++m) { ops.g[m] = xn; ops.x[m] = src[m]; res.dw[m] = dst[m]; }
  return launch_pg<OUT, TIN>(ops, 3, N, partial, pg_ctas(sms, (N + kPgRows - 1) / kPgRows, 3), res, nullptr, true, st, "qkv_weight_grads");
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
