// a12, second half: the output projection out = out_pre W^T + b (example/hept.py:80, nn.Linear(H*D, D)) and its
// backward.  The library GEMMs this replaces were 12 % of a tracking-60k step: the 24 x 60000 x 192 weight-gradient
// product alone took 132 us in cuBLAS (one long reduction, poorly split), the two 60000 x 192 x 24 products 29 us each.
// All three are HBM-sized problems (46 MB of out_pre / d_out_pre, 0.28 GFMA).  For the shipped (H, D) = (8, 24) shape they run
// on the legacy tensor path (mma.sync m16n8k8, 3xTF32, fp32 accumulation -- mma_tf32.cuh):
//   out_linear_fwd        the narrow row kernel of attn_block.cu (K = 192 -> 24 outputs, + bias)          17 us
//   out_linear_bwd_input  the wide row kernel of attn_block.cu (K = 24 -> 192 outputs)                    16 us
//   out_linear_bwd_params dW = g^T out_pre, db = sum_n g: per-CTA partials over slabs of hits (below),    16 + 6 us
//                         then a fixed-order reduction over CTAs (deterministic, no floating-point atomics)
// (60k hits, ncu launch list.)  Other shapes take the CUDA-core kernels below: a hit per lane, weights broadcast from shared
// memory (44 / 41 us at 60k hits; bound by the L1 tag stage -- a lane-per-hit LDG.128 / STG.128 touches 32 lines).
#include "common.cuh"
#include "mma_tf32.cuh"

namespace hept {

// the tensor-path row kernels of attn_block.cu, for the shipped (H, D) = (8, 24) shape
int out_linear_fwd_rows(const float* out_pre, const float* w, const float* b, int N, float* out, cudaStream_t st);
int out_linear_bwd_input_rows(const float* g, const float* w, int N, float* dx, cudaStream_t st);

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

// weight / bias gradient, stage 1, on the legacy tensor path (mma_tf32.cuh: mma.sync m16n8k8, 3xTF32 with fp32 accumulation).
// The CUDA-core version of this product (8 x 8 register tiles over staged slabs) was bound by its shared-memory wavefronts
// (70 % of the LSU data pipe, profiles/r2a_ncu_prof_front_cudacore.csv: 33.6 us); an mma needs an eighth of the operand loads per
// product (15.5 us).  P[c][j] = sum_n x[n][c] g[n][j]: M = the IN columns of x, N = the
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

// CTAs per operand pair: two resident CTAs per SM over all pairs
static int pg_ctas(int sms, int slabs, int pairs) {
  int ctas = 2 * sms / pairs;
  if (ctas > slabs) ctas = slabs;
  return ctas > kOlMaxCtas / pairs ? kOlMaxCtas / pairs : ctas;     // the workspace holds kOlMaxCtas partials
}
// partial sums of `pairs` products g[m]^T x[m] -> partial (pairs, ctas, (OUT + 1) IN), then their fixed-order sums -> res
template <int OUT, int TIN>
static int launch_pg(const PgOperands& ops, int pairs, int N, float* partial, int ctas, const PgResults& res, float* db,
                     bool transposed, cudaStream_t st, const char* what) {
  const int entries = (OUT + 1) * TIN;
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
  if constexpr (OUT == 24) if (IN == OUT * 8)      // the shipped H = 8, D = 24 shape: the tensor-path row kernels (attn_block.cu)
    return out_linear_fwd_rows(x, w, b, s->N, out, st);
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
    bool rows = false;
    if constexpr (OUT == 24) if (IN == OUT * 8) {
      rows = true;
      if (int rc = out_linear_bwd_input_rows(g, w, s->N, dx, st)) return rc;
    }
    if (!rows) {
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
    return launch_pg<OUT, TIN>(ops, 1, s->N, partial, pg_ctas(sms, (s->N + kPmRows - 1) / kPmRows, 1), res, db, false, st, "out_linear_bwd");
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
  return launch_pg<OUT, TIN>(ops, 3, N, partial, pg_ctas(sms, (N + kPmRows - 1) / kPmRows, 3), res, nullptr, true, st, "qkv_weight_grads");
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
  HEPT_REQUIRE(aligned16({d_out, weight, out_pre, d_out_pre, d_weight, workspace}), HEPT_EINVAL,
               "out_linear_bwd: array pointers must be 16-byte aligned");
  HEPT_REQUIRE(workspace_bytes >= hept_out_linear_bwd_workspace_bytes(s), HEPT_EWORKSPACE,
               "out_linear_bwd: workspace needs %zu bytes", hept_out_linear_bwd_workspace_bytes(s));
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (s->D == 24) return launch_ol_bwd<24, 4>(s, d_out, weight, out_pre, d_out_pre, d_weight, d_bias, partial, st);
  if (s->D == 8) return launch_ol_bwd<8, 1>(s, d_out, weight, out_pre, d_out_pre, d_weight, d_bias, partial, st);
  set_error("out_linear_bwd: D=%d not compiled in", s->D);
  return HEPT_EUNSUPPORTED;
}
