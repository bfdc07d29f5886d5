// The caller's side of the boundary: the front of the reference's Attn block,
//     x_normed = norm1(x);  q, k, v = w_q(x_normed), w_k(x_normed), w_v(x_normed)        example/transformer.py:157-158
// (src/models/baselines/transformer.py:209-212), forward and backward, so that a hit enters the library as its 96-byte
// activation row instead of three 768-byte q / k / v rows: with host-resident inputs the call is no longer a PCIe benchmark
// (2 520 -> 312 bytes per hit over the link), and on the device the three projections stop being library GEMM launches.
// q, k, v are still materialised in HBM (they are what the hash and tile kernels gather).
//
// The three products of a call (hits x 24 by 24 x 576, its transpose, and the 576 x hits by hits x 24 weight gradient) are
// 0.83 GFMA each against 138 MB of q / k / v traffic: on CUDA cores they sit on the FMA pipe's floor (23 us at 128 FMA per
// clock and SM; FFMA2 occupies the pipe for two cycles, it only saves issue slots) and, in practice, on shared-memory
// wavefronts -- the first versions measured 54 / 80 / 3 x 40 us.  They run on the legacy tensor path instead (mma.sync
// m16n8k8, 3xTF32, fp32 accumulation: 2.2 clocks per mma and SM measured, tools/micro/mma_rate.cu, i.e. an 18 us floor per
// product next to the 21 us HBM floor): 40 / 43 / 42 us.  A tcgen05 pipeline would lift the arithmetic floor, not the HBM one.
//
//   qkv_weights_t        Wt (3, DM, OW) = the three (OW, DM) weights transposed once per call, saved for the backward
//   ln_qkv_fwd           LayerNorm + the three projections: a warp per 16-hit tile, weights as pre-split fragments in shared memory
//   ln_qkv_bwd_input     dxn = dq Wq + dk Wk + dv Wv the same way, gradient rows read straight from global memory three
//                        32-column blocks ahead, then the LayerNorm backward of the row -> dx, and the CTA's partial sums of
//                        d gamma / d beta (fixed order: deterministic)
//   ln_params_reduce     fixed-order sum of those partials
//   weight gradients     dW_m = dq_m^T xn: the out_linear parameter-gradient kernel (out_linear.cu) with the roles of the
//                        operands exchanged, the three matrices in one launch, result written transposed
#include "common.cuh"
#include "mma_tf32.cuh"

namespace hept {

int qkv_weight_grads(const float* xn, const float* dq, const float* dk, const float* dv, int N, int H, int D, float* dwq,
                     float* dwk, float* dwv, float* partial, size_t partial_floats, cudaStream_t st);
size_t qkv_weight_grads_partial_floats(int H, int D);

// the tensor-path kernels: a warp per 16-hit tile, 16 warps per persistent CTA, gradient blocks in flight per warp
constexpr int kAbMmaWarps = 16, kAbMmaThreads = 32 * kAbMmaWarps, kAbMmaTile = 16, kAbMmaAhead = 3;

// Wt[m][j][c] = W_m[c][j]
__global__ void __launch_bounds__(256) qkv_weights_t_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                                            const float* __restrict__ wv, int DM, int OW, float* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = DM * OW;
  if (i >= 3 * per) return;
  const int m = i / per, r = i - m * per;
  const int j = r / OW, c = r - j * OW;
  const float* w = m == 0 ? wq : (m == 1 ? wk : wv);
  wt[i] = __ldg(w + (size_t)c * DM + j);
}

// LayerNorm of the rows (two-pass variance, eps inside the square root like torch.nn.functional.layer_norm), xn kept for the
// backward, then q / k / v (hits x OW) = xn (hits x DM) W_m^T on the legacy tensor path (mma_tf32.cuh).  One persistent CTA per
// SM; a warp owns 16-hit tiles as in the backward kernel below (same tile order):
//   A  a quad's lane t holds the six activations 6 t .. 6 t + 5 of its two hits (k-slot t of k-step s is column 6 t + 2 s,
//      k-slot t + 4 is 6 t + 2 s + 1); the LayerNorm sums go over the quad by shuffles; xn is split into tf32 hi / lo once.
//   B  the three weights, split ONCE per CTA and parked in shared memory in fragment order, s_b[m][n-tile][k-step][lane] =
//      {b0 hi, b0 lo, b1 hi, b1 lo}.  Fragment column j of the n-tile pair (p, q) is output column 16 p + 4 (j / 2) + 2 q + j % 2,
//      so that a thread's C fragments of a pair are four adjacent outputs: one 16-byte store per hit and pair.
// NMAT = 3, LN: the block's front.  NMAT = 1, no LayerNorm, WT (the weight is (DM, OW): element (column, k) at w[k OW + column]):
// out_linear's input gradient d out_pre = d out W (out_linear.cu).
template <int DM, int OW, int NMAT>
constexpr size_t rows_wide_smem_bytes() { return sizeof(uint4) * NMAT * (OW / 8) * (DM / 8) * 32; }

template <int DM, int OW, int NMAT, bool LN, bool WT>
__global__ void __launch_bounds__(kAbMmaThreads, 1) rows_wide_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, const float* __restrict__ wq,
                                                                       const float* __restrict__ wk, const float* __restrict__ wv,
                                                                       int N, float eps, float* __restrict__ xn_out,
                                                                       float* __restrict__ q, float* __restrict__ k, float* __restrict__ v) {
  constexpr int NT = OW / 8, KS = DM / 8, PERL = DM / 4;           // n-tiles per matrix, k-steps, activations per lane
  static_assert(DM % 8 == 0 && OW % 32 == 0 && PERL == 2 * KS, "tile shape");
  extern __shared__ __align__(16) uint4 s_b[];                     // (NMAT, NT, KS, 32)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  {
    constexpr int TOTAL = NMAT * NT * KS * 32, PER = (TOTAL + kAbMmaThreads - 1) / kAbMmaThreads;
    float2 w[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {                                // all of a thread's (L2-resident) loads in flight at once
      const int i = tid + u * kAbMmaThreads;
      const int l = i & 31, ks = (i >> 5) % KS, nt = (i / (32 * KS)) % NT, m = i / (32 * KS * NT);
      const int j = l >> 2, col = 16 * (nt >> 1) + 4 * (j >> 1) + 2 * (nt & 1) + (j & 1);
      const float* wm = m == 0 ? wq : (m == 1 ? wk : wv);
      const int k0 = PERL * (l & 3) + 2 * ks;
      if (i >= TOTAL) w[u] = make_float2(0.f, 0.f);
      else if (WT) w[u] = make_float2(__ldg(wm + (size_t)k0 * OW + col), __ldg(wm + (size_t)(k0 + 1) * OW + col));
      else w[u] = ldg2(wm + (size_t)col * DM + k0);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = tid + u * kAbMmaThreads;
      uint4 f;
      split_tf32(w[u].x, f.x, f.y);
      split_tf32(w[u].y, f.z, f.w);
      if (i < TOTAL) s_b[i] = f;
    }
  }
  __syncthreads();
  float gam[PERL], bet[PERL];
#pragma unroll
  for (int u = 0; u < PERL; ++u) { gam[u] = LN ? __ldg(gamma + PERL * t + u) : 1.f; bet[u] = LN ? __ldg(beta + PERL * t + u) : 0.f; }
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
#pragma unroll 1
  for (int tile = blockIdx.x + gridDim.x * warp; tile < tiles; tile += gridDim.x * kAbMmaWarps) {
    const int rows[2] = {tile * kAbMmaTile + gq, tile * kAbMmaTile + gq + 8};
    uint32_t ah[KS][4], al[KS][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                  // LayerNorm of the quad's hit: two-pass variance, eps inside the root
      const bool live = rows[h] < N;
      float xv[PERL];
#pragma unroll
      for (int u2 = 0; u2 < PERL / 2; ++u2) {
        const float2 tt = live ? ldg2(x + (size_t)rows[h] * DM + PERL * t + 2 * u2) : make_float2(0.f, 0.f);
        xv[2 * u2] = tt.x; xv[2 * u2 + 1] = tt.y;
      }
      if (LN) {
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < PERL; ++u) sum += xv[u];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / DM);
        float var = 0.f;
#pragma unroll
        for (int u = 0; u < PERL; ++u) { const float dlt = xv[u] - mean; var = fmaf(dlt, dlt, var); }
        var += __shfl_xor_sync(0xffffffffu, var, 1);
        var += __shfl_xor_sync(0xffffffffu, var, 2);
        const float rstd = 1.f / sqrtf(var * (1.f / DM) + eps);
#pragma unroll
        for (int u = 0; u < PERL; ++u) xv[u] = fmaf((xv[u] - mean) * rstd, gam[u], bet[u]);
        if (live) {
#pragma unroll
          for (int u2 = 0; u2 < PERL / 2; ++u2)
            *reinterpret_cast<float2*>(xn_out + (size_t)rows[h] * DM + PERL * t + 2 * u2) = make_float2(xv[2 * u2], xv[2 * u2 + 1]);
        }
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {                            // a0 / a1 = rows g / g + 8 at k-slot t, a2 / a3 at k-slot t + 4
        split_tf32(xv[2 * ks], ah[ks][h], al[ks][h]);
        split_tf32(xv[2 * ks + 1], ah[ks][2 + h], al[ks][2 + h]);
      }
    }
#pragma unroll 1
    for (int m = 0; m < NMAT; ++m) {
      float* __restrict__ out = m == 0 ? q : (m == 1 ? k : v);
      const uint4* bm = s_b + (size_t)m * NT * KS * 32 + lane;
#pragma unroll 2
      for (int n4 = 0; n4 < NT / 4; ++n4) {                        // four n-tiles (two pairs) at a time: twelve independent mma chains
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint4 b[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) b[c] = bm[((4 * n4 + c) * KS + ks) * 32];
#pragma unroll
          for (int c = 0; c < 4; ++c) mma_tf32(acc[c], al[ks], b[c].x, b[c].z);     // the small terms first
#pragma unroll
          for (int c = 0; c < 4; ++c) mma_tf32(acc[c], ah[ks], b[c].y, b[c].w);
#pragma unroll
          for (int c = 0; c < 4; ++c) mma_tf32(acc[c], ah[ks], b[c].x, b[c].z);
        }
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int col = 16 * (2 * n4 + pr) + 4 * t;
          if (rows[0] < N) *reinterpret_cast<float4*>(out + (size_t)rows[0] * OW + col) = make_float4(acc[2 * pr][0], acc[2 * pr][1], acc[2 * pr + 1][0], acc[2 * pr + 1][1]);
          if (rows[1] < N) *reinterpret_cast<float4*>(out + (size_t)rows[1] * OW + col) = make_float4(acc[2 * pr][2], acc[2 * pr][3], acc[2 * pr + 1][2], acc[2 * pr + 1][3]);
        }
      }
    }
  }
}

// dxn (hits x DM) = [dq | dk | dv] (hits x 3 OW) * [Wq; Wk; Wv] (3 OW x DM) on the legacy tensor path (mma_tf32.cuh), then the
// LayerNorm backward of the row and the CTA's partial sums of d gamma / d beta.
// One persistent CTA of 16 warps per SM; a warp owns 16-hit tiles (tile = blockIdx.x + gridDim.x (warp + 16 j): every SM gets
// the same number of tiles to within one) and needs no CTA barrier inside the loop (13 warps, which would fill the last round
// of a 60 000-hit event's 25.3 tiles per SM exactly, measured 5 % slower than 16: the tensor pipe wants the warps).
//   B  the three weights, split into tf32 hi / lo ONCE per CTA and parked in shared memory in fragment order:
//      s_b[k-step][n-tile][lane] = {b0 hi, b0 lo, b1 hi, b1 lo} -- one conflict-free 128-bit load per (k-step, n-tile)
//   A  the gradient rows, read straight from global memory: within a block of 32 columns lane t of a row's quad holds the
//      columns 4 t .. 4 t + 3 (k-slot t of k-steps 0 .. 3) and 16 + 4 t .. (k-slot t + 4) -- K is the contraction index, any
//      bijection will do, and this one makes every load a full 16-byte vector, a row's quad cover whole sectors.  Three
//      blocks (12 vectors per thread) are in flight ahead of the one being multiplied.
//   C  a thread ends with 6 of the 24 dxn of two hits (the C fragment); the four lanes of a quad share a hit exactly as in
//      the forward kernel; the LayerNorm backward runs on the fragment, the sums over the quad by shuffles.
// NMAT = 3, LN: the block's front.  NMAT = 1, no LayerNorm: out_linear's forward out = out_pre W^T + b (out_linear.cu; `wt` is
// the (DM, OW) weight itself, `gamma` its bias, `dx` the output).
template <int DM, int OW, int NMAT>
constexpr size_t rows_narrow_smem_bytes() { return sizeof(uint4) * (NMAT * OW / 8) * (DM / 8) * 32; }

template <int DM, int OW, int NMAT, bool LN>
__global__ void __launch_bounds__(kAbMmaThreads, 1) rows_narrow_kernel(const float* __restrict__ dq, const float* __restrict__ dk,
                                                                              const float* __restrict__ dv, const float* __restrict__ wt,
                                                                              const float* __restrict__ x, const float* __restrict__ gamma,
                                                                              int N, float eps, float* __restrict__ dx,
                                                                              float* __restrict__ partial) {
  constexpr int NT = DM / 8, KB = OW / 32, KS = NMAT * OW / 8;    // n-tiles, 32-column blocks per matrix, k-steps in all
  static_assert(DM % 8 == 0 && OW % 32 == 0, "tile shape");
  static_assert(KS * NT * 32 * 4 >= kAbMmaWarps * 2 * DM, "the fragment store doubles as the reduction buffer");
  extern __shared__ __align__(16) uint4 s_b[];                     // (KS, NT, 32)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  {
    constexpr int TOTAL = KS * NT * 32, PER = (TOTAL + kAbMmaThreads - 1) / kAbMmaThreads;
    float w0[PER], w1[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {                                // all of a thread's (L2-resident) loads in flight at once
      const int i = tid + u * kAbMmaThreads;
      const int l = i & 31, nt = (i >> 5) % NT, ks = i / (32 * NT);
      const int m = ks / (OW / 8), kk = ks - m * (OW / 8);
      const int k0 = 32 * (kk >> 2) + 4 * (l & 3) + (kk & 3), n = 8 * nt + (l >> 2);
      const float* wrow = wt + ((size_t)m * DM + n) * OW;          // Wt[m][n][k] = W_m[k][n]
      w0[u] = i < TOTAL ? __ldg(wrow + k0) : 0.f;
      w1[u] = i < TOTAL ? __ldg(wrow + k0 + 16) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = tid + u * kAbMmaThreads;
      uint4 v;
      split_tf32(w0[u], v.x, v.y);
      split_tf32(w1[u], v.z, v.w);
      if (i < TOTAL) s_b[i] = v;
    }
  }
  float gam[NT][2], dgam[NT][2], dbet[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) { gam[nt][e] = __ldg(gamma + 8 * nt + 2 * t + e); dgam[nt][e] = 0.f; dbet[nt][e] = 0.f; }
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
  // dead rows (beyond N, or beyond the warp's last tile) read row 0 and are dropped below
  auto row_offset = [&](int r) { return (size_t)(r < N ? r : 0) * OW + 4 * t; };
  float4 ring[kAbMmaAhead][4];      // [block in flight][row a slot t, row b slot t, row a slot t + 4, row b slot t + 4]
  auto load_block = [&](int blk, size_t oa, size_t ob, float4 (&dst)[4]) {
    const int m = blk / KB, kb = blk - m * KB;
    const float* src = m == 0 ? dq : (m == 1 ? dk : dv);
    dst[0] = ldg4(src + oa + 32 * kb);
    dst[1] = ldg4(src + ob + 32 * kb);
    dst[2] = ldg4(src + oa + 32 * kb + 16);
    dst[3] = ldg4(src + ob + 32 * kb + 16);
  };
  int tile = blockIdx.x + gridDim.x * warp;
  size_t oa = row_offset(tile * kAbMmaTile + gq), ob = row_offset(tile * kAbMmaTile + gq + 8);
  if (tile < tiles) {               // the first tile's first blocks fly while the fragment store settles
#pragma unroll
    for (int pre = 0; pre < kAbMmaAhead; ++pre) load_block(pre, oa, ob, ring[pre]);
  }
  __syncthreads();
#pragma unroll 1
  for (; tile < tiles; tile += gridDim.x * kAbMmaWarps) {
    const int ra = tile * kAbMmaTile + gq, rb = ra + 8;
    const bool la = ra < N, lb = rb < N;
    const int next = tile + gridDim.x * kAbMmaWarps;
    const size_t na = row_offset(next * kAbMmaTile + gq), nb = row_offset(next * kAbMmaTile + gq + 8);
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int blk = 0; blk < NMAT * KB; ++blk) {
      float4 cur[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) cur[e] = ring[blk % kAbMmaAhead][e];
      // the slot just read takes the block three ahead -- of this tile, or of the warp's next one (its first blocks are
      // on their way while this tile's LayerNorm backward runs)
      if (blk + kAbMmaAhead < NMAT * KB) load_block(blk + kAbMmaAhead, oa, ob, ring[blk % kAbMmaAhead]);
      else if (next < tiles) load_block(blk + kAbMmaAhead - NMAT * KB, na, nb, ring[blk % kAbMmaAhead]);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = s == 0 ? cur[e].x : (s == 1 ? cur[e].y : (s == 2 ? cur[e].z : cur[e].w));
          split_tf32(v, ah[e], al[e]);
        }
        uint4 b[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = s_b[((blk * 4 + s) * NT + nt) * 32 + lane];
        // the small terms first; product kind outermost, so that consecutive mma write different accumulators
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], al, b[nt].x, b[nt].z);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, b[nt].y, b[nt].w);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, b[nt].x, b[nt].z);
      }
    }
    oa = na;
    ob = nb;
    if constexpr (!LN) {                                             // out = acc + bias
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h == 0 ? ra : rb;
        if (h == 0 ? la : lb) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            *reinterpret_cast<float2*>(dx + (size_t)r * DM + 8 * nt + 2 * t) = make_float2(acc[nt][2 * h] + gam[nt][0], acc[nt][2 * h + 1] + gam[nt][1]);
        }
      }
    } else {
      // LayerNorm backward of the two hits: y = xhat gamma + beta, xhat = (x - mean) rstd
      //   dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dxn gamma;  d gamma += dxn xhat;  d beta += dxn
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h == 0 ? ra : rb;
        const bool live = h == 0 ? la : lb;
        float xv[NT][2], dxn[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float2 tt = live ? ldg2(x + (size_t)r * DM + 8 * nt + 2 * t) : make_float2(0.f, 0.f);
          xv[nt][0] = tt.x; xv[nt][1] = tt.y;
          dxn[nt][0] = acc[nt][2 * h]; dxn[nt][1] = acc[nt][2 * h + 1];
        }
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) sum += xv[nt][0] + xv[nt][1];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / DM);
        float vs = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) { const float dlt = xv[nt][e] - mean; vs = fmaf(dlt, dlt, vs); }
        vs += __shfl_xor_sync(0xffffffffu, vs, 1);
        vs += __shfl_xor_sync(0xffffffffu, vs, 2);
        const float rstd = 1.f / sqrtf(vs * (1.f / DM) + eps);
        float xh[NT][2], gg[NT][2], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            xh[nt][e] = (xv[nt][e] - mean) * rstd;
            gg[nt][e] = dxn[nt][e] * gam[nt][e];
            m1 += gg[nt][e];
            m2 = fmaf(gg[nt][e], xh[nt][e], m2);
          }
        m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
        m1 += __shfl_xor_sync(0xffffffffu, m1, 2);
        m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
        m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
        m1 *= 1.f / DM;
        m2 *= 1.f / DM;
        if (live) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              o[e] = rstd * (gg[nt][e] - m1 - xh[nt][e] * m2);
              dgam[nt][e] = fmaf(dxn[nt][e], xh[nt][e], dgam[nt][e]);
              dbet[nt][e] += dxn[nt][e];
            }
            *reinterpret_cast<float2*>(dx + (size_t)r * DM + 8 * nt + 2 * t) = make_float2(o[0], o[1]);
          }
        }
      }
    }
  }
  if constexpr (LN) {
    // the CTA's sums: the eight row lanes of a column in butterfly order, then the warps in order (fixed: deterministic)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int sh = 4; sh < 32; sh <<= 1) {
          dgam[nt][e] += __shfl_xor_sync(0xffffffffu, dgam[nt][e], sh);
          dbet[nt][e] += __shfl_xor_sync(0xffffffffu, dbet[nt][e], sh);
        }
    __syncthreads();                                                 // every warp is done with the fragment store
    float* s_red = reinterpret_cast<float*>(s_b);                    // (warps, 2 DM)
    if (gq == 0) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s_red[warp * 2 * DM + 8 * nt + 2 * t + e] = dgam[nt][e];
          s_red[warp * 2 * DM + DM + 8 * nt + 2 * t + e] = dbet[nt][e];
        }
    }
    __syncthreads();
    if (tid < 2 * DM) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kAbMmaWarps; ++w) s += s_red[w * 2 * DM + tid];
      partial[(size_t)blockIdx.x * 2 * DM + tid] = s;
    }
  }
}

// d gamma / d beta = fixed-order sum over the CTAs' partials (ctas, 2 DM): one CTA, tree over 256 strided sums
__global__ void __launch_bounds__(256) ln_params_reduce_kernel(const float* __restrict__ partial, int ctas, int DM,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[256];
  const int e = blockIdx.x;                  // entry in [0, 2 DM)
  float s = 0.f;
  for (int b = threadIdx.x; b < ctas; b += 256) s += partial[(size_t)b * 2 * DM + e];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (e < DM) dgamma[e] = red[0];
    else dbeta[e - DM] = red[0];
  }
}

// (also used by layer_norm.cu)
void ln_params_reduce_launch(const float* partial, int ctas, int DM, float* dgamma, float* dbeta, cudaStream_t st) {
  ln_params_reduce_kernel<<<2 * DM, 256, 0, st>>>(partial, ctas, DM, dgamma, dbeta);
}

template <int DM, int OW>
static int launch_qkv_fwd(const float* x, const float* gamma, const float* beta, const float* wq, const float* wk, const float* wv,
                          int N, float eps, float* wt, float* xn, float* q, float* k, float* v, cudaStream_t st) {
  qkv_weights_t_kernel<<<(3 * DM * OW + 255) / 256, 256, 0, st>>>(wq, wk, wv, DM, OW, wt);
  HEPT_CHECK_LAUNCH("qkv_weights_t");
  const size_t smem = rows_wide_smem_bytes<DM, OW, 3>();
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rows_wide_kernel<DM, OW, 3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "attn_qkv_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "attn_qkv_fwd: cannot read the SM count");
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
  rows_wide_kernel<DM, OW, 3, true, false><<<sms < tiles ? sms : tiles, kAbMmaThreads, smem, st>>>(x, gamma, beta, wq, wk, wv, N, eps, xn, q, k, v);
  HEPT_CHECK_LAUNCH("ln_qkv_fwd");
  return HEPT_OK;
}

// rows of the (ctas, 2 DM) buffer of LayerNorm parameter partials: one per persistent CTA, i.e. per SM
constexpr int kAbMaxCtas = 1024;

template <int DM, int OW>
static int launch_qkv_bwd(const float* x, const float* xn, const float* gamma, const float* wt, const float* dq, const float* dk,
                          const float* dv, int N, int H, int D, float eps, float* dx, float* dgamma, float* dbeta, float* dwq,
                          float* dwk, float* dwv, float* ws, size_t ws_floats, cudaStream_t st) {
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "attn_qkv_bwd: cannot read the SM count");
  HEPT_REQUIRE(sms <= kAbMaxCtas, HEPT_EUNSUPPORTED, "attn_qkv_bwd: %d SMs, at most %d supported", sms, kAbMaxCtas);
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
  const int ctas = sms < tiles ? sms : tiles;
  const size_t ln_floats = (size_t)kAbMaxCtas * 2 * DM;
  HEPT_REQUIRE(ws_floats >= ln_floats + qkv_weight_grads_partial_floats(H, D), HEPT_EWORKSPACE, "attn_qkv_bwd: workspace too small");
  const size_t smem = rows_narrow_smem_bytes<DM, OW, 3>();
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rows_narrow_kernel<DM, OW, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "attn_qkv_bwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured.mark();
  }
  rows_narrow_kernel<DM, OW, 3, true><<<ctas, kAbMmaThreads, smem, st>>>(dq, dk, dv, wt, x, gamma, N, eps, dx, ws);
  HEPT_CHECK_LAUNCH("ln_qkv_bwd_input");
  ln_params_reduce_kernel<<<2 * DM, 256, 0, st>>>(ws, ctas, DM, dgamma, dbeta);
  HEPT_CHECK_LAUNCH("ln_params_reduce");
  return qkv_weight_grads(xn, dq, dk, dv, N, H, D, dwq, dwk, dwv, ws + ln_floats, ws_floats - ln_floats, st);
}

// out_linear (out_linear.cu) on the same two kernels, for the shipped H = 8, D = 24 shape: out = out_pre W^T + b and
// d out_pre = d out W, W (24, 192)
int out_linear_fwd_rows(const float* out_pre, const float* w, const float* b, int N, float* out, cudaStream_t st) {
  constexpr int DM = 24, OW = 192;
  const size_t smem = rows_narrow_smem_bytes<DM, OW, 1>();
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rows_narrow_kernel<DM, OW, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "out_linear_fwd: cannot read the SM count");
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
  rows_narrow_kernel<DM, OW, 1, false><<<sms < tiles ? sms : tiles, kAbMmaThreads, smem, st>>>(out_pre, nullptr, nullptr, w, nullptr, b, N, 0.f,
                                                                                             out, nullptr);
  HEPT_CHECK_LAUNCH("out_linear_fwd");
  return HEPT_OK;
}

int out_linear_bwd_input_rows(const float* g, const float* w, int N, float* dx, cudaStream_t st) {
  constexpr int DM = 24, OW = 192;
  const size_t smem = rows_wide_smem_bytes<DM, OW, 1>();
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rows_wide_kernel<DM, OW, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "out_linear_bwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured.mark();
  }
  const int sms = sm_count();
  HEPT_REQUIRE(sms > 0, HEPT_ECUDA, "out_linear_bwd: cannot read the SM count");
  const int tiles = (N + kAbMmaTile - 1) / kAbMmaTile;
  rows_wide_kernel<DM, OW, 1, false, true><<<sms < tiles ? sms : tiles, kAbMmaThreads, smem, st>>>(g, nullptr, nullptr, w, nullptr, nullptr, N, 0.f,
                                                                                                 nullptr, dx, nullptr, nullptr);
  HEPT_CHECK_LAUNCH("out_linear_bwd_input");
  return HEPT_OK;
}

}  // namespace hept

using namespace hept;

extern "C" int hept_attn_qkv_supported(int32_t H, int32_t D) { return H == 8 && D == 24; }

extern "C" size_t hept_attn_qkv_bwd_workspace_bytes(int32_t N, int32_t H, int32_t D) {
  if (N <= 0 || H <= 0 || D <= 0) return 0;
  return sizeof(float) * ((size_t)kAbMaxCtas * 2 * D + qkv_weight_grads_partial_floats(H, D));
}

extern "C" int hept_attn_qkv_fwd(const float* x, const float* norm_weight, const float* norm_bias, const float* w_q,
                                 const float* w_k, const float* w_v, int32_t N, int32_t H, int32_t D, float eps, float* wt,
                                 float* x_normed, float* q, float* k, float* v, void* stream) {
  HEPT_REQUIRE(x && norm_weight && norm_bias && w_q && w_k && w_v && wt && x_normed && q && k && v && N > 0, HEPT_EINVAL,
               "attn_qkv_fwd: bad argument");
  HEPT_REQUIRE(hept_attn_qkv_supported(H, D), HEPT_EUNSUPPORTED, "attn_qkv_fwd: (H=%d, D=%d) not compiled in", H, D);
  HEPT_REQUIRE(aligned16({x, w_q, w_k, w_v, wt, x_normed, q, k, v}), HEPT_EINVAL, "attn_qkv_fwd: array pointers must be 16-byte aligned");
  return launch_qkv_fwd<24, 192>(x, norm_weight, norm_bias, w_q, w_k, w_v, N, eps, wt, x_normed, q, k, v, (cudaStream_t)stream);
}

extern "C" int hept_attn_qkv_bwd(const float* x, const float* x_normed, const float* norm_weight, const float* wt,
                                 const float* dq, const float* dk, const float* dv, int32_t N, int32_t H, int32_t D, float eps,
                                 float* dx, float* d_norm_weight, float* d_norm_bias, float* d_w_q, float* d_w_k, float* d_w_v,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  HEPT_REQUIRE(x && x_normed && norm_weight && wt && dq && dk && dv && dx && d_norm_weight && d_norm_bias && d_w_q && d_w_k &&
                   d_w_v && workspace && N > 0,
               HEPT_EINVAL, "attn_qkv_bwd: bad argument");
  HEPT_REQUIRE(hept_attn_qkv_supported(H, D), HEPT_EUNSUPPORTED, "attn_qkv_bwd: (H=%d, D=%d) not compiled in", H, D);
  HEPT_REQUIRE(aligned16({x, x_normed, wt, dq, dk, dv, dx, d_w_q, d_w_k, d_w_v, workspace}), HEPT_EINVAL,
               "attn_qkv_bwd: array pointers must be 16-byte aligned");
  HEPT_REQUIRE(workspace_bytes >= hept_attn_qkv_bwd_workspace_bytes(N, H, D), HEPT_EWORKSPACE, "attn_qkv_bwd: workspace needs %zu bytes",
               hept_attn_qkv_bwd_workspace_bytes(N, H, D));
  return launch_qkv_bwd<24, 192>(x, x_normed, norm_weight, wt, dq, dk, dv, N, H, D, eps, dx, d_norm_weight, d_norm_bias, d_w_q,
                                 d_w_k, d_w_v, (float*)workspace, workspace_bytes / sizeof(float), (cudaStream_t)stream);
}
