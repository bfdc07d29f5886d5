// a8-a11 fused: gather the sorted rows of each block (sort_to_buckets, example/hept_utils.py:74-92),
// evaluate the block-local RBF-kernel attention (qkv_res, example/hept.py:7-18) and write numerator
// and normaliser straight back to ORIGINAL hit order (invert_permutation + unsort_from_buckets,
// example/hept_utils.py:50-61,95-97; example/hept.py:76-78) — the (T,H,nb,B,B) score tensor the
// reference materialises never exists.  a12 (OR-combine over the T tables, example/hept.py:79) is the
// small streaming kernel at the bottom.
#include "tile.cuh"

namespace hept {

template <int D, int C, int B, int G, int R, int MINB>
__global__ void __launch_bounds__((TileLayout<D, C, B, G, R>::THREADS), MINB)
    block_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                          const float* __restrict__ coords, const float* __restrict__ scale,
                          const int32_t* __restrict__ positions, int N, int H, int T, int raw_size,
                          float* __restrict__ stage) {
  using L = TileLayout<D, C, B, G, R>;
  constexpr int E = L::E;
  extern __shared__ float4 smem[];
  float4* ks = smem;                             // [G*B][8]   k' rows: k'[0..E), nk
  float4* vs = smem + G * B * L::ROW_CHUNKS;     // [G*B][D/4] value rows

  const int h = blockIdx.y / T, t = blockIdx.y % T, th = t * H + h;  // tables of one head run back to back: its q/k/v slices stay in L2
  const int nb = N / B;
  const int blk0 = blockIdx.x * G;
  const int32_t* qpos = positions + (size_t)th * N;
  const int32_t* kpos = positions + ((size_t)T * H + th) * N;
  const float* scale_h = scale + h * C;
  const int tid = threadIdx.x;

  gather_streamed_rows<L, false>(k, k, v, nullptr, nullptr, coords, scale_h, kpos, kpos, blk0, nb, h, H, raw_size, ks, vs);
  __syncthreads();

  // ---- each lane owns R query rows of one block ----------------------------------------------------
  if (tid >= L::LANES) return;
  const int g = tid / L::LPB, pp = tid - g * L::LPB;
  const int blk = blk0 + g;
  if (blk >= nb) return;

  float a[R][E], nq[R], o[R][D], l[R];
  int nrow[R];
  {
    const int n0 = __ldg(kpos + (size_t)blk * B + (B - 1));
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = pp + r * L::LPB;
      nrow[r] = i < B ? __ldg(qpos + (size_t)blk * B + i) : -1;
      load_resident_row<L>(q, k, coords, scale_h, nrow[r], n0, h, H, raw_size, a[r], nq[r]);
      l[r] = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) o[r][d] = 0.f;
    }
  }

  const float4* krow = ks + (size_t)g * B * L::ROW_CHUNKS;
  const float4* vrow = vs + (size_t)g * B * L::VCH;
#pragma unroll 1
  for (int j = 0; j < B; ++j, krow += L::ROW_CHUNKS, vrow += L::VCH) {
    float s[R], nk = 0.f, unused = 0.f;
    float4 keep[L::USED_CHUNKS];
    dot_rows<L>(krow, a, s, nk, unused, keep);
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float tt = (s[r] + nq[r]) + nk;        // canonical order, see tile.cuh
      p[r] = exp2_fast(fminf(tt * kLog2e, 0.f));  // exp(min(S, 0)), example/hept.py:12
      l[r] += p[r];
    }
#pragma unroll
    for (int c = 0; c < L::VCH; ++c) {
      const float4 vv = vrow[c];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        o[r][4 * c + 0] = fmaf(p[r], vv.x, o[r][4 * c + 0]);
        o[r][4 * c + 1] = fmaf(p[r], vv.y, o[r][4 * c + 1]);
        o[r][4 * c + 2] = fmaf(p[r], vv.z, o[r][4 * c + 2]);
        o[r][4 * c + 3] = fmaf(p[r], vv.w, o[r][4 * c + 3]);
      }
    }
  }

  // ---- scatter back to original order: stage (H, N, T, 32) ----------------------------------------
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (nrow[r] < 0) continue;
    float4* dst = reinterpret_cast<float4*>(stage + (((size_t)h * N + nrow[r]) * T + t) * kStageRow);
#pragma unroll
    for (int c = 0; c < L::VCH; ++c) dst[c] = make_float4(o[r][4 * c], o[r][4 * c + 1], o[r][4 * c + 2], o[r][4 * c + 3]);
    dst[L::VCH] = make_float4(l[r] + 1e-20f, 0.f, 0.f, 0.f);  // denom = rowsum + 1e-20, example/hept.py:14
    if constexpr ((L::VCH + 1) % 2 == 1) dst[L::VCH + 1] = make_float4(0.f, 0.f, 0.f, 0.f);  // keep sectors whole
  }
}

// OR-combine: thread <-> (n, h, 16-byte chunk of the D outputs); reads the T staged rows of (h, n).
template <int D>
__global__ void __launch_bounds__(256) or_combine_kernel(const float* __restrict__ stage, int N, int H, int T,
                                                         float* __restrict__ out_pre, float* __restrict__ den_sum) {
  constexpr int VCH = D / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * H * VCH) return;
  const int c = (int)(idx % VCH);
  const size_t nh = idx / VCH;
  const int h = (int)(nh % H);
  const size_t n = nh / H;
  const float* rows = stage + ((size_t)h * N + n) * T * kStageRow;
  float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
  float den = 0.f;
  if (T <= 4) {   // every load in flight before the first add (T is a run-time value: the plain loop serialises them)
    float4 x[4];
    float dn[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      x[t] = t < T ? ldg4(rows + t * kStageRow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      dn[t] = t < T ? __ldg(rows + t * kStageRow + D) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {  // sum over tables in table order, like Tensor.sum(dim=0)
      if (t < T) {
        num.x += x[t].x; num.y += x[t].y; num.z += x[t].z; num.w += x[t].w;
        den += dn[t];
      }
    }
  } else {
    for (int t = 0; t < T; ++t) {
      const float4 x = ldg4(rows + t * kStageRow + 4 * c);
      num.x += x.x; num.y += x.y; num.z += x.z; num.w += x.w;
      den += __ldg(rows + t * kStageRow + D);
    }
  }
  float4 y = make_float4(num.x / den, num.y / den, num.z / den, num.w / den);
  *reinterpret_cast<float4*>(out_pre + (n * H + h) * D + 4 * c) = y;
  if (c == 0) den_sum[n * H + h] = den;
}

template <int D, int C, int B, int G, int R, int MINB>
static int launch_fwd(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                      const float* scale, const int32_t* positions, float* stage, cudaStream_t st) {
  using L = TileLayout<D, C, B, G, R>;
  auto kern = block_attn_fwd_kernel<D, C, B, G, R, MINB>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_BYTES);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "block_attn_fwd: cannot reserve %zu B of shared memory: %s",
                 L::SMEM_BYTES, cudaGetErrorString(e));
    configured.mark();
  }
  const int nb = s->N / s->B;
  dim3 grid((nb + G - 1) / G, s->T * s->H);
  kern<<<grid, L::THREADS, L::SMEM_BYTES, st>>>(q, k, v, coords, scale, positions, s->N, s->H, s->T, s->raw_size, stage);
  HEPT_CHECK_LAUNCH("block_attn_fwd");
  return HEPT_OK;
}

int block_attention_fwd_tc(const hept_shape* s, const float* q, const float* k, const float* v, const float* hat_coords,
                           const int32_t* positions, float* stage, cudaStream_t st);

}  // namespace hept

using namespace hept;

extern "C" int hept_shape_supported(int32_t D, int32_t C, int32_t B) {
  return (D == 24 && (C == 6 || C == 4) && (B == 100 || B == 64 || B == 128)) || (D == 8 && C == 6 && B == 10);
}
namespace hept {
// the tcgen05 tiles keep two (forward) or four (backward) score accumulators of B columns in the 512 TMEM columns:
// blocks of up to 112 hits; larger blocks run on the fp32 CUDA-core tiles whatever engine is selected
bool tc_tiles_supported(int D, int C, int B) { return hept_shape_supported(D, C, B) && B <= 112; }
}

extern "C" int hept_block_attention_fwd(const hept_shape* s, const float* q, const float* k, const float* v,
                                        const float* coords, const float* scale, const float* hat_coords,
                                        const int32_t* positions, float* stage, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(q && k && v && coords && scale && positions && stage, HEPT_EINVAL, "block_attention_fwd: null pointer");
  HEPT_REQUIRE((long long)s->T * s->H <= 65535, HEPT_EINVAL, "block_attention_fwd: T*H too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (engine() == 1 && tc_tiles_supported(s->D, s->C, s->B)) {
    HEPT_REQUIRE(hat_coords, HEPT_EINVAL, "block_attention_fwd: the tcgen05 engine needs hat_coords (hept_hat_coords)");
    return block_attention_fwd_tc(s, q, k, v, hat_coords, positions, stage, st);
  }
  if (s->D == 24 && s->C == 6 && s->B == 100) return launch_fwd<24, 6, 100, 5, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 100) return launch_fwd<24, 4, 100, 5, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 24 && s->C == 6 && s->B == 64) return launch_fwd<24, 6, 64, 8, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 64) return launch_fwd<24, 4, 64, 8, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 24 && s->C == 6 && s->B == 128) return launch_fwd<24, 6, 128, 4, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 24 && s->C == 4 && s->B == 128) return launch_fwd<24, 4, 128, 4, 2, 2>(s, q, k, v, coords, scale, positions, stage, st);
  if (s->D == 8 && s->C == 6 && s->B == 10) return launch_fwd<8, 6, 10, 4, 2, 1>(s, q, k, v, coords, scale, positions, stage, st);
  set_error("block_attention_fwd: (D=%d, C=%d, B=%d) not compiled in", s->D, s->C, s->B);
  return HEPT_EUNSUPPORTED;
}

extern "C" int hept_or_combine(const hept_shape* s, const float* stage, float* out_pre, float* den_sum, void* stream) {
  if (int rc = validate_shape(s)) return rc;
  HEPT_REQUIRE(stage && out_pre && den_sum, HEPT_EINVAL, "or_combine: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)s->N * s->H * (s->D / 4);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (s->D == 24) or_combine_kernel<24><<<grid, 256, 0, st>>>(stage, s->N, s->H, s->T, out_pre, den_sum);
  else if (s->D == 8) or_combine_kernel<8><<<grid, 256, 0, st>>>(stage, s->N, s->H, s->T, out_pre, den_sum);
  else {
    set_error("or_combine: D=%d not compiled in", s->D);
    return HEPT_EUNSUPPORTED;
  }
  HEPT_CHECK_LAUNCH("or_combine");
  return HEPT_OK;
}
