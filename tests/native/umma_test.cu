// Self-test of the tcgen05 building blocks (umma.cuh) on exactly representable data: an SS MMA with both
// operands K-major (S = A B^T, 128 x 112 x 32) followed by a TS MMA whose A operand is read back from
// TMEM and whose B operand is MN-major (O = S V, 128 x 32 x 112).  Used by tests/test_gpu_umma.py.
//
// TEST CODE: built into its own library (tests/native/libhept_umma_test.so, `make -C tests/native`), never linked into
// the product library libhept_sm100.so; it shares only the header-only helpers of hept_b200/csrc (umma.cuh, common.cuh).
#include <stdarg.h>

#include <type_traits>

#include "../../hept_b200/csrc/common.cuh"
#include "../../hept_b200/csrc/umma.cuh"

namespace hept {

// the two hooks of common.cuh's macros that the product library defines in abi.cu
static thread_local char g_test_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_test_error, sizeof(g_test_error), fmt, ap);
  va_end(ap);
}
void count_launch(int) {}

constexpr int kTM = 128, kTN = 112, kTK = 32, kTV = 32;

// KM_B32 == true: the K-major operands A and Bm are stored with the SWIZZLE_128B_BASE32B pattern as well (the pattern
// the MN-major operand needs), i.e. one shared-memory image of a row-per-hit tile serves both roles.
// O_out (128, 96): columns [0, 32) = S V from an N = 32 MMA; columns [32, 96) = S [V | V2] from ONE N = 64 MMA per k-step whose
// MN-major B operand spans two 32-column tiles LBO bytes apart (V2 = 2 V + 1 elementwise).
template <bool KM_B32>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                               const float* __restrict__ V, float* __restrict__ S_out,
                                                               float* __restrict__ O_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  uint8_t* sA = smem;                       // 128 rows x 128 B
  uint8_t* sB = sA + kTM * 128;             // 112 rows x 128 B
  uint8_t* sV = sB + kTN * 128;             // 112 rows x 128 B   (row = k, 32 floats along n)
  uint8_t* sV2 = sV + kTN * 128;            // the second 32-column tile of the N = 64 operand
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < kTM * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sA + (KM_B32 ? umma::sw128b32_offset(r, c) : umma::sw128_offset(r, c))) =
        *reinterpret_cast<const float4*>(A + r * kTK + 4 * c);
  }
  for (int idx = tid; idx < kTN * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sB + (KM_B32 ? umma::sw128b32_offset(r, c) : umma::sw128_offset(r, c))) =
        *reinterpret_cast<const float4*>(Bm + r * kTK + 4 * c);
    const float4 vv = *reinterpret_cast<const float4*>(V + r * kTV + 4 * c);
    *reinterpret_cast<float4*>(sV + umma::sw128b32_offset(r, c)) = vv;
    *reinterpret_cast<float4*>(sV2 + umma::sw128b32_offset(r, c)) = make_float4(2.f * vv.x + 1.f, 2.f * vv.y + 1.f, 2.f * vv.z + 1.f, 2.f * vv.w + 1.f);
  }
  if (tid == 0) umma::mbar_init(&mbar, 1);
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base_slot);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_tf32(kTM, kTN, false, false);
#pragma unroll
    for (int k = 0; k < kTK / 8; ++k) {
      const uint32_t lay = KM_B32 ? umma::kLayoutSw128Base32 : umma::kLayoutSw128;
      const uint64_t da = umma::smem_desc(umma::smem_u32(sA) + 32 * k, 1024, 16, lay);
      const uint64_t db = umma::smem_desc(umma::smem_u32(sB) + 32 * k, 1024, 16, lay);
      umma::mma_ss(tS, da, db, idesc, k > 0);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();

  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
  for (int c0 = 0; c0 < kTN; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tS + lane_base + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) S_out[tid * kTN + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();

  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_tf32(kTM, kTV, false, true);
#pragma unroll 1
    for (int k = 0; k < kTN / 8; ++k) {
      const uint64_t db = umma::smem_desc(umma::smem_u32(sV) + 1024 * k, 512, 1024, umma::kLayoutSw128Base32);
      umma::mma_ts(tO, tS + 8 * k, db, idesc, k > 0);
      const uint64_t db2 = umma::smem_desc(umma::smem_u32(sV) + 1024 * k, 512, kTN * 128, umma::kLayoutSw128Base32);
      umma::mma_ts(tO + 32, tS + 8 * k, db2, umma::idesc_tf32(kTM, 64, false, true), k > 0);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 1);
  umma::fence_after_sync();
#pragma unroll 1
  for (int c0 = 0; c0 < 3 * kTV; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tO + lane_base + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) O_out[tid * 3 * kTV + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

// Is the tf32 MMA bitwise symmetric under an exchange of its operands?  S = X Y^T and S' = Y X^T on the same
// 112 x 32 fp32 matrices (rows 112..127 of the M side are zero), each accumulated over the same four K = 8 steps.
__global__ void __launch_bounds__(128, 1) umma_symmetry_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                               float* __restrict__ S_xy, float* __restrict__ S_yx) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  uint8_t* sX = smem;                       // 128 rows x 128 B (rows >= 112 zero)
  uint8_t* sY = sX + kTM * 128;             // 128 rows x 128 B
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < kTM * 8; idx += 128) {
    const int r = idx >> 3, c = idx & 7;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sX + umma::sw128_offset(r, c)) = r < kTN ? *reinterpret_cast<const float4*>(X + r * kTK + 4 * c) : z;
    *reinterpret_cast<float4*>(sY + umma::sw128_offset(r, c)) = r < kTN ? *reinterpret_cast<const float4*>(Y + r * kTK + 4 * c) : z;
  }
  if (tid == 0) umma::mbar_init(&mbar, 1);
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base_slot);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_tf32(kTM, kTN, false, false);
#pragma unroll
    for (int k = 0; k < kTK / 8; ++k) {
      const uint64_t dx = umma::smem_desc_sw128(umma::smem_u32(sX) + 32 * k, 1024, 16);
      const uint64_t dy = umma::smem_desc_sw128(umma::smem_u32(sY) + 32 * k, 1024, 16);
      umma::mma_ss(tmem, dx, dy, idesc, k > 0);          // S_xy[i][j] = x_i . y_j
      umma::mma_ss(tmem + 128, dy, dx, idesc, k > 0);    // S_yx[j][i] = y_j . x_i
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
  for (int c0 = 0; c0 < kTN; c0 += 16) {
    float a[16], b[16];
    umma::tmem_ld16(tmem + lane_base + c0, a);
    umma::tmem_ld16(tmem + 128 + lane_base + c0, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) { S_xy[tid * kTN + c0 + i] = a[i]; S_yx[tid * kTN + c0 + i] = b[i]; }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

}  // namespace hept

using namespace hept;

extern "C" const char* hept_debug_last_error(void) { return hept::g_test_error; }

extern "C" int hept_debug_umma_symmetry(const float* X, const float* Y, float* S_xy, float* S_yx, void* stream) {
  HEPT_REQUIRE(X && Y && S_xy && S_yx, HEPT_EINVAL, "umma_symmetry: null pointer");
  const size_t smem = (size_t)2 * kTM * 128 + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(umma_symmetry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "umma_symmetry: %s", cudaGetErrorString(e));
    configured.mark();
  }
  umma_symmetry_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(X, Y, S_xy, S_yx);
  HEPT_CHECK_LAUNCH("umma_symmetry");
  return HEPT_OK;
}

extern "C" int hept_debug_umma_selftest(const float* A, const float* Bm, const float* V, float* S_out, float* O_out,
                                        int kmajor_base32, void* stream) {
  HEPT_REQUIRE(A && Bm && V && S_out && O_out, HEPT_EINVAL, "umma_selftest: null pointer");
  const size_t smem = (size_t)(kTM + 3 * kTN) * 128 + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "umma_selftest: %s", cudaGetErrorString(e));
    configured.mark();
  }
  if (kmajor_base32) umma_selftest_kernel<true><<<1, 128, smem, (cudaStream_t)stream>>>(A, Bm, V, S_out, O_out);
  else umma_selftest_kernel<false><<<1, 128, smem, (cudaStream_t)stream>>>(A, Bm, V, S_out, O_out);
  HEPT_CHECK_LAUNCH("umma_selftest");
  return HEPT_OK;
}

// ---- timing probe: cycles for a burst of tcgen05.mma (tf32, M = 128) from issue to mbarrier completion -----------
// mode 0: `count` TS MMAs (A from TMEM, N = 32) accumulating into ONE accumulator
// mode 1: the same MMAs alternating between TWO accumulators
// mode 2: `count` SS MMAs, N = 112, one accumulator
// mode 3: `count` SS MMAs, N = 112, alternating between two accumulators
// mode 4: TS MMAs, N = 32, round-robin over FOUR accumulators
namespace hept {
__global__ void __launch_bounds__(128, 1) umma_timing_kernel(int mode, int count, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 5 * 128 * 32; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) umma::mbar_init(&mbar, 1);
  if (warp == 0) umma::tmem_alloc<512>(&tmem_base_slot);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  {  // defined A operand in TMEM
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    for (int c0 = 0; c0 < 512; c0 += 16) umma::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c0, z);
    umma::tmem_wait_st();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  uint32_t phase = 0;
  for (int rep = 0; rep < 3; ++rep) {   // the last repetition is reported
    long long t0 = 0;
    if (tid == 0) {
      const uint32_t sA = umma::smem_u32(smem), sB = sA + 128 * 128, sV = sB + 128 * 128;
      constexpr uint32_t id_ts = umma::idesc_tf32(128, 32, false, true);
      constexpr uint32_t id_ss = umma::idesc_tf32(128, 112, false, false);
      const uint64_t dA = umma::smem_desc_sw128(sA, 1024, 16), dB = umma::smem_desc_sw128(sB, 1024, 16);
      const uint64_t dV = umma::smem_desc(sV, 512, 1024, umma::kLayoutSw128Base32);
      const uint32_t acc0 = tmem + 256;
      t0 = clock64();
      // straight-line issue (descriptor = base + constant), so the burst is paced by the tensor pipe, not by the loop
      auto burst = [&](auto mode_c) {
        constexpr int M = decltype(mode_c)::value;
#pragma unroll
        for (int i = 0; i < 52; ++i) {
          if (i >= count) break;
          if constexpr (M >= 9 && M <= 11) {       // the backward's TS product: N = 64 (A hi) + N = 32 (A lo) per k-step, at its TMEM columns:
            // 9: A at 224 / 336, D at 448 (dV);  10: A at 0 / 112, D at 448 (dK);  11: A at 0 / 112, D at 224 (dQ)
            constexpr uint32_t a_hi = M == 9 ? 224 : 0, a_lo = M == 9 ? 336 : 112, dcol = M == 11 ? 224 : 448;
            const uint64_t dW = umma::smem_desc(sA, 512, 13 * 1024, umma::kLayoutSw128Base32);
            const int kk = (i >> 1) % 13;
            if ((i & 1) == 0) umma::mma_ts(tmem + dcol, tmem + a_hi + 8 * kk, dW + 64 * kk, umma::idesc_tf32(128, 64, false, true), i >= 2);
            else umma::mma_ts(tmem + dcol + 32, tmem + a_lo + 8 * kk, dW + 64 * kk, id_ts, true);
          } else if constexpr (M >= 5 && M <= 7) {        // TS, one accumulator, N = 64 / 96 / 128: B spans N / 32 row-per-k tiles, LBO apart
            constexpr int NN = M == 5 ? 64 : (M == 6 ? 96 : 128);
            const uint64_t dW = umma::smem_desc(sA, 512, 13 * 1024, umma::kLayoutSw128Base32);
            umma::mma_ts(acc0, tmem + 8 * (i % 13), dW + 64 * (i % 13), umma::idesc_tf32(128, NN, false, true), i >= 1);
          } else if constexpr (M == 8) {           // SS, N = 32
            umma::mma_ss(acc0, dA + 2 * (i & 3), dB + 2 * (i & 3), umma::idesc_tf32(128, 32, false, false), i >= 1);
          } else if constexpr (M == 0 || M == 1 || M == 4) {
            constexpr int nacc = M == 0 ? 1 : (M == 1 ? 2 : 4);
            umma::mma_ts(acc0 + 32 * (i % nacc), tmem + 8 * (i % 13), dV + 64 * (i % 13), id_ts, i >= nacc);
          } else {
            umma::mma_ss(acc0 + (M == 3 ? 112 * (i & 1) : 0), dA + 2 * (i & 3), dB + 2 * (i & 3), id_ss, i >= (M == 3 ? 2 : 1));
          }
        }
      };
      switch (mode) {
        case 0: burst(std::integral_constant<int, 0>{}); break;
        case 1: burst(std::integral_constant<int, 1>{}); break;
        case 2: burst(std::integral_constant<int, 2>{}); break;
        case 3: burst(std::integral_constant<int, 3>{}); break;
        case 4: burst(std::integral_constant<int, 4>{}); break;
        case 5: burst(std::integral_constant<int, 5>{}); break;
        case 6: burst(std::integral_constant<int, 6>{}); break;
        case 7: burst(std::integral_constant<int, 7>{}); break;
        case 9: burst(std::integral_constant<int, 9>{}); break;
        case 10: burst(std::integral_constant<int, 10>{}); break;
        case 11: burst(std::integral_constant<int, 11>{}); break;
        default: burst(std::integral_constant<int, 8>{}); break;
      }
      umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    if (tid == 0) cycles[0] = clock64() - t0;
    __syncthreads();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}
}  // namespace hept

extern "C" int hept_debug_umma_timing(int mode, int count, long long* cycles, void* stream) {
  HEPT_REQUIRE(cycles && count > 0, HEPT_EINVAL, "umma_timing: bad argument");
  const size_t smem = (size_t)5 * 128 * 128 + 1024;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(hept::umma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HEPT_REQUIRE(e == cudaSuccess, HEPT_ECUDA, "umma_timing: %s", cudaGetErrorString(e));
    configured.mark();
  }
  hept::umma_timing_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, count, cycles);
  HEPT_CHECK_LAUNCH("umma_timing");
  return HEPT_OK;
}
