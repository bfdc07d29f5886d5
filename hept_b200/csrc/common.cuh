// Shared device/host helpers for the hept_b200 sm_100a library.
#pragma once
#include <initializer_list>
#include <cstdint>

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/hept_b200.h"

namespace hept {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int bwd_stage_mask();
int engine();  // 0 = fp32 SIMT tiles, 1 = tcgen05 tiles
int bwd_variant();
bool tc_tiles_supported(int D, int C, int B);   // blocks of up to 112 hits fit the TMEM layout of the tcgen05 tiles
int hash_project_impl(const hept_shape* s, const float* q, const float* k, const float* coords, const float* scale,
                      const float* alpha, float* proj, float* span, void* workspace, size_t workspace_bytes, float* hat,
                      bool* hat_done, void* stream);

// Per-DEVICE one-time setup.  cudaFuncSetAttribute opt-ins and SM counts belong to the device that is current when a
// call arrives (the host side makes the tensors' device current around every call), not to the process: a second GPU in
// the same process needs its own opt-in and its own grid size.  Two threads racing through the same first use both
// configure (idempotent) before either marks it done.
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  static int device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
  }
  bool needed() const { return !((done.load(std::memory_order_acquire) >> device()) & 1ull); }
  void mark() { done.fetch_or(1ull << device(), std::memory_order_release); }
};
// SM count of the current device (cached per device); 0 if it cannot be read
inline int sm_count() {
  static std::atomic<int> cache[64];
  const int dev = DeviceOnce::device();
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 0;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

#define HEPT_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::hept::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

#define HEPT_CHECK_LAUNCH(name)                                                            \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      ::hept::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));         \
      return HEPT_ECUDA;                                                                   \
    }                                                                                      \
    ::hept::count_launch();                                                                \
  } while (0)

inline int validate_shape(const hept_shape* s) {
  HEPT_REQUIRE(s != nullptr, HEPT_EINVAL, "shape is null");
  HEPT_REQUIRE(s->N > 0 && s->H > 0 && s->D > 0 && s->C > 1 && s->T > 0 && s->B > 0, HEPT_EINVAL,
               "non-positive dimension (N=%d H=%d D=%d C=%d T=%d B=%d)", s->N, s->H, s->D, s->C, s->T, s->B);
  HEPT_REQUIRE(s->N % s->B == 0, HEPT_EINVAL, "N=%d is not a multiple of block_size=%d", s->N, s->B);
  HEPT_REQUIRE(s->raw_size >= 0 && s->raw_size <= s->N, HEPT_EINVAL, "raw_size=%d outside [0, N=%d]", s->raw_size,
               s->N);
  return HEPT_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kStageRow = 32;  // floats per staged (hit, head, table) row: 128 bytes

// float -> uint32 whose unsigned order equals the float order (-0.0 canonicalised to +0.0 first).
__device__ __forceinline__ uint32_t ordered_bits(float x) {
  x = x + 0.0f;  // -0.0 + 0.0 == +0.0 under round-to-nearest
  uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

// 16-byte asynchronous global -> shared copy that allocates in L1 (.ca): neighbouring rows share sectors
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// the same, L2 only (.cg): streamed rows nobody else on the SM reads
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// one 256-bit store (STG.E.256, sm_100): a thread's 32 contiguous bytes leave as ONE full sector instead of two half-sector
// writes that the L2 has to merge; dst must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* dst, const float4 a, const float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
// the kernels read and write rows as 16-byte vectors: every array pointer of the ABI must be 16-byte aligned
inline bool aligned16(std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs) if (reinterpret_cast<uintptr_t>(p) & 15u) return false;
  return true;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// 2^x, x <= 0.  ex2.approx.ftz: max relative error 2^-22; inputs below -126 flush to 0.
__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Loads `COUNT` consecutive floats (COUNT % 4 == 0 -> float4 loads, else float2 / scalar) into dst.
template <int COUNT>
__device__ __forceinline__ void load_row(const float* __restrict__ src, float* dst) {
  if constexpr (COUNT % 4 == 0) {
#pragma unroll
    for (int i = 0; i < COUNT / 4; ++i) {
      float4 t = ldg4(src + 4 * i);
      dst[4 * i + 0] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
    }
  } else if constexpr (COUNT % 2 == 0) {
#pragma unroll
    for (int i = 0; i < COUNT / 2; ++i) {
      float2 t = ldg2(src + 2 * i);
      dst[2 * i + 0] = t.x; dst[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < COUNT; ++i) dst[i] = __ldg(src + i);
  }
}

}  // namespace hept
