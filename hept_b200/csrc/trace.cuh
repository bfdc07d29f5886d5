// Pipeline timeline probe for the persistent tile kernels, compiled in only with -DHEPT_TRACE (make TRACE=1 builds
// libhept_sm100_trace.so next to the product library; tools/pipeline_trace.py reads it).  One lane of one warp per
// role of CTA 0 stamps clock64() at its hand-off points for the first kTraceTiles tiles.
#pragma once

#include <cuda_runtime.h>

namespace hept {
constexpr int kTraceTiles = 64, kTraceEvents = 32;
#ifdef HEPT_TRACE
static __device__ long long* g_trace_ptr;
__device__ __forceinline__ unsigned __smid_for_trace() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define HEPT_TRACE_EVENT(ev, it)                                                                      \
  do {                                                                                                \
    if (blockIdx.x == 0 && (it) < ::hept::kTraceTiles && (threadIdx.x & 31) == 0 && g_trace_ptr)      \
      g_trace_ptr[(ev) * ::hept::kTraceTiles + (it)] = clock64();                                     \
  } while (0)
// per-CTA stamp (slot 0 = start, 1 = end), stored behind the event table
#define HEPT_TRACE_CTA(slot)                                                                          \
  do {                                                                                                \
    if ((threadIdx.x & 31) == 0 && g_trace_ptr)                                                       \
      g_trace_ptr[::hept::kTraceEvents * ::hept::kTraceTiles + 4 * blockIdx.x + (slot)] =            \
          (slot) == 2 ? (long long)__smid_for_trace() : clock64();                                    \
  } while (0)
#define HEPT_TRACE_SETTER(name)                                                                       \
  extern "C" int name(long long* buf) {                                                               \
    return cudaMemcpyToSymbol(::hept::g_trace_ptr, &buf, sizeof(buf)) == cudaSuccess ? 0 : -3;        \
  }
#else
#define HEPT_TRACE_EVENT(ev, it) do { } while (0)
#define HEPT_TRACE_CTA(slot) do { } while (0)
#define HEPT_TRACE_SETTER(name)
#endif
}  // namespace hept
