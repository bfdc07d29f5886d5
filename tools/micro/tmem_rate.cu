// Thread-side TMEM throughput: 8 warps (2 per lane quadrant) each store / load REPS x 2 chunks of 8 columns, as the tile
// kernels' epilogues do.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I hept_b200/csrc -o /tmp/tmem_rate tools/micro/tmem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

template <int MODE>   // 0: stores only, 1: loads only (one wait at the end), 2: load pair + wait + store pair per chunk
__global__ void __launch_bounds__(256) k(long long* out, int reps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot, lane_base = (uint32_t)((warp & 3) * 32) << 16, part = warp >> 2;
  uint32_t v[8], w[8];
  for (int i = 0; i < 8; ++i) { v[i] = lane + i; w[i] = lane * 3 + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int ci = 0; ci < 7; ++ci) {
      const uint32_t col = 8 * (part + 2 * ci);
      if (MODE == 0) { st8(tmem + lane_base + col, v); st8(tmem + lane_base + 112 + col, w); }
      if (MODE == 1) { ld8(tmem + lane_base + col, v); ld8(tmem + lane_base + 112 + col, w); }
      if (MODE == 2) {
        ld8(tmem + lane_base + col, v); ld8(tmem + lane_base + 112 + col, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) { v[i] += w[i]; }
        st8(tmem + lane_base + col, v); st8(tmem + lane_base + 112 + col, w);
      }
    }
    if (MODE == 0 || MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (MODE == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + w[i];
  if (s == 0xdeadbeef) out[1] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const char* names[3] = {"14 x STTM.x8 per warp, 8 warps (114 KB)", "14 x LDTM.x8 per warp, 8 warps (114 KB)", "7 x (2 LDTM, wait, 2 STTM) per warp, 8 warps"};
  for (int m = 0; m < 3; ++m) {
    for (int it = 0; it < 2; ++it) {
      if (m == 0) k<0><<<1, 256>>>(out, 100);
      if (m == 1) k<1><<<1, 256>>>(out, 100);
      if (m == 2) k<2><<<1, 256>>>(out, 100);
      cudaDeviceSynchronize();
    }
    printf("%s: %lld cycles per round -> %.1f B/clk  (%s)\n", names[m], out[0], 114688.0 * (m == 2 ? 2 : 1) / out[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
