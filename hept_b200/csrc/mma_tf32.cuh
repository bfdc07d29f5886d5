// 3xTF32 products on the legacy tensor path (mma.sync m16n8k8): x = hi + lo with hi, lo in tf32; a product keeps
// lo*hi + hi*lo + hi*hi with fp32 accumulation (the dropped lo*lo term is 2^-22 relative).  Used by the skinny products around
// the attention call (K or N = 24): they are bound by the FMA pipe / shared-memory wavefronts on CUDA cores, and too small per
// tile for a tcgen05 pipeline to pay.
//
// Fragment layout (PTX ISA, m16n8k8 .tf32), g = lane / 4, t = lane % 4:
//   A (16 x 8, row):  a0 = A[g][t]      a1 = A[g + 8][t]      a2 = A[g][t + 4]      a3 = A[g + 8][t + 4]
//   B (8 x 8, col):   b0 = B[t][g]      b1 = B[t + 4][g]
//   C (16 x 8):       c0 = C[g][2 t]    c1 = C[g][2 t + 1]    c2 = C[g + 8][2 t]    c3 = C[g + 8][2 t + 1]
#pragma once
#include <cstdint>

namespace hept {

// round to nearest (ties away) at bit 13 on the bit pattern: cvt.rna.tf32.f32 without its NaN / infinity guards, which cost
// three more instructions per value (a finite |x| above 0x7f7ff000 rounds to infinity, as its products would overflow)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace hept
