// Thin inline-PTX layer over the sm_100a tensor-core path (tcgen05 + TMEM + mbarrier), sized for the
// HEPT block tiles: one CTA, cta_group::1, kind::tf32, M = 128.
//
// Shared-memory operand layouts (both with the 128-byte swizzle, 16-byte chunk index ^= row & 7,
// tile bases 1024-byte aligned):
//   K-major  (A, and B for Q K^T): row r of the operand is 128 bytes = 32 fp32 along K; 8 rows = one
//            1024-byte atom; SBO = 1024.  One MMA consumes K = 8 (32 bytes): advance the start address
//            by 32 bytes inside the atom.
//   MN-major (B for P V): row k of the operand (a key) is 128 bytes = 32 fp32 along N.  For 32-bit
//            MN-major operands the only legal layout is SWIZZLE_128B_BASE32B: 32-byte chunk index ^= row & 3,
//            atoms of 4 k-rows (512 bytes), SBO = 512; one MMA (K = 8) = two atoms: advance by 1024 bytes.
// Descriptor bit layout follows cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hept {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor, SWIZZLE_128B -------------------------------------------------
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);            // start address, 16-byte units, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;      // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;      // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;                           // layout type
  return d;
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  return smem_desc(smem_addr, sbo_bytes, lbo_bytes, kLayoutSw128);
}

// ---- instruction descriptor: tf32 x tf32 -> f32, dense ---------------------------------------------
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                     // c_format = F32
         | (2u << 7)                   // a_format = TF32
         | (2u << 10)                  // b_format = TF32
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Warp-uniform variants: the whole warp executes the call, elect.sync picks the one lane that issues (always the same
// lane for a full mask, which is what tcgen05.commit needs: it tracks the MMAs of the issuing thread).  Keeps the
// issue loop free of divergence, so descriptors stay in uniform registers.
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* mbar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
      :: "r"(smem_u32(mbar)) : "memory");
}
// true on exactly one lane of a converged warp (always the same lane).  Branching on it tells ptxas that a single
// lane runs the guarded block, so tcgen05 instructions inside need no per-lane serialisation loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}\n" : "=r"(p));
  return p != 0;
}
// make the mbarrier observe completion of all MMAs issued so far by this thread
__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

// ---- TMEM allocation (one warp, power of two >= 32 columns) ----------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM <-> registers: 32 lanes x 32 bit, 16 consecutive columns per thread -----------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r) :: "memory");
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
         "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
         "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
         "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
         "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                  "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                  "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
// loads without the trailing wait, so several can be in flight; call tmem_wait_ld() before using the values
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, with the loaded registers as in/out operands: no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]),
                 "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
// after tmem_wait_ld(): an empty statement with the loaded registers as in/out operands -- volatile statements keep their
// order, so no use of the registers can be scheduled above the wait
__device__ __forceinline__ void tmem_tie(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_tie(uint32_t& a, uint32_t& b) { asm volatile("" : "+r"(a), "+r"(b) :: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: a descriptor or protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// The same with a suspend-time hint: the hardware parks the warp until the phase completes instead of returning after
// its short default limit, so a waiting warp stops spending issue slots on TRYWAIT / BRA pairs.  Waking up from the
// parked state is slower, though (measured, any hint value): good for the forward tiles (two tiles in flight, -2 %),
// bad for the backward tiles, whose eleven hand-offs per tile are all on the critical path (+11 %).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
    if (done) return;
  }
  __trap();
}

// one non-blocking probe of the phase with parity `parity`
__device__ __forceinline__ bool mbar_test(uint64_t* mbar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
  return done != 0;
}
// per-warpgroup register budget (all four warps of the warpgroup execute it)
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REGS)); }
// dynamic shared memory rounded up to 1024 B without leaving the shared address space (a cast through uintptr_t
// makes every later access a generic LD/ST)
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
// element `I` of a float4, I known at compile time
template <int I>
__device__ __forceinline__ float& elem(float4& v) {
  if constexpr (I == 0) return v.x;
  else if constexpr (I == 1) return v.y;
  else if constexpr (I == 2) return v.z;
  else return v.w;
}
// plain arrive (count 1), CTA scope, release semantics
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// named barrier among `count` threads (count % 32 == 0); id 0 is __syncthreads
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }

// ---- 16-byte asynchronous global -> shared copy; src_bytes == 0 zero-fills the destination -------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- tf32 split: x = hi + lo with hi = round-to-nearest tf32 --------------------------------------
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// hi = x rounded to tf32 (round half away, like cvt.rna; x finite), lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
// hi = x truncated to tf32, lo = x - hi (exact): one instruction less; the tensor core drops the low 13 bits of lo
// either way
__device__ __forceinline__ void trunc_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a SWIZZLE_128B_BASE32B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128b32_offset(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)(((((chunk >> 1) ^ (row & 3)) << 1) | (chunk & 1)) << 4);
}
// byte offset of 16-byte chunk `chunk` of row `row` inside a 128-byte-swizzled tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

}  // namespace umma
}  // namespace hept
