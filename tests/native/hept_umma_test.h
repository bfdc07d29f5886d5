/* Test-only entry points of tests/native/libhept_umma_test.so (NOT part of the product ABI include/hept_b200.h). */
#ifndef HEPT_UMMA_TEST_H
#define HEPT_UMMA_TEST_H
#ifdef __cplusplus
extern "C" {
#endif

const char* hept_debug_last_error(void);

/* self-test of the tcgen05 / TMEM building blocks (hept_b200/csrc/umma.cuh): S = A B^T with A (128,32),
 * Bm (112,32) both K-major, then O = S V with S read back from TMEM and V (112,32) MN-major.
 * Outputs S_out (128,112), O_out (128,96).  kmajor_base32 != 0 stores the K-major operands with the MN-major
 * operand's swizzle (SWIZZLE_128B_BASE32B). */
int hept_debug_umma_selftest(const float* A, const float* Bm, const float* V, float* S_out, float* O_out,
                             int kmajor_base32, void* stream);

/* S_xy = X Y^T and S_yx = Y X^T (X, Y (112,32) fp32; outputs (128,112), rows >= 112 zero) on the tensor core: probes
 * whether the tf32 MMA is bitwise symmetric under an exchange of its operands. */
int hept_debug_umma_symmetry(const float* X, const float* Y, float* S_xy, float* S_yx, void* stream);

/* cycles (clock64 of the issuing thread) from the first tcgen05.mma of a burst of `count` tf32 M = 128 MMAs to the
 * completion of its commit; see tools/umma_timing.py for the modes.  cycles is a device pointer to one int64. */
int hept_debug_umma_timing(int mode, int count, long long* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif
