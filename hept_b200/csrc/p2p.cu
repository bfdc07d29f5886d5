// the one collective of the path — the all-reduce of the parameter gradients of the data-parallel training
// configuration — as ONE kernel over NVLink peer memory instead of an NCCL launch.  The gradient bucket (1.3 MB for the tracking
// model, 57 KB for one attention module) is latency-bound: a one-shot all-reduce (every rank reads every peer's bucket directly
// and sums in RANK order, so all ranks get the same bits) costs two flag round trips over NVLink, where a ring pays 2 (W - 1).
//
// Memory: every rank owns a symmetric buffer [flags | data] mapped into all peers (torch symmetric memory provides the mapping:
// plumbing; hept_b200/sharding.py), `bufs` is the device array of the W base pointers.  A CTA owns a slice of the data and
// synchronises with the same CTA of every peer through monotonically increasing sequence numbers:
//   flags[(phase * B + cta) * W + src]   written by rank `src`, read by the owner
//   phase 0: "my slice is final"  -> peers may read it          phase 1: "I have read your slice" -> the owner may overwrite it
// Waits are bounded: a peer that never arrives sets *err instead of hanging the GPU.
#include "common.cuh"

namespace hept {

constexpr int kP2pThreads = 512, kP2pMaxCtas = 64, kP2pMaxWorld = 16;
constexpr size_t kP2pFlagBytes = 2 * kP2pMaxCtas * kP2pMaxWorld * sizeof(uint32_t);      // 8 KB in front of the data

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// signal `seq` to every peer's flag slot of (phase, cta, me), then wait until every peer has signalled mine
__device__ __forceinline__ void cross_gpu_barrier(uint8_t* const* bufs, int rank, int world, int phase, int ctas, uint32_t seq, int* err) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    const size_t slot = (size_t)(phase * ctas + blockIdx.x) * kP2pMaxWorld;
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(bufs[peer]) + slot + rank, seq);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(bufs[rank]) + slot + peer;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      if (++spins > (1ll << 26)) { atomicExch(err, 1 + phase); break; }      // ~ a second: report, do not hang
      __nanosleep(20);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kP2pThreads) p2p_allreduce_kernel(uint8_t* const* __restrict__ bufs, int rank, int world, long long n4,
                                                                    uint32_t seq, float scale, float4* __restrict__ scratch,
                                                                    int* __restrict__ err) {
  const int ctas = gridDim.x;
  const long long per = (n4 + ctas - 1) / ctas;
  const long long lo = blockIdx.x * per, hi = min(n4, lo + per);
  cross_gpu_barrier(bufs, rank, world, 0, ctas, seq, err);                  // every rank's slice is final and visible
  for (long long i = lo + threadIdx.x; i < hi; i += kP2pThreads) {
    float4 v[kP2pMaxWorld];
#pragma unroll
    for (int p = 0; p < kP2pMaxWorld; ++p)                                   // every peer's load in flight before the first add
      if (p < world) v[p] = ld_peer4(reinterpret_cast<const float*>(bufs[p] + kP2pFlagBytes) + 4 * i);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kP2pMaxWorld; ++p)                                   // rank order: the same bits on every rank
      if (p < world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
    scratch[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
  }
  cross_gpu_barrier(bufs, rank, world, 1, ctas, seq, err);                  // every peer has read my slice: it may change now
  float4* mine = reinterpret_cast<float4*>(bufs[rank] + kP2pFlagBytes);
  for (long long i = lo + threadIdx.x; i < hi; i += kP2pThreads) mine[i] = scratch[i];
}

}  // namespace hept

using namespace hept;

extern "C" size_t hept_p2p_flag_bytes(void) { return kP2pFlagBytes; }

extern "C" int hept_p2p_allreduce(void* const* bufs_dev, int32_t rank, int32_t world, int64_t n_floats, uint32_t seq, float scale,
                                  float* scratch, int32_t* err, void* stream) {
  HEPT_REQUIRE(bufs_dev && scratch && err, HEPT_EINVAL, "p2p_allreduce: null pointer");
  HEPT_REQUIRE(world >= 1 && world <= kP2pMaxWorld && rank >= 0 && rank < world && n_floats > 0 && n_floats % 4 == 0 && seq != 0, HEPT_EINVAL,
               "p2p_allreduce: bad argument (rank=%d world=%d n=%lld seq=%u)", rank, world, (long long)n_floats, seq);
  const long long n4 = n_floats / 4;
  int ctas = (int)((n4 + kP2pThreads - 1) / kP2pThreads);                 // one 16-byte element per thread up to 64 CTAs
  ctas = ctas < 1 ? 1 : (ctas > kP2pMaxCtas ? kP2pMaxCtas : ctas);
  p2p_allreduce_kernel<<<ctas, kP2pThreads, 0, (cudaStream_t)stream>>>((uint8_t* const*)bufs_dev, rank, world, n4, seq, scale,
                                                                       (float4*)scratch, err);
  HEPT_CHECK_LAUNCH("p2p_allreduce");
  return HEPT_OK;
}
