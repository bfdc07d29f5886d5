// Issue rate of the legacy tensor path on this GPU: mma.sync m16n8k8 tf32 with independent accumulators.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_rate tools/micro/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(512) rate_kernel(float* out, int iters) {
  float acc[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b0 = threadIdx.x, b1 = threadIdx.x * 3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int warps, int sms, float* out) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<CHAINS><<<sms, 32 * warps>>>(out, 100);
  cudaEventRecord(e0);
  rate_kernel<CHAINS><<<sms, 32 * warps>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mma_per_sm = (double)iters * CHAINS * warps;
  printf("chains %d warps %2d: %.3f ms, %.2f clk per mma per SM (at %d MHz nominal), %.1f TFMA/s\n", CHAINS, warps, ms,
         ms * 1e-3 * khz * 1e3 / mma_per_sm, khz / 1000, mma_per_sm * sms * 1024 / (ms * 1e-3) * 1e-12);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 512);
  for (int warps : {4, 8, 16}) { run<1>(warps, sms, out); run<3>(warps, sms, out); run<8>(warps, sms, out); }
  return 0;
}
