// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
  float a[16]; for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float x = out[0] + 1.0001f, y = 0.9999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  unsigned long long a[8];
  for (int i = 0; i < 8; ++i) { float2 t = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); a[i] = *reinterpret_cast<unsigned long long*>(&t); }
  float2 xv = make_float2(out[0] + 1.0001f, out[0] + 1.0002f), yv = make_float2(0.9999f, 0.9998f);
  unsigned long long x = *reinterpret_cast<unsigned long long*>(&xv), y = *reinterpret_cast<unsigned long long*>(&yv);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
  }
  float s = 0; for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMemset(d, 0, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_ffma<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * 148 * 8 * 256;
    printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaEventRecord(e0); k_ffma2<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
