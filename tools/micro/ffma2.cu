// FFMA vs FFMA2 issue-rate microbenchmark (sm_100a). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int CH> __global__ void k_ffma(float* out, float a, float b, int iters) {
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], a, b);
  }
  float s = 0; for (int c = 0; c < CH; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __global__ void k_ffma2(float* out, unsigned long long a, unsigned long long b, int iters) {
  unsigned long long acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = ffma2(acc[c], a, b);
  }
  unsigned long long s = 0; for (int c = 0; c < CH; ++c) s ^= acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 512 * 4);
  const int iters = 20000, blocks = 148 * 4, threads = 512;   // 16 warps / scheduler
  unsigned long long one = 0x3f8000003f800000ull;
  float t1 = timeit([&] { k_ffma<8><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
  float t2 = timeit([&] { k_ffma2<8><<<blocks, threads>>>(out, one, one, iters); });
  float t3 = timeit([&] { k_ffma2<2><<<blocks, threads>>>(out, one, one, iters); });
  double n = (double)blocks * threads * iters * 8;
  printf("FFMA  x8 chains: %.3f ms  -> %.1f TFLOP/s (fp32 FMA lanes)\n", t1, 2 * n / t1 / 1e9);
  printf("FFMA2 x8 chains: %.3f ms  -> %.1f TFLOP/s\n", t2, 4 * n / t2 / 1e9);
  printf("FFMA2 x2 chains: %.3f ms  -> %.1f TFLOP/s\n", t3, 4 * n / 4 / t3 / 1e9);
  return 0;
}
