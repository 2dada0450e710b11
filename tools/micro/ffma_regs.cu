// FFMA / FFMA2 throughput with THREE REGISTER operands (distinct registers), as in a register-blocked matrix product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_regs ffma_regs.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// NY distinct y registers, NC distinct c registers, 4 accumulator chains; all operands live in registers.
template <int NY> __global__ void k_ffma(float* out, const float* in, int iters) {
  float y[NY], c[8], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < NY; ++j) y[j] = in[threadIdx.x + 32 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = in[threadIdx.x + 7 * j + 1];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NY; ++j) acc[j & 3] = fmaf(c[j & 7], y[j], acc[j & 3]);
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] += 1e-9f;   // keep c loop-variant (cheap vs NY FMAs)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}
template <int NY> __global__ void k_ffma2(float* out, const unsigned long long* in, int iters) {
  unsigned long long y[NY], c[8], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < NY; ++j) y[j] = in[threadIdx.x + 32 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = in[threadIdx.x + 7 * j + 1];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NY; ++j) acc[j & 3] = ffma2(c[j & 7], y[j], acc[j & 3]);
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] += 0x100000001ull;
  }
  unsigned long long s = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; float* in; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
  const int iters = 4000;
  for (int warps : {1, 2, 4, 8}) {   // warps per scheduler (SMSP)
    const int blocks = 148, threads = 128 * warps;
    float t1 = timeit([&] { k_ffma<64><<<blocks, threads>>>(out, in, iters); });
    float t2 = timeit([&] { k_ffma2<32><<<blocks, threads>>>(out, (unsigned long long*)in, iters); });
    double n1 = (double)blocks * threads * iters * 64, n2 = (double)blocks * threads * iters * 32 * 2;
    printf("warps/SMSP %d: FFMA 3-reg %.1f TFMA/s   FFMA2 3-reg %.1f TFMA/s  (peak 148*128*1.965e9 = 37.2)\n", warps, n1 / t1 / 1e9, n2 / t2 / 1e9);
  }
  return 0;
}
