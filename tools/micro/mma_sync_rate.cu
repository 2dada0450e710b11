// Issue rate of the warp-level mma.sync on sm_100a (the legacy tensor path): m16n8k8 tf32 and m16n8k16 bf16, fp32 accumulate.
// Each warp keeps 8 independent accumulator tiles; prints MMA instructions per clock per SM for 1..16 warps per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a --cudart shared -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void rate_kernel(float* out, int iters, long long* cycles) {
  float acc[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
  unsigned a0 = threadIdx.x * 2654435761u, a1 = a0 ^ 0x3f800000u, a2 = a0 + 7u, a3 = a1 + 11u, b0 = a0 * 3u, b1 = a1 * 5u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += acc[t][0] + acc[t][1] + acc[t][2] + acc[t][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4096;
  for (int kind = 0; kind < 2; ++kind)
    for (int warps = 1; warps <= 16; warps *= 2) {
      if (kind == 0) rate_kernel<0><<<148, warps * 32>>>(out, iters, cyc); else rate_kernel<1><<<148, warps * 32>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per_clk = (double)iters * 8 * warps / mx;
      printf("%s warps/SM=%2d: %.3f mma/clk/SM = %.0f MAC/clk/SM (%s)\n", kind == 0 ? "m16n8k8 tf32 " : "m16n8k16 bf16", warps, per_clk,
             per_clk * (kind == 0 ? 1024 : 2048), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
