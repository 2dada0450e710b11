// What does CC.CF hold after sub.cc / subc.cc, as seen by a following addc?  (ranking loop of csrc/adjacency.cu)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a --cudart shared -o cc_polarity cc_polarity.cu && ./cc_polarity
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const unsigned* in, unsigned* out) {
  unsigned xlo = in[0], xhi = in[1], ylo = in[2], yhi = in[3], r = 100, r2 = 100;
  asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(r) : "r"(xlo), "r"(ylo), "r"(xhi), "r"(yhi));
  asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\tsubc.u32 %0, %0, 0;\n\t}" : "+r"(r2) : "r"(xlo), "r"(ylo), "r"(xhi), "r"(yhi));
  out[0] = r; out[1] = r2;
}
int main() {
  unsigned *in, *out, h[4], o[2];
  cudaMalloc(&in, 16); cudaMalloc(&out, 8);
  const unsigned cases[4][4] = {{5, 7, 9, 7}, {9, 7, 5, 7}, {5, 7, 5, 7}, {5, 8, 9, 7}};   // (xlo, xhi, ylo, yhi)
  for (auto& c : cases) {
    for (int i = 0; i < 4; ++i) h[i] = c[i];
    cudaMemcpy(in, h, 16, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(in, out);
    cudaMemcpy(o, out, 8, cudaMemcpyDeviceToHost);
    const unsigned long long x = ((unsigned long long)c[1] << 32) | c[0], y = ((unsigned long long)c[3] << 32) | c[2];
    printf("x %s y: addc -> %u (100 + borrow would be %u), subc -> %u (100 - borrow would be %u)\n", x < y ? "<" : (x == y ? "==" : ">"), o[0], 100 + (x < y), o[1], 100 - (x < y));
  }
  return 0;
}
