// Small HBM-bound helpers of the VQA hot path: fused dropout, weight-norm (fwd/bwd), bias / broadcast reductions,
// gate backward, and the library's error plumbing.  All are single-pass, vectorised where alignment allows.
#include "common.cuh"
#include "../../include/vqa_b200.h"

thread_local char g_vqa_err[512] = "";

extern "C" const char* vqa_last_error(void) { return g_vqa_err; }
extern "C" int vqa_abi_version(void) { return VQA_ABI_VERSION; }

namespace vqa {

// ------------------------------------------------------------------------------------------ dropout
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                     float p, float scale, unsigned long long seed,
                                                     unsigned long long offset, const unsigned long long* __restrict__ step_ptr, int vec) {
  const Philox rng(seed);
  if (step_ptr) offset += *step_ptr * 16ull;     // device-side step counter: CUDA-graph replays draw fresh masks
  const long long ngroups = (n + 3) >> 2;
  for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (long long)gridDim.x * blockDim.x) {
    const uint4 r = rng((unsigned long long)gidx, offset);
    const long long e = gidx << 2;
    if (vec && e + 3 < n) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x + e));
      v.x = u32_to_unit(r.x) >= p ? v.x * scale : 0.f;
      v.y = u32_to_unit(r.y) >= p ? v.y * scale : 0.f;
      v.z = u32_to_unit(r.z) >= p ? v.z * scale : 0.f;
      v.w = u32_to_unit(r.w) >= p ? v.w * scale : 0.f;
      *reinterpret_cast<float4*>(y + e) = v;
    } else {
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
      for (int j = 0; j < 4 && e + j < n; ++j) y[e + j] = u32_to_unit(rr[j]) >= p ? x[e + j] * scale : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------ weight norm
// one warp per row
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                             float* __restrict__ w, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* vr = v + (long long)row * cols;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) { const float a = vr[c]; ss = fmaf(a, a, ss); }
  ss = warp_sum(ss);
  const float s = g[row] / sqrtf(ss);
  float* wr = w + (long long)row * cols;
  for (int c = lane; c < cols; c += 32) wr[c] = vr[c] * s;
}

__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                             const float* __restrict__ g, float* __restrict__ dv,
                                                             float* __restrict__ dg, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* vr = v + (long long)row * cols;
  const float* dr = dw + (long long)row * cols;
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < cols; c += 32) { const float a = vr[c]; ss = fmaf(a, a, ss); dot = fmaf(dr[c], a, dot); }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float norm = sqrtf(ss);
  const float gr = g[row];
  if (lane == 0) dg[row] = dot / norm;
  const float a = gr / norm, b = gr * dot / (norm * ss);   // dv = g/||v|| * dw - g*dot/||v||^3 * v
  float* o = dv + (long long)row * cols;
  for (int c = lane; c < cols; c += 32) o[c] = a * dr[c] - b * vr[c];
}

// ------------------------------------------------------------------------------------------ reductions
// stage 1: block b sums rows [b*rpb, (b+1)*rpb) for every column -> scratch[b, c]; stage 2 sums the blocks in order.
__global__ void __launch_bounds__(256) colsum_stage1(const float* __restrict__ x, long long ldx, float* __restrict__ scratch,
                                                    long long rows, int cols, long long rpb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += x[r * ldx + c];
  scratch[(long long)blockIdx.y * cols + c] = acc;
}
__global__ void __launch_bounds__(256) colsum_stage2(const float* __restrict__ scratch, float* __restrict__ out, int nblk, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int b = 0; b < nblk; ++b) acc += scratch[(long long)b * cols + c];
  out[c] = acc;
}

__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int seg_len, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const float* p = x + (long long)blockIdx.y * seg_len * cols + c;
  float acc = 0.f;
  for (int i = 0; i < seg_len; ++i) acc += p[(long long)i * cols];
  out[(long long)blockIdx.y * cols + c] = acc;
}

__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ dhq, const float* __restrict__ q,
                                                      const float* __restrict__ pooled, float* __restrict__ dpooled,
                                                      float* __restrict__ dq, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = dhq[i], qq = q[i], pp = pooled[i];
    dpooled[i] = (pp > 0.f && qq > 0.f) ? d * qq : 0.f;
    dq[i] = qq > 0.f ? d * pp : 0.f;
  }
}

}  // namespace vqa
using namespace vqa;

extern "C" int vqa_dropout_f32(const float* x, float* y, long long n, float p, unsigned long long seed,
                               unsigned long long offset, const unsigned long long* step_ptr, cudaStream_t stream) {
  VQA_CHECK_ARG(x && y && n >= 0, "vqa_dropout_f32: bad arguments");
  VQA_CHECK_ARG(p >= 0.f && p < 1.f, "vqa_dropout_f32: p must be in [0,1), got %f", p);
  if (n == 0) return VQA_OK;
  const long long groups = (n + 3) / 4;
  const int blocks = (int)min((long long)kNumSMs * 16, (groups + 255) / 256);
  dropout_kernel<<<blocks, 256, 0, stream>>>(x, y, n, p, 1.f / (1.f - p), seed, offset, step_ptr, aligned16(x) && aligned16(y));
  VQA_LAUNCH_CHECK("dropout_kernel");
  return VQA_OK;
}

extern "C" int vqa_weight_norm_fwd_f32(const float* v, const float* g, float* w, int rows, int cols, cudaStream_t stream) {
  VQA_CHECK_ARG(v && g && w && rows > 0 && cols > 0, "vqa_weight_norm_fwd_f32: bad arguments");
  weight_norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(v, g, w, rows, cols);
  VQA_LAUNCH_CHECK("weight_norm_fwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_weight_norm_bwd_f32(const float* dw, const float* v, const float* g, float* dv, float* dg, int rows,
                                       int cols, cudaStream_t stream) {
  VQA_CHECK_ARG(dw && v && g && dv && dg && rows > 0 && cols > 0, "vqa_weight_norm_bwd_f32: bad arguments");
  weight_norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(dw, v, g, dv, dg, rows, cols);
  VQA_LAUNCH_CHECK("weight_norm_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_colsum_f32(const float* x, long long ldx, float* out, float* scratch, long long rows, int cols,
                              cudaStream_t stream) {
  VQA_CHECK_ARG(x && out && scratch && rows > 0 && cols > 0 && ldx >= cols, "vqa_colsum_f32: bad arguments");
  const int nblk = (int)min(256LL, (rows + 63) / 64);
  const long long rpb = (rows + nblk - 1) / nblk;
  dim3 grid((cols + 255) / 256, nblk);
  colsum_stage1<<<grid, 256, 0, stream>>>(x, ldx, scratch, rows, cols, rpb);
  VQA_LAUNCH_CHECK("colsum_stage1");
  colsum_stage2<<<(cols + 255) / 256, 256, 0, stream>>>(scratch, out, nblk, cols);
  VQA_LAUNCH_CHECK("colsum_stage2");
  return VQA_OK;
}

extern "C" int vqa_segment_sum_f32(const float* x, float* out, int segments, int seg_len, int cols, cudaStream_t stream) {
  VQA_CHECK_ARG(x && out && segments > 0 && seg_len > 0 && cols > 0, "vqa_segment_sum_f32: bad arguments");
  dim3 grid((cols + 255) / 256, segments);
  segment_sum_kernel<<<grid, 256, 0, stream>>>(x, out, seg_len, cols);
  VQA_LAUNCH_CHECK("segment_sum_kernel");
  return VQA_OK;
}

extern "C" int vqa_gate_bwd_f32(const float* dhq, const float* q, const float* pooled, float* dpooled, float* dq,
                                long long n, cudaStream_t stream) {
  VQA_CHECK_ARG(dhq && q && pooled && dpooled && dq && n > 0, "vqa_gate_bwd_f32: bad arguments");
  const int blocks = (int)min((long long)kNumSMs * 8, (n + 255) / 256);
  gate_bwd_kernel<<<blocks, 256, 0, stream>>>(dhq, q, pooled, dpooled, dq, n);
  VQA_LAUNCH_CHECK("gate_bwd_kernel");
  return VQA_OK;
}
