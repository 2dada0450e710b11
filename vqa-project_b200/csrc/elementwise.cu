// Small HBM-bound helpers of the VQA hot path: fused dropout, weight-norm (fwd/bwd), bias / broadcast reductions,
// gate backward, and the library's error plumbing.  All are single-pass, vectorised where alignment allows.
#include "common.cuh"
#include <cuda_bf16.h>
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

// Fused weight norm + operand split: columns [c0, c1) of w = v * g / ||v|| written directly as (hi, lo) bf16 planes - the
// fp32 effective weight is never materialised (it only ever fed vqa_split_bf16_f32).  One warp per row, float4 loads.
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__global__ void __launch_bounds__(256) weight_norm_split_kernel(const float* __restrict__ v, const float* __restrict__ g, int rows, int cols,
                                                               int c0, int c1, __nv_bfloat16* __restrict__ hi,
                                                               __nv_bfloat16* __restrict__ lo, long long ldp) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* vr = v + (long long)row * cols;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int c = lane * 4; c < cols; c += 128) {             // cols % 4 == 0 (host check)
    const float4 a = *reinterpret_cast<const float4*>(vr + c);
    s0 = fmaf(a.x, a.x, s0); s1 = fmaf(a.y, a.y, s1); s2 = fmaf(a.z, a.z, s2); s3 = fmaf(a.w, a.w, s3);
  }
  const float ss = warp_sum((s0 + s1) + (s2 + s3));
  const float s = g[row] / sqrtf(ss);
  __nv_bfloat16* hr = hi + (long long)row * ldp;
  __nv_bfloat16* lr = lo ? lo + (long long)row * ldp : nullptr;
  const int n = c1 - c0;
  for (int c = lane * 8; c < n; c += 256) {
    float x[8];
    if (c + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(vr + c0 + c), b = *reinterpret_cast<const float4*>(vr + c0 + c + 4);
      x[0] = a.x * s; x[1] = a.y * s; x[2] = a.z * s; x[3] = a.w * s; x[4] = b.x * s; x[5] = b.y * s; x[6] = b.z * s; x[7] = b.w * s;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = c + e < n ? vr[c0 + c + e] * s : 0.f;     // plane padding columns stay zero
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = pack2_bf16(x[2 * e], x[2 * e + 1]);
      l[e] = pack2_bf16(x[2 * e] - __uint_as_float(h[e] << 16), x[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u));
    }
    *reinterpret_cast<uint4*>(hr + c) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lr) *reinterpret_cast<uint4*>(lr + c) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <bool VEC>   // VEC: rows are 16-byte aligned (cols % 4 == 0) -> float4 accesses
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                             const float* __restrict__ g, float* __restrict__ dv,
                                                             float* __restrict__ dg, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* vr = v + (long long)row * cols;
  const float* dr = dw + (long long)row * cols;
  float ss = 0.f, dot = 0.f;
  if (VEC) {
    float s1 = 0.f, d1 = 0.f;
    for (int c = lane * 4; c < cols; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(vr + c), d = *reinterpret_cast<const float4*>(dr + c);
      ss = fmaf(a.x, a.x, ss); s1 = fmaf(a.y, a.y, s1); ss = fmaf(a.z, a.z, ss); s1 = fmaf(a.w, a.w, s1);
      dot = fmaf(d.x, a.x, dot); d1 = fmaf(d.y, a.y, d1); dot = fmaf(d.z, a.z, dot); d1 = fmaf(d.w, a.w, d1);
    }
    ss += s1; dot += d1;
  } else {
    for (int c = lane; c < cols; c += 32) { const float a = vr[c]; ss = fmaf(a, a, ss); dot = fmaf(dr[c], a, dot); }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float norm = sqrtf(ss);
  const float gr = g[row];
  if (lane == 0) dg[row] = dot / norm;
  const float a = gr / norm, b = gr * dot / (norm * ss);   // dv = g/||v|| * dw - g*dot/||v||^3 * v
  float* o = dv + (long long)row * cols;
  if (VEC) {
    for (int c = lane * 4; c < cols; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(vr + c), d = *reinterpret_cast<const float4*>(dr + c);
      *reinterpret_cast<float4*>(o + c) = make_float4(a * d.x - b * x.x, a * d.y - b * x.y, a * d.z - b * x.z, a * d.w - b * x.w);
    }
  } else {
    for (int c = lane; c < cols; c += 32) o[c] = a * dr[c] - b * vr[c];
  }
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
// vector variant: 64 column quads x 4 row lanes per block, float4 loads, 4 independent accumulators per thread; the four row
// lanes are combined in a fixed order through shared memory (deterministic, like stage 2)
__global__ void __launch_bounds__(256) colsum_stage1_v4(const float* __restrict__ x, long long ldx, float* __restrict__ scratch,
                                                       long long rows, int cols, long long rpb) {
  __shared__ float4 part[4][64];
  const int cq = blockIdx.x * 64 + threadIdx.x, ry = threadIdx.y;
  const long long r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  if (cq * 4 < cols) {
    const float* px = x + (long long)cq * 4;
    long long r = r0 + ry;
    for (; r + 12 < r1; r += 16) {
      const float4 v0 = *reinterpret_cast<const float4*>(px + r * ldx), v1 = *reinterpret_cast<const float4*>(px + (r + 4) * ldx),
                   v2 = *reinterpret_cast<const float4*>(px + (r + 8) * ldx), v3 = *reinterpret_cast<const float4*>(px + (r + 12) * ldx);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < r1; r += 4) {
      const float4 v0 = *reinterpret_cast<const float4*>(px + r * ldx);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
  }
  part[ry][threadIdx.x] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
  __syncthreads();
  if (ry == 0 && cq * 4 < cols) {
    const float4 p0 = part[0][threadIdx.x], p1 = part[1][threadIdx.x], p2 = part[2][threadIdx.x], p3 = part[3][threadIdx.x];
    *reinterpret_cast<float4*>(scratch + (long long)blockIdx.y * cols + cq * 4) =
        make_float4((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z), (p0.w + p1.w) + (p2.w + p3.w));
  }
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

extern "C" int vqa_weight_norm_split_f32(const float* v, const float* g, int rows, int cols, int c0, int c1, void* hi, void* lo,
                                         long long ldp, cudaStream_t stream) {
  VQA_CHECK_ARG(v && g && hi && rows > 0 && cols > 0 && 0 <= c0 && c0 < c1 && c1 <= cols, "vqa_weight_norm_split_f32: bad arguments");
  VQA_CHECK_ARG((cols & 3) == 0 && (c0 & 3) == 0 && aligned16(v), "vqa_weight_norm_split_f32: v needs 16-byte aligned rows (cols %% 4 == 0) and c0 %% 4 == 0");
  VQA_CHECK_ARG((ldp & 7) == 0 && ldp >= ((c1 - c0 + 7) & ~7) && aligned16(hi) && (!lo || aligned16(lo)), "vqa_weight_norm_split_f32: planes need ld %% 8 == 0 and ld >= round8(c1 - c0)");
  weight_norm_split_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(v, g, rows, cols, c0, c1, reinterpret_cast<__nv_bfloat16*>(hi),
                                                              reinterpret_cast<__nv_bfloat16*>(lo), ldp);
  VQA_LAUNCH_CHECK("weight_norm_split_kernel");
  return VQA_OK;
}

extern "C" int vqa_weight_norm_bwd_f32(const float* dw, const float* v, const float* g, float* dv, float* dg, int rows,
                                       int cols, cudaStream_t stream) {
  VQA_CHECK_ARG(dw && v && g && dv && dg && rows > 0 && cols > 0, "vqa_weight_norm_bwd_f32: bad arguments");
  if ((cols & 3) == 0 && aligned16(dw) && aligned16(v) && aligned16(dv))
    weight_norm_bwd_kernel<true><<<(rows + 7) / 8, 256, 0, stream>>>(dw, v, g, dv, dg, rows, cols);
  else
    weight_norm_bwd_kernel<false><<<(rows + 7) / 8, 256, 0, stream>>>(dw, v, g, dv, dg, rows, cols);
  VQA_LAUNCH_CHECK("weight_norm_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_colsum_f32(const float* x, long long ldx, float* out, float* scratch, long long rows, int cols,
                              cudaStream_t stream) {
  VQA_CHECK_ARG(x && out && scratch && rows > 0 && cols > 0 && ldx >= cols, "vqa_colsum_f32: bad arguments");
  const int nblk = (int)min(256LL, (rows + 63) / 64);
  const long long rpb = (rows + nblk - 1) / nblk;
  if ((cols & 3) == 0 && (ldx & 3) == 0 && aligned16(x) && aligned16(scratch)) {
    dim3 grid((cols / 4 + 63) / 64, nblk), block(64, 4);
    colsum_stage1_v4<<<grid, block, 0, stream>>>(x, ldx, scratch, rows, cols, rpb);
  } else {
    dim3 grid((cols + 255) / 256, nblk);
    colsum_stage1<<<grid, 256, 0, stream>>>(x, ldx, scratch, rows, cols, rpb);
  }
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
