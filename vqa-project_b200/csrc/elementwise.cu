// Small HBM-bound helpers of the VQA hot path: fused dropout, weight-norm (fwd/bwd), bias / broadcast reductions,
// gate backward, and the library's error plumbing.  All are single-pass, vectorised where alignment allows.
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/vqa_b200.h"

thread_local char g_vqa_err[512] = "";

extern "C" const char* vqa_last_error(void) { return g_vqa_err; }
int g_vqa_sm_budget = vqa::kNumSMs;
extern "C" int vqa_set_sm_budget(int n_sms) {
  const int old = g_vqa_sm_budget;
  g_vqa_sm_budget = n_sms < 8 ? 8 : (n_sms > vqa::kNumSMs ? vqa::kNumSMs : n_sms);
  return old;
}
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

// One CTA per row, the row (and dw) held in registers: v and dw are read ONCE.  NV float4 per thread: cols <= 1024 * NV.
__device__ __forceinline__ float block_sum_256(float x, float* sh) {
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sh[w];                  // fixed order: deterministic
  __syncthreads();
  return t;
}
template <int NV>
__global__ void __launch_bounds__(256) weight_norm_bwd_row_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                                 const float* __restrict__ g, float* __restrict__ dv,
                                                                 float* __restrict__ dg, int cols) {
  __shared__ float sh[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* vr = v + (long long)row * cols;
  const float* dr = dw + (long long)row * cols;
  float4 a[NV], d[NV];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int c = (u * 256 + tid) * 4;
    if (c < cols) { a[u] = *reinterpret_cast<const float4*>(vr + c); d[u] = *reinterpret_cast<const float4*>(dr + c); }
    else { a[u] = make_float4(0.f, 0.f, 0.f, 0.f); d[u] = a[u]; }
    ss += (a[u].x * a[u].x + a[u].y * a[u].y) + (a[u].z * a[u].z + a[u].w * a[u].w);
    dot += (d[u].x * a[u].x + d[u].y * a[u].y) + (d[u].z * a[u].z + d[u].w * a[u].w);
  }
  ss = block_sum_256(ss, sh);
  dot = block_sum_256(dot, sh);
  const float norm = sqrtf(ss), gr = g[row];
  if (tid == 0) dg[row] = dot / norm;
  const float ca = gr / norm, cb = gr * dot / (norm * ss);   // dv = g/||v|| * dw - g*dot/||v||^3 * v
  float* o = dv + (long long)row * cols;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int c = (u * 256 + tid) * 4;
    if (c < cols) *reinterpret_cast<float4*>(o + c) = make_float4(ca * d[u].x - cb * a[u].x, ca * d[u].y - cb * a[u].y, ca * d[u].z - cb * a[u].z, ca * d[u].w - cb * a[u].w);
  }
}
template <int NV>
__global__ void __launch_bounds__(256) weight_norm_split_row_kernel(const float* __restrict__ v, const float* __restrict__ g, int cols, int c0,
                                                                   int c1, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                                   long long ldp) {
  __shared__ float sh[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* vr = v + (long long)row * cols;
  float4 a[NV];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int c = (u * 256 + tid) * 4;
    a[u] = c < cols ? *reinterpret_cast<const float4*>(vr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += (a[u].x * a[u].x + a[u].y * a[u].y) + (a[u].z * a[u].z + a[u].w * a[u].w);
  }
  ss = block_sum_256(ss, sh);
  const float s = g[row] / sqrtf(ss);
  __nv_bfloat16* hr = hi + (long long)row * ldp;
  __nv_bfloat16* lr = lo ? lo + (long long)row * ldp : nullptr;
  const int n8 = (c1 - c0 + 7) & ~7;                        // plane columns incl. zero padding; (c1 - c0) % 4 == 0 and c0 % 4 == 0
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int c = (u * 256 + tid) * 4;                       // column of v; plane column c - c0
    if (c >= c0 && c < c1) {
      const uint32_t h0 = pack2_bf16(a[u].x * s, a[u].y * s), h1 = pack2_bf16(a[u].z * s, a[u].w * s);
      *reinterpret_cast<uint2*>(hr + (c - c0)) = make_uint2(h0, h1);
      if (lr) *reinterpret_cast<uint2*>(lr + (c - c0)) = make_uint2(pack2_bf16(a[u].x * s - __uint_as_float(h0 << 16), a[u].y * s - __uint_as_float(h0 & 0xFFFF0000u)),
                                                                    pack2_bf16(a[u].z * s - __uint_as_float(h1 << 16), a[u].w * s - __uint_as_float(h1 & 0xFFFF0000u)));
    } else if (c >= c1 && c - c0 < n8) {                   // the padding quad of the last 8-column group stays zero (TMA reads it)
      *reinterpret_cast<uint2*>(hr + (c - c0)) = make_uint2(0u, 0u);
      if (lr) *reinterpret_cast<uint2*>(lr + (c - c0)) = make_uint2(0u, 0u);
    }
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
                                                       long long rows, int cols, long long rpb, float* __restrict__ out,
                                                       int* __restrict__ counters) {
  __shared__ float4 part[4][64];
  const int cq = blockIdx.x * 64 + threadIdx.x, ry = threadIdx.y;
  const long long r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  if (cq * 4 < cols) {
    const float* px = x + (long long)cq * 4;
    long long r = r0 + ry;
    for (; r + 28 < r1; r += 32) {                           // 8 loads (128 B) in flight per thread
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldcs(reinterpret_cast<const float4*>(px + (r + 4 * i) * ldx));
      a0.x += v[0].x; a0.y += v[0].y; a0.z += v[0].z; a0.w += v[0].w;
      a1.x += v[1].x; a1.y += v[1].y; a1.z += v[1].z; a1.w += v[1].w;
      a2.x += v[2].x; a2.y += v[2].y; a2.z += v[2].z; a2.w += v[2].w;
      a3.x += v[3].x; a3.y += v[3].y; a3.z += v[3].z; a3.w += v[3].w;
      a0.x += v[4].x; a0.y += v[4].y; a0.z += v[4].z; a0.w += v[4].w;
      a1.x += v[5].x; a1.y += v[5].y; a1.z += v[5].z; a1.w += v[5].w;
      a2.x += v[6].x; a2.y += v[6].y; a2.z += v[6].z; a2.w += v[6].w;
      a3.x += v[7].x; a3.y += v[7].y; a3.z += v[7].z; a3.w += v[7].w;
    }
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
  if (counters == nullptr) return;                          // two-kernel mode: colsum_stage2 follows
  // single-kernel mode: the LAST row block of this column block to finish adds the partials, in block order (deterministic)
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && ry == 0) {
    const int ticket = atomicAdd(&counters[blockIdx.x], 1);
    last = ticket == (int)gridDim.y - 1;
    if (last) counters[blockIdx.x] = 0;                     // self-resetting: the buffer stays zero between launches
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the four row lanes each add a quarter of the row blocks' partials (in block order), then lane 0 adds the four quarters in order
  {
    const int nb_ = (int)gridDim.y, q = (nb_ + 3) >> 2, b1 = min(nb_, (ry + 1) * q);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cq * 4 < cols) {
      int b = ry * q;
      for (; b + 8 <= b1; b += 8) {                         // loads batched, sums in block order
        float4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __ldcg(reinterpret_cast<const float4*>(scratch + (long long)(b + i) * cols + cq * 4));
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc.x += t[i].x; acc.y += t[i].y; acc.z += t[i].z; acc.w += t[i].w; }
      }
      for (; b < b1; ++b) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(scratch + (long long)b * cols + cq * 4));
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
    }
    part[ry][threadIdx.x] = acc;
  }
  __syncthreads();
  if (ry == 0 && cq * 4 < cols) {
    const float4 p0 = part[0][threadIdx.x], p1 = part[1][threadIdx.x], p2 = part[2][threadIdx.x], p3 = part[3][threadIdx.x];
    *reinterpret_cast<float4*>(out + cq * 4) =
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
  __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(lo);
  const bool rowk = ((c1 - c0) & 3) == 0 || c1 == cols;     // quads never straddle c1
  if (rowk && cols > 256 && cols <= 1024) weight_norm_split_row_kernel<1><<<rows, 256, 0, stream>>>(v, g, cols, c0, c1, ph, pl, ldp);
  else if (rowk && cols > 1024 && cols <= 2048) weight_norm_split_row_kernel<2><<<rows, 256, 0, stream>>>(v, g, cols, c0, c1, ph, pl, ldp);
  else if (rowk && cols > 2048 && cols <= 4096) weight_norm_split_row_kernel<4><<<rows, 256, 0, stream>>>(v, g, cols, c0, c1, ph, pl, ldp);
  else weight_norm_split_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(v, g, rows, cols, c0, c1, ph, pl, ldp);
  VQA_LAUNCH_CHECK("weight_norm_split_kernel");
  return VQA_OK;
}

extern "C" int vqa_weight_norm_bwd_f32(const float* dw, const float* v, const float* g, float* dv, float* dg, int rows,
                                       int cols, cudaStream_t stream) {
  VQA_CHECK_ARG(dw && v && g && dv && dg && rows > 0 && cols > 0, "vqa_weight_norm_bwd_f32: bad arguments");
  const bool vec = (cols & 3) == 0 && aligned16(dw) && aligned16(v) && aligned16(dv);
  if (vec && cols > 256 && cols <= 1024)
    weight_norm_bwd_row_kernel<1><<<rows, 256, 0, stream>>>(dw, v, g, dv, dg, cols);
  else if (vec && cols > 1024 && cols <= 2048)
    weight_norm_bwd_row_kernel<2><<<rows, 256, 0, stream>>>(dw, v, g, dv, dg, cols);
  else if (vec && cols > 2048 && cols <= 4096)
    weight_norm_bwd_row_kernel<4><<<rows, 256, 0, stream>>>(dw, v, g, dv, dg, cols);
  else if (vec)
    weight_norm_bwd_kernel<true><<<(rows + 7) / 8, 256, 0, stream>>>(dw, v, g, dv, dg, rows, cols);
  else
    weight_norm_bwd_kernel<false><<<(rows + 7) / 8, 256, 0, stream>>>(dw, v, g, dv, dg, rows, cols);
  VQA_LAUNCH_CHECK("weight_norm_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_colsum_f32(const float* x, long long ldx, float* out, float* scratch, long long rows, int cols,
                              int* counters, cudaStream_t stream) {
  VQA_CHECK_ARG(x && out && scratch && rows > 0 && cols > 0 && ldx >= cols, "vqa_colsum_f32: bad arguments");
  int nblk = (int)min(256LL, (rows + 63) / 64);
  long long rpb = (rows + nblk - 1) / nblk;
  if ((cols & 3) == 0 && (ldx & 3) == 0 && aligned16(x) && aligned16(scratch)) {
    // one resident wave: 8 blocks of 256 threads per SM; a second, nearly empty wave would cost as much as the first
    const int cblk = (cols / 4 + 63) / 64;
    const int one_wave = (kNumSMs * 8) / cblk;
    if (nblk > one_wave) nblk = one_wave < 1 ? 1 : one_wave;
    rpb = (rows + nblk - 1) / nblk;
    nblk = (int)((rows + rpb - 1) / rpb);
    dim3 grid(cblk, nblk), block(64, 4);
    const bool fused = counters != nullptr && aligned16(out);
    colsum_stage1_v4<<<grid, block, 0, stream>>>(x, ldx, scratch, rows, cols, rpb, out, fused ? counters : nullptr);
    VQA_LAUNCH_CHECK("colsum_stage1_v4");
    if (fused) return VQA_OK;
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
