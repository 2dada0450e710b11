// Question encoder pieces (sm_100a): embedding gather / scatter-add and the GRU cell pointwise kernels.
//
// The reference encodes the question with nn.Embedding + pack_padded_sequence + nn.GRU and keeps the final hidden
// state of every sequence (sparse_graph_model.py:117-121).  Here the recurrence runs over PADDED time-major steps:
// all B sequences advance together for T = max length steps, a sequence past its own length keeps its state
// (h_t = h_{t-1} for t >= len_b), which yields exactly the packed-sequence result with no host-side packing, no
// data-dependent shapes (the whole train step becomes CUDA-graph capturable) and lets the six matrix products of the
// cell run as a few large tcgen05 GEMMs (gemm_bf16s.cu): GI = E W_ih^T for all steps at once, GH_t = h_{t-1} W_hh^T
// per step, and in backward dW_ih, dW_hh, dE as single products over all T*B rows.
//
// Gate order and formulas are torch's (r, z, n):  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n),
// h' = (1 - z) * n + z * h,  with gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh.
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/vqa_b200.h"

namespace vqa {

__device__ __forceinline__ uint32_t gru_pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[4]) {
  float h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
  *reinterpret_cast<uint2*>(hi) = make_uint2(gru_pack_bf16(h[0], h[1]), gru_pack_bf16(h[2], h[3]));
  if (lo) *reinterpret_cast<uint2*>(lo) = make_uint2(gru_pack_bf16(v[0] - h[0], v[1] - h[1]), gru_pack_bf16(v[2] - h[2], v[3] - h[3]));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// E[(t*B + b), :] = W[question[b, t], :] as split planes (time-major rows); one warp per output row.
__global__ void __launch_bounds__(256) embed_gather_split_kernel(const long long* __restrict__ question, long long ldq,
                                                                const float* __restrict__ W, int emb, long long vocab,
                                                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                                long long ldp, int B, int T, int* __restrict__ err) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T * B) return;
  const int t = row / B, b = row - t * B;
  long long tok = question[(long long)b * ldq + t];
  if (tok < 0 || tok >= vocab) {            // nn.Embedding raises: flag it (the host raises at its next sync point), never read out of range
    if (lane == 0 && err) atomicOr(err, 1);
    tok = 0;
  }
  const float* src = W + tok * emb;
  for (int c = lane * 4; c < (int)ldp; c += 128) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = c + e < emb ? src[c + e] : 0.f;
    store_split4(hi + (long long)row * ldp + c, lo ? lo + (long long)row * ldp + c : nullptr, v);
  }
}

// dW[question[b,t], :] += dE[(t*B + b), :] for t < len_b (rows past the sequence end carry zero gradient).
__global__ void __launch_bounds__(256) embed_scatter_add_kernel(const float* __restrict__ dE, long long ldd,
                                                               const long long* __restrict__ question, long long ldq,
                                                               const int* __restrict__ len, float* __restrict__ dW, int emb,
                                                               long long vocab, int B, int T) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T * B) return;
  const int t = row / B, b = row - t * B;
  if (t >= len[b]) return;
  const long long tok = question[(long long)b * ldq + t];
  if (tok < 0 || tok >= vocab) return;
  const float* src = dE + (long long)row * ldd;
  float* dst = dW + tok * emb;
  for (int c = lane; c < emb; c += 32) atomicAdd(dst + c, src[c]);
}

// One GRU step for all B sequences; 4 hidden units per thread.
//   gi: (B, 3H) rows of GI_t (bias b_ih already added by the GEMM epilogue), ld = ldgi
//   gh: (B, 3H) = h_{t-1} W_hh^T WITHOUT the bias (b_hh is added here), or NULL at t = 0 (h = 0)
//   gates: (B, 4H) saved for backward: r | z | n | gh_n
__global__ void __launch_bounds__(256) gru_cell_fwd_kernel(const float* __restrict__ gi, long long ldgi, const float* __restrict__ gh,
                                                          const float* __restrict__ b_hh, const float* __restrict__ h_prev,
                                                          const int* __restrict__ len, int t, float* __restrict__ h_out,
                                                          __nv_bfloat16* __restrict__ h_hi, __nv_bfloat16* __restrict__ h_lo,
                                                          long long ldp, float* __restrict__ gates, int B, int H) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;     // float4 index over B * H/4
  const int H4 = H >> 2;
  if (q >= B * H4) return;
  const int b = q / H4, j = (q - b * H4) << 2;
  const float* gib = gi + (long long)b * ldgi;
  const float4 ir = *reinterpret_cast<const float4*>(gib + j), iz = *reinterpret_cast<const float4*>(gib + H + j),
               in_ = *reinterpret_cast<const float4*>(gib + 2 * H + j);
  float4 hr, hz, hn, hp;
  if (gh) {
    const float* ghb = gh + (long long)b * 3 * H;
    hr = *reinterpret_cast<const float4*>(ghb + j); hz = *reinterpret_cast<const float4*>(ghb + H + j); hn = *reinterpret_cast<const float4*>(ghb + 2 * H + j);
    hp = *reinterpret_cast<const float4*>(h_prev + (long long)b * H + j);
  } else {
    hr = hz = hn = hp = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {   // gh excludes the bias (the per-step product is a split-K accumulation): add b_hh here
    const float4 br = __ldg(reinterpret_cast<const float4*>(b_hh + j)), bz = __ldg(reinterpret_cast<const float4*>(b_hh + H + j)),
                 bn = __ldg(reinterpret_cast<const float4*>(b_hh + 2 * H + j));
    hr.x += br.x; hr.y += br.y; hr.z += br.z; hr.w += br.w;
    hz.x += bz.x; hz.y += bz.y; hz.z += bz.z; hz.w += bz.w;
    hn.x += bn.x; hn.y += bn.y; hn.z += bn.z; hn.w += bn.w;
  }
  const bool active = t < len[b];
  const float air[4] = {ir.x, ir.y, ir.z, ir.w}, aiz[4] = {iz.x, iz.y, iz.z, iz.w}, ain[4] = {in_.x, in_.y, in_.z, in_.w};
  const float ahr[4] = {hr.x, hr.y, hr.z, hr.w}, ahz[4] = {hz.x, hz.y, hz.z, hz.w}, ahn[4] = {hn.x, hn.y, hn.z, hn.w};
  const float ahp[4] = {hp.x, hp.y, hp.z, hp.w};
  float r[4], z[4], n[4], h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    r[e] = sigmoidf_(air[e] + ahr[e]);
    z[e] = sigmoidf_(aiz[e] + ahz[e]);
    n[e] = tanhf(ain[e] + r[e] * ahn[e]);
    const float hnew = (1.f - z[e]) * n[e] + z[e] * ahp[e];
    h[e] = active ? hnew : ahp[e];
  }
  *reinterpret_cast<float4*>(h_out + (long long)b * H + j) = make_float4(h[0], h[1], h[2], h[3]);
  store_split4(h_hi + (long long)b * ldp + j, h_lo ? h_lo + (long long)b * ldp + j : nullptr, h);
  float* g = gates + (long long)b * 4 * H;
  *reinterpret_cast<float4*>(g + j) = make_float4(r[0], r[1], r[2], r[3]);
  *reinterpret_cast<float4*>(g + H + j) = make_float4(z[0], z[1], z[2], z[3]);
  *reinterpret_cast<float4*>(g + 2 * H + j) = make_float4(n[0], n[1], n[2], n[3]);
  *reinterpret_cast<float4*>(g + 3 * H + j) = hn;
}

// Backward of one step.  dh: gradient w.r.t. h_t.  Writes dgi (B,3H), dgh (B,3H) as fp32 and split planes, and
// dh_part = the direct part of the gradient w.r.t. h_{t-1} (dh*z when active, dh itself past the sequence end);
// the caller adds dgh W_hh with the GEMM.
__global__ void __launch_bounds__(256) gru_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gates,
                                                          const float* __restrict__ h_prev, const int* __restrict__ len, int t,
                                                          float* __restrict__ dgi, float* __restrict__ dgh,
                                                          __nv_bfloat16* __restrict__ dgi_hi, __nv_bfloat16* __restrict__ dgi_lo,
                                                          __nv_bfloat16* __restrict__ dgh_hi, __nv_bfloat16* __restrict__ dgh_lo,
                                                          long long ldp, float* __restrict__ dh_part, int B, int H) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int H4 = H >> 2;
  if (q >= B * H4) return;
  const int b = q / H4, j = (q - b * H4) << 2;
  const float4 d4 = *reinterpret_cast<const float4*>(dh + (long long)b * H + j);
  const float* g = gates + (long long)b * 4 * H;
  const float4 r4 = *reinterpret_cast<const float4*>(g + j), z4 = *reinterpret_cast<const float4*>(g + H + j),
               n4 = *reinterpret_cast<const float4*>(g + 2 * H + j), m4 = *reinterpret_cast<const float4*>(g + 3 * H + j);
  const float4 p4 = h_prev ? *reinterpret_cast<const float4*>(h_prev + (long long)b * H + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool active = t < len[b];
  const float d[4] = {d4.x, d4.y, d4.z, d4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
  const float n[4] = {n4.x, n4.y, n4.z, n4.w}, ghn[4] = {m4.x, m4.y, m4.z, m4.w}, hp[4] = {p4.x, p4.y, p4.z, p4.w};
  float gr[4], gz[4], gn[4], gm[4], dp[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float dn_pre = d[e] * (1.f - z[e]) * (1.f - n[e] * n[e]);
    const float dz_pre = d[e] * (hp[e] - n[e]) * z[e] * (1.f - z[e]);
    const float dr_pre = dn_pre * ghn[e] * r[e] * (1.f - r[e]);
    gr[e] = active ? dr_pre : 0.f;
    gz[e] = active ? dz_pre : 0.f;
    gn[e] = active ? dn_pre : 0.f;
    gm[e] = active ? dn_pre * r[e] : 0.f;
    dp[e] = active ? d[e] * z[e] : d[e];
  }
  float* gi = dgi + (long long)b * 3 * H;
  float* gh = dgh + (long long)b * 3 * H;
  *reinterpret_cast<float4*>(gi + j) = make_float4(gr[0], gr[1], gr[2], gr[3]);
  *reinterpret_cast<float4*>(gi + H + j) = make_float4(gz[0], gz[1], gz[2], gz[3]);
  *reinterpret_cast<float4*>(gi + 2 * H + j) = make_float4(gn[0], gn[1], gn[2], gn[3]);
  *reinterpret_cast<float4*>(gh + j) = make_float4(gr[0], gr[1], gr[2], gr[3]);
  *reinterpret_cast<float4*>(gh + H + j) = make_float4(gz[0], gz[1], gz[2], gz[3]);
  *reinterpret_cast<float4*>(gh + 2 * H + j) = make_float4(gm[0], gm[1], gm[2], gm[3]);
  const long long po = (long long)b * ldp + j;
  store_split4(dgi_hi + po, dgi_lo ? dgi_lo + po : nullptr, gr);
  store_split4(dgi_hi + po + H, dgi_lo ? dgi_lo + po + H : nullptr, gz);
  store_split4(dgi_hi + po + 2 * H, dgi_lo ? dgi_lo + po + 2 * H : nullptr, gn);
  store_split4(dgh_hi + po, dgh_lo ? dgh_lo + po : nullptr, gr);
  store_split4(dgh_hi + po + H, dgh_lo ? dgh_lo + po + H : nullptr, gz);
  store_split4(dgh_hi + po + 2 * H, dgh_lo ? dgh_lo + po + 2 * H : nullptr, gm);
  *reinterpret_cast<float4*>(dh_part + (long long)b * H + j) = make_float4(dp[0], dp[1], dp[2], dp[3]);
}

}  // namespace vqa
using namespace vqa;

extern "C" int vqa_embed_gather_split(const long long* question, long long ldq, const float* W, long long vocab, int emb,
                                      void* hi, void* lo, long long ldp, int B, int T, int* err, cudaStream_t stream) {
  VQA_CHECK_ARG(question && W && hi && B > 0 && T > 0 && emb > 0 && vocab > 0, "vqa_embed_gather_split: bad arguments");
  VQA_CHECK_ARG((ldp & 7) == 0 && ldp >= emb && aligned16(hi) && (!lo || aligned16(lo)), "vqa_embed_gather_split: planes need ld %% 8 == 0, ld >= emb");
  embed_gather_split_kernel<<<(T * B + 7) / 8, 256, 0, stream>>>(question, ldq, W, emb, vocab, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                reinterpret_cast<__nv_bfloat16*>(lo), ldp, B, T, err);
  VQA_LAUNCH_CHECK("embed_gather_split_kernel");
  return VQA_OK;
}

extern "C" int vqa_embed_scatter_add_f32(const float* dE, long long ldd, const long long* question, long long ldq, const int* len,
                                         float* dW, long long vocab, int emb, int B, int T, cudaStream_t stream) {
  VQA_CHECK_ARG(dE && question && len && dW && B > 0 && T > 0 && emb > 0, "vqa_embed_scatter_add_f32: bad arguments");
  embed_scatter_add_kernel<<<(T * B + 7) / 8, 256, 0, stream>>>(dE, ldd, question, ldq, len, dW, emb, vocab, B, T);
  VQA_LAUNCH_CHECK("embed_scatter_add_kernel");
  return VQA_OK;
}

extern "C" int vqa_gru_cell_fwd_f32(const float* gi, long long ldgi, const float* gh, const float* b_hh, const float* h_prev,
                                    const int* len, int t, float* h_out, void* h_hi, void* h_lo, long long ldp, float* gates,
                                    int B, int H, cudaStream_t stream) {
  VQA_CHECK_ARG(gi && b_hh && len && h_out && h_hi && gates && (gh == nullptr || h_prev), "vqa_gru_cell_fwd_f32: null pointer");
  VQA_CHECK_ARG(B > 0 && H > 0 && (H & 3) == 0 && (ldgi & 3) == 0 && (ldp & 7) == 0, "vqa_gru_cell_fwd_f32: H and leading dimensions must be multiples of 4 (H=%d)", H);
  VQA_CHECK_ARG(aligned16(gi) && aligned16(b_hh) && aligned16(h_out) && aligned16(gates) && (!gh || (aligned16(gh) && aligned16(h_prev))), "vqa_gru_cell_fwd_f32: pointers must be 16-byte aligned");
  const int n = B * (H >> 2);
  gru_cell_fwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gi, ldgi, gh, b_hh, h_prev, len, t, h_out, reinterpret_cast<__nv_bfloat16*>(h_hi),
                                                          reinterpret_cast<__nv_bfloat16*>(h_lo), ldp, gates, B, H);
  VQA_LAUNCH_CHECK("gru_cell_fwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_gru_cell_bwd_f32(const float* dh, const float* gates, const float* h_prev, const int* len, int t, float* dgi,
                                    float* dgh, void* dgi_hi, void* dgi_lo, void* dgh_hi, void* dgh_lo, long long ldp,
                                    float* dh_part, int B, int H, cudaStream_t stream) {
  VQA_CHECK_ARG(dh && gates && len && dgi && dgh && dgi_hi && dgh_hi && dh_part, "vqa_gru_cell_bwd_f32: null pointer");
  VQA_CHECK_ARG(B > 0 && H > 0 && (H & 3) == 0 && (ldp & 7) == 0, "vqa_gru_cell_bwd_f32: H must be a multiple of 4 (H=%d)", H);
  const int n = B * (H >> 2);
  gru_cell_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dh, gates, h_prev, len, t, dgi, dgh, reinterpret_cast<__nv_bfloat16*>(dgi_hi),
                                                          reinterpret_cast<__nv_bfloat16*>(dgi_lo), reinterpret_cast<__nv_bfloat16*>(dgh_hi),
                                                          reinterpret_cast<__nv_bfloat16*>(dgh_lo), ldp, dh_part, B, H);
  VQA_LAUNCH_CHECK("gru_cell_bwd_kernel");
  return VQA_OK;
}
