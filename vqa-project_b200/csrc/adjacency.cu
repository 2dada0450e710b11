// Fused graph-learner tail (sm_100a): per-image adjacency A = h h^T held in shared memory, per-row top-nb
// neighbourhood selection and softmax with warp primitives, and the matching backward.
//
// Replaces layers.py:193-195 (torch.matmul(h, h^T)) and sparse_graph_model.py:225-227 (torch.topk + a Python loop of
// K softmax launches, executed twice per forward with identical inputs).  HBM-bound by design: h is read once,
// A / idx / alpha are written once (A must be materialised because Model.forward returns it).
#include "common.cuh"
#include "../../include/vqa_b200.h"

namespace vqa {

constexpr int ADJ_THREADS = 256;
constexpr int ADJ_WARPS = ADJ_THREADS / 32;
constexpr int CH = 64;          // feature columns staged per chunk
constexpr int CHP = CH + 4;     // padded row stride: 68 mod 32 = 4 -> conflict-free LDS.128 across rows

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Top-nb of each row of the K x K matrix in shared memory (row stride KP) + softmax over the selected values.
// One warp per row, <= 4 entries per lane (K <= 128).  rank(j) = #{j' : A[j'] > A[j] or (A[j'] == A[j] and j' < j)};
// entry j is selected iff rank < nb and is emitted at slot rank -> output is in descending-value order, ties go to
// the lower index.  The reference's topk(sorted=False) order is unspecified; callers compare index SETS.
// NE = ceil(K / 32) entries per lane: the rank loop costs K * NE compare-and-count steps per row, so it is instantiated per NE
template <int NE>
__device__ __forceinline__ void topk_softmax_rows_ne(const float* As, int KP, int K, int nb, int* __restrict__ idx_out,
                                                     float* __restrict__ alpha_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < K; i += ADJ_WARPS) {
    const float* row = As + i * KP;
    float x[NE];
    int rank[NE];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int j = lane + 32 * e;
      x[e] = j < K ? row[j] : -INFINITY;
      rank[e] = 0;
      mx = fmaxf(mx, x[e]);
    }
    mx = warp_max(mx);
#pragma unroll 4
    for (int jj = 0; jj < K; ++jj) {
      const float y = row[jj];
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int j = lane + 32 * e;
        rank[e] += (y > x[e] || (y == x[e] && jj < j)) ? 1 : 0;
      }
    }
    float ex[NE], s = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int j = lane + 32 * e;
      const bool sel = j < K && rank[e] < nb;
      ex[e] = sel ? expf(x[e] - mx) : 0.f;
      s += ex[e];
    }
    s = warp_sum(s);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int j = lane + 32 * e;
      if (j < K && rank[e] < nb) {
        idx_out[i * nb + rank[e]] = j;
        alpha_out[i * nb + rank[e]] = ex[e] / s;
      }
    }
  }
}
__device__ void topk_softmax_rows(const float* As, int KP, int K, int nb, int* __restrict__ idx_out,
                                  float* __restrict__ alpha_out) {
  if (K <= 32) topk_softmax_rows_ne<1>(As, KP, K, nb, idx_out, alpha_out);
  else if (K <= 64) topk_softmax_rows_ne<2>(As, KP, K, nb, idx_out, alpha_out);
  else if (K <= 96) topk_softmax_rows_ne<3>(As, KP, K, nb, idx_out, alpha_out);
  else topk_softmax_rows_ne<4>(As, KP, K, nb, idx_out, alpha_out);
}

// MAXT: upper-triangle 4x4 tiles owned per thread (1 for K <= 64 incl. the split-C groups, up to 3 for K <= 128)
template <int MAXT>
__global__ void __launch_bounds__(ADJ_THREADS)
adjacency_topk_fwd_kernel(const float* __restrict__ h, float* __restrict__ adj, int* __restrict__ idx,
                          float* __restrict__ alpha, int K, int C, int nb, int G) {
  extern __shared__ __align__(16) float sm[];
  const int nt = (K + 3) >> 2, K4 = nt * 4, KP = K + 1;
  const int ntiles = nt * (nt + 1) / 2;
  float* hs = sm;                              // [2][K4][CHP]
  float* red = sm + 2 * K4 * CHP;              // [G][K][KP]; red[0] becomes A
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* hb = h + (long long)b * K * C;

  for (int v = tid; v < 2 * K4 * CHP; v += ADJ_THREADS) hs[v] = 0.f;   // padding rows must be finite

  int ti[MAXT], tj[MAXT], g = 0;
  bool valid[MAXT];
#pragma unroll
  for (int tt = 0; tt < MAXT; ++tt) {
    int tile;
    if (MAXT == 1) { tile = tid % ntiles; g = tid / ntiles; valid[tt] = g < G; }
    else { tile = tid + tt * ADJ_THREADS; valid[tt] = tile < ntiles; }
    int rem = valid[tt] ? tile : 0, a = 0;
    while (rem >= nt - a) { rem -= nt - a; ++a; }
    ti[tt] = a; tj[tt] = a + rem;
  }
  float acc[MAXT][4][4];
#pragma unroll
  for (int tt = 0; tt < MAXT; ++tt)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) acc[tt][r][s] = 0.f;
  __syncthreads();

  const int nch = (C + CH - 1) / CH;
  auto load_chunk = [&](int ch) {
    const int c0 = ch * CH, cw4 = (min(CH, C - c0)) >> 2;
    float* dst = hs + (ch & 1) * K4 * CHP;
    for (int v = tid; v < K * cw4; v += ADJ_THREADS) {
      const int r = v / cw4, c4 = v - r * cw4;
      cp_async16(dst + r * CHP + c4 * 4, hb + (long long)r * C + c0 + c4 * 4);
    }
    cp_async_commit();
  };
  load_chunk(0);
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) { load_chunk(ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* buf = hs + (ch & 1) * K4 * CHP;
    const int cw4 = (min(CH, C - ch * CH)) >> 2;
#pragma unroll
    for (int tt = 0; tt < MAXT; ++tt) {
      if (!valid[tt]) continue;
      const float* pa = buf + ti[tt] * 4 * CHP;
      const float* pb = buf + tj[tt] * 4 * CHP;
      for (int c4 = g; c4 < cw4; c4 += G) {
        float4 a[4], bb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a[r] = *reinterpret_cast<const float4*>(pa + r * CHP + c4 * 4);
          bb[r] = *reinterpret_cast<const float4*>(pb + r * CHP + c4 * 4);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            float t = acc[tt][r][s];
            t = fmaf(a[r].x, bb[s].x, t); t = fmaf(a[r].y, bb[s].y, t);
            t = fmaf(a[r].z, bb[s].z, t); t = fmaf(a[r].w, bb[s].w, t);
            acc[tt][r][s] = t;
          }
      }
    }
    __syncthreads();
  }
  // partial tiles -> red[g] (both triangles: A is symmetric by construction)
#pragma unroll
  for (int tt = 0; tt < MAXT; ++tt) {
    if (!valid[tt]) continue;
    float* rg = red + g * K * KP;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int i = ti[tt] * 4 + r, j = tj[tt] * 4 + s;
        if (i < K && j < K) { rg[i * KP + j] = acc[tt][r][s]; rg[j * KP + i] = acc[tt][r][s]; }
      }
  }
  __syncthreads();
  float* ab = adj + (long long)b * K * K;
  for (int v = tid; v < K * K; v += ADJ_THREADS) {
    const int i = v / K, j = v - i * K;
    float s = red[i * KP + j];
    for (int gg = 1; gg < G; ++gg) s += red[gg * K * KP + i * KP + j];   // fixed order -> deterministic
    red[i * KP + j] = s;
    ab[v] = s;
  }
  __syncthreads();
  topk_softmax_rows(red, KP, K, nb, idx + (long long)b * K * nb, alpha + (long long)b * K * nb);
}

__global__ void __launch_bounds__(ADJ_THREADS)
topk_softmax_kernel(const float* __restrict__ adj, int* __restrict__ idx, float* __restrict__ alpha, int K, int nb) {
  extern __shared__ __align__(16) float sm[];
  const int KP = K + 1, b = blockIdx.x;
  const float* ab = adj + (long long)b * K * K;
  for (int v = threadIdx.x; v < K * K; v += ADJ_THREADS) { const int i = v / K; sm[i * KP + (v - i * K)] = ab[v]; }
  __syncthreads();
  topk_softmax_rows(sm, KP, K, nb, idx + (long long)b * K * nb, alpha + (long long)b * K * nb);
}

// dalpha -> dv (softmax bwd) -> sparse dA in smem -> S = dA + dA^T (+ dadj + dadj^T) -> dh = (S h) * (h > 0)
__global__ void __launch_bounds__(ADJ_THREADS)
adjacency_topk_bwd_kernel(const float* __restrict__ h, const int* __restrict__ idx, const float* __restrict__ alpha,
                          const float* __restrict__ dalpha, const float* __restrict__ dadj, float* __restrict__ dh,
                          int K, int C, int nb, int CW) {
  extern __shared__ __align__(16) float sm[];
  const int KP = (K + 3) & ~3, K4 = KP;
  float* S = sm;                    // [K4][KP]
  float* hs = sm + K4 * KP;         // [K][CW]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int v = tid; v < K4 * KP; v += ADJ_THREADS) S[v] = 0.f;
  __syncthreads();
  for (int i = warp; i < K; i += ADJ_WARPS) {
    const long long base = ((long long)b * K + i) * nb;
    float dot = 0.f;
    for (int m = lane; m < nb; m += 32) dot = fmaf(alpha[base + m], dalpha[base + m], dot);
    dot = warp_sum(dot);
    for (int m = lane; m < nb; m += 32) S[i * KP + idx[base + m]] = alpha[base + m] * (dalpha[base + m] - dot);
  }
  __syncthreads();
  if (dadj) {
    const float* db = dadj + (long long)b * K * K;
    for (int v = tid; v < K * K; v += ADJ_THREADS) { const int i = v / K; S[i * KP + (v - i * K)] += db[v]; }
    __syncthreads();
  }
  for (int v = tid; v < K * K; v += ADJ_THREADS) {      // symmetrise in place: the pair (i,j), i<=j has one owner
    const int i = v / K, j = v - i * K;
    if (i <= j) { const float s = S[i * KP + j] + S[j * KP + i]; S[i * KP + j] = s; S[j * KP + i] = s; }
  }
  __syncthreads();

  const float* hb = h + (long long)b * K * C;
  float* ob = dh + (long long)b * K * C;
  const int ngroups = K4 >> 2;
  for (int c0 = 0; c0 < C; c0 += CW) {
    const int cw = min(CW, C - c0), cw4 = cw >> 2;
    for (int v = tid; v < K * cw4; v += ADJ_THREADS) {
      const int r = v / cw4, c4 = v - r * cw4;
      cp_async16(hs + r * CW + c4 * 4, hb + (long long)r * C + c0 + c4 * 4);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int ncb = (cw + 127) >> 7;
    for (int item = warp; item < ngroups * ncb; item += ADJ_WARPS) {
      const int rg = (item % ngroups) * 4, col = (item / ngroups) * 128 + lane * 4;
      if (col < cw) {
        float4 acc[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < K; ++j) {
          const float4 hv = *reinterpret_cast<const float4*>(hs + j * CW + col);
          const float4 sv = *reinterpret_cast<const float4*>(S + j * KP + rg);   // S symmetric: S[j][rg..rg+3]
          acc[0].x = fmaf(sv.x, hv.x, acc[0].x); acc[0].y = fmaf(sv.x, hv.y, acc[0].y); acc[0].z = fmaf(sv.x, hv.z, acc[0].z); acc[0].w = fmaf(sv.x, hv.w, acc[0].w);
          acc[1].x = fmaf(sv.y, hv.x, acc[1].x); acc[1].y = fmaf(sv.y, hv.y, acc[1].y); acc[1].z = fmaf(sv.y, hv.z, acc[1].z); acc[1].w = fmaf(sv.y, hv.w, acc[1].w);
          acc[2].x = fmaf(sv.z, hv.x, acc[2].x); acc[2].y = fmaf(sv.z, hv.y, acc[2].y); acc[2].z = fmaf(sv.z, hv.z, acc[2].z); acc[2].w = fmaf(sv.z, hv.w, acc[2].w);
          acc[3].x = fmaf(sv.w, hv.x, acc[3].x); acc[3].y = fmaf(sv.w, hv.y, acc[3].y); acc[3].z = fmaf(sv.w, hv.z, acc[3].z); acc[3].w = fmaf(sv.w, hv.w, acc[3].w);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = rg + r;
          if (i < K) {
            const float4 hv = *reinterpret_cast<const float4*>(hs + i * CW + col);
            float4 o;
            o.x = hv.x > 0.f ? acc[r].x : 0.f; o.y = hv.y > 0.f ? acc[r].y : 0.f;
            o.z = hv.z > 0.f ? acc[r].z : 0.f; o.w = hv.w > 0.f ? acc[r].w : 0.f;
            *reinterpret_cast<float4*>(ob + (long long)i * C + c0 + col) = o;
          }
        }
      }
    }
    __syncthreads();
  }
}

static int adj_check(int B, int K, int C, int nb, const char* who) {
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128, "%s: need 0 < K <= 128 (got B=%d K=%d)", who, B, K);
  VQA_CHECK_ARG(nb > 0 && nb <= K, "%s: neighbourhood size must be in [1, K] (nb=%d K=%d)", who, nb, K);
  VQA_CHECK_ARG(C > 0 && (C & 3) == 0, "%s: feature dim must be a positive multiple of 4 (C=%d)", who, C);
  return VQA_OK;
}

}  // namespace vqa
using namespace vqa;

extern "C" int vqa_adjacency_topk_fwd_f32(const float* h, float* adjacency, int* idx, float* alpha, int B, int K, int C,
                                          int nb, cudaStream_t stream) {
  VQA_CHECK_ARG(h && adjacency && idx && alpha, "vqa_adjacency_topk_fwd_f32: null pointer");
  if (int rc = adj_check(B, K, C, nb, "vqa_adjacency_topk_fwd_f32")) return rc;
  VQA_CHECK_ARG(aligned16(h), "vqa_adjacency_topk_fwd_f32: h must be 16-byte aligned");
  const int nt = (K + 3) / 4, K4 = nt * 4, ntiles = nt * (nt + 1) / 2;
  int G = 1, maxt = (ntiles + ADJ_THREADS - 1) / ADJ_THREADS;
  if (maxt == 1) { G = ADJ_THREADS / ntiles; if (G > 8) G = 8; if (G < 1) G = 1; }
  const size_t smem = (size_t)(2 * K4 * CHP + G * K * (K + 1)) * sizeof(float);
  auto run = [&](auto kern) -> int {
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, ADJ_THREADS, smem, stream>>>(h, adjacency, idx, alpha, K, C, nb, G);
    VQA_LAUNCH_CHECK("adjacency_topk_fwd_kernel");
    return VQA_OK;
  };
  if (maxt == 1) return run(adjacency_topk_fwd_kernel<1>);
  if (maxt == 2) return run(adjacency_topk_fwd_kernel<2>);
  return run(adjacency_topk_fwd_kernel<3>);
}

extern "C" int vqa_topk_softmax_f32(const float* adjacency, int* idx, float* alpha, int B, int K, int nb, cudaStream_t stream) {
  VQA_CHECK_ARG(adjacency && idx && alpha, "vqa_topk_softmax_f32: null pointer");
  if (int rc = adj_check(B, K, 4, nb, "vqa_topk_softmax_f32")) return rc;
  const size_t smem = (size_t)K * (K + 1) * sizeof(float);
  VQA_CUDA(cudaFuncSetAttribute(topk_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_softmax_kernel<<<B, ADJ_THREADS, smem, stream>>>(adjacency, idx, alpha, K, nb);
  VQA_LAUNCH_CHECK("topk_softmax_kernel");
  return VQA_OK;
}

extern "C" int vqa_adjacency_topk_bwd_f32(const float* h, const int* idx, const float* alpha, const float* dalpha,
                                          const float* dadj, float* dh, int B, int K, int C, int nb, cudaStream_t stream) {
  VQA_CHECK_ARG(h && idx && alpha && dalpha && dh, "vqa_adjacency_topk_bwd_f32: null pointer");
  if (int rc = adj_check(B, K, C, nb, "vqa_adjacency_topk_bwd_f32")) return rc;
  VQA_CHECK_ARG(aligned16(h) && aligned16(dh), "vqa_adjacency_topk_bwd_f32: h/dh must be 16-byte aligned");
  const int KP = (K + 3) & ~3;
  int CW = (int)((96 * 1024) / (K * 4) / 128) * 128;
  if (CW < 128) CW = 128;
  if (CW > ((C + 127) & ~127)) CW = (C + 127) & ~127;
  const size_t smem = (size_t)(KP * KP + K * CW) * sizeof(float);
  VQA_CUDA(cudaFuncSetAttribute(adjacency_topk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  adjacency_topk_bwd_kernel<<<B, ADJ_THREADS, smem, stream>>>(h, idx, alpha, dalpha, dadj, dh, K, C, nb, CW);
  VQA_LAUNCH_CHECK("adjacency_topk_bwd_kernel");
  return VQA_OK;
}
