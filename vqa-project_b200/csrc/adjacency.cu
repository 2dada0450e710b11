// Fused graph-learner tail (sm_100a): per-image adjacency A = h h^T held in shared memory, per-row top-nb
// neighbourhood selection and softmax with warp primitives, and the matching backward.
//
// Replaces layers.py:193-195 (torch.matmul(h, h^T)) and sparse_graph_model.py:225-227 (torch.topk + a Python loop of
// K softmax launches, executed twice per forward with identical inputs).  HBM-bound by design: h is read once,
// A / idx / alpha are written once (A must be materialised because Model.forward returns it).
#include "common.cuh"
#include "../../include/vqa_b200.h"

namespace vqa {

constexpr int ADJ_THREADS = 256;            // backward kernel / stand-alone top-k
constexpr int ADJ_WARPS = ADJ_THREADS / 32;
constexpr int ADJ_MAXW = 12;                // forward kernel: 8..12 warps, chosen so that the upper-triangle tiles divide evenly
constexpr int CH = 64;          // feature columns staged per chunk
constexpr int CHP = CH + 4;     // padded row stride: 68 mod 32 = 4 -> the (row g, column t) fragment loads of mma.m16n8k8 hit 32 distinct banks

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- 3xTF32 on the warp-level tensor-core path (mma.sync.m16n8k8, fp32 accumulate).  A K x K Gram matrix per image is far too
// small for a tcgen05 tile pipeline (TMEM allocation, mbarrier ring, ~130-cycle single-thread issue per MMA: measured in
// graphconv_mma.cu), but it does not belong on the FFMA pipe either: the previous 4x4-register-tile version needed 68 k warp
// instructions per image, half of its shared-memory wavefronts bank-conflicted (profiles/r01c_adjacency_ncu_full.txt, 0.11 of
// the HBM roofline).  mma.sync issues at 0.5 / clk / SM on B200 (profiles/r02_mma_sync_rate.txt) - 512 tf32 MAC / clk / SM, a
// quarter of tcgen05 - which is ~10x what this kernel needs.  x = hi + lo with hi = tf32(x) (round to nearest), lo = x - hi
// (exact in fp32; the tensor core reads its upper 19 bits); lo*hi + hi*lo + hi*hi accumulates to ~1e-7 of fp32 FMA results
// (the adjacency needs fp32-grade products, never bf16).
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// one 16x8 output tile += A(16x8) . B(8x8), three passes, small terms first
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_tf32(a[e], ah[e], al[e]);
#pragma unroll
  for (int e = 0; e < 2; ++e) split_tf32(b[e], bh[e], bl[e]);
  mma_tf32(c, al[0], al[1], al[2], al[3], bh[0], bh[1]);
  mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
  mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
}

// Order-preserving map fp32 -> uint32 (larger float <=> larger key); -0.0 is folded onto +0.0 so that equal floats have equal keys.
__device__ __forceinline__ uint32_t sort_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  return __uint_as_float(u);
}

// Top-nb of each row of the K x K matrix of SORT KEYS in shared memory (row stride KP) + softmax over the selected values.
// rank(j) = #{j' : A[j'] > A[j] or (A[j'] == A[j] and j' < j)}; entry j is selected iff rank < nb and is emitted at slot rank ->
// output in descending-value order, ties to the lower index (the reference's topk(sorted=False) order is unspecified; callers
// compare index SETS).  The K^3 comparisons per image are the bulk of this kernel's instructions, so: every lane owns EPL entries
// of a row and LPR lanes share a row (32 / LPR rows per warp pass - with one row per warp, K = 36 used 36 of 64 entry slots); one
// comparison is a single 64-bit integer compare of (key, 255 - index) pairs + a predicated add instead of two float compares, an
// index compare and their combination.
template <int LPR, int EPL>
__device__ __forceinline__ void topk_softmax_rows_t(const uint32_t* Ak, int KP, int K, int nb, int* __restrict__ idx_out,
                                                    float* __restrict__ alpha_out, int nwarps) {
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = lane % LPR, rsub = lane / LPR;
  for (int row0 = warp * RPW; row0 < K; row0 += nwarps * RPW) {
    const int i = row0 + rsub;
    const bool active = i < K;
    const uint32_t* row = Ak + (active ? i : K - 1) * KP;
    uint32_t kx[EPL], jx[EPL], rank[EPL];
    uint32_t kmax = 0u;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int j = l + LPR * e;
      kx[e] = j < K ? row[j] : 0u;
      kmax = kx[e] > kmax ? kx[e] : kmax;
      jx[e] = (uint32_t)(255 - j);
      rank[e] = 0u;
    }
#pragma unroll 4
    for (int jj = 0; jj < K; ++jj) {
      const uint32_t ky = row[jj], jy = (uint32_t)(255 - jj);
      // rank += ((ky, jy) > (kx, jx)) as 64-bit pairs: the borrow of (kx, jx) - (ky, jy), three instructions per comparison
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\taddc.u32 %0, %0, 0;\n\t}"
            : "+r"(rank[e]) : "r"(jx[e]), "r"(jy), "r"(kx[e]), "r"(ky));
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(0xffffffffu, kmax, o); kmax = t > kmax ? t : kmax; }
    const float mx = key_value(kmax);
    float ex[EPL], s = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int j = l + LPR * e;
      const bool sel = j < K && (int)rank[e] < nb;
      ex[e] = sel ? expf(key_value(kx[e]) - mx) : 0.f;
      s += ex[e];
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (active) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int j = l + LPR * e;
        if (j < K && (int)rank[e] < nb) {
          idx_out[i * nb + (int)rank[e]] = j;
          alpha_out[i * nb + (int)rank[e]] = ex[e] / s;
        }
      }
    }
  }
}
__device__ void topk_softmax_rows(const uint32_t* Ak, int KP, int K, int nb, int* __restrict__ idx_out, float* __restrict__ alpha_out,
                                  int nwarps) {
  if (K <= 32) topk_softmax_rows_t<4, 8>(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
  else if (K <= 40) topk_softmax_rows_t<8, 5>(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
  else if (K <= 64) topk_softmax_rows_t<8, 8>(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
  else if (K <= 112) topk_softmax_rows_t<16, 7>(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
  else topk_softmax_rows_t<16, 8>(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
}

// One CTA per image.  The upper-triangle 16 x 8 tiles of A = h h^T (tile (mi, nj), nj >= 2 mi) are dealt round-robin to the warps,
// MAXT per warp; h streams through shared memory in 64-column chunks (cp.async, double buffered), every warp takes its A / B
// fragments straight from the chunk (both are rows of h: B[k][n] = h[n][k]).  Elements i <= j are mirrored into the K x K matrix
// in shared memory (A symmetric bit for bit), which is written to HBM once and then ranked in place.
template <int MAXT>
__global__ void __launch_bounds__(ADJ_MAXW * 32)
adjacency_topk_fwd_kernel(const float* __restrict__ h, float* __restrict__ adj, int* __restrict__ idx,
                          float* __restrict__ alpha, int K, int C, int nb, int MTl, int NTl, int ntiles) {
  extern __shared__ __align__(16) float sm[];
  const int KR = MTl * 16, KP = K | 1;
  float* hs = sm;                              // [2][KR][CHP]
  float* As = sm + 2 * KR * CHP;               // [K][KP]
  const int b = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* hb = h + (long long)b * K * C;

  for (int v = tid; v < 2 * KR * CHP; v += nthreads) hs[v] = 0.f;   // rows >= K (and a ragged last k-step) must read as zero

  int tmi[MAXT], tnj[MAXT];
  bool tv[MAXT];
  float acc[MAXT][4];
#pragma unroll
  for (int s = 0; s < MAXT; ++s) {
    const int tile = warp + s * nwarps;
    tv[s] = tile < ntiles;
    int rem = tv[s] ? tile : 0, mi = 0;
    while (rem >= NTl - 2 * mi) { rem -= NTl - 2 * mi; ++mi; }
    tmi[s] = mi; tnj[s] = 2 * mi + rem;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[s][e] = 0.f;
  }
  __syncthreads();

  const int nch = (C + CH - 1) / CH;
  auto load_chunk = [&](int ch) {
    const int c0 = ch * CH, cw4 = (min(CH, C - c0)) >> 2;
    float* dst = hs + (ch & 1) * KR * CHP;
    for (int v = tid; v < K * cw4; v += nthreads) {
      const int r = v / cw4, c4 = v - r * cw4;
      cp_async16(dst + r * CHP + c4 * 4, hb + (long long)r * C + c0 + c4 * 4);
    }
    cp_async_commit();
  };
  load_chunk(0);
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) { load_chunk(ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    const int cw = min(CH, C - ch * CH);
    float* wbuf = hs + (ch & 1) * KR * CHP;
    if ((cw & 7) && ch >= 2)                   // ragged last k-step (C % 8 == 4): columns [cw, cw + 4) still hold chunk ch - 2
      for (int r = tid; r < K; r += nthreads) *reinterpret_cast<float4*>(wbuf + r * CHP + cw) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const float* buf = wbuf;
    const int nks = (cw + 7) >> 3;
    for (int ks = 0; ks < nks; ++ks) {
#pragma unroll
      for (int s = 0; s < MAXT; ++s) {
        if (!tv[s]) continue;
        const float* pa = buf + (tmi[s] * 16 + g) * CHP + ks * 8 + t;
        const float* pb = buf + (tnj[s] * 8 + g) * CHP + ks * 8 + t;
        const float a[4] = {pa[0], pa[8 * CHP], pa[4], pa[8 * CHP + 4]};
        const float bb[2] = {pb[0], pb[4]};
        mma_3xtf32(acc[s], a, bb);
      }
    }
    __syncthreads();
  }
  // accumulator fragment: c0,c1 = (row g, columns 2t, 2t+1), c2,c3 = (row g + 8, same columns)
#pragma unroll
  for (int s = 0; s < MAXT; ++s) {
    if (!tv[s]) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = tmi[s] * 16 + g + (e >> 1) * 8, j = tnj[s] * 8 + 2 * t + (e & 1);
      if (i <= j && j < K) { As[i * KP + j] = acc[s][e]; As[j * KP + i] = acc[s][e]; }
    }
  }
  __syncthreads();
  float* ab = adj + (long long)b * K * K;
  uint32_t* Ak = reinterpret_cast<uint32_t*>(As);
  for (int v = tid; v < K * K; v += nthreads) {
    const int i = v / K, j = v - i * K;
    const float a = As[i * KP + j];
    ab[v] = a;
    Ak[i * KP + j] = sort_key(a);              // ranked as integers from here on
  }
  __syncthreads();
  topk_softmax_rows(Ak, KP, K, nb, idx + (long long)b * K * nb, alpha + (long long)b * K * nb, nwarps);
}

__global__ void __launch_bounds__(ADJ_THREADS)
topk_softmax_kernel(const float* __restrict__ adj, int* __restrict__ idx, float* __restrict__ alpha, int K, int nb) {
  extern __shared__ __align__(16) float sm[];
  uint32_t* Ak = reinterpret_cast<uint32_t*>(sm);
  const int KP = K | 1, b = blockIdx.x;
  const float* ab = adj + (long long)b * K * K;
  for (int v = threadIdx.x; v < K * K; v += ADJ_THREADS) { const int i = v / K; Ak[i * KP + (v - i * K)] = sort_key(ab[v]); }
  __syncthreads();
  topk_softmax_rows(Ak, KP, K, nb, idx + (long long)b * K * nb, alpha + (long long)b * K * nb, ADJ_WARPS);
}

// dalpha -> dv (softmax bwd) -> sparse dA in smem -> S = dA + dA^T (+ dadj + dadj^T) -> dh = (S h) * (h > 0)
// The product runs on the same 3xTF32 mma.sync path as the forward: D[i][c] = sum_j S[i][j] h[j][c] with A = S (row-major, split
// once into tf32 hi / lo planes in shared memory), B[k = j][n = c] = h[j][c] straight from the staged chunk, one 8-column n-tile per
// warp and 64-column chunk, MTC m-tiles of accumulators per warp.  Strides: S rows = 4 (mod 32) words, h rows = 8 (mod 32) words ->
// both fragment load patterns touch 32 distinct banks.
constexpr int HBP = CH + 8;      // backward: padded row stride of the staged h chunk (72 = 8 mod 32)

template <int MTC>
__global__ void __launch_bounds__(ADJ_THREADS)
adjacency_topk_bwd_kernel(const float* __restrict__ h, const int* __restrict__ idx, const float* __restrict__ alpha,
                          const float* __restrict__ dalpha, const float* __restrict__ dadj, float* __restrict__ dh,
                          int K, int C, int nb, int SP) {
  extern __shared__ __align__(16) float sm[];
  const int MTl = (K + 15) >> 4, KT = (K + 7) >> 3, SR = MTl * 16, HR = KT * 8;
  float* S = sm;                                            // [SR][SP] fp32, then the lo plane
  uint32_t* Sh = reinterpret_cast<uint32_t*>(sm + SR * SP); // [SR][SP] tf32 hi plane
  float* hs = sm + 2 * SR * SP;                             // [2][HR][HBP]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int v = tid; v < SR * SP; v += ADJ_THREADS) S[v] = 0.f;
  for (int v = tid; v < 2 * HR * HBP; v += ADJ_THREADS) hs[v] = 0.f;     // rows >= K of the staged chunks must read as zero
  __syncthreads();
  for (int i = warp; i < K; i += ADJ_WARPS) {
    const long long base = ((long long)b * K + i) * nb;
    float dot = 0.f;
    for (int m = lane; m < nb; m += 32) dot = fmaf(alpha[base + m], dalpha[base + m], dot);
    dot = warp_sum(dot);
    for (int m = lane; m < nb; m += 32) S[i * SP + idx[base + m]] = alpha[base + m] * (dalpha[base + m] - dot);
  }
  __syncthreads();
  if (dadj) {
    const float* db = dadj + (long long)b * K * K;
    for (int v = tid; v < K * K; v += ADJ_THREADS) { const int i = v / K; S[i * SP + (v - i * K)] += db[v]; }
    __syncthreads();
  }
  for (int v = tid; v < K * K; v += ADJ_THREADS) {      // symmetrise in place: the pair (i,j), i<=j has one owner
    const int i = v / K, j = v - i * K;
    if (i <= j) { const float s = S[i * SP + j] + S[j * SP + i]; S[i * SP + j] = s; S[j * SP + i] = s; }
  }
  __syncthreads();
  for (int v = tid; v < SR * SP; v += ADJ_THREADS) {    // hi / lo planes of S, once per image
    uint32_t hi, lo;
    split_tf32(S[v], hi, lo);
    Sh[v] = hi;
    S[v] = __uint_as_float(lo);
  }

  const float* hb = h + (long long)b * K * C;
  float* ob = dh + (long long)b * K * C;
  const int nch = (C + CH - 1) / CH;
  auto load_chunk = [&](int ch) {
    const int c0 = ch * CH, cw4 = (min(CH, C - c0)) >> 2;
    float* dst = hs + (ch & 1) * HR * HBP;
    for (int v = tid; v < K * cw4; v += ADJ_THREADS) {
      const int r = v / cw4, c4 = v - r * cw4;
      cp_async16(dst + r * HBP + c4 * 4, hb + (long long)r * C + c0 + c4 * 4);
    }
    cp_async_commit();
  };
  load_chunk(0);
  const uint32_t* Sl = reinterpret_cast<const uint32_t*>(S);
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) { load_chunk(ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                                     // chunk ch landed (and, first time round, the S planes are complete)
    const float* buf = hs + (ch & 1) * HR * HBP;
    const int cw = min(CH, C - ch * CH);
    const int n0 = warp * 8;                             // my n-tile of this chunk (8 warps x 8 columns = CH)
    if (n0 < cw) {                                       // warp-uniform
      float acc[MTC][4];
#pragma unroll
      for (int mi = 0; mi < MTC; ++mi)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mi][e] = 0.f;
      for (int kt = 0; kt < KT; ++kt) {
        const float* pb = buf + (kt * 8 + t) * HBP + n0 + g;
        uint32_t bh[2], bl[2];
        split_tf32(pb[0], bh[0], bl[0]);
        split_tf32(pb[4 * HBP], bh[1], bl[1]);
#pragma unroll
        for (int mi = 0; mi < MTC; ++mi) {
          if (mi < MTl) {
            const int o = (mi * 16 + g) * SP + kt * 8 + t;
            const uint32_t ah0 = Sh[o], ah1 = Sh[o + 8 * SP], ah2 = Sh[o + 4], ah3 = Sh[o + 8 * SP + 4];
            const uint32_t al0 = Sl[o], al1 = Sl[o + 8 * SP], al2 = Sl[o + 4], al3 = Sl[o + 8 * SP + 4];
            mma_tf32(acc[mi], al0, al1, al2, al3, bh[0], bh[1]);
            mma_tf32(acc[mi], ah0, ah1, ah2, ah3, bl[0], bl[1]);
            mma_tf32(acc[mi], ah0, ah1, ah2, ah3, bh[0], bh[1]);
          }
        }
      }
      const int col = n0 + 2 * t;                        // columns col, col + 1 of the chunk (cw % 4 == 0: both valid or both not)
      if (col < cw) {
#pragma unroll
        for (int mi = 0; mi < MTC; ++mi) {
          if (mi < MTl) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int i = mi * 16 + g + hf * 8;
              if (i < K) {
                const float2 hv = *reinterpret_cast<const float2*>(buf + i * HBP + col);   // ReLU mask of the forward: h > 0
                float2 o;
                o.x = hv.x > 0.f ? acc[mi][2 * hf] : 0.f;
                o.y = hv.y > 0.f ? acc[mi][2 * hf + 1] : 0.f;
                *reinterpret_cast<float2*>(ob + (long long)i * C + ch * CH + col) = o;
              }
            }
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
  const int MT = (K + 15) / 16, NT = (K + 7) / 8;
  int ntiles = 0;
  for (int mi = 0; mi < MT; ++mi) ntiles += NT - 2 * mi;
  int nw = 8, best = 1 << 30;                               // fewest idle tile slots, then fewest warps
  for (int w = 8; w <= ADJ_MAXW; ++w) {
    const int waste = (ntiles + w - 1) / w * w - ntiles;
    if (waste < best) { best = waste; nw = w; }
  }
  const int maxt = (ntiles + nw - 1) / nw;
  const size_t smem = (size_t)(2 * MT * 16 * CHP + K * (K | 1)) * sizeof(float);
  auto run = [&](auto kern) -> int {
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, nw * 32, smem, stream>>>(h, adjacency, idx, alpha, K, C, nb, MT, NT, ntiles);
    VQA_LAUNCH_CHECK("adjacency_topk_fwd_kernel");
    return VQA_OK;
  };
  if (maxt == 1) return run(adjacency_topk_fwd_kernel<1>);
  if (maxt == 2) return run(adjacency_topk_fwd_kernel<2>);
  if (maxt <= 4) return run(adjacency_topk_fwd_kernel<4>);
  if (maxt <= 6) return run(adjacency_topk_fwd_kernel<6>);
  return run(adjacency_topk_fwd_kernel<9>);
}

extern "C" int vqa_topk_softmax_f32(const float* adjacency, int* idx, float* alpha, int B, int K, int nb, cudaStream_t stream) {
  VQA_CHECK_ARG(adjacency && idx && alpha, "vqa_topk_softmax_f32: null pointer");
  if (int rc = adj_check(B, K, 4, nb, "vqa_topk_softmax_f32")) return rc;
  const size_t smem = (size_t)K * (K | 1) * sizeof(float);
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
  const int MT = (K + 15) / 16, KT = (K + 7) / 8;
  int SP = KT * 8;                                           // row stride of the S planes: >= 8 KT words and = 4 (mod 32)
  SP += (4 - SP % 32 + 32) % 32;
  const size_t smem = (size_t)(2 * MT * 16 * SP + 2 * KT * 8 * HBP) * sizeof(float);
  auto run = [&](auto kern) -> int {
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, ADJ_THREADS, smem, stream>>>(h, idx, alpha, dalpha, dadj, dh, K, C, nb, SP);
    VQA_LAUNCH_CHECK("adjacency_topk_bwd_kernel");
    return VQA_OK;
  };
  if (MT <= 3) return run(adjacency_topk_bwd_kernel<3>);
  if (MT <= 4) return run(adjacency_topk_bwd_kernel<4>);
  return run(adjacency_topk_bwd_kernel<8>);
}
