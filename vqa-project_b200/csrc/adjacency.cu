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
constexpr int FW_NBUF = 4;      // K > 64 forward: chunks in flight
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
// quarter of tcgen05 - which is ~10x what this kernel needs.  x = hi + lo with hi = x truncated to tf32 (one LOP3; cvt.rna.tf32 is
// emulated with four instructions on sm_100a - measured in the first version of this kernel), lo = x - hi (exact in fp32; the
// tensor core reads its upper 19 bits); lo*hi + hi*lo + hi*hi reproduces fp32 FMA results to ~1e-6 relative
// (the adjacency needs fp32-grade products, never bf16).
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// Stage columns [c0, c0 + cw) of the K rows of one image (row stride C floats) into shared memory rows of `stride` floats.
__device__ __forceinline__ void stage_chunk(float* dst, int stride, const float* src, int C, int K, int c0, int cw, int tid, int nthreads) {
  if (cw == CH) {                              // 16 sixteen-byte pieces per row: shifts, no divisions
    for (int v = tid; v < K * 16; v += nthreads) {
      const int r = v >> 4, c4 = v & 15;
      cp_async16(dst + r * stride + c4 * 4, src + (long long)r * C + c0 + c4 * 4);
    }
  } else {
    const int cw4 = cw >> 2;
    for (int v = tid; v < K * cw4; v += nthreads) {
      const int r = v / cw4, c4 = v - r * cw4;
      cp_async16(dst + r * stride + c4 * 4, src + (long long)r * C + c0 + c4 * 4);
    }
  }
  cp_async_commit();
}
// The same for a whole 64-column chunk with a compile-time row bound: thread (r0 = tid / 16, piece = tid % 16) copies rows r0, r0 + 16,
// ... - pointer increments only, no index arithmetic (the generic loop above costs ~25 instructions per 16-byte piece).
template <int ROWS16>
__device__ __forceinline__ void stage_chunk_t(float* dst, int stride, const float* src, int C, int K, int c0, int tid) {
  const int r0 = tid >> 4, c4 = tid & 15;
  const float* s = src + (long long)r0 * C + c0 + c4 * 4;
  float* d = dst + r0 * stride + c4 * 4;
#pragma unroll
  for (int m = 0; m < ROWS16; ++m) {
    if (r0 + 16 * m < K) cp_async16(d + m * 16 * stride, s + (long long)m * 16 * C);
  }
  cp_async_commit();
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
      // nrank -= ((ky, jy) > (kx, jx)) as 64-bit pairs = the borrow of (kx, jx) - (ky, jy): a three-instruction borrow chain that ptxas
      // folds to 2.5 (one IADD3.X takes the borrows of two comparisons).  The chain stays inside the sub family: CC.CF after sub.cc
      // is the hardware carry (NOT borrow) as far as a following addc is concerned - measured, tools/micro/cc_polarity.cu.
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %1, %2;\n\tsubc.cc.u32 t, %3, %4;\n\tsubc.u32 %0, %0, 0;\n\t}"
            : "+r"(rank[e]) : "r"(jx[e]), "r"(jy), "r"(kx[e]), "r"(ky));
    }
#pragma unroll
    for (int e = 0; e < EPL; ++e) rank[e] = 0u - rank[e];          // the loop counted downwards
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(0xffffffffu, kmax, o); kmax = t > kmax ? t : kmax; }
    const float mx = key_value(kmax);
    float ex[EPL], s = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int j = l + LPR * e;
      const bool sel = j < K && (int)rank[e] < nb;
      ex[e] = sel ? ex2_fast((key_value(kx[e]) - mx) * 1.4426950408889634f) : 0.f;   // exp(x - max), |rel err| ~ 2^-22
      s += ex[e];
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s = 1.f / s;                               // s >= 1 (the maximum contributes exp(0))
    if (active) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int j = l + LPR * e;
        if (j < K && (int)rank[e] < nb) {
          idx_out[i * nb + (int)rank[e]] = j;
          alpha_out[i * nb + (int)rank[e]] = ex[e] * s;
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

// Tail shared by both forward kernels: A (fp32, complete, symmetric) in shared memory -> HBM once, -> sort keys in place -> ranking.
__device__ __forceinline__ void adjacency_finish(float* As, int KP, int K, int nb, float* __restrict__ ab, int* __restrict__ idx_out,
                                                 float* __restrict__ alpha_out, int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* Ak = reinterpret_cast<uint32_t*>(As);
  for (int i = warp; i < K; i += nwarps)
    for (int j = lane; j < K; j += 32) {
      const float a = As[i * KP + j];
      ab[i * K + j] = a;
      Ak[i * KP + j] = sort_key(a);            // ranked as integers from here on
    }
  __syncthreads();
  topk_softmax_rows(Ak, KP, K, nb, idx_out, alpha_out, nwarps);
}

// K <= 64 (MT <= 4 m-tiles): one CTA of 8 warps per image, split-K.  h streams through shared memory in 64-column chunks
// (cp.async, double buffered); warp w owns k-step w of every chunk and computes ALL upper-triangle 16 x 8 tiles (mi, nj >= 2 mi)
// for it from 2 MT row-group fragments loaded and split once (A and B fragments are both rows of h: B[k][n] = h[n][k], and the
// A fragment of m-tile mi is the pair of row-group fragments 2 mi, 2 mi + 1).  The eight partial sums are then added into the
// K x K matrix in shared memory in warp order (deterministic), elements i <= j mirrored (A symmetric bit for bit).
template <int MT, int NBUF>
__global__ void __launch_bounds__(ADJ_THREADS)
adjacency_topk_fwd_sk_kernel(const float* __restrict__ h, float* __restrict__ adj, int* __restrict__ idx,
                             float* __restrict__ alpha, int K, int C, int nb) {
  extern __shared__ __align__(16) float sm[];
  constexpr int KR = MT * 16, NF = 2 * MT;
  const int KP = K | 1, NT = (K + 7) >> 3;
  float* hs = sm;                              // [NBUF][KR][CHP]: NBUF chunks in flight (the whole image at C = 512 when NBUF = 8) -
  float* As = sm + NBUF * KR * CHP;            // [K][KP]   with one 9 KB chunk in flight per CTA the kernel was HBM-LATENCY bound
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* hb = h + (long long)b * K * C;
  const int nch = (C + CH - 1) / CH;
  auto stage_sk = [&](float* dst, int c) {     // 256 threads = 16 rows x 16 pieces per sweep
    if (C - c * CH >= CH) stage_chunk_t<MT>(dst, CHP, hb, C, K, c * CH, tid);
    else stage_chunk(dst, CHP, hb, C, K, c * CH, C - c * CH, tid, ADJ_THREADS);
  };
  for (int c = 0; c < NBUF - 1; ++c) {         // prologue: NBUF - 1 chunks on their way (empty groups keep the group count uniform)
    if (c < nch) stage_sk(hs + c * KR * CHP, c);
    else cp_async_commit();
  }
  for (int c = 0; c < NBUF; ++c)               // rows >= K of every buffer read as zero
    for (int v = tid; v < (KR - K) * CHP; v += ADJ_THREADS) hs[(c * KR + K) * CHP + v] = 0.f;

  float acc[MT][NF][4];                        // tile (mi, nj) lives in acc[mi][nj]; slots nj < 2 mi are never touched (and cost nothing)
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int nj = 0; nj < NF; ++nj)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][nj][e] = 0.f;

  for (int ch = 0; ch < nch; ++ch) {
    const int nx = ch + NBUF - 1;              // its buffer held chunk ch - 1, which every warp left behind at the barrier below
    if (nx < nch) stage_sk(hs + (nx % NBUF) * KR * CHP, nx);
    else cp_async_commit();
    cp_async_wait<NBUF - 1>();                 // all but the NBUF - 1 youngest groups have landed: chunk ch is here
    __syncthreads();
    const float* buf = hs + (ch % NBUF) * KR * CHP;
    const int cw = min(CH, C - ch * CH);
    if (warp * 8 < cw) {                        // my k-step of this chunk (a ragged last k-step, C % 8 == 4, reads 4 valid columns)
      const bool half = warp * 8 + 4 >= cw;     // ... and nothing beyond them
      uint32_t fh[NF][2], fl[NF][2];
#pragma unroll
      for (int r = 0; r < NF; ++r) {
        const float* pr = buf + (r * 8 + g) * CHP + warp * 8 + t;
        split_tf32(pr[0], fh[r][0], fl[r][0]);
        split_tf32(half ? 0.f : pr[4], fh[r][1], fl[r][1]);
      }
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int nj = 2 * mi; nj < NF; ++nj) {
          if (nj < NT) {                        // warp-uniform: the last column tile may be all padding
            mma_tf32(acc[mi][nj], fl[2 * mi][0], fl[2 * mi + 1][0], fl[2 * mi][1], fl[2 * mi + 1][1], fh[nj][0], fh[nj][1]);
            mma_tf32(acc[mi][nj], fh[2 * mi][0], fh[2 * mi + 1][0], fh[2 * mi][1], fh[2 * mi + 1][1], fl[nj][0], fl[nj][1]);
            mma_tf32(acc[mi][nj], fh[2 * mi][0], fh[2 * mi + 1][0], fh[2 * mi][1], fh[2 * mi + 1][1], fh[nj][0], fh[nj][1]);
          }
        }
    }
    __syncthreads();
  }
  // Fixed-order tree reduction of the eight k-slices (deterministic): warps [h, 2h) park their fragments in shared memory (the chunk
  // buffers are idle now), warps [0, h) add them, h = 4, 2, 1 - every step parallel inside the warps that take part.  (Adding the
  // slices into As one warp after the other was 54 % of the kernel's time: a chain of dependent shared-memory read-modify-writes.)
  {
    constexpr int TS = MT * (MT + 1);          // tile slots (mi, nj >= 2 mi)
    float4* red = reinterpret_cast<float4*>(hs);
#pragma unroll
    for (int hw = 4; hw >= 1; hw >>= 1) {
      if (warp >= hw && warp < 2 * hw) {
        int slot = 0;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int nj = 2 * mi; nj < NF; ++nj, ++slot)
            red[((warp - hw) * TS + slot) * 32 + lane] = make_float4(acc[mi][nj][0], acc[mi][nj][1], acc[mi][nj][2], acc[mi][nj][3]);
      }
      __syncthreads();
      if (warp < hw) {
        int slot = 0;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
          for (int nj = 2 * mi; nj < NF; ++nj, ++slot) {
            const float4 v = red[(warp * TS + slot) * 32 + lane];
            acc[mi][nj][0] += v.x; acc[mi][nj][1] += v.y; acc[mi][nj][2] += v.z; acc[mi][nj][3] += v.w;
          }
      }
      __syncthreads();
    }
  }
  if (warp == 0) {
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
      for (int nj = 2 * mi; nj < NF; ++nj)
#pragma unroll
        for (int e = 0; e < 4; ++e) {          // c0,c1 = (row g, columns 2t, 2t+1), c2,c3 = (row g + 8, same columns)
          const int i = mi * 16 + g + (e >> 1) * 8, j = nj * 8 + 2 * t + (e & 1);
          if (i <= j && j < K) { As[i * KP + j] = acc[mi][nj][e]; As[j * KP + i] = acc[mi][nj][e]; }
        }
  }
  __syncthreads();
  adjacency_finish(As, KP, K, nb, adj + (long long)b * K * K, idx + (long long)b * K * nb, alpha + (long long)b * K * nb, ADJ_WARPS);
}

// K > 64: the 36 .. 72 upper-triangle tiles do not fit one warp's registers; they are dealt round-robin to 8 .. 12 warps, MAXT per
// warp, each warp walks all k-steps for its tiles with fragments taken straight from the staged chunk.
template <int MAXT>
__global__ void __launch_bounds__(ADJ_MAXW * 32)
adjacency_topk_fwd_kernel(const float* __restrict__ h, float* __restrict__ adj, int* __restrict__ idx,
                          float* __restrict__ alpha, int K, int C, int nb, int MTl, int NTl, int ntiles) {
  extern __shared__ __align__(16) float sm[];
  const int KR = MTl * 16, KP = K | 1;
  float* hs = sm;                              // [FW_NBUF][KR][CHP]
  float* As = sm + FW_NBUF * KR * CHP;         // [K][KP]
  const int b = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* hb = h + (long long)b * K * C;

  for (int v = tid; v < FW_NBUF * KR * CHP; v += nthreads) hs[v] = 0.f;   // rows >= K (and a ragged last k-step) must read as zero

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
  for (int c = 0; c < FW_NBUF - 1; ++c) {
    if (c < nch) stage_chunk(hs + c * KR * CHP, CHP, hb, C, K, c * CH, min(CH, C - c * CH), tid, nthreads);
    else cp_async_commit();
  }
  for (int ch = 0; ch < nch; ++ch) {
    const int nx = ch + FW_NBUF - 1;
    if (nx < nch) stage_chunk(hs + (nx % FW_NBUF) * KR * CHP, CHP, hb, C, K, nx * CH, min(CH, C - nx * CH), tid, nthreads);
    else cp_async_commit();
    cp_async_wait<FW_NBUF - 1>();
    const int cw = min(CH, C - ch * CH);
    float* wbuf = hs + (ch % FW_NBUF) * KR * CHP;
    if ((cw & 7) && ch >= FW_NBUF)             // ragged last k-step (C % 8 == 4): columns [cw, cw + 4) still hold an older chunk
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
  adjacency_finish(As, KP, K, nb, adj + (long long)b * K * K, idx + (long long)b * K * nb, alpha + (long long)b * K * nb, nwarps);
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

// dalpha -> dv (softmax bwd) -> sparse dA in smem -> S = dA + dA^T (+ dadj + dadj^T) -> dh = (S h) * (h > 0), K <= 64: fp32 FFMA on
// 4-row x 4-column register tiles.  Measured in-step at B = 512 (profiles/r02_adjacency_kernels.md): 54 us at K = 36 and 86 us at
// K = 51, against 52 / 124 us for the mma.sync formulation below (its S fragments live in registers: at K = 51 that leaves two
// CTAs of eight warps per SM) - so this kernel keeps K <= 64 and the tensor-core one takes K > 64 (639 vs 819 us per 1024 images).
__global__ void __launch_bounds__(ADJ_THREADS)
adjacency_topk_bwd_ffma_kernel(const float* __restrict__ h, const int* __restrict__ idx, const float* __restrict__ alpha,
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


// The same product for K > 64 on the tensor cores:
// The product runs on the same 3xTF32 mma.sync path as the forward: D[i][c] = sum_j S[i][j] h[j][c] with A = S (row-major, split
// once into tf32 hi / lo planes in shared memory), B[k = j][n = c] = h[j][c] straight from the staged chunk, one 8-column n-tile per
// warp and 64-column chunk, MTC m-tiles of accumulators per warp.  Strides: S rows = 4 (mod 32) words, h rows = 8 (mod 32) words ->
// both fragment load patterns touch 32 distinct banks.
constexpr int HBP = CH + 8;      // backward: padded row stride of the staged h chunk (72 = 8 mod 32)

// Warp (mi, cg): m-tile mi of the K output rows, column group cg; its S fragments (KTC k-steps x hi / lo) stay in registers for the
// whole image, so the streaming loop is 2 LDS + 2 splits + 3 MMAs per (n-tile, k-step).
template <int KTC, int CG, int BW_NBUF>   // KTC >= ceil(K / 8); CG column groups; BW_NBUF chunks of h in flight
__global__ void __launch_bounds__(ADJ_MAXW * 32)
adjacency_topk_bwd_kernel(const float* __restrict__ h, const int* __restrict__ idx, const float* __restrict__ alpha,
                          const float* __restrict__ dalpha, const float* __restrict__ dadj, float* __restrict__ dh,
                          int K, int C, int nb, int SP) {
  extern __shared__ __align__(16) float sm[];
  const int ADJ_THREADS = blockDim.x, ADJ_WARPS = ADJ_THREADS >> 5;        // (shadow the file-level constants: MT * CG warps here)
  const int MTl = (K + 15) >> 4, KT = (K + 7) >> 3, SR = MTl * 16, HR = KT * 8;
  float* S = sm;                                            // [SR][SP] fp32, then the lo plane
  uint32_t* Sh = reinterpret_cast<uint32_t*>(sm + SR * SP); // [SR][SP] tf32 hi plane
  float* hs = sm + 2 * SR * SP;                             // [BW_NBUF][HR][HBP]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int v = tid; v < 2 * SR * SP; v += ADJ_THREADS) S[v] = 0.f;          // both planes (Sh follows S)
  for (int c = 0; c < BW_NBUF; ++c)                         // rows >= K of the staged chunks read as zero
    for (int v = tid; v < (HR - K) * HBP; v += ADJ_THREADS) hs[(c * HR + K) * HBP + v] = 0.f;
  const float* hb = h + (long long)b * K * C;
  float* ob = dh + (long long)b * K * C;
  const int nch = (C + CH - 1) / CH;
  for (int c = 0; c < BW_NBUF - 1; ++c) {                   // the first chunks travel while S is being built
    if (c < nch) stage_chunk(hs + c * HR * HBP, HBP, hb, C, K, c * CH, min(CH, C - c * CH), tid, ADJ_THREADS);
    else cp_async_commit();
  }
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
    for (int i = warp; i < K; i += ADJ_WARPS)
      for (int j = lane; j < K; j += 32) S[i * SP + j] += db[i * K + j];
    __syncthreads();
  }
  // symmetrise and split in one pass: the pair (i, j), i <= j has one owner, who leaves the lo plane in S and the tf32 hi plane in Sh
  for (int i = warp; i < K; i += ADJ_WARPS)
    for (int j = i + lane; j < K; j += 32) {
      uint32_t hi, lo;
      split_tf32(S[i * SP + j] + S[j * SP + i], hi, lo);
      Sh[i * SP + j] = hi; Sh[j * SP + i] = hi;
      S[i * SP + j] = __uint_as_float(lo); S[j * SP + i] = __uint_as_float(lo);
    }

  const uint32_t* Sl = reinterpret_cast<const uint32_t*>(S);
  const int mi = warp / CG, cg = warp - mi * CG;
  uint32_t ah[KTC][4], al[KTC][4];
  for (int ch = 0; ch < nch; ++ch) {
    const int nx = ch + BW_NBUF - 1;
    if (nx < nch) stage_chunk(hs + (nx % BW_NBUF) * HR * HBP, HBP, hb, C, K, nx * CH, min(CH, C - nx * CH), tid, ADJ_THREADS);
    else cp_async_commit();
    cp_async_wait<BW_NBUF - 1>();
    __syncthreads();                                     // chunk ch landed (and, first time round, the S planes are complete)
    if (ch == 0) {
#pragma unroll
      for (int kt = 0; kt < KTC; ++kt) {
        const int o = (mi * 16 + g) * SP + (kt < KT ? kt : 0) * 8 + t;
        ah[kt][0] = Sh[o]; ah[kt][1] = Sh[o + 8 * SP]; ah[kt][2] = Sh[o + 4]; ah[kt][3] = Sh[o + 8 * SP + 4];
        al[kt][0] = Sl[o]; al[kt][1] = Sl[o + 8 * SP]; al[kt][2] = Sl[o + 4]; al[kt][3] = Sl[o + 8 * SP + 4];
      }
    }
    const float* buf = hs + (ch % BW_NBUF) * HR * HBP;
    const int cw = min(CH, C - ch * CH);
    for (int n0 = cg * 8; n0 < cw; n0 += CG * 8) {       // my n-tiles of this chunk (warp-uniform bounds)
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* pb = buf + t * HBP + n0 + g;
#pragma unroll
      for (int kt = 0; kt < KTC; ++kt) {
        if (kt < KT) {
          uint32_t bh[2], bl[2];
          split_tf32(pb[kt * 8 * HBP], bh[0], bl[0]);
          split_tf32(pb[(kt * 8 + 4) * HBP], bh[1], bl[1]);
          mma_tf32(acc, al[kt][0], al[kt][1], al[kt][2], al[kt][3], bh[0], bh[1]);
          mma_tf32(acc, ah[kt][0], ah[kt][1], ah[kt][2], ah[kt][3], bl[0], bl[1]);
          mma_tf32(acc, ah[kt][0], ah[kt][1], ah[kt][2], ah[kt][3], bh[0], bh[1]);
        }
      }
      const int col = n0 + 2 * t;                        // columns col, col + 1 of the chunk (cw % 4 == 0: both valid or both not)
      if (col < cw) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int i = mi * 16 + g + hf * 8;
          if (i < K) {
            const float2 hv = *reinterpret_cast<const float2*>(buf + i * HBP + col);   // ReLU mask of the forward: h > 0
            float2 o;
            o.x = hv.x > 0.f ? acc[2 * hf] : 0.f;
            o.y = hv.y > 0.f ? acc[2 * hf + 1] : 0.f;
            *reinterpret_cast<float2*>(ob + (long long)i * C + ch * CH + col) = o;
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
  if (MT <= 4) {                                           // K <= 64: split-K over the 8 warps, every warp holds all tiles
    const int nbuf = MT <= 3 ? 8 : 4;
    const size_t smem_sk = (size_t)(nbuf * MT * 16 * CHP + K * (K | 1)) * sizeof(float);
    auto run_sk = [&](auto kern) -> int {
      VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sk));
      kern<<<B, ADJ_THREADS, smem_sk, stream>>>(h, adjacency, idx, alpha, K, C, nb);
      VQA_LAUNCH_CHECK("adjacency_topk_fwd_sk_kernel");
      return VQA_OK;
    };
    if (MT == 1) return run_sk(adjacency_topk_fwd_sk_kernel<1, 8>);
    if (MT == 2) return run_sk(adjacency_topk_fwd_sk_kernel<2, 8>);
    if (MT == 3) return run_sk(adjacency_topk_fwd_sk_kernel<3, 8>);
    return run_sk(adjacency_topk_fwd_sk_kernel<4, 4>);
  }
  int ntiles = 0;
  for (int mi = 0; mi < MT; ++mi) ntiles += NT - 2 * mi;
  int nw = 8, best = 1 << 30;                               // fewest idle tile slots, then fewest warps
  for (int w = 8; w <= ADJ_MAXW; ++w) {
    const int waste = (ntiles + w - 1) / w * w - ntiles;
    if (waste < best) { best = waste; nw = w; }
  }
  const int maxt = (ntiles + nw - 1) / nw;
  const size_t smem = (size_t)(FW_NBUF * MT * 16 * CHP + K * (K | 1)) * sizeof(float);
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
  if (K <= 64) {                                            // fp32 FFMA register tiles (see the kernel's header for the measurements)
    const int KPf = (K + 3) & ~3;
    int CW = (int)((96 * 1024) / (K * 4) / 128) * 128;
    if (CW < 128) CW = 128;
    if (CW > ((C + 127) & ~127)) CW = (C + 127) & ~127;
    const size_t smem_f = (size_t)(KPf * KPf + K * CW) * sizeof(float);
    VQA_CUDA(cudaFuncSetAttribute(adjacency_topk_bwd_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    adjacency_topk_bwd_ffma_kernel<<<B, ADJ_THREADS, smem_f, stream>>>(h, idx, alpha, dalpha, dadj, dh, K, C, nb, CW);
    VQA_LAUNCH_CHECK("adjacency_topk_bwd_ffma_kernel");
    return VQA_OK;
  }
  const int MT = (K + 15) / 16, KT = (K + 7) / 8;
  int SP = KT * 8;                                           // row stride of the S planes: >= 8 KT words and = 4 (mod 32)
  SP += (4 - SP % 32 + 32) % 32;
  const int nbuf = MT <= 4 ? 4 : 2;                         // K > 64: the S planes alone take 118 .. 135 KB
  const size_t smem = (size_t)(2 * MT * 16 * SP + nbuf * KT * 8 * HBP) * sizeof(float);
  auto run = [&](auto kern, int cg) -> int {
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, MT * cg * 32, smem, stream>>>(h, idx, alpha, dalpha, dadj, dh, K, C, nb, SP);
    VQA_LAUNCH_CHECK("adjacency_topk_bwd_kernel");
    return VQA_OK;
  };
  // (k-steps held in registers, column groups)
  if (KT <= 13) return run(adjacency_topk_bwd_kernel<13, 1, 2>, 1);        // K <= 104: 7 x 1 (K = 100)
  return run(adjacency_topk_bwd_kernel<16, 1, 2>, 1);                      // K <= 128: 8 x 1

}
