// Dense projections on the bf16 tensor-core path with fp32-grade results: "split-bf16" tcgen05 GEMM (sm_100a).
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),   A = A_hi + A_lo,  B = B_hi + B_lo   (bf16 planes in HBM)
//
// Every fp32 operand x is stored as two bf16 planes hi = bf16(x), lo = bf16(x - hi) (16+ mantissa bits, produced by
// vqa_split_bf16_f32 or directly by the producing kernel's epilogue).  PASSES = 3 issues lo*hi + hi*lo + hi*hi into
// one fp32 TMEM accumulator (relative error ~2^-17 per product: well inside the 1e-3 parity budget, where single-pass
// TF32 measured 2.9e-3 is not); PASSES = 1 uses the hi planes only (plain bf16 tensor-core GEMM, stated tolerance).
// Versus the TF32x3 kernel (gemm_tcgen05.cu) this halves the tensor-pipe time per pass (kind::f16, K = 16 per
// instruction), needs no in-kernel operand transform (no generic-proxy smem traffic at all: TMA -> smem -> UMMA) and
// moves the same 4 bytes per operand element.
//
// Replaces the same reference call sites as gemm_tcgen05.cu: the per-kernel conv Linears (layers.py:140-142, ONE
// projection), the classifier (sparse_graph_model.py:154-157), the graph-learner backward, and every dX / dW product.
//
// Structure: one 128 x BN output tile per CTA, 6 warps: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA
// issuer, warps 2-5 = epilogue (TMEM -> registers -> global; bias / row-broadcast / ReLU / mask / split-plane output).
// Operands are K-major (contraction contiguous) or MN-major (dW = dY^T X: contraction strided); both arrive by TMA
// with the 128-byte swizzle in the canonical UMMA layouts.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <mutex>
#include <cstring>
#include <cstdlib>

#include "../../include/vqa_b200.h"

namespace vqa {
namespace sb {

constexpr int BM = 128;
constexpr int BK = 64;                  // bf16 elements per k-block: one 128-byte swizzle row
constexpr int A_TILE = BM * 128;        // bytes per plane
constexpr int THREADS = 192;
constexpr int THREADS_EPI8 = 32 * 10;    // with eight epilogue warps instead of four (see gemm_bf16s_persistent_kernel)
constexpr int SMEM_BUDGET = 232448 - 1024 - 256;

struct Params {
  float* C; long long ldc;
  __nv_bfloat16* Chi; __nv_bfloat16* Clo; long long ldcs;      // optional split-plane copy of the output
  int M, N, Kc, a_mn, b_mn;
  const float* bias;
  const float* rowb; long long ldrb; int group;
  const float* aux; long long ldaux;                            // mask source, fp32 ...
  const __nv_bfloat16* auxh; long long ldauxh;                  // ... or the hi plane of a split tensor
  float aux_scale;
  int flags, kb_per_split, num_kb;
  const int* tile_gate; int gate_t;                            // optional: the CTA of row tile m exits when tile_gate[m] <= gate_t
};

template <int BN, int PASSES>
struct Cfg {
  static constexpr int PLANES = PASSES == 3 ? 2 : 1;
  static constexpr int B_TILE = BN * 128;
  static constexpr int PLANE = A_TILE + B_TILE;               // [A | B] of one plane
  static constexpr int STAGE = PLANE * PLANES;
  static constexpr int BIAS_S = 8 * BN * 4;                   // one copy of the tile's bias columns per epilogue warp (up to 8 of them)
  static constexpr int S_ = (SMEM_BUDGET - BIAS_S) / STAGE;
  // BN = 64 is the small-problem tile (few CTAs, latency-bound): 2 x 48 KB stages so that 2 CTAs are resident per SM
  static constexpr int S = BN == 64 ? (PLANES == 2 ? 2 : 4) : (S_ > 8 ? 8 : S_);
  static constexpr int SMEM = S * STAGE + 1024 + 256 + BIAS_S;
};

// UMMA shared-memory descriptor, 16-bit operands, 128-byte swizzle.
//   K-major : rows of 128 B (64 k), 8-row swizzle atoms 1024 B apart (SBO); LBO unused.
//   MN-major: rows of 128 B (64 mn) indexed by k, 8-k atoms 1024 B apart (SBO), 64-element MN chunks 8192 B apart
//             (LBO = one [64 k x 64 mn] TMA box).
__device__ __forceinline__ uint64_t umma_desc16(uint32_t saddr, int mn_major) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(mn_major ? (8192 >> 4) : 1) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;                                            // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// 256-bit global stores (sm_100): a lane's 32 contiguous bytes leave as ONE full sector instead of two half-sector requests.  The
// epilogue stores one output row per lane, so every request is its own sector anyway; with 16-byte stores the short-contraction
// products were bound by the L1 -> crossbar request path (GI: 5.5 M sector requests and 264 MB of write traffic for 88 MB).
__device__ __forceinline__ void st_v8_f32(float* ptr, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
               "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_v8_b32(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w) : "memory");
}

// epilogue for 32 consecutive columns [cb, cb+32) of one output row
struct Epi {
  const Params& p;
  float* crow; __nv_bfloat16* hrow; __nv_bfloat16* lrow; const float* rb; const float* ax; const __nv_bfloat16* axh;
  const float* bs;      // the bias, staged in shared memory by the caller and offset so that bs[col] is column col (NULL: no bias)
  bool vec_ok, relu, atomic, c32, h32;
  __device__ __forceinline__ Epi(const Params& p_, int row, const float* bias_smem) : p(p_), bs(bias_smem) {
    c32 = p.C && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 31) == 0);
    h32 = p.Chi && ((p.ldcs & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.Chi) & 31) == 0) && (!p.Clo || (reinterpret_cast<uintptr_t>(p.Clo) & 31) == 0);
    relu = p.flags & VQA_GEMM_RELU; atomic = p.flags & VQA_GEMM_ATOMIC_ADD;
    crow = p.C ? p.C + (long long)row * p.ldc : nullptr;
    hrow = p.Chi ? p.Chi + (long long)row * p.ldcs : nullptr;
    lrow = p.Clo ? p.Clo + (long long)row * p.ldcs : nullptr;
    rb = p.rowb ? p.rowb + (long long)(row / p.group) * p.ldrb : nullptr;
    ax = p.aux ? p.aux + (long long)row * p.ldaux : nullptr;
    axh = p.auxh ? p.auxh + (long long)row * p.ldauxh : nullptr;
    vec_ok = ((p.N & 7) == 0) &&
             (!p.C || (((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0))) &&
             (!p.Chi || (((p.ldcs & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.Chi) & 15) == 0) && (!p.Clo || (reinterpret_cast<uintptr_t>(p.Clo) & 15) == 0))) &&
             (!ax || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0))) &&
             (!axh || (((p.ldauxh & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.auxh) & 15) == 0))) &&
             (!rb || (((p.ldrb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.rowb) & 15) == 0))) &&
             (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
  }
  __device__ __forceinline__ float one(float v, int col) const {
    if (rb) v += rb[col];
    if (p.bias) v += p.bias[col];
    if (relu) v = fmaxf(v, 0.f);
    if (ax) v = ax[col] > 0.f ? v * p.aux_scale : 0.f;
    if (axh) v = __bfloat162float(axh[col]) > 0.f ? v * p.aux_scale : 0.f;
    return v;
  }
  __device__ __forceinline__ void store32(int cb, const float* r) const {
    if (vec_ok && !atomic && !(rb && ax)) {
      // Loads first, arithmetic and stores second: with the loads interleaved group by group (and an early exit between the
      // groups that keeps the compiler from hoisting them) every 8 columns waited for a global-memory round trip of their own -
      // measured on GI = E W_ih^T (K = 300): 45 % of all stall samples sat on the FADD behind the bias load, and draining a
      // 128 x 256 tile took ~20 us against 4 us to accumulate it.  The bias now comes from shared memory (staged per tile before
      // the accumulator is waited for); the row broadcast / mask values of all 32 columns are requested in one go.
      const int ng = min(4, (p.N - cb + 7) >> 3);           // 8-column groups inside N (cb < N)
      const float* src = rb ? rb : ax;                      // (never both here)
      float t[32];
      uint4 mh[4];
      uint4 keep_h = make_uint4(0u, 0u, 0u, 0u), keep_l = keep_h;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = cb + 8 * g;
        if (g < ng) {
          if (src) {
            const float4 a0 = *reinterpret_cast<const float4*>(src + col), a1 = *reinterpret_cast<const float4*>(src + col + 4);
            t[8 * g] = a0.x; t[8 * g + 1] = a0.y; t[8 * g + 2] = a0.z; t[8 * g + 3] = a0.w;
            t[8 * g + 4] = a1.x; t[8 * g + 5] = a1.y; t[8 * g + 6] = a1.z; t[8 * g + 7] = a1.w;
          }
          if (axh) mh[g] = *reinterpret_cast<const uint4*>(axh + col);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = cb + 8 * g, j = 8 * g;
        if (g >= ng) continue;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = r[j + e];
        if (rb) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += t[j + e];
        }
        if (bs) {
          const float4 a0 = *reinterpret_cast<const float4*>(bs + col), a1 = *reinterpret_cast<const float4*>(bs + col + 4);
          v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w; v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if (ax) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = t[j + e] > 0.f ? v[e] * p.aux_scale : 0.f;
        }
        if (axh) {
          const uint32_t w[4] = {mh[g].x, mh[g].y, mh[g].z, mh[g].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {     // bf16 > 0  <=>  sign clear and magnitude bits non-zero
            const uint32_t lo16 = w[e] & 0xFFFFu, hi16 = w[e] >> 16;
            v[2 * e] = (lo16 != 0 && lo16 < 0x8000u) ? v[2 * e] * p.aux_scale : 0.f;
            v[2 * e + 1] = (hi16 != 0 && hi16 < 0x8000u) ? v[2 * e + 1] * p.aux_scale : 0.f;
          }
        }
        if (crow) {
          if (c32) st_v8_f32(crow + col, v);
          else {
            *reinterpret_cast<float4*>(crow + col) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(crow + col + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
        if (hrow) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
          const uint4 ph = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
          const uint4 pl = make_uint4(pack_bf16(v[0] - h[0], v[1] - h[1]), pack_bf16(v[2] - h[2], v[3] - h[3]),
                                      pack_bf16(v[4] - h[4], v[5] - h[5]), pack_bf16(v[6] - h[6], v[7] - h[7]));
          if (h32 && !(g & 1) && g + 1 < ng) { keep_h = ph; keep_l = pl; }           // first half of a 16-column pair: wait for the second
          else if (h32 && (g & 1)) {
            st_v8_b32(hrow + col - 8, keep_h, ph);
            if (lrow) st_v8_b32(lrow + col - 8, keep_l, pl);
          } else {
            *reinterpret_cast<uint4*>(hrow + col) = ph;
            if (lrow) *reinterpret_cast<uint4*>(lrow + col) = pl;
          }
        }
      }
    } else if (atomic && crow && !hrow && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {           // split-K / accumulate: 16-byte vector reductions (red.global.add.v4.f32)
        const int col = cb + j;
        if (col >= p.N) break;
        atomicAdd(reinterpret_cast<float4*>(crow + col), make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]));
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const int col = cb + j;
        if (col >= p.N) break;
        const float v = one(r[j], col);
        if (crow) { if (atomic) atomicAdd(crow + col, v); else crow[col] = v; }
        if (hrow) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          hrow[col] = h;
          if (lrow) lrow[col] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
      }
    }
  }
};

// The bias columns [n0, n0 + BN) of a tile, copied by each epilogue warp into its own shared-memory row BEFORE it waits for the
// accumulator; returns the pointer offset so that [col] addresses column col (NULL without a bias).
template <int BN>
__device__ __forceinline__ const float* stage_bias(const Params& p, float* warp_row, int lane, int n0) {
  if (!p.bias) return nullptr;
  __syncwarp();                                              // the previous tile's reads are done
#pragma unroll
  for (int i = 0; i < BN / 32; ++i) {
    const int col = n0 + lane + 32 * i;
    warp_row[lane + 32 * i] = col < p.N ? __ldg(p.bias + col) : 0.f;
  }
  __syncwarp();
  return warp_row - n0;
}

struct Maps { CUtensorMap a_hi, a_lo, b_hi, b_lo; };

#ifdef VQA_GEMM_TRACE
// Tracing build only (make EXTRA=-DVQA_GEMM_TRACE BUILD=build_trace OUT=../vqa_b200/libvqa_trace.so; tools/gemm_timeline.py): clock64
// stamps of the phases of each CTA of gemm_bf16s_kernel (the one-tile-per-CTA kernel the recurrence products run on).
__device__ long long g_gemm_trace[8 * 512];
#define GEMM_STAMP(COND, I) do { if ((COND)) { const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
    if (cta_ < 512) g_gemm_trace[cta_ * 8 + (I)] = clock64(); } } while (0)
#else
#define GEMM_STAMP(COND, I) do { } while (0)
#endif

// ---- thread-block-cluster helpers (CL = 2: the two CTAs of a pair own vertically adjacent output tiles and share B) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) in every CTA of cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// MMA-completion arrive on the barrier at this offset in every CTA of cta_mask (a stage is refilled by both producers)
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

template <int BN, int PASSES, int CL>
__global__ void __launch_bounds__(BN >= 128 ? THREADS_EPI8 : THREADS)
gemm_bf16s_kernel(const __grid_constant__ Maps tm, const Params p) {
  using C = Cfg<BN, PASSES>;
  constexpr int S = C::S;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * C::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  // Row-tile gate (padded recurrences): a 128-row tile whose sequences have all ended contributes nothing - its rows of the
  // output are never read (forward) or would only receive zeros (backward accumulate) - so the whole CTA leaves before it
  // touches a barrier.  Uniform per CTA; never combined with CTA pairs (host side).
  if (CL == 1 && p.tile_gate != nullptr && p.tile_gate[blockIdx.y] <= p.gate_t) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  GEMM_STAMP(threadIdx.x == 0, 0);
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_stop = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_stop - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a_hi); tma_prefetch_desc(&tm.b_hi);
    if (PASSES == 3) { tma_prefetch_desc(&tm.a_lo); tma_prefetch_desc(&tm.b_lo); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }   // CL = 2: both CTAs' MMAs release a stage
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();          // the peer's barriers are initialised before any multicast can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GEMM_STAMP(threadIdx.x == 0, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S, ph = (i / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], C::STAGE);
        const int k0 = (kb_begin + i) * BK;
#pragma unroll
        for (int pl = 0; pl < C::PLANES; ++pl) {
          uint8_t* a_dst = smem + s * C::STAGE + pl * C::PLANE;
          uint8_t* b_dst = a_dst + A_TILE;
          const CUtensorMap* ma = pl ? &tm.a_lo : &tm.a_hi;
          const CUtensorMap* mb = pl ? &tm.b_lo : &tm.b_hi;
          if (!p.a_mn) {
            tma_load_2d(a_dst, ma, &full[s], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * 8192, ma, &full[s], m0 + c * 64, k0);
          }
          if (CL == 2) {
            // this CTA fetches its half of the shared B tile and multicasts it into both CTAs of the pair
            if (!p.b_mn) {
              tma_load_2d_mc(b_dst + crank * (BN / 2) * 128, mb, &full[s], k0, n0 + (int)crank * (BN / 2), (uint16_t)3);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 128; ++c) {
                const int cc = (int)crank * (BN / 128) + c;
                tma_load_2d_mc(b_dst + cc * 8192, mb, &full[s], n0 + cc * 64, k0, (uint16_t)3);
              }
            }
          } else if (!p.b_mn) {
            tma_load_2d(b_dst, mb, &full[s], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, mb, &full[s], n0 + c * 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    // instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t a_step = p.a_mn ? 2048 : 32, b_step = p.b_mn ? 2048 : 32;   // bytes per K = 16
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S, ph = (i / S) & 1;
      mbar_wait(&full[s], ph);
      GEMM_STAMP(lane == 0 && i == 0, 2);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + s * C::STAGE), b_hi = a_hi + A_TILE;
        const uint32_t a_lo = a_hi + C::PLANE, b_lo = a_lo + A_TILE;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t dah = umma_desc16(a_hi + ks * a_step, p.a_mn), dbh = umma_desc16(b_hi + ks * b_step, p.b_mn);
          const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
          if (PASSES == 3) {
            const uint64_t dal = umma_desc16(a_lo + ks * a_step, p.a_mn), dbl = umma_desc16(b_lo + ks * b_step, p.b_mn);
            tc_mma<1>(tmem_base, dal, dbh, idesc, acc);   // small terms first
            tc_mma<1>(tmem_base, dah, dbl, idesc, 1u);
            tc_mma<1>(tmem_base, dah, dbh, idesc, 1u);
          } else {
            tc_mma<1>(tmem_base, dah, dbh, idesc, acc);
          }
        }
        if (CL == 2) tc_commit_mc(&empty[s], (uint16_t)3);   // the slot is refilled by both producers: release it in both CTAs
        else tc_commit(&empty[s]);                  // smem slot reusable once these MMAs retire
        if (i == nkb - 1) tc_commit(acc_full);      // accumulator complete
        GEMM_STAMP(i == nkb - 1, 3);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ warps 2-5: epilogue
    const float* bs = stage_bias<BN>(p, reinterpret_cast<float*>(smem + S * C::STAGE + 256) + (warp - 2) * BN, lane, n0);
    mbar_wait(acc_full, 0);
    GEMM_STAMP(threadIdx.x == 64, 4);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const bool epi8 = blockDim.x == THREADS_EPI8;     // eight epilogue warps: two per lane quarter, half of the tile's columns each
    const int c_begin = epi8 ? ((warp - 2) >> 2) * (BN / 2) : 0, c_end = epi8 ? c_begin + BN / 2 : BN;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    // Plain fp32 destination (optionally + bias, ReLU, or a split-K / accumulate reduction): the drain of a one-tile CTA is never
    // hidden, and stored one row per lane it is bound by the request rate of 32 scattered sectors per instruction (3.3 us of the
    // recurrence product's 12 us, tools/gemm_timeline.py).  The pipeline's shared memory is free once the accumulator is complete
    // (every stage fill has been consumed), so the warp parks its 32 rows there and writes them out row by row: every store
    // instruction covers whole 128-byte lines of one or two rows.
    const bool coalesced = BN >= 128 && p.C && !p.Chi && !p.rowb && !p.aux && !p.auxh && ((p.N & 3) == 0) && ((p.ldc & 3) == 0) &&
                           ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    if (coalesced) {
      const int W = c_end - c_begin;                          // this warp's columns: 64, 128 or 256
      float* st = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (W + 4);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        if (n0 + c0 >= p.N) break;            // warp-uniform
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tc_wait_ld();
        float* mine = st + lane * (W + 4) + (c0 - c_begin);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(mine + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
      __syncwarp();
      const int lpr = W >= 128 ? 32 : W / 4;                  // lanes per row
      const bool relu = p.flags & VQA_GEMM_RELU, atomic = p.flags & VQA_GEMM_ATOMIC_ADD;
      for (int seg = 0; seg < W; seg += 128) {
        const int cl = seg + (lane % lpr) * 4, col = n0 + c_begin + cl;
        if (col >= p.N) continue;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bs && !atomic) b4 = *reinterpret_cast<const float4*>(bs + col);
        for (int rr = lane / lpr; rr < 32; rr += 32 / lpr) {
          const int orow = m0 + q * 32 + rr;
          if (orow >= p.M) break;
          float4 v = *reinterpret_cast<const float4*>(st + rr * (W + 4) + cl);
          float* dst = p.C + (long long)orow * p.ldc + col;
          if (atomic) { atomicAdd(reinterpret_cast<float4*>(dst), v); continue; }
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
    } else {
      const Epi epi(p, row_ok ? row : 0, bs);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        if (n0 + c0 >= p.N) break;            // warp-uniform
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tc_wait_ld();
        if (row_ok) epi.store32(n0 + c0, reinterpret_cast<const float*>(r));
      }
    }
    GEMM_STAMP(threadIdx.x == 64, 5);
  }
  tc_fence_before();
  __syncthreads();
  GEMM_STAMP(threadIdx.x == 0, 6);
  if (CL == 2) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// ------------------------------------------------------------------------------------------ persistent variant
// Same tile maths, operands and epilogue; what changes is the schedule.  One CTA (or CTA pair) per SM walks the output tiles
// unit = cluster + i * #clusters (n fastest, so concurrently running CTAs share their A rows in L2) with TWO TMEM
// accumulators: the epilogue of tile i (TMEM -> registers -> HBM, up to ~10 us with a mask and split-plane output) runs
// while the producer and the MMA issuer are already inside tile i+1, and barrier setup / TMEM allocation / descriptor
// prefetch happen once per CTA instead of once per tile.  With one tile per CTA the epilogue and the prologue were serial:
// 54 % of the tensor peak on dG1 (K = 1024, mask + planes), 30-35 % on the short-K products (K = 300 / 512).
struct PSched { int tiles_n, tiles_m, units_m, splits, units, nclusters; };

// Epilogue warps: four (one per TMEM lane quarter) or eight (two per quarter, each half of the tile's columns), chosen by the block
// size the host launches with.  A drain is a latency-bound instruction stream (one warp per scheduler has nothing to hide its
// loads and conversions behind): eight warps drain a tile in about half the time, which is what the short-contraction products
// need (their tile is accumulated in 4-6 us and drained in 10-20); the long ones hide the drain anyway and ran 7 % slower with
// eight (measured in round 1), so they keep four.
template <int BN, int PASSES, int CL>
__global__ void __launch_bounds__(THREADS_EPI8, 1)
gemm_bf16s_persistent_kernel(const __grid_constant__ Maps tm, const Params p, const PSched sc) {
  using C = Cfg<BN, PASSES>;
  constexpr int S = C::S;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * C::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;          // [2]
  uint64_t* acc_empty = bars + 2 * S + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cluster = blockIdx.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a_hi); tma_prefetch_desc(&tm.b_hi);
    if (PASSES == 3) { tma_prefetch_desc(&tm.a_lo); tma_prefetch_desc(&tm.b_lo); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
      for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], (blockDim.x >> 5) - 2); }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (split z, row tile(s), column tile); every role decodes the same sequence
  auto decode = [&](int u, int& n0, int& m0, int& kb_begin, int& nkb, bool& live) {
    const int tn = u % sc.tiles_n, r = u / sc.tiles_n, um = r % sc.units_m, z = r / sc.units_m;
    const int mt = um * CL + (int)crank;
    n0 = tn * BN; m0 = mt * BM;
    kb_begin = z * p.kb_per_split;
    nkb = min(p.num_kb, kb_begin + p.kb_per_split) - kb_begin;
    live = !(CL == 1 && p.tile_gate != nullptr && p.tile_gate[mt] <= p.gate_t);   // row-tile gate: skipped by every role alike
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int u = cluster; u < sc.units; u += sc.nclusters) {
        int n0, m0, kb_begin, nkb; bool live;
        decode(u, n0, m0, kb_begin, nkb, live);
        if (!live) continue;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], C::STAGE);
          const int k0 = (kb_begin + i) * BK;
#pragma unroll
          for (int pl = 0; pl < C::PLANES; ++pl) {
            uint8_t* a_dst = smem + s * C::STAGE + pl * C::PLANE;
            uint8_t* b_dst = a_dst + A_TILE;
            const CUtensorMap* ma = pl ? &tm.a_lo : &tm.a_hi;
            const CUtensorMap* mb = pl ? &tm.b_lo : &tm.b_hi;
            if (!p.a_mn) {
              tma_load_2d(a_dst, ma, &full[s], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(a_dst + c * 8192, ma, &full[s], m0 + c * 64, k0);
            }
            if (CL == 2) {
              if (!p.b_mn) {
                tma_load_2d_mc(b_dst + crank * (BN / 2) * 128, mb, &full[s], k0, n0 + (int)crank * (BN / 2), (uint16_t)3);
              } else {
#pragma unroll
                for (int c = 0; c < BN / 128; ++c) {
                  const int cc = (int)crank * (BN / 128) + c;
                  tma_load_2d_mc(b_dst + cc * 8192, mb, &full[s], n0 + cc * 64, k0, (uint16_t)3);
                }
              }
            } else if (!p.b_mn) {
              tma_load_2d(b_dst, mb, &full[s], k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) tma_load_2d(b_dst + c * 8192, mb, &full[s], n0 + c * 64, k0);
            }
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread), accumulator t & 1 for the t-th live tile
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_step = p.a_mn ? 2048 : 32, b_step = p.b_mn ? 2048 : 32;   // bytes per K = 16
      int s = 0, t = 0; uint32_t ph = 0;
      for (int u = cluster; u < sc.units; u += sc.nclusters) {
        int n0, m0, kb_begin, nkb; bool live;
        decode(u, n0, m0, kb_begin, nkb, live);
        if (!live) continue;
        const int acc = t & 1;
        mbar_wait(&acc_empty[acc], ((t >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator (tile t - 2)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * C::STAGE), b_hi = a_hi + A_TILE;
          const uint32_t a_lo = a_hi + C::PLANE, b_lo = a_lo + A_TILE;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t dah = umma_desc16(a_hi + ks * a_step, p.a_mn), dbh = umma_desc16(b_hi + ks * b_step, p.b_mn);
            const uint32_t accum = (i > 0 || ks > 0) ? 1u : 0u;
            if (PASSES == 3) {
              const uint64_t dal = umma_desc16(a_lo + ks * a_step, p.a_mn), dbl = umma_desc16(b_lo + ks * b_step, p.b_mn);
              tc_mma<1>(d_tmem, dal, dbh, idesc, accum);   // small terms first
              tc_mma<1>(d_tmem, dah, dbl, idesc, 1u);
              tc_mma<1>(d_tmem, dah, dbh, idesc, 1u);
            } else {
              tc_mma<1>(d_tmem, dah, dbh, idesc, accum);
            }
          }
          if (CL == 2) tc_commit_mc(&empty[s], (uint16_t)3);
          else tc_commit(&empty[s]);
          if (i == nkb - 1) tc_commit(&acc_full[acc]);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        ++t;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ warps 2-5 (2-9): epilogue of tile t while tile t + 1 is being accumulated
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const bool epi8 = blockDim.x == THREADS_EPI8;
    const int c_begin = epi8 ? ((warp - 2) >> 2) * (BN / 2) : 0, c_end = epi8 ? c_begin + BN / 2 : BN;   // this warp's columns of the tile
    int t = 0;
    for (int u = cluster; u < sc.units; u += sc.nclusters) {
      int n0, m0, kb_begin, nkb; bool live;
      decode(u, n0, m0, kb_begin, nkb, live);
      if (!live) continue;
      const int acc = t & 1;
      const float* bs = stage_bias<BN>(p, reinterpret_cast<float*>(smem + S * C::STAGE + 256) + (warp - 2) * BN, lane, n0);
      if (lane == 0) mbar_wait(&acc_full[acc], (t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const Epi epi(p, row_ok ? row : 0, bs);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        if (n0 + c0 >= p.N) break;            // warp-uniform
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tc_wait_ld();
        if (row_ok) epi.store32(n0 + c0, reinterpret_cast<const float*>(r));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      ++t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
  }
}

// ------------------------------------------------------------------------------------------ fused GRU step (forward)
// One recurrence step of the question encoder (sparse_graph_model.py:117-121, torch gate order r,z,n) as ONE kernel:
//   gh = h_{t-1} W_hh^T  (3-pass split-bf16, M = batch rows, K = H)  ->  cell  ->  h_t (fp32 + planes) and the saved gates.
// W_hh rows (and gi, b_hh columns) are given in UNIT-BLOCK order: block u holds [r | z | n] of hidden units 32u .. 32u+31
// (96 rows), so the 128 x 96 accumulator of CTA (u, row tile) contains all three gate pre-activations of its 32 units and the
// cell is evaluated straight out of TMEM - no (B,3H) gh round trip through HBM, no separate cell launch.
// Outputs are written in the ORIGINAL layouts (h at unit j, gates r|z|n|gh_n at j, H+j, 2H+j, 3H+j), so the backward kernels
// are untouched.  Row tiles whose questions have all ended exit at once (tile_gate); their h rows are never read again (the
// caller gathers every question's state at its own last step).
constexpr int G_BN = 96, G_S = 3;
constexpr int G_STAGE = 2 * (A_TILE + G_BN * 128);          // hi + lo planes of [A 128 x 64 | B 96 x 64]
constexpr int G_SMEM = G_S * G_STAGE + 1024 + 256;
struct GruParams {
  const float* gi; long long ldgi; const float* b_hh; const float* h_prev; const int* len; int t;
  float* h_out; __nv_bfloat16* hout_hi; __nv_bfloat16* hout_lo; long long ldp; float* gates; const int* tile_gate;
  int B, H, num_kb;
};
__device__ __forceinline__ float gru_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// The cell for the 128 x 96 accumulator of CTA (u, row tile m0): shared by the per-step and the whole-sequence kernels.
__device__ __forceinline__ void gru_cell_from_tmem(const GruParams& p, uint32_t tmem_base, uint64_t* acc_full, uint32_t acc_parity, int warp,
                                                   int lane, int u, int m0) {
    // ------------------------------------------------------------ warps 2-5: the cell, one batch row per thread.  The row's input
    // projections and previous state (128 floats) are fetched into registers WHILE the product is being accumulated.
    const int q = warp & 3, b = m0 + q * 32 + lane;
    const bool row_ok = b < p.B;
    const int H = p.H, j0 = u * 32;
    const long long br_ = row_ok ? b : 0;
    const bool active = row_ok && p.t < p.len[br_];
    const float* gib = p.gi + br_ * p.ldgi + u * G_BN;
    const float* bh = p.b_hh + u * G_BN;
    float4 gi4[24], hp4[8];
#pragma unroll
    for (int v = 0; v < 24; ++v) gi4[v] = *reinterpret_cast<const float4*>(gib + v * 4);
#pragma unroll
    for (int v = 0; v < 8; ++v) hp4[v] = *reinterpret_cast<const float4*>(p.h_prev + br_ * H + j0 + v * 4);
    if (lane == 0) mbar_wait(acc_full, acc_parity);
    __syncwarp();
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint32_t ar[8], az[8], an[8];
      tc_ld_32x8(lane_base + (uint32_t)c, ar);
      tc_ld_32x8(lane_base + (uint32_t)(32 + c), az);
      tc_ld_32x8(lane_base + (uint32_t)(64 + c), an);
      tc_wait_ld();
      float hv[8], rr[8], zz[8], nn[8], mm[8];
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 ir = gi4[(c + e) >> 2], iz = gi4[8 + ((c + e) >> 2)], in_ = gi4[16 + ((c + e) >> 2)], h4 = hp4[(c + e) >> 2];
        const float4 br = __ldg(reinterpret_cast<const float4*>(bh + c + e)), bz = __ldg(reinterpret_cast<const float4*>(bh + 32 + c + e)),
                     bn = __ldg(reinterpret_cast<const float4*>(bh + 64 + c + e));
        const float air[4] = {ir.x, ir.y, ir.z, ir.w}, aiz[4] = {iz.x, iz.y, iz.z, iz.w}, ain[4] = {in_.x, in_.y, in_.z, in_.w};
        const float abr[4] = {br.x, br.y, br.z, br.w}, abz[4] = {bz.x, bz.y, bz.z, bz.w}, abn[4] = {bn.x, bn.y, bn.z, bn.w};
        const float ahp[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float ghr = __uint_as_float(ar[e + w]) + abr[w], ghz = __uint_as_float(az[e + w]) + abz[w], ghn = __uint_as_float(an[e + w]) + abn[w];
          const float r = gru_sigmoid(air[w] + ghr), z = gru_sigmoid(aiz[w] + ghz);
          const float n = tanhf(ain[w] + r * ghn);
          const float hnew = (1.f - z) * n + z * ahp[w];
          hv[e + w] = active ? hnew : ahp[w];
          rr[e + w] = r; zz[e + w] = z; nn[e + w] = n; mm[e + w] = ghn;
        }
      }
      if (row_ok) {
        float* ho = p.h_out + (long long)b * H + j0 + c;
        *reinterpret_cast<float4*>(ho) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4*>(ho + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        uint32_t hh[4], ll[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hh[e] = pack_bf16(hv[2 * e], hv[2 * e + 1]);
          ll[e] = pack_bf16(hv[2 * e] - __uint_as_float(hh[e] << 16), hv[2 * e + 1] - __uint_as_float(hh[e] & 0xFFFF0000u));
        }
        const long long po = (long long)b * p.ldp + j0 + c;
        *reinterpret_cast<uint4*>(p.hout_hi + po) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4*>(p.hout_lo + po) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        float* g = p.gates + (long long)b * 4 * H + j0 + c;
        *reinterpret_cast<float4*>(g) = make_float4(rr[0], rr[1], rr[2], rr[3]);          *reinterpret_cast<float4*>(g + 4) = make_float4(rr[4], rr[5], rr[6], rr[7]);
        *reinterpret_cast<float4*>(g + H) = make_float4(zz[0], zz[1], zz[2], zz[3]);      *reinterpret_cast<float4*>(g + H + 4) = make_float4(zz[4], zz[5], zz[6], zz[7]);
        *reinterpret_cast<float4*>(g + 2 * H) = make_float4(nn[0], nn[1], nn[2], nn[3]);  *reinterpret_cast<float4*>(g + 2 * H + 4) = make_float4(nn[4], nn[5], nn[6], nn[7]);
        *reinterpret_cast<float4*>(g + 3 * H) = make_float4(mm[0], mm[1], mm[2], mm[3]);  *reinterpret_cast<float4*>(g + 3 * H + 4) = make_float4(mm[4], mm[5], mm[6], mm[7]);
      }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
gru_step_kernel(const __grid_constant__ Maps tm, const GruParams p) {
  constexpr int S = G_S, PLANE = A_TILE + G_BN * 128;
  if (p.tile_gate != nullptr && p.tile_gate[blockIdx.y] <= p.t) return;   // every question of this row tile has ended
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * G_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x, m0 = blockIdx.y * BM;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm.a_hi); tma_prefetch_desc(&tm.b_hi); tma_prefetch_desc(&tm.a_lo); tma_prefetch_desc(&tm.b_lo); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < p.num_kb; ++i) {
        const int s = i % S, ph = (i / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], G_STAGE);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          uint8_t* a_dst = smem + s * G_STAGE + pl * PLANE;
          tma_load_2d(a_dst, pl ? &tm.a_lo : &tm.a_hi, &full[s], i * BK, m0);
          tma_load_2d(a_dst + A_TILE, pl ? &tm.b_lo : &tm.b_hi, &full[s], i * BK, u * G_BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(G_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < p.num_kb; ++i) {
        const int s = i % S, ph = (i / S) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * G_STAGE), b_hi = a_hi + A_TILE;
        const uint32_t a_lo = a_hi + PLANE, b_lo = a_lo + A_TILE;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t dah = umma_desc16(a_hi + ks * 32, 0), dbh = umma_desc16(b_hi + ks * 32, 0);
          const uint64_t dal = umma_desc16(a_lo + ks * 32, 0), dbl = umma_desc16(b_lo + ks * 32, 0);
          tc_mma<1>(tmem_base, dal, dbh, idesc, (i > 0 || ks > 0) ? 1u : 0u);   // small terms first (same order as gemm_bf16s_kernel)
          tc_mma<1>(tmem_base, dah, dbl, idesc, 1u);
          tc_mma<1>(tmem_base, dah, dbh, idesc, 1u);
        }
        tc_commit(&empty[s]);
        if (i == p.num_kb - 1) tc_commit(acc_full);
      }
    }
    __syncwarp();
  } else {
    gru_cell_from_tmem(p, tmem_base, acc_full, 0u, warp, lane, u, m0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// The whole recurrence in ONE launch: the step kernel's body inside a loop over t, barriers / TMEM set up once, a grid-wide
// barrier (monotonic counter, cooperative launch: all CTAs are co-resident) between steps instead of a kernel boundary.
struct GruSeq { int T; unsigned* counter; long long gi_step, hp_step, ho_step, hpl_step, gates_step; };

__global__ void __launch_bounds__(THREADS, 1)
gru_seq_kernel(const __grid_constant__ Maps tm, const GruParams p0, const GruSeq sq) {
  constexpr int S = G_S, PLANE = A_TILE + G_BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * G_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x, m0 = blockIdx.y * BM;
  const unsigned nctas = gridDim.x * gridDim.y;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm.a_hi); tma_prefetch_desc(&tm.b_hi); tma_prefetch_desc(&tm.a_lo); tma_prefetch_desc(&tm.b_lo); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int gate = p0.tile_gate ? p0.tile_gate[blockIdx.y] : 0x7fffffff;   // this row tile is live while t < gate (a prefix of the steps)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(G_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  int it = 0;                                              // k-blocks issued / consumed so far (ring position), per role

  for (int t = 0; t < sq.T; ++t) {
    if (t < gate) {
      if (warp == 0) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.global;" ::: "memory");   // h_{t-1} planes were written with generic stores by other CTAs
          for (int i = 0; i < p0.num_kb; ++i, ++it) {
            const int s = it % S, ph = (it / S) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], G_STAGE);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              uint8_t* a_dst = smem + s * G_STAGE + pl * PLANE;
              tma_load_2d(a_dst, pl ? &tm.a_lo : &tm.a_hi, &full[s], i * BK, t * p0.B + m0);   // A map covers all (T+1)*B rows of the h planes
              tma_load_2d(a_dst + A_TILE, pl ? &tm.b_lo : &tm.b_hi, &full[s], i * BK, u * G_BN);
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          for (int i = 0; i < p0.num_kb; ++i, ++it) {
            const int s = it % S, ph = (it / S) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * G_STAGE), b_hi = a_hi + A_TILE;
            const uint32_t a_lo = a_hi + PLANE, b_lo = a_lo + A_TILE;
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              const uint64_t dah = umma_desc16(a_hi + ks * 32, 0), dbh = umma_desc16(b_hi + ks * 32, 0);
              const uint64_t dal = umma_desc16(a_lo + ks * 32, 0), dbl = umma_desc16(b_lo + ks * 32, 0);
              tc_mma<1>(tmem_base, dal, dbh, idesc, (i > 0 || ks > 0) ? 1u : 0u);
              tc_mma<1>(tmem_base, dah, dbl, idesc, 1u);
              tc_mma<1>(tmem_base, dah, dbh, idesc, 1u);
            }
            tc_commit(&empty[s]);
            if (i == p0.num_kb - 1) tc_commit(acc_full);
          }
        }
        __syncwarp();
      } else {
        GruParams pt = p0;                                   // this step's slices
        pt.t = t;
        pt.gi = p0.gi + (long long)t * sq.gi_step;
        pt.h_prev = p0.h_prev + (long long)t * sq.hp_step;
        pt.h_out = p0.h_out + (long long)t * sq.ho_step;
        pt.hout_hi = p0.hout_hi + (long long)t * sq.hpl_step;
        pt.hout_lo = p0.hout_lo + (long long)t * sq.hpl_step;
        pt.gates = p0.gates + (long long)t * sq.gates_step;
        gru_cell_from_tmem(pt, tmem_base, acc_full, (uint32_t)(t & 1), warp, lane, u, m0);   // live steps are a prefix: t-th use of acc_full
        tc_fence_before();
      }
    }
    // ---- grid-wide barrier: every CTA (live or not) has finished step t, its h_t rows are visible device-wide
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(sq.counter, 1u);
      const unsigned target = (unsigned)(t + 1) * nctas;
      const long long t0 = clock64();
      while (*reinterpret_cast<volatile unsigned*>(sq.counter) < target)
        if (clock64() - t0 > 4000000000LL) { printf("vqa_b200: GRU grid barrier timed out (block %d,%d step %d)\n", blockIdx.x, blockIdx.y, t); __trap(); }
      __threadfence();
    }
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

// plane stored either (rows = MN, cols = K contiguous) [k-major] or (rows = K, cols = MN contiguous) [mn-major]
static int make_map(CUtensorMap* tm, const void* ptr, long long ld, int mn_extent, int k_extent, int mn_major, int tile_mn) {
  auto enc = get_encode();
  if (!enc) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) { dims[0] = k_extent; dims[1] = mn_extent; box[0] = BK; box[1] = tile_mn; }
  else           { dims[0] = mn_extent; dims[1] = k_extent; box[0] = 64; box[1] = BK; }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled(bf16 plane) failed with CUresult %d (ld=%lld mn=%d k=%d)", (int)r, ld, mn_extent, k_extent);
  return VQA_OK;
}

template <int BN, int PASSES, int CL>
static int launch(const Maps& tm, const Params& p, int splits, cudaStream_t st) {
  using C = Cfg<BN, PASSES>;
  int mt = (p.M + BM - 1) / BM;
  if (CL == 2) mt = (mt + 1) & ~1;                 // pairs of vertically adjacent tiles; a padding tile only sees zero-filled rows
  static const bool one_tile_per_cta = getenv("VQA_GEMM_ONE_TILE_PER_CTA") != nullptr;   // A/B switch for measurements
  PSched sc;
  sc.tiles_n = (p.N + BN - 1) / BN; sc.tiles_m = mt; sc.units_m = mt / CL; sc.splits = splits;
  sc.units = sc.tiles_n * sc.units_m * splits;
  const int cap = g_vqa_sm_budget / CL;            // one CTA per SM (the double-buffered accumulator takes 2 * BN TMEM columns)
  sc.nclusters = sc.units < cap ? sc.units : cap;
  // A short contraction with a plain fp32 destination is drained faster than it could be overlapped: the one-tile kernel parks the
  // tile in its (free) pipeline shared memory and writes whole lines, the persistent kernel's epilogue stores one row per lane.
  static const int short_kb = getenv("VQA_GEMM_SHORT_KB") ? atoi(getenv("VQA_GEMM_SHORT_KB")) : 16;
  const bool coalescible = p.C && !p.Chi && !p.rowb && !p.aux && !p.auxh && ((p.N & 3) == 0) && ((p.ldc & 3) == 0) &&
                           ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
  const bool short_plain = coalescible && p.kb_per_split <= short_kb;
  if (!one_tile_per_cta && !short_plain && BN >= 128 && sc.units > sc.nclusters) {   // more tiles than SMs: walk them persistently, epilogue overlapped
    // (BN = 64 is the small-problem tile: two co-resident CTAs per SM with shallow rings already overlap each other)
    static bool attr_set_p = false;
    if (!attr_set_p) {
      VQA_CUDA(cudaFuncSetAttribute(gemm_bf16s_persistent_kernel<BN, PASSES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
      attr_set_p = true;
    }
    cudaLaunchConfig_t cfg = {};
    // eight epilogue warps when a tile is accumulated faster than four warps drain it (contraction of at most 8 k-blocks = 512 per split; VQA_GEMM_EPI8_MAX_KB overrides for measurements)
    static const int epi8_max_kb = getenv("VQA_GEMM_EPI8_MAX_KB") ? atoi(getenv("VQA_GEMM_EPI8_MAX_KB")) : 8;
    const bool epi8 = BN >= 128 && p.kb_per_split <= epi8_max_kb;
    cfg.gridDim = dim3(sc.nclusters * CL); cfg.blockDim = dim3(epi8 ? THREADS_EPI8 : THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    VQA_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16s_persistent_kernel<BN, PASSES, CL>, tm, p, sc));
    VQA_LAUNCH_CHECK("gemm_bf16s_persistent_kernel");
    return VQA_OK;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VQA_CUDA(cudaFuncSetAttribute(gemm_bf16s_kernel<BN, PASSES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, mt, splits);
  // one tile per CTA: the drain is never hidden, so the wide tiles always get eight epilogue warps (the recurrence product's 128 x 128
  // tile: 3.5 us of its 12 us with four, tools/gemm_timeline.py); VQA_GEMM_EPI8=0 switches back for measurements
  static const bool epi8_on = !(getenv("VQA_GEMM_EPI8") && atoi(getenv("VQA_GEMM_EPI8")) == 0);
  const int threads = (BN >= 128 && epi8_on) ? THREADS_EPI8 : THREADS;
  if (CL == 1) {
    gemm_bf16s_kernel<BN, PASSES, CL><<<grid, threads, C::SMEM, st>>>(tm, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 2; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    VQA_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16s_kernel<BN, PASSES, CL>, tm, p));
  }
  VQA_LAUNCH_CHECK("gemm_bf16s_kernel");
  return VQA_OK;
}

// ------------------------------------------------------------------------------------------ fp32 -> (hi, lo) planes
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ hi,
                                                   __nv_bfloat16* __restrict__ lo, long long ldp, long long rows, int cols, int vec) {
  const int groups = (cols + 7) >> 3;
  const long long total = rows * groups;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / groups;
    const int c = (int)(g - r * groups) << 3;
    const float* src = x + r * ldx + c;
    float v[8];
    if (vec && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = c + e < cols ? src[e] : 0.f;
    }
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    if (lo)
      *reinterpret_cast<uint4*>(lo + r * ldp + c) = make_uint4(pack_bf16(v[0] - h[0], v[1] - h[1]), pack_bf16(v[2] - h[2], v[3] - h[3]),
                                                               pack_bf16(v[4] - h[4], v[5] - h[5]), pack_bf16(v[6] - h[6], v[7] - h[7]));
  }
}

// fused inverted dropout + split: hi/lo planes of dropout(x) in one pass (x read once, no fp32 intermediate).
// Mask: Philox4x32-10(seed; counter = (row * ceil(cols/4) + col/4, offset)), one call per 4 consecutive columns.
__global__ void __launch_bounds__(256) dropout_split_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, long long ldp, long long rows, int cols,
                                                           float p, float scale, unsigned long long seed, unsigned long long offset,
                                                           const unsigned long long* __restrict__ step_ptr, int vec) {
  const Philox rng(seed);
  if (step_ptr) offset += *step_ptr * 16ull;
  const int groups = (cols + 7) >> 3, q4 = (cols + 3) >> 2;
  const long long total = rows * groups;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / groups;
    const int c = (int)(g - r * groups) << 3;
    const float* src = x + r * ldx + c;
    float v[8];
    if (vec && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = c + e < cols ? src[e] : 0.f;
    }
    const uint4 r0 = rng((unsigned long long)(r * q4 + (c >> 2)), offset), r1 = rng((unsigned long long)(r * q4 + (c >> 2) + 1), offset);
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = u32_to_unit(rr[e]) >= p ? v[e] * scale : 0.f;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    if (lo)
      *reinterpret_cast<uint4*>(lo + r * ldp + c) = make_uint4(pack_bf16(v[0] - h[0], v[1] - h[1]), pack_bf16(v[2] - h[2], v[3] - h[3]),
                                                               pack_bf16(v[4] - h[4], v[5] - h[5]), pack_bf16(v[6] - h[6], v[7] - h[7]));
  }
}

}  // namespace sb
}  // namespace vqa

using namespace vqa;

extern "C" int vqa_split_bf16_f32(const float* x, long long ldx, void* hi, void* lo, long long ldp, long long rows, int cols,
                                  cudaStream_t stream) {
  VQA_CHECK_ARG(x && hi && rows > 0 && cols > 0, "vqa_split_bf16_f32: bad arguments");
  VQA_CHECK_ARG((ldp & 7) == 0 && ldp >= ((cols + 7) & ~7) && aligned16(hi) && (!lo || aligned16(lo)),
                "vqa_split_bf16_f32: planes need 16-byte aligned rows with ld %% 8 == 0 and ld >= cols rounded up to 8 (ldp=%lld cols=%d)", ldp, cols);
  const int vec = aligned16(x) && (ldx & 3) == 0;
  const long long groups = rows * ((cols + 7) >> 3);
  long long blocks = (groups + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sb::split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ldp, rows, cols, vec);
  VQA_LAUNCH_CHECK("split_kernel");
  return VQA_OK;
}

extern "C" int vqa_gemm_bf16s(const void* A_hi, const void* A_lo, long long lda, int a_mn_major, const void* B_hi,
                              const void* B_lo, long long ldb, int b_mn_major, float* C, long long ldc, void* C_hi, void* C_lo,
                              long long ldcs, int M, int N, int Kc, const float* bias, const float* rowbcast, long long ldrb,
                              int group, const float* aux, long long ldaux, const void* aux_hi, long long ldauxh,
                              float aux_scale, int flags, int passes, int split_k, int tile_n, const int* tile_gate, int gate_t,
                              cudaStream_t stream) {
  const char* who = "vqa_gemm_bf16s";
  VQA_CHECK_ARG(A_hi && B_hi && (C || C_hi), "%s: null operand", who);
  VQA_CHECK_ARG(passes == 1 || passes == 3, "%s: passes must be 1 (bf16) or 3 (split-bf16, fp32-grade), got %d", who, passes);
  VQA_CHECK_ARG(passes == 1 || (A_lo && B_lo), "%s: 3-pass mode needs the lo planes", who);
  VQA_CHECK_ARG(M > 0 && N > 0 && Kc > 0, "%s: empty problem M=%d N=%d K=%d", who, M, N, Kc);
  VQA_CHECK_ARG(aligned16(A_hi) && aligned16(B_hi) && (!A_lo || aligned16(A_lo)) && (!B_lo || aligned16(B_lo)), "%s: planes must be 16-byte aligned for TMA", who);
  VQA_CHECK_ARG((lda & 7) == 0 && (ldb & 7) == 0, "%s: plane leading dimensions must be multiples of 8 bf16 (TMA 16-byte strides), got lda=%lld ldb=%lld", who, lda, ldb);
  VQA_CHECK_ARG(lda >= (a_mn_major ? M : Kc) && ldb >= (b_mn_major ? N : Kc) && (!C || ldc >= N) && (!C_hi || ldcs >= N), "%s: leading dimension smaller than the row length", who);
  VQA_CHECK_ARG(!rowbcast || group > 0, "%s: rowbcast needs group > 0", who);
  VQA_CHECK_ARG(!(aux && aux_hi), "%s: give the mask either as fp32 or as a bf16 hi plane, not both", who);
  const int num_kb = (Kc + sb::BK - 1) / sb::BK;
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > num_kb) splits = num_kb;
  int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;   // no empty split
  if (splits > 1 || (flags & VQA_GEMM_ACCUMULATE)) {
    VQA_CHECK_ARG(C && !C_hi && !bias && !rowbcast && !aux && !aux_hi && !(flags & VQA_GEMM_RELU), "%s: split-K / accumulate support the plain fp32 epilogue only", who);
    flags |= VQA_GEMM_ATOMIC_ADD;      // caller zero-fills (split-K) or pre-loads (accumulate) C
  }
  int bn = tile_n;
  if (bn == 0) {                       // widest tile that still gives the 148 SMs enough CTAs
    const long long mt = (M + sb::BM - 1) / sb::BM;
    auto ctas = [&](int w) { return mt * ((N + w - 1) / w) * splits; };
    bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
    if (bn == 256 && ctas(256) < 120) bn = 128;
    if (bn == 128 && ctas(128) < 120) bn = 64;
  }
  VQA_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "%s: tile_n must be 64, 128 or 256", who);

  // CTA pairs (cluster 1x2x1) with the B tile multicast to both: the big projections are L2->SM bandwidth bound (96 KB per
  // k-block and CTA at BN = 256), sharing B cuts that to 64 KB.  Worth it only when there are plenty of tile rows.
  const int mtiles = (M + sb::BM - 1) / sb::BM;
  int cl = (bn == 256 && mtiles >= 8 && !(flags & VQA_GEMM_NO_CLUSTER) && !tile_gate) ? 2 : 1;
  sb::Maps tm;
  memset(&tm, 0, sizeof(tm));
  const int b_box = cl == 2 ? bn / 2 : bn;
  int rc = sb::make_map(&tm.a_hi, A_hi, lda, M, Kc, a_mn_major, sb::BM);
  if (!rc) rc = sb::make_map(&tm.b_hi, B_hi, ldb, N, Kc, b_mn_major, b_box);
  if (!rc && passes == 3) rc = sb::make_map(&tm.a_lo, A_lo, lda, M, Kc, a_mn_major, sb::BM);
  if (!rc && passes == 3) rc = sb::make_map(&tm.b_lo, B_lo, ldb, N, Kc, b_mn_major, b_box);
  if (rc) return rc;
  sb::Params p{C, ldc, reinterpret_cast<__nv_bfloat16*>(C_hi), reinterpret_cast<__nv_bfloat16*>(C_lo), ldcs, M, N, Kc,
               a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, bias, rowbcast, ldrb, group, aux, ldaux,
               reinterpret_cast<const __nv_bfloat16*>(aux_hi), ldauxh, aux_scale, flags, per, num_kb, tile_gate, gate_t};
#define VQA_DISPATCH(BN_) (passes == 3 ? sb::launch<BN_, 3, 1>(tm, p, splits, stream) : sb::launch<BN_, 1, 1>(tm, p, splits, stream))
  if (bn == 256 && cl == 2) return passes == 3 ? sb::launch<256, 3, 2>(tm, p, splits, stream) : sb::launch<256, 1, 2>(tm, p, splits, stream);
  if (bn == 256) return VQA_DISPATCH(256);
  if (bn == 128) return VQA_DISPATCH(128);
  return VQA_DISPATCH(64);
#undef VQA_DISPATCH
}

extern "C" int vqa_dropout_split_f32(const float* x, long long ldx, void* hi, void* lo, long long ldp, long long rows, int cols, float p,
                                     unsigned long long seed, unsigned long long offset, const unsigned long long* step_ptr,
                                     cudaStream_t stream) {
  VQA_CHECK_ARG(x && hi && rows > 0 && cols > 0, "vqa_dropout_split_f32: bad arguments");
  VQA_CHECK_ARG(p >= 0.f && p < 1.f, "vqa_dropout_split_f32: p must be in [0,1)");
  VQA_CHECK_ARG((ldp & 7) == 0 && ldp >= ((cols + 7) & ~7) && aligned16(hi) && (!lo || aligned16(lo)),
                "vqa_dropout_split_f32: planes need 16-byte aligned rows with ld %% 8 == 0 and ld >= cols rounded up to 8");
  const int vec = aligned16(x) && (ldx & 3) == 0;
  const long long groups = rows * ((cols + 7) >> 3);
  long long blocks = (groups + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sb::dropout_split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, ldx, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo),
                                                                ldp, rows, cols, p, 1.f / (1.f - p), seed, offset, step_ptr, vec);
  VQA_LAUNCH_CHECK("dropout_split_kernel");
  return VQA_OK;
}

extern "C" int vqa_gru_step_fused(const void* hprev_hi, const void* hprev_lo, long long ldh, const void* Whh_hi, const void* Whh_lo,
                                  long long ldw, const float* gi, long long ldgi, const float* b_hh, const float* h_prev, const int* len,
                                  int t, float* h_out, void* hout_hi, void* hout_lo, long long ldp, float* gates, const int* tile_gate,
                                  int B, int H, cudaStream_t stream) {
  const char* who = "vqa_gru_step_fused";
  VQA_CHECK_ARG(hprev_hi && hprev_lo && Whh_hi && Whh_lo && gi && b_hh && h_prev && len && h_out && hout_hi && hout_lo && gates, "%s: null pointer", who);
  VQA_CHECK_ARG(B > 0 && H > 0 && H % 32 == 0, "%s: H must be a multiple of 32 (H=%d)", who, H);
  VQA_CHECK_ARG((ldh & 7) == 0 && (ldw & 7) == 0 && (ldp & 7) == 0 && ldh >= H && ldw >= H && ldp >= H && (ldgi & 3) == 0 && ldgi >= 3 * H, "%s: bad leading dimensions", who);
  VQA_CHECK_ARG(aligned16(hprev_hi) && aligned16(hprev_lo) && aligned16(Whh_hi) && aligned16(Whh_lo) && aligned16(gi) && aligned16(b_hh) &&
                aligned16(h_prev) && aligned16(h_out) && aligned16(hout_hi) && aligned16(hout_lo) && aligned16(gates), "%s: pointers must be 16-byte aligned", who);
  sb::Maps tm;
  memset(&tm, 0, sizeof(tm));
  int rc = sb::make_map(&tm.a_hi, hprev_hi, ldh, B, H, 0, sb::BM);
  if (!rc) rc = sb::make_map(&tm.a_lo, hprev_lo, ldh, B, H, 0, sb::BM);
  if (!rc) rc = sb::make_map(&tm.b_hi, Whh_hi, ldw, 3 * H, H, 0, sb::G_BN);
  if (!rc) rc = sb::make_map(&tm.b_lo, Whh_lo, ldw, 3 * H, H, 0, sb::G_BN);
  if (rc) return rc;
  sb::GruParams p{gi, ldgi, b_hh, h_prev, len, t, h_out, reinterpret_cast<__nv_bfloat16*>(hout_hi), reinterpret_cast<__nv_bfloat16*>(hout_lo),
                  ldp, gates, tile_gate, B, H, (H + sb::BK - 1) / sb::BK};
  static bool attr_set = false;
  if (!attr_set) {
    VQA_CUDA(cudaFuncSetAttribute(sb::gru_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sb::G_SMEM));
    attr_set = true;
  }
  dim3 grid(H / 32, (B + sb::BM - 1) / sb::BM);
  sb::gru_step_kernel<<<grid, sb::THREADS, sb::G_SMEM, stream>>>(tm, p);
  VQA_LAUNCH_CHECK("gru_step_kernel");
  return VQA_OK;
}

extern "C" int vqa_gru_seq_fused(const void* H_hi, const void* H_lo, long long ldh, const void* Whh_hi, const void* Whh_lo, long long ldw,
                                 const float* gi, long long ldgi, const float* b_hh, float* Hall, const int* len, float* gates,
                                 const int* tile_gate, unsigned* counter, int T, int B, int H, cudaStream_t stream) {
  const char* who = "vqa_gru_seq_fused";
  VQA_CHECK_ARG(H_hi && H_lo && Whh_hi && Whh_lo && gi && b_hh && Hall && len && gates && counter, "%s: null pointer", who);
  VQA_CHECK_ARG(T > 0 && B > 0 && H > 0 && H % 32 == 0, "%s: H must be a multiple of 32 (H=%d)", who, H);
  VQA_CHECK_ARG((ldh & 7) == 0 && (ldw & 7) == 0 && ldh >= H && ldw >= H && (ldgi & 3) == 0 && ldgi >= 3 * H, "%s: bad leading dimensions", who);
  VQA_CHECK_ARG(aligned16(H_hi) && aligned16(H_lo) && aligned16(Whh_hi) && aligned16(Whh_lo) && aligned16(gi) && aligned16(b_hh) && aligned16(Hall) && aligned16(gates),
                "%s: pointers must be 16-byte aligned", who);
  dim3 grid(H / 32, (B + sb::BM - 1) / sb::BM);
  if ((long long)grid.x * grid.y > kNumSMs)
    return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: %u x %u CTAs cannot be co-resident on %d SMs (the grid barrier needs that): launch the steps one by one", who, grid.x, grid.y, kNumSMs);
  sb::Maps tm;
  memset(&tm, 0, sizeof(tm));
  int rc = sb::make_map(&tm.a_hi, H_hi, ldh, (T + 1) * B, H, 0, sb::BM);      // rows t*B .. : h_{t-1}; rows (t+1)*B .. : h_t
  if (!rc) rc = sb::make_map(&tm.a_lo, H_lo, ldh, (T + 1) * B, H, 0, sb::BM);
  if (!rc) rc = sb::make_map(&tm.b_hi, Whh_hi, ldw, 3 * H, H, 0, sb::G_BN);
  if (!rc) rc = sb::make_map(&tm.b_lo, Whh_lo, ldw, 3 * H, H, 0, sb::G_BN);
  if (rc) return rc;
  __nv_bfloat16* hh = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(H_hi));
  __nv_bfloat16* hl = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(H_lo));
  sb::GruParams p{gi, ldgi, b_hh, Hall, len, 0, Hall + (long long)B * H, hh + (long long)B * ldh, hl + (long long)B * ldh, ldh, gates, tile_gate,
                  B, H, (H + sb::BK - 1) / sb::BK};
  sb::GruSeq sq{T, counter, (long long)B * ldgi, (long long)B * H, (long long)B * H, (long long)B * ldh, (long long)B * 4 * H};
  static bool attr_set = false;
  if (!attr_set) {
    VQA_CUDA(cudaFuncSetAttribute(sb::gru_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sb::G_SMEM));
    attr_set = true;
  }
  VQA_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(sb::THREADS); cfg.dynamicSmemBytes = sb::G_SMEM; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  VQA_CUDA(cudaLaunchKernelEx(&cfg, sb::gru_seq_kernel, tm, p, sq));
  VQA_LAUNCH_CHECK("gru_seq_kernel");
  return VQA_OK;
}

#ifdef VQA_GEMM_TRACE
extern "C" int vqa_debug_gemm_trace(long long* out, int n_ctas) {
  if (n_ctas > 512) n_ctas = 512;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, sb::g_gemm_trace, (size_t)n_ctas * 8 * sizeof(long long)) == cudaSuccess ? VQA_OK : VQA_ERR_CUDA;
}
#endif
