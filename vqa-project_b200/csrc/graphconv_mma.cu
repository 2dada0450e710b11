// Fused graph convolution with the per-image aggregate on the tensor cores (sm_100a), project-first formulation.
//
//   out[b,i, chunk k] = act( sum_j M_k[b][i,j] * Y[b, j, chunk k] ),   M_k[i, idx[i,m]] = w[i,m,k] * alpha[i,m]
//
// The reference forms this product with torch.bmm over materialised neighbourhoods (layers.py:136-137, after
// sparse_graph_model.py:161-195 expanded + gathered (B,K,nb,F)).  Here the K x K coefficient matrix of every Gaussian
// kernel is built in shared memory from the box centres, the top-nb ids and alpha (Gaussian weights only for the
// B*K*nb selected edges: sparse_graph_model.py:244-269, layers.py:100-125), and each [K rows x 128 columns] tile of Y
// is multiplied by it with tcgen05.mma.  Operands are the split-bf16 planes the projections already produce
// (gemm_bf16s.cu): three passes lo*hi + hi*lo + hi*hi accumulate in fp32 TMEM (fp32-grade, ~1e-5), so the arithmetic
// costs ~6 us per launch and the kernel is purely HBM-bound: Y is read once by TMA, the result written once by TMA.
//
// Why not CUDA cores: measured on B200, register-operand FFMA/FFMA2 peak at 21 TFMA/s (57 % of nominal), which puts the
// dense-in-register formulation at >= 65 us and the shared-memory gather at >= 67 us for layer 1 (B=512) against 46 us
// of HBM time -- neither can reach the roofline target.  (Both remain in graphconv.cu as the generic fallback.)
//
// Tile mapping (per CTA = image b, slab of column tiles):  D[128 columns c, NP rows i] (+)= A[c, j] * B[i, j]
//   A = Y^T tile, MN-major (c contiguous), rows j = 0..KP-1 straight from a 2-D TMA box (rows past K belong to the next
//       image or are zero-filled; their coefficients are zero), 128-byte swizzle;
//   B = M_k, K-major (j contiguous), written by the prologue threads in the same swizzled layout, hi and lo planes;
//   D in TMEM: lane = column c, TMEM column = node i  ->  the epilogue thread of lane c owns one output column for all
//       nodes: ReLU / dropout / max-pool + gate are per-thread, and a warp writes 32 consecutive columns of a node row.
//
// Kernels in this file: edge_coef_kernel (everything that depends only on the selected edges, once per layer and step),
// agg_persistent_kernel (the streaming aggregate, forward / pooled forward / transposed backward: one warp-specialised CTA
// per SM), agg_kernel (the same per (image, slab) CTA, used when no precomputed edge coefficients are passed or the shapes do
// not fit the persistent kernel's shared-memory plan), pool_bwd_data_kernel (backward data path of the max-pooled layer),
// edge_p_kernel (edge products of the backward, persistent) and edge_finish_kernel (what depends on them).
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <mutex>
#include <cstring>
#include "../../include/vqa_b200.h"

namespace vqa {
namespace gm {

constexpr int THREADS = 192;
constexpr int MT = 128;                // columns per M-tile
constexpr int MAX_NK = 64;
#define GM_TWO_PI_F 6.28318530717958647692f
#define GM_EPS_F 1e-14f

enum { AGG_FWD = 0, AGG_FWD_POOL = 1, AGG_BWD = 2 };

struct Maps { CUtensorMap in_hi, in_lo, out_hi, out_lo; };

struct AggParams {
  const int* idx; const float* alpha; const float* boxes; long long ldbox; const float* gauss;
  const float* coef; const unsigned* eoff;
  const float* q; float* pooled; long long* argmax; float* hq;
  int B, K, KP, nb, nk, out_dim, D, nkc, tiles_per_cta, ntiles, nstage, flags, with_lo;
  float drop_scale; unsigned drop_thresh16; unsigned long long seed, offset; const unsigned long long* step_ptr;
  int off_coef, coef_plane, off_stage, stage_bytes, off_out, out_plane, off_misc, off_bars, tmem_cols;
};

// One Gaussian kernel value with precomputed cr = -0.5*log2(e)/(eps + sigma_rho^2), ct likewise for theta:
//   exp(-0.5 (rho-mr)^2 / (eps+sr^2)) * exp(-0.5 d_theta^2 / (eps+st^2)) = 2^( (rho-mr)^2 cr + d_theta^2 ct )
// (layers.py:109-117; one ex2 instead of two exp and two divisions; |rel err| ~ 2^-22).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gauss_val(float rho, float theta, float mr, float cr, float mt, float ct) {
  const float d = rho - mr;
  const float a1 = fabsf(theta - mt);
  const float mn = fminf(a1, fabsf(GM_TWO_PI_F - a1));
  const float g = ex2_approx(fmaf(d * d, cr, mn * mn * ct));
  return (g != g) ? 0.f : g;      // NaN -> 0 BEFORE the kernel-axis normalisation (layers.py:120)
}

// UMMA shared-memory descriptors (16-bit operands, 128-byte swizzle)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {            // rows of 128 B, 8-row atoms 1024 B apart
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {   // 64-element MN chunks lbo apart
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// lowbias32 avalanche hash (2 multiplies): one call per output element for the fused dropout
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// byte offset of element (n, k) inside one K-major, 128B-swizzled coefficient plane with NP rows
__device__ __forceinline__ uint32_t coef_off(int n, int k, int NP) {
  const int chunk = k >> 6, kc = k & 63;
  return (uint32_t)(chunk * NP * 128 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kc >> 3) ^ (n & 7))) << 4) + ((kc & 7) << 1));
}

template <int MODE>
__global__ void __launch_bounds__(THREADS)
agg_kernel(const __grid_constant__ Maps tm, const AggParams p) {
  extern __shared__ uint8_t gsm_raw[];
  uint8_t* sm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, K = p.K, KP = p.KP, NP = p.KP, nb = p.nb, nk = p.nk;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, p.ntiles - t0);
  const int k_lo = (t0 * MT) / p.D;
  const int nkc = ((t0 + nt) * MT - 1) / p.D - k_lo + 1;
  const int S = p.nstage;
  const int planes = p.with_lo ? 2 : 1;

  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + p.off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  uint8_t* idx8 = sm + p.off_misc;
  float* cen = reinterpret_cast<float*>(sm + p.off_misc + ((K * nb + 15) & ~15));
  float* gs = cen + 2 * ((K + 1) & ~1);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.in_hi);
    if (p.with_lo) tma_prefetch_desc(&tm.in_lo);
    if (MODE != AGG_FWD_POOL) { tma_prefetch_desc(&tm.out_hi); if (p.with_lo) tma_prefetch_desc(&tm.out_lo); }
  }
  // ---- stage neighbour ids / box centres / Gaussian parameters
  for (int v = tid; v < K * nb; v += THREADS) idx8[v] = (uint8_t)p.idx[(long long)b * K * nb + v];
  for (int i = tid; i < K; i += THREADS) {
    const float* bx = p.boxes + ((long long)b * K + i) * p.ldbox;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    cen[2 * i] = x1 + 0.5f * (x2 - x1);                 // sparse_graph_model.py:106-108
    cen[2 * i + 1] = y1 + 0.5f * (y2 - y1);
  }
  for (int k = tid; k < nk; k += THREADS) {
    const float sr = p.gauss[nk + k], st = p.gauss[3 * nk + k];
    gs[k] = p.gauss[k];
    gs[nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + sr * sr);
    gs[2 * nk + k] = p.gauss[2 * nk + k];
    gs[3 * nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + st * st);
  }
  // zero the coefficient planes of this slab
  {
    uint4* cz = reinterpret_cast<uint4*>(sm + p.off_coef);
    const int n16 = nkc * planes * p.coef_plane / 16;
    for (int v = tid; v < n16; v += THREADS) cz[v] = make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- first input tiles in flight while the coefficients are built
  auto issue = [&](int t) {
    const int s = t % S;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(planes * 2 * KP * 128));
    uint8_t* dst = sm + p.off_stage + (size_t)s * p.stage_bytes;
    const int c0 = (t0 + t) * MT;
    for (int pl = 0; pl < planes; ++pl) {
      const CUtensorMap* m = pl ? &tm.in_lo : &tm.in_hi;
      tma_load_2d(dst + (pl * 2 + 0) * KP * 128, m, &full[s], c0, b * K);
      tma_load_2d(dst + (pl * 2 + 1) * KP * 128, m, &full[s], c0 + 64, b * K);
    }
  };
  if (warp == 0 && lane == 0)
    for (int t = 0; t < S && t < nt; ++t) issue(t);

  // ---- per-edge Gaussian weights -> coefficient matrices (hi / lo bf16, swizzled K-major UMMA layout)
  for (int e = tid; e < K * nb; e += THREADS) {
    const int i = e / nb, j = idx8[e];
    const float dx = cen[2 * i] - cen[2 * j], dy = cen[2 * i + 1] - cen[2 * j + 1];   // centre_i - centre_j (sparse_graph_model.py:258-259)
    const float rho = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float theta = atan2f(dx, dy);                                              // x FIRST (sparse_graph_model.py:264-265)
    float Ssum = 0.f;
    for (int k = 0; k < nk; ++k) Ssum += gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
    const float a_over_S = __fdiv_rn(p.alpha ? p.alpha[(long long)b * K * nb + e] : 1.f, Ssum);   // Ssum == 0 -> inf/NaN, as the reference
    const int n = MODE == AGG_BWD ? j : i, kc = MODE == AGG_BWD ? i : j;   // bwd: dY[j] += c * dO[i]  (transposed matrix)
    const uint32_t off = coef_off(n, kc, NP);
    for (int kk = 0; kk < nkc; ++kk) {
      const int k = k_lo + kk;
      const float c = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]) * a_over_S;
      const __nv_bfloat16 h = __float2bfloat16_rn(c);
      uint8_t* base = sm + p.off_coef + (size_t)(kk * planes) * p.coef_plane;
      *reinterpret_cast<__nv_bfloat16*>(base + off) = h;
      if (p.with_lo) *reinterpret_cast<__nv_bfloat16*>(base + p.coef_plane + off) = __float2bfloat16_rn(c - __bfloat162float(h));
    }
  }
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (remaining tiles)
    if (lane == 0) {
      for (int t = S; t < nt; ++t) {
        mbar_wait(&empty[t % S], ((t / S) & 1) ^ 1);
        issue(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
    for (int t = 0; t < nt; ++t) {
      const int s = t % S, acc = t & 1;
      const int kk = ((t0 + t) * MT) / p.D - k_lo;
      mbar_wait(&tempty[acc], ((t >> 1) & 1) ^ 1);
      mbar_wait(&full[s], (t / S) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(sm + p.off_stage + (size_t)s * p.stage_bytes), a_lo = a_hi + 2 * KP * 128;
        const uint32_t b_hi = smem_u32(sm + p.off_coef + (size_t)(kk * planes) * p.coef_plane), b_lo = b_hi + p.coef_plane;
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NP);
        const uint32_t lbo = (uint32_t)KP * 128;
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint32_t ao = ks * 2048, bo = (ks >> 2) * NP * 128 + (ks & 3) * 32;
          const uint64_t dah = desc_mnmajor(a_hi + ao, lbo), dbh = desc_kmajor(b_hi + bo);
          if (p.with_lo) {
            const uint64_t dal = desc_mnmajor(a_lo + ao, lbo), dbl = desc_kmajor(b_lo + bo);
            tc_mma<1>(d_tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
            tc_mma<1>(d_tmem, dah, dbl, idesc, 1u);
            tc_mma<1>(d_tmem, dah, dbh, idesc, 1u);
          } else {
            tc_mma<1>(d_tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit(&empty[s]);
        tc_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ warps 2-5: epilogue, thread = output column
    const int q4 = warp & 3;                              // TMEM lane quarter this warp may access
    const int cl = q4 * 32 + lane;                        // column inside the M-tile
    const unsigned long long rng_off = p.offset + (p.step_ptr ? *p.step_ptr * 16ull : 0ull);
    const uint32_t key = hash32((uint32_t)p.seed ^ hash32((uint32_t)(p.seed >> 32) ^ hash32((uint32_t)rng_off * 0x9E3779B1u + 0x85EBCA77u)));
    const bool leader = (warp == 2 && lane == 0);
    for (int t = 0; t < nt; ++t) {
      const int acc = t & 1;
      const int col = (t0 + t) * MT + cl;                 // global output column
      __nv_bfloat16* st_hi = reinterpret_cast<__nv_bfloat16*>(sm + p.off_out + (size_t)(t & 1) * planes * p.out_plane);
      __nv_bfloat16* st_lo = st_hi + p.out_plane / 2;
      if (MODE != AGG_FWD_POOL) {
        if (leader) bulk_wait_read<1>();                  // the stores that used this staging buffer two tiles ago have read it
        epi_bar_sync();
      }
      mbar_wait(&tfull[acc], (t >> 1) & 1);
      tc_fence_after();
      float best = -1.f; int barg = 0;
      const bool drop = MODE == AGG_FWD && p.drop_thresh16 != 0;
      const bool relu = MODE == AGG_FWD && (p.flags & VQA_GC_RELU);
      // dropout: one counter hash per (pair of node rows, column); low / high 16 bits decide the even / odd row
      const uint32_t ctr0 = ((uint32_t)b * (uint32_t)((K + 1) >> 1)) * (uint32_t)p.out_dim + (uint32_t)col;
      for (int i0 = 0; i0 < K; i0 += 16) {
        uint32_t r[16];
        tc_ld_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * NP + i0), r);
        tc_wait_ld();
        const bool full16 = i0 + 16 <= K;
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int i = i0 + e;
          if (full16 || i < K) {
            float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]);
            const bool ok1 = full16 || i + 1 < K;
            if (MODE == AGG_FWD_POOL) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);
              if (v0 > best) { best = v0; barg = i; }     // strict > : first index on ties
              if (ok1 && v1 > best) { best = v1; barg = i + 1; }
            } else {
              if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              if (drop) {
                const uint32_t rnd = hash32((ctr0 + (uint32_t)(i >> 1) * (uint32_t)p.out_dim) ^ key);
                v0 = (rnd & 0xFFFFu) >= p.drop_thresh16 ? v0 * p.drop_scale : 0.f;
                v1 = (rnd >> 16) >= p.drop_thresh16 ? v1 * p.drop_scale : 0.f;
              }
              const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
              st_hi[i * MT + cl] = h0;
              if (ok1) st_hi[(i + 1) * MT + cl] = h1;
              if (p.with_lo) {
                st_lo[i * MT + cl] = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                if (ok1) st_lo[(i + 1) * MT + cl] = __float2bfloat16_rn(v1 - __bfloat162float(h1));
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);                          // accumulator drained
      if (MODE == AGG_FWD_POOL) {
        const long long o = (long long)b * p.out_dim + col;
        p.pooled[o] = best;
        p.argmax[o] = barg;
        p.hq[o] = fmaxf(p.q[o], 0.f) * best;
      } else {
        fence_proxy_async();                              // staging writes -> visible to the TMA store
        epi_bar_sync();
        if (leader) {
          tma_store_2d(&tm.out_hi, st_hi, (t0 + t) * MT, b * K);
          if (p.with_lo) tma_store_2d(&tm.out_lo, st_lo, (t0 + t) * MT, b * K);
          bulk_commit();
        }
      }
    }
    if (MODE != AGG_FWD_POOL && leader) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// Epilogue work for two consecutive node rows (i, i+1) of one output column: optional ReLU, optional dropout (rnd: one
// 32-bit hash, low / high half decide row i / i+1; thresh_hi = threshold << 16), split into hi / lo bf16 and staged.
template <bool WITH_LO, bool DROP, bool RELU>
__device__ __forceinline__ void epi_pair(float v0, float v1, bool ok1, uint32_t rnd, uint32_t thresh_hi, __nv_bfloat16* hi0,
                                         __nv_bfloat16* lo0) {
  if (RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
  if (DROP) {                                             // the 1/(1-p) scale is folded into the coefficient matrix
    v0 = (rnd << 16) >= thresh_hi ? v0 : 0.f;
    v1 = rnd >= thresh_hi ? v1 : 0.f;
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);   // one F2FP for both conversions
  const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
  *reinterpret_cast<uint16_t*>(hi0) = (uint16_t)hb;
  if (ok1) *reinterpret_cast<uint16_t*>(hi0 + MT) = (uint16_t)(hb >> 16);
  if (WITH_LO) {
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xFFFF0000u));
    const uint32_t lb = *reinterpret_cast<const uint32_t*>(&l);
    *reinterpret_cast<uint16_t*>(lo0) = (uint16_t)lb;
    if (ok1) *reinterpret_cast<uint16_t*>(lo0 + MT) = (uint16_t)(lb >> 16);
  }
}

// ------------------------------------------------------------------------------------------ edge coefficients
// One thread per selected edge e = (b, i, m), j = idx[b,i,m]:  polar pseudo-coordinates of centre_i - centre_j
// (sparse_graph_model.py:106-108,258-267), the nk Gaussian weights normalised over the KERNEL axis (layers.py:109-123) and
// the edge weight alpha (sparse_graph_model.py:239-240):  coef[e][k] = w[e][k] * alpha[e].  Also the byte offsets of the
// edge's entry inside a swizzled K-major coefficient plane, forward (row i, column j) in the low half-word and transposed
// (row j, column i) in the high one.  Computed once per layer and step; the forward aggregate and the backward-data
// aggregate both read it, so no transcendental is evaluated inside the streaming kernels.
__global__ void __launch_bounds__(256)
edge_coef_kernel(const int* __restrict__ idx, const float* __restrict__ alpha, const float* __restrict__ boxes, long long ldbox,
                 const float* __restrict__ gauss, float* __restrict__ coef, unsigned* __restrict__ eoff, long long n_edges, int K, int nb,
                 int nk, int KP) {
  __shared__ float gs[4 * MAX_NK];
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const float sr = gauss[nk + k], st = gauss[3 * nk + k];
    gs[k] = gauss[k];
    gs[nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + sr * sr);
    gs[2 * nk + k] = gauss[2 * nk + k];
    gs[3 * nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + st * st);
  }
  __syncthreads();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const long long row = e / nb;                       // b * K + i
  const int i = (int)(row % K), j = idx[e];
  const float* bi = boxes + row * ldbox;
  const float* bj = boxes + (row - i + j) * ldbox;
  const float xi = bi[0] + 0.5f * (bi[2] - bi[0]), yi = bi[1] + 0.5f * (bi[3] - bi[1]);   // sparse_graph_model.py:106-108
  const float xj = bj[0] + 0.5f * (bj[2] - bj[0]), yj = bj[1] + 0.5f * (bj[3] - bj[1]);
  const float dx = xi - xj, dy = yi - yj;                                                   // centre_i - centre_j (:258-259)
  const float rho = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  const float theta = atan2f(dx, dy);                                                       // x FIRST (:264-265)
  float Ssum = 0.f;
  for (int k = 0; k < nk; ++k) Ssum += gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
  const float a_over_S = __fdiv_rn(alpha ? alpha[e] : 1.f, Ssum);    // Ssum == 0 -> inf/NaN, as the reference
  float* c = coef + e * nk;
  for (int k = 0; k < nk; ++k) c[k] = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]) * a_over_S;
  eoff[e] = coef_off(i, j, KP) | (coef_off(j, i, KP) << 16);
}

// ------------------------------------------------------------------------------------------ backward of the pooled layer, data path
// Layer 2 ends in a max over the nodes (sparse_graph_model.py:150): its upstream gradient is one value per (image, column)
// at node a = argmax[b,c], so the transposed aggregate  dY[b,j,c] = sum_i M_k[i,j] dO[b,i,c]  collapses to
//   dY[b,j,c] = coef[b,a,m,k(c)] * dpooled[b,c]  if j = idx[b,a,m] for some m,   else 0
// - no contraction at all.  One CTA per image: ids -> inverse table pos[a][j] = m, coefficients, arg-max nodes and upstream
// values staged in shared memory; every thread then produces 8 consecutive columns of one output row per iteration and
// writes them as hi / lo bf16 with 16-byte stores (pure streaming write of the planes, nothing is zero-filled or scattered).
constexpr int PB_THREADS = 256;
__global__ void __launch_bounds__(PB_THREADS)
pool_bwd_data_kernel(const float* __restrict__ dpooled, const long long* __restrict__ argmax, const int* __restrict__ idx,
                     const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy_hi, __nv_bfloat16* __restrict__ dy_lo, long long ldy,
                     int K, int nb, int nk, int out_dim, int D, int stage_coef) {
  extern __shared__ __align__(16) uint8_t pb_sm[];
  float* dp_s = reinterpret_cast<float*>(pb_sm);                          // [out_dim]
  uint8_t* a_s = reinterpret_cast<uint8_t*>(dp_s + out_dim);              // [out_dim]
  int8_t* pos = reinterpret_cast<int8_t*>(a_s + ((out_dim + 15) & ~15));  // [K][K]: slot of j in node a's neighbour list, -1 if absent
  float* cf = reinterpret_cast<float*>(pos + ((K * K + 15) & ~15));       // [K*nb][nk] (only if stage_coef)
  const int b = blockIdx.x, tid = threadIdx.x, E = K * nb;
  const float* cg = coef + (long long)b * E * nk;
  for (int v = tid; v < out_dim; v += PB_THREADS) {
    dp_s[v] = dpooled[(long long)b * out_dim + v];
    a_s[v] = (uint8_t)argmax[(long long)b * out_dim + v];
  }
  for (int v = tid; v < K * K; v += PB_THREADS) pos[v] = -1;
  if (stage_coef)
    for (int v = tid; v < E * nk / 4; v += PB_THREADS) reinterpret_cast<float4*>(cf)[v] = reinterpret_cast<const float4*>(cg)[v];
  __syncthreads();
  for (int e = tid; e < E; e += PB_THREADS) pos[(e / nb) * K + idx[(long long)b * E + e]] = (int8_t)(e % nb);   // distinct ids per node
  __syncthreads();
  const float* cfp = stage_coef ? cf : cg;
  const int cpr = out_dim >> 3;                                           // 16-byte chunks per row
  for (int v = tid; v < K * cpr; v += PB_THREADS) {
    const int j = v / cpr, c0 = (v - j * cpr) * 8;
    const int k = c0 / D;                                                 // D % 8 == 0: the 8 columns share their kernel
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int a = a_s[c0 + u], m = pos[a * K + j];
      x[u] = m >= 0 ? cfp[(a * nb + m) * nk + k] * dp_s[c0 + u] : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
      h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * e] - __uint_as_float(h[e] << 16), x[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u));
      l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const long long go = ((long long)b * K + j) * ldy + c0;
    *reinterpret_cast<uint4*>(dy_hi + go) = make_uint4(h[0], h[1], h[2], h[3]);
    if (dy_lo) *reinterpret_cast<uint4*>(dy_lo + go) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------ persistent aggregate
// Same maths and tile mapping as agg_kernel, restructured so that no phase waits for another: one persistent CTA per SM
// walks a contiguous range of work items (image, slab of kernels) and every phase has its own warps, all handshakes are
// mbarriers polled by ONE lane per warp:
//   warp 0       TMA producer (Y tiles, multi-stage ring across items)
//   warps 1-3    MMA issuers, tile g belongs to issuer g % 3 (issuing one of these small MMAs costs a single thread ~130
//                cycles - measured - so one issuer cannot keep up with HBM); warp 1 owns the TMEM allocation (4 accumulators)
//   warps 4-15   epilogue: three warps per TMEM lane quarter split the node rows; ReLU / dropout / hi-lo split / staging
//                (latency-bound instruction streams: more warps, not fewer instructions, is what buys throughput here)
//   warps 16-19  coefficient builders, one item ahead: scatter the precomputed edge coefficients (edge_coef_kernel) of the
//                item's kernels into the swizzled hi/lo planes (double-buffered); next item's values prefetched in registers
//   warp 20      TMA stores of the staged output tiles (double-buffered staging, mbarrier handshake with the epilogue)
constexpr int P_EG = 3;                                // epilogue row groups (4 warps each)
constexpr int P_EPI = P_EG * 128, P_BLD = 128, P_NI = 3;
constexpr int P_W_BLD = 4 + 4 * P_EG, P_W_ST = P_W_BLD + 4;     // first builder warp, store warp
constexpr int P_THREADS = 32 * (P_W_ST + 1);
constexpr int P_CH = 5, P_MAXKC = 4;                   // builder: edges per thread per chunk, kernels per item

struct Agg2Params {
  const float* coef; const unsigned* eoff;
  const float* q; float* pooled; long long* argmax; float* hq;
  int B, K, KP, nb, nk, out_dim, D, nkc, tpk, slabs, nitems, nstage, nacc, ni, rpg, flags, with_lo;
  float drop_scale; unsigned drop_thresh16; unsigned long long seed, offset; const unsigned long long* step_ptr;
  int off_coef, coef_plane, coef_buf, off_stage, stage_bytes, off_out, out_plane, off_pool, off_bars, tmem_cols;
  // Large node counts (K = 100: one coefficient buffer of one kernel is 57 KB, one input stage 57 KB, one output tile 51 KB): the
  // plan with two coefficient buffers and separate output staging does not fit 227 KB.  cshift = 0: ONE coefficient buffer (the
  // builders wait for the previous item's MMAs); alias = 1: the epilogue writes the output tile into the input stage its MMAs have
  // just finished reading, the store warp hands the stage back to the producer once the TMA store has read it.
  int cshift, alias;
};

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// FR: node rows per epilogue warp handled by straight-line code (the common case rpg == FR; anything else takes the
// bounds-checked path)
template <int MODE, bool WITH_LO, bool DROP, int FR>
__global__ void __launch_bounds__(P_THREADS, 1)
agg_persistent_kernel(const __grid_constant__ Maps tm, const Agg2Params p) {
  extern __shared__ uint8_t gsm_raw[];
  uint8_t* sm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, KP = p.KP, NP = p.KP, S = p.nstage, NACC = p.nacc;
  constexpr int planes = WITH_LO ? 2 : 1;
  const int tiles_per_item = p.nkc * p.tpk;
  const int it0 = (int)((long long)blockIdx.x * p.nitems / gridDim.x), it1 = (int)((long long)(blockIdx.x + 1) * p.nitems / gridDim.x);

  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + p.off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = tfull + 8;
  uint64_t* cfull = tempty + 8;
  uint64_t* cempty = cfull + 2;
  uint64_t* sfull = cempty + 2;
  uint64_t* sempty = sfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 2);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 8; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], P_EPI / 32); }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&cfull[a], 1); mbar_init(&cempty[a], p.ni);
        mbar_init(&sfull[a], P_EPI / 32); mbar_init(&sempty[a], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.in_hi);
    if (p.with_lo) tma_prefetch_desc(&tm.in_lo);
    if (MODE != AGG_FWD_POOL) { tma_prefetch_desc(&tm.out_hi); if (p.with_lo) tma_prefetch_desc(&tm.out_lo); }
  }
  // rows [K, KP) of every staged Y tile are never written by TMA (the boxes have K rows): zero them once, so that the
  // K-padding of the contraction multiplies zero coefficients with zeros (stale shared memory could hold NaN patterns)
  if (KP > K) {
    const int pad16 = (KP - K) * 8, nbox = S * planes * 2;
    for (int v = tid; v < nbox * pad16; v += P_THREADS) {
      const int bx = v / pad16, w = v - bx * pad16;
      *reinterpret_cast<uint4*>(sm + p.off_stage + (size_t)bx * KP * 128 + K * 128 + w * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
 // the builders' first reads (edge coefficients of the first two images) go to L2 while the CTA sets itself up
  auto prefetch_image = [&](int b, int t, int nthreads) {
    const int E = K * p.nb;
    const char* c0 = reinterpret_cast<const char*>(p.coef + (long long)b * E * p.nk);
    const char* e0 = reinterpret_cast<const char*>(p.eoff + (long long)b * E);
    for (int o = t * 128; o < E * p.nk * 4; o += nthreads * 128) prefetch_l2(c0 + o);
    for (int o = t * 128; o < E * 4; o += nthreads * 128) prefetch_l2(e0 + o);
  };
  if (warp >= P_W_BLD && warp < P_W_ST && it0 < it1) {
    const int b0 = it0 / p.slabs;
    prefetch_image(b0, tid - P_W_BLD * 32, P_BLD);
    if ((b0 + 1) * p.slabs < it1) prefetch_image(b0 + 1, tid - P_W_BLD * 32, P_BLD);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0, s = 0; uint32_t sph = 0;                 // ring position / phase kept incrementally (no divisions in the loops)
      int b = it0 / p.slabs, sl = it0 - b * p.slabs;
      for (int it = it0; it < it1; ++it) {
        for (int t = 0; t < tiles_per_item; ++t, ++g) {
          mbar_wait(&empty[s], sph ^ 1);
          mbar_arrive_expect_tx(&full[s], (uint32_t)(planes * 2 * K * 128));   // boxes hold exactly the image's K rows
          uint8_t* dst = sm + p.off_stage + (size_t)s * p.stage_bytes;
          const int c0 = (sl * tiles_per_item + t) * MT;
          for (int pl = 0; pl < planes; ++pl) {
            const CUtensorMap* m = pl ? &tm.in_lo : &tm.in_hi;
            tma_load_2d(dst + (pl * 2 + 0) * KP * 128, m, &full[s], c0, b * K);
            tma_load_2d(dst + (pl * 2 + 1) * KP * 128, m, &full[s], c0 + 64, b * K);
          }
          if (++s == S) { s = 0; sph ^= 1; }
        }
        if (++sl == p.slabs) { sl = 0; ++b; }
      }
    }
  } else if (warp <= P_NI) {
    // ------------------------------------------------------------ MMA issuers (one thread per warp runs the whole loop)
    // p.ni <= P_NI issuers are active and p.ni <= #stages: with several consumers on one ring, the issuer of tile g may only wait on
    // full[g % S] once tile g - S has been loaded (else the parity test is one phase off); its own previous tile g - ni guarantees that
    if (lane == 0 && warp - 1 < p.ni) {
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
      const uint32_t lbo = (uint32_t)KP * 128;
      const int me = warp - 1;
      int g = 0, n = 0, s = 0, acc = 0, owner = 0; uint32_t sph = 0, aph = 0;
      auto next_tile = [&]() {
        if (++s == S) { s = 0; sph ^= 1; }
        if (++acc == NACC) { acc = 0; aph ^= 1; }
        if (++owner == p.ni) owner = 0;
      };
      for (int it = it0; it < it1; ++it, ++n) {
        const int cb = n & p.cshift;
        bool waited = false, mine = false;
        int kk = 0, tk = 0;                               // kernel of the item this tile belongs to
        for (int t = 0; t < tiles_per_item; ++t, ++g, next_tile()) {
          const int kcur = kk;
          if (++tk == p.tpk) { tk = 0; ++kk; }
          if (owner != me) continue;
          if (!waited) { mbar_wait(&cfull[cb], (n >> p.cshift) & 1); waited = true; }
          mbar_wait(&tempty[acc], aph ^ 1);
          mbar_wait(&full[s], sph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(sm + p.off_stage + (size_t)s * p.stage_bytes), a_lo = a_hi + 2 * KP * 128;
          const uint32_t b_hi = smem_u32(sm + p.off_coef + (size_t)cb * p.coef_buf + (size_t)(kcur * planes) * p.coef_plane), b_lo = b_hi + p.coef_plane;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NP);
          for (int ks = 0; ks < KP / 16; ++ks) {
            const uint32_t ao = ks * 2048, bo = (ks >> 2) * NP * 128 + (ks & 3) * 32;
            const uint64_t dah = desc_mnmajor(a_hi + ao, lbo), dbh = desc_kmajor(b_hi + bo);
            if (WITH_LO) {
              const uint64_t dal = desc_mnmajor(a_lo + ao, lbo), dbl = desc_kmajor(b_lo + bo);
              tc_mma<1>(d_tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
              tc_mma<1>(d_tmem, dah, dbl, idesc, 1u);
              tc_mma<1>(d_tmem, dah, dbh, idesc, 1u);
            } else {
              tc_mma<1>(d_tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
            }
          }
          if (!p.alias) tc_commit(&empty[s]);                       // (alias: the stage now receives the output tile; the store warp frees it)
          tc_commit(&tfull[acc]);
          mine = t + p.ni >= tiles_per_item;                        // my last tile of this item
          if (mine) tc_commit(&cempty[cb]);                          // coefficient buffer free once every issuer's MMAs of the item retire
        }
        if (!mine) mbar_arrive(&cempty[cb]);                        // no tile of this item was mine
      }
    }
    __syncwarp();
  } else if (warp < P_W_BLD) {
    // ------------------------------------------------------------ epilogue, thread = output column, P_EG warps per lane quarter
    const int q4 = warp & 3, grp = (warp - 4) >> 2;
    const int r0 = grp * p.rpg, r1 = min(K, r0 + p.rpg);   // my node rows
    const int cl = q4 * 32 + lane;
    const unsigned long long rng_off = p.offset + (p.step_ptr ? *p.step_ptr * 16ull : 0ull);
    const uint32_t key = hash32((uint32_t)p.seed ^ hash32((uint32_t)(p.seed >> 32) ^ hash32((uint32_t)rng_off * 0x9E3779B1u + 0x85EBCA77u)));
    constexpr bool RELU = MODE == AGG_FWD;                 // layer-1 forward always applies the ReLU (flags checked on the host)
    const uint32_t thresh_hi = p.drop_thresh16 << 16;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    int g = 0, acc = 0, sg = 0; uint32_t aph = 0;
    int b = it0 / p.slabs, sl = it0 - b * p.slabs;
    for (int it = it0; it < it1; ++it) {
      for (int t = 0; t < tiles_per_item; ++t, ++g) {
        const int buf = g & 1;
        const int col0 = (sl * tiles_per_item + t) * MT, col = col0 + cl;
        __nv_bfloat16* st_hi = reinterpret_cast<__nv_bfloat16*>(p.alias ? sm + p.off_stage + (size_t)sg * p.stage_bytes
                                                                         : sm + p.off_out + (size_t)buf * planes * p.out_plane);
        if (++sg == S) sg = 0;
        __nv_bfloat16* st_lo = st_hi + p.out_plane / 2;
        float best = -1.f; int barg = 0;
        // dropout: one counter hash per (pair of node rows, column)
        const uint32_t ctr0 = (((uint32_t)b * (uint32_t)((K + 1) >> 1)) * (uint32_t)p.out_dim + (uint32_t)col) ^ key;
        auto chunk = [&](const uint32_t (&r)[16], int i0) {       // rows [i0, min(i0 + 16, r1))
          if (MODE == AGG_FWD_POOL) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float v = fmaxf(__uint_as_float(r[e]), 0.f);
              if (i0 + e < r1 && v > best) { best = v; barg = i0 + e; }   // strict > : first index on ties
            }
          } else {
            __nv_bfloat16* h0 = st_hi + i0 * MT + cl;
            __nv_bfloat16* l0 = st_lo + i0 * MT + cl;
            const uint32_t cbase = ctr0 + (uint32_t)(i0 >> 1) * (uint32_t)p.out_dim;
            if (i0 + FR == r1) {                          // the whole group in one chunk: straight-line code, no bounds checks
#pragma unroll
              for (int e = 0; e < FR; e += 2) {
                const uint32_t rnd = DROP ? hash32(cbase + (uint32_t)(e >> 1) * (uint32_t)p.out_dim) : 0u;
                epi_pair<WITH_LO, DROP, RELU>(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), true, rnd, thresh_hi, h0 + e * MT, l0 + e * MT);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                if (i0 + e < r1) {
                  const uint32_t rnd = DROP ? hash32(cbase + (uint32_t)(e >> 1) * (uint32_t)p.out_dim) : 0u;
                  epi_pair<WITH_LO, DROP, RELU>(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), i0 + e + 1 < r1, rnd, thresh_hi, h0 + e * MT, l0 + e * MT);
                }
              }
            }
          }
        };
        float qv = 0.f;
        if (MODE == AGG_FWD_POOL && grp == 0) qv = p.q[(long long)b * p.out_dim + col];   // in flight during the wait
        if (lane == 0) {                                  // one poller per warp
          mbar_wait(&tfull[acc], aph);
          if (MODE != AGG_FWD_POOL) mbar_wait(&sempty[buf], ((g >> 1) & 1) ^ 1);   // the store of two tiles ago is done (alias mode: it
                                                                                   // keeps the sfull / sempty phases from lapping)
        }
        __syncwarp();
        tc_fence_after();
        {
          for (int i0 = r0; i0 < r1; i0 += 32) {
            uint32_t ra[16], rb[16];
            const bool two = i0 + 16 < r1;                // both TMEM loads in flight before the wait
            tc_ld_32x16(lane_base + (uint32_t)(acc * NP + i0), ra);
            if (two) tc_ld_32x16(lane_base + (uint32_t)(acc * NP + i0 + 16), rb);
            tc_wait_ld();
            chunk(ra, i0);
            if (two) chunk(rb, i0 + 16);
          }
        }
        tc_fence_before();
        if (MODE == AGG_FWD_POOL) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);       // accumulator drained: one arrival per epilogue warp
          float* pool_v = reinterpret_cast<float*>(sm + p.off_pool) + (g & 1) * 2 * P_EG * MT;   // scratch double-buffered by tile
          int* pool_a = reinterpret_cast<int*>(pool_v + P_EG * MT);                             // parity: one barrier per tile
          if (grp > 0) { pool_v[grp * MT + cl] = best; pool_a[grp * MT + cl] = barg; }
          named_bar_sync(1, P_EPI);
          if (grp == 0) {
#pragma unroll
            for (int o = 1; o < P_EG; ++o) {              // groups hold increasing row ranges: strict > keeps the first index on ties
              const float ov = pool_v[o * MT + cl]; const int oa = pool_a[o * MT + cl];
              if (ov > best) { best = ov; barg = oa; }
            }
            const long long o = (long long)b * p.out_dim + col;
            p.pooled[o] = best;
            p.argmax[o] = barg;
            p.hq[o] = fmaxf(qv, 0.f) * best;
          }
        } else {
          fence_proxy_async();                            // staging writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) { mbar_arrive(&tempty[acc]); mbar_arrive(&sfull[buf]); }
        }
        if (++acc == NACC) { acc = 0; aph ^= 1; }
      }
      if (++sl == p.slabs) { sl = 0; ++b; }
    }
  } else if (warp < P_W_ST) {
    // ------------------------------------------------------------ coefficient builders
    const int bt = tid - P_W_BLD * 32;
    const int E = K * p.nb, nkc = p.nkc, nk = p.nk;
    float pc[P_CH][P_MAXKC]; unsigned po[P_CH];
    auto load_chunk = [&](int it, int ch) {               // edge values of item `it`, chunk `ch`, into registers
      const int b = it / p.slabs, k_lo = (it - b * p.slabs) * nkc;
#pragma unroll
      for (int u = 0; u < P_CH; ++u) {
        const int e = (ch * P_CH + u) * P_BLD + bt;
        if (e < E) {
          const long long ge = (long long)b * E + e;
          po[u] = p.eoff[ge];
#pragma unroll
          for (int kk = 0; kk < P_MAXKC; ++kk)
            if (kk < nkc) pc[u][kk] = p.coef[ge * nk + k_lo + kk];
        }
      }
    };
    const int nchunks = (E + P_CH * P_BLD - 1) / (P_CH * P_BLD);
    if (it0 < it1) load_chunk(it0, 0);
    int n = 0;
    for (int it = it0; it < it1; ++it, ++n) {
      const int cb = n & p.cshift;
      if (lane == 0) mbar_wait(&cempty[cb], ((n >> p.cshift) & 1) ^ 1);   // every issuer's MMAs of the item that used this buffer are done
      __syncwarp();
      uint8_t* cbase = sm + p.off_coef + (size_t)cb * p.coef_buf;
      {
        uint4* cz = reinterpret_cast<uint4*>(cbase);
        const int n16 = nkc * planes * p.coef_plane / 16;
        for (int v = bt; v < n16; v += P_BLD) cz[v] = make_uint4(0u, 0u, 0u, 0u);
        named_bar_sync(2, P_BLD);
        for (int ch = 0; ch < nchunks; ++ch) {
          if (ch > 0) load_chunk(it, ch);
#pragma unroll
          for (int u = 0; u < P_CH; ++u) {
            const int e = (ch * P_CH + u) * P_BLD + bt;
            if (e < E) {
              const uint32_t off = MODE == AGG_BWD ? (po[u] >> 16) : (po[u] & 0xFFFFu);   // bwd: dY[j] += c * dO[i]  (transposed matrix)
#pragma unroll
              for (int kk = 0; kk < P_MAXKC; ++kk) {
                if (kk < nkc) {
                  // the dropout scale 1/(1-p) rides on the coefficients (ReLU commutes with a positive scale)
                  const float c = DROP ? pc[u][kk] * p.drop_scale : pc[u][kk];
                  const __nv_bfloat16 h = __float2bfloat16_rn(c);
                  uint8_t* base = cbase + (size_t)(kk * planes) * p.coef_plane;
                  *reinterpret_cast<__nv_bfloat16*>(base + off) = h;
                  if (WITH_LO) *reinterpret_cast<__nv_bfloat16*>(base + p.coef_plane + off) = __float2bfloat16_rn(c - __bfloat162float(h));
                }
              }
            }
          }
        }
      }
      fence_proxy_async();                                // generic-proxy smem writes -> visible to the tensor core
      named_bar_sync(2, P_BLD);
      if (bt == 0) mbar_arrive(&cfull[cb]);
      if (it + 1 < it1) load_chunk(it + 1, 0);            // next item's values in flight while waiting for its buffer
      {                                                   // two images ahead -> L2 (an LDG to cold HBM lines takes microseconds
        const int b = it / p.slabs;                       // while the TMA streams saturate the memory system)
        if (it == b * p.slabs && (b + 2) * p.slabs < it1) prefetch_image(b + 2, bt, P_BLD);
      }
    }
  } else if (MODE != AGG_FWD_POOL) {
    // ------------------------------------------------------------ last warp: TMA stores of the staged tiles
    if (!p.alias) {
      if (lane == 0) {
        int g = 0;
        int b = it0 / p.slabs, sl = it0 - b * p.slabs;
        for (int it = it0; it < it1; ++it) {
          for (int t = 0; t < tiles_per_item; ++t, ++g) {
            const int buf = g & 1;
            const int col0 = (sl * tiles_per_item + t) * MT;
            const uint8_t* st_hi = sm + p.off_out + (size_t)buf * planes * p.out_plane;
            mbar_wait(&sfull[buf], (g >> 1) & 1);
            tma_store_2d(&tm.out_hi, st_hi, col0, b * K);
            if (WITH_LO) tma_store_2d(&tm.out_lo, st_hi + p.out_plane, col0, b * K);
            bulk_commit();
            bulk_wait_read<0>();                            // ~18 KB of shared memory read: a few hundred cycles
            mbar_arrive(&sempty[buf]);                      // staging buffer free again
          }
          if (++sl == p.slabs) { sl = 0; ++b; }
        }
        bulk_wait_read<0>();
      }
    } else {
      // alias mode: the tile sits in the input stage of its own Y tile.  After the store has read it, the K-padding rows of the
      // stage's boxes (which the output overwrote, and which the next TMA load - K rows per box - will not) are zeroed again,
      // then the stage goes back to the producer.
      int g = 0, sg = 0;
      int b = it0 / p.slabs, sl = it0 - b * p.slabs;
      const int pad16 = (KP - K) * 8, nbox = planes * 2;
      for (int it = it0; it < it1; ++it) {
        for (int t = 0; t < tiles_per_item; ++t, ++g) {
          const int buf = g & 1;
          const int col0 = (sl * tiles_per_item + t) * MT;
          uint8_t* st = sm + p.off_stage + (size_t)sg * p.stage_bytes;
          if (lane == 0) {
            mbar_wait(&sfull[buf], (g >> 1) & 1);
            tma_store_2d(&tm.out_hi, st, col0, b * K);
            if (WITH_LO) tma_store_2d(&tm.out_lo, st + p.out_plane, col0, b * K);
            bulk_commit();
            bulk_wait_read<0>();
          }
          __syncwarp();
          for (int v = lane; v < nbox * pad16; v += 32) {
            const int bx = v / pad16, w = v - bx * pad16;
            *reinterpret_cast<uint4*>(st + (size_t)bx * KP * 128 + K * 128 + w * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&sempty[buf]); mbar_arrive(&empty[sg]); }   // stage free: the producer may load the next Y tile into it
          if (++sg == S) sg = 0;
        }
        if (++sl == p.slabs) { sl = 0; ++b; }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------ backward: edge products
// P[i,m,k] = < dO[i, chunk k], Y[idx[i,m], chunk k] >  and everything that depends on it.
//
// Dense per (image, kernel): Pd = dO_k Y_k^T  (K x K, contraction over the D columns of chunk k) on the tensor cores:
//   A = dO tile [i, c] and B = Y tile [j, c], both K-major straight from TMA (rows past K of either tile only produce
//   rows / columns of Pd that are never read).  The selected entries Pd[i, idx[i,m]] go to a (B, nk, K*nb) scratch
//   (L2-resident: 9 MB at the VQA2 shapes), and edge_finish_kernel turns them into dalpha[i,m] = sum_k w_k P_k and the
//   per-image partial sums of the four Gaussian-parameter gradients.  dO and Y are read exactly once.
// POOLED upstream (layer 2): dO[i, c] = (argmax[c] == i) ? dpooled[c] : 0 is synthesised straight into the A tile.
//
// The tensor core takes ~100 cycles for one M=128, K=16 instruction whatever N is (measured with the tracing build below: with one
// image per instruction the MMA issuer was busy 78 % of the kernel and everything else waited for it), so G = floor(128 / K)
// consecutive images share an instruction: their dO rows are stacked in the A tile and their Y rows in the B tile -- one TMA box
// each, the rows of consecutive images are contiguous -- and only the G diagonal K x K blocks of the product are read back.
//
// One persistent warp-specialised CTA per SM walks a contiguous range of (image group, kernel) units with one TMA stream across
// unit and group boundaries:
//   warp 0    TMA producer (item = 64 columns of a unit: the dO and Y boxes of K rows, hi and lo planes)
//   warp 1    MMA issuer (accumulators double-buffered in TMEM by unit parity)
//   warps 2-5 read Pd out of TMEM and pick the selected entries
//   warps 6-9 (pooled upstream only) build the A tiles
// (Round 2 trace, profiles/r02_edge_trace_before_*.txt: the previous one-CTA-per-image kernel ran in two lock-step waves,
// all CTAs streaming and then all CTAs finishing their edges with HBM idle, 97 us against 46 us of HBM time.)
struct PMaps { CUtensorMap d_hi, d_lo, y_hi, y_lo; };
struct PParams {
  const int* idx; const float* alpha; const float* boxes; long long ldbox; const float* gauss;
  const float* dpooled; const long long* argmax;           // pooled upstream (else NULL)
  float* dalpha; float* partial;                           // (B,K,nb) or NULL ; (B, 4*nk)
  float* pacc;                                             // (B, nk, K*nb) selected products
  int B, K, G, GK, NP, nb, nk, out_dim, D, nblk, nstage, with_lo;   // G images per unit, GK = G * K stacked rows, NP = round16(GK)
  int off_stage, a_plane, b_plane, stage_bytes, off_pd, off_idx, off_prev, off_ring, off_bars, tmem_cols;
};
constexpr int EP_THREADS = 320;
constexpr int EP_MAX_G = 4;
constexpr int EP_PF = 6;                // pooled A-tile builders: items of look-ahead for argmax / dpooled
#ifdef VQA_EDGE_TRACE
// Tracing build only (make EXTRA=-DVQA_EDGE_TRACE BUILD=build_trace OUT=../vqa_b200/libvqa_trace.so; tools/edge_trace.py): cycles each
// role of a CTA spent waiting on each of its barriers -- the role that never waits is the bottleneck.
__device__ long long g_edge_trace[8 * 256];
#define EDGE_T0() const long long t0_ = clock64()
#define EDGE_WAIT(SLOT, STMT) do { const long long w0_ = clock64(); STMT; tr_[SLOT] += clock64() - w0_; } while (0)
#define EDGE_DECL() long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define EDGE_FLUSH(SLOT, COND) do { if ((COND) && blockIdx.x < 256) g_edge_trace[blockIdx.x * 8 + (SLOT)] = tr_[SLOT]; } while (0)
#else
#define EDGE_T0() do { } while (0)
#define EDGE_WAIT(SLOT, STMT) STMT
#define EDGE_DECL() do { } while (0)
#define EDGE_FLUSH(SLOT, COND) do { } while (0)
#endif
constexpr int EF_THREADS = 576;          // edge_finish_kernel: at most this many threads (one per edge) per image

template <bool POOLED>
__global__ void __launch_bounds__(EP_THREADS, 1)
edge_p_kernel(const __grid_constant__ PMaps tm, const PParams p) {
  extern __shared__ uint8_t gsm_raw[];
  uint8_t* sm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, G = p.G, GK = p.GK, NP = p.NP, nb = p.nb, nk = p.nk, S = p.nstage;
  const int planes = p.with_lo ? 2 : 1;
  const int nblk = p.nblk;                                 // 64-column blocks per kernel chunk = pipeline items per unit
  const long long units = (long long)((p.B + G - 1) / G) * nk;
  const int u0 = (int)(units * blockIdx.x / gridDim.x), nu = (int)(units * (blockIdx.x + 1) / gridDim.x) - u0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + p.off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  float* Pd = reinterpret_cast<float*>(sm + p.off_pd);     // [GK][K+1]: the diagonal blocks
  uint8_t* idx8 = sm + p.off_idx;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&full[s], POOLED ? 129 : 1); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.y_hi);
    if (p.with_lo) tma_prefetch_desc(&tm.y_lo);
    if (!POOLED) { tma_prefetch_desc(&tm.d_hi); if (p.with_lo) tma_prefetch_desc(&tm.d_lo); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  EDGE_DECL();
  EDGE_T0();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(planes * GK * 128) * (POOLED ? 1u : 2u);   // (rows past the last image are zero-filled and counted)
      int it = 0;
      for (int n = 0; n < nu; ++n) {
        const int u = u0 + n, grp = u / nk, k = u - grp * nk, b = grp * G;
        for (int blk = 0; blk < nblk; ++blk, ++it) {
          const int s = it % S;
          EDGE_WAIT(1, mbar_wait(&empty[s], ((it / S) & 1) ^ 1));
          mbar_arrive_expect_tx(&full[s], tx);
          uint8_t* st = sm + p.off_stage + (size_t)s * p.stage_bytes;
          const int c0 = (k * nblk + blk) * 64;
          for (int pl = 0; pl < planes; ++pl) {
            if (!POOLED) tma_load_2d(st + pl * p.a_plane, pl ? &tm.d_lo : &tm.d_hi, &full[s], c0, b * K);
            tma_load_2d(st + planes * p.a_plane + pl * p.b_plane, pl ? &tm.y_lo : &tm.y_hi, &full[s], c0, b * K);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: Pd(unit) += A_blk B_blk^T
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int it = 0;
    for (int n = 0; n < nu; ++n) {
      const int acc = n & 1;
      for (int blk = 0; blk < nblk; ++blk, ++it) {
        const int s = it % S;
        if (blk == 0) EDGE_WAIT(3, mbar_wait(&tempty[acc], ((n >> 1) & 1) ^ 1));
        EDGE_WAIT(2, mbar_wait(&full[s], (it / S) & 1));
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(sm + p.off_stage + (size_t)s * p.stage_bytes), a_lo = a_hi + p.a_plane;
          const uint32_t b_hi = a_hi + planes * p.a_plane, b_lo = b_hi + p.b_plane;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NP);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dah = desc_kmajor(a_hi + ks * 32), dbh = desc_kmajor(b_hi + ks * 32);
            const uint32_t accum = (blk > 0 || ks > 0) ? 1u : 0u;
            if (p.with_lo) {
              const uint64_t dal = desc_kmajor(a_lo + ks * 32), dbl = desc_kmajor(b_lo + ks * 32);
              tc_mma<1>(d_tmem, dal, dbh, idesc, accum);
              tc_mma<1>(d_tmem, dah, dbl, idesc, 1u);
              tc_mma<1>(d_tmem, dah, dbh, idesc, 1u);
            } else {
              tc_mma<1>(d_tmem, dah, dbh, idesc, accum);
            }
          }
          tc_commit(&empty[s]);
          if (blk == nblk - 1) tc_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ warps 2-5: Pd out of TMEM, selected entries -> scratch
    // Thread = stacked row (TMEM lane).  Its row of the diagonal block and its neighbour ids live in row-private shared memory,
    // so the phases below need no barrier between them.
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;                         // image t = row / K of the group, node i = row - t K
    const int nedge = K * nb;
    const int t_row = row / K, i_row = row - t_row * K, c_row = t_row * K;   // this row's diagonal block: columns [c_row, c_row + K)
    const int w_lo = ((q4 * 32) / K) * K, w_hi = min((min(q4 * 32 + 31, GK - 1) / K) * K + K, GK);   // columns this warp needs
    float* Prow = Pd + row * (K + 1);
    uint8_t* irow = idx8 + row * nb;
    int cur_b = -1;
    for (int n = 0; n < nu; ++n) {
      const int u = u0 + n, grp = u / nk, k = u - grp * nk, b = grp * G, acc = n & 1;
      const bool live = row < GK && b + t_row < p.B;
      if (b != cur_b) {
        if (live) {
          const int* src = p.idx + ((long long)(b + t_row) * K + i_row) * nb;
          for (int m = 0; m < nb; ++m) irow[m] = (uint8_t)src[m];
        }
        cur_b = b;
      }
      EDGE_WAIT(4, mbar_wait(&tfull[acc], (n >> 1) & 1));
      tc_fence_after();
      if (q4 * 32 < GK) {                                   // warp-uniform: the .sync.aligned loads need the whole warp
        for (int j0 = w_lo & ~15; j0 < w_hi; j0 += 16) {
          uint32_t r[16];
          tc_ld_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * NP + j0), r);
          tc_wait_ld();
          if (row < GK) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int j = j0 + e - c_row;
              if (j >= 0 && j < K) Prow[j] = __uint_as_float(r[e]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (live) {
        float* out = p.pacc + ((long long)(b + t_row) * nk + k) * nedge + i_row * nb;
        if ((nb & 3) == 0) {
          for (int m = 0; m < nb; m += 4) {
            const uchar4 j4 = *reinterpret_cast<const uchar4*>(irow + m);
            *reinterpret_cast<float4*>(out + m) = make_float4(Prow[j4.x], Prow[j4.y], Prow[j4.z], Prow[j4.w]);
          }
        } else {
          for (int m = 0; m < nb; ++m) out[m] = Prow[irow[m]];
        }
      }
    }
  } else if (POOLED) {
    // ------------------------------------------------------------ warps 6-9: A tile = dpooled[c] at row argmax[c], zero elsewhere
    // Entry w = (image t = w / 64 of the group, column c = w % 64) always belongs to thread w % 128 and lands in column c of the
    // rows of image t: entries never share a byte, so each thread keeps its own entries of a stage up to date -- clear where the
    // entry was the last time this stage was used, write where it is now -- and the tile is zeroed only once.
    const int bt = tid - 192;                               // 0..127
    uint32_t* prev = reinterpret_cast<uint32_t*>(sm + p.off_prev);   // [S][256]: offset of each entry in its stage, ~0 = none
    for (int s = 0; s < S; ++s) {
      uint8_t* a_hi = sm + p.off_stage + (size_t)s * p.stage_bytes;
      for (int w = bt; w < planes * (p.a_plane >> 4); w += 128) *reinterpret_cast<uint4*>(a_hi + w * 16) = make_uint4(0u, 0u, 0u, 0u);
      prev[s * 256 + bt] = ~0u; prev[s * 256 + 128 + bt] = ~0u;
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");          // the zero fill is shared work: nobody writes an entry before all of it has landed
    // argmax / dpooled of an item come from HBM (~1 us under load): each thread copies its two entries EP_PF - 1 items ahead into
    // a private slot of a small ring with cp.async, so their latency never sits on the chain stage-free -> tile-ready.
    long long* aring = reinterpret_cast<long long*>(sm + p.off_ring);           // [EP_PF][256]
    float* vring = reinterpret_cast<float*>(sm + p.off_ring + EP_PF * 256 * 8); // [EP_PF][256]
    const int nit = nu * nblk;
    auto issue = [&](int itn) {
      if (itn < nit) {
        const int n = itn / nblk, blk = itn - n * nblk;
        const int u = u0 + n, grp = u / nk, k = u - grp * nk, b = grp * G;
        const int gn = min(G, p.B - b), slot = (itn % EP_PF) * 256;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int w = bt + h * 128, t = w >> 6;
          if (t < gn) {
            const long long o = (long long)(b + t) * p.out_dim + (long long)(k * nblk + blk) * 64 + (w & 63);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(aring + slot + w)), "l"(p.argmax + o) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(vring + slot + w)), "l"(p.dpooled + o) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // (an empty group keeps the count uniform)
    };
    for (int i = 0; i < EP_PF - 1; ++i) issue(i);
    for (int it = 0; it < nit; ++it) {
      issue(it + EP_PF - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(EP_PF - 1) : "memory");
      int arg[2]; float v[2];
      {
        const int n = it / nblk, grp = (u0 + n) / nk, b = grp * G, gn = min(G, p.B - b), slot = (it % EP_PF) * 256;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int w = bt + h * 128, t = w >> 6;
          arg[h] = -1; v[h] = 0.f;
          if (t < gn) { arg[h] = t * K + (int)aring[slot + w]; v[h] = vring[slot + w]; }
        }
      }
      const int s = it % S;
      EDGE_WAIT(5, mbar_wait(&empty[s], ((it / S) & 1) ^ 1));
      uint8_t* a_hi = sm + p.off_stage + (size_t)s * p.stage_bytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t* slot = prev + s * 256 + h * 128 + bt;
        const uint32_t po = *slot;
        if (po != ~0u) {
          *reinterpret_cast<uint16_t*>(a_hi + po) = 0;
          if (p.with_lo) *reinterpret_cast<uint16_t*>(a_hi + p.a_plane + po) = 0;
        }
        uint32_t no = ~0u;
        if (arg[h] >= 0) {
          no = coef_off(arg[h], bt & 63, 128);              // K-major 128B-swizzled tile with 128-row pitch layout
          const __nv_bfloat16 hi = __float2bfloat16_rn(v[h]);
          *reinterpret_cast<__nv_bfloat16*>(a_hi + no) = hi;
          if (p.with_lo) *reinterpret_cast<__nv_bfloat16*>(a_hi + p.a_plane + no) = __float2bfloat16_rn(v[h] - __bfloat162float(hi));
        }
        *slot = no;
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  }
  EDGE_FLUSH(1, tid == 0); EDGE_FLUSH(2, tid == 32); EDGE_FLUSH(3, tid == 32); EDGE_FLUSH(4, tid == 64); EDGE_FLUSH(5, tid == 192);
  tc_fence_before();
  __syncthreads();
#ifdef VQA_EDGE_TRACE
  if (tid == 0 && blockIdx.x < 256) { g_edge_trace[blockIdx.x * 8] = clock64() - t0_; g_edge_trace[blockIdx.x * 8 + 7] = nu; }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// Edge finish, one CTA per image, one thread per edge: dalpha[i,m] = sum_k w_k P_k and the per-image partial sums of the
// gradients of the four Gaussian parameters (mean_rho, precision_rho, mean_theta, precision_theta; layers.py:100-125
// differentiated).  With g_k the Gaussian value, S = sum_k g_k, w_k = g_k / S and gam_k = w_k (alpha P_k - alpha dalpha):
//   d/d mean_rho_k   = sum_e gam_k (rho - mr_k) / vr_k            d/d sigma_rho_k   = sum_e gam_k (rho - mr_k)^2 sr_k / vr_k^2
//   d/d mean_theta_k = sum_e gam_k x_k / vt_k                     d/d sigma_theta_k = sum_e gam_k del_k^2 st_k / vt_k^2
// where del = min(|df|, |2 pi - |df||), df = theta - mt_k, and x = -del * d del / d mt = df on the first branch of the min and
// -(2 pi - |df|) sign(df) on the second; the per-kernel factors are applied once per image, after the sums.  A NaN gam (S == 0,
// as the reference) poisons its four sums like autograd would.  The sums run in a fixed order (each thread's 32 terms go to its
// row of a shared-memory table; a warp sums 32 rows of a column, then the warps in turn): same bits every run.
__global__ void __launch_bounds__(EF_THREADS, 2)
edge_finish_kernel(const PParams p) {
  extern __shared__ float red[];                           // [blockDim][33]
  __shared__ float cen[2 * 128];
  __shared__ float gs[8 * MAX_NK];                         // mean_rho | cr | mean_theta | ct | 1/vr | sr/vr^2 | 1/vt | st/vt^2
  __shared__ float part[EF_THREADS / 32][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int b = blockIdx.x, K = p.K, nb = p.nb, nk = p.nk;
  for (int i = tid; i < K; i += blockDim.x) {
    const float* bx = p.boxes + ((long long)b * K + i) * p.ldbox;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    cen[2 * i] = x1 + 0.5f * (x2 - x1);
    cen[2 * i + 1] = y1 + 0.5f * (y2 - y1);
  }
  for (int k = tid; k < nk; k += blockDim.x) {
    const float sr = p.gauss[nk + k], st = p.gauss[3 * nk + k];
    const float vr = GM_EPS_F + sr * sr, vt = GM_EPS_F + st * st;
    gs[k] = p.gauss[k];
    gs[nk + k] = -0.5f * 1.4426950408889634f / vr;
    gs[2 * nk + k] = p.gauss[2 * nk + k];
    gs[3 * nk + k] = -0.5f * 1.4426950408889634f / vt;
    gs[4 * nk + k] = 1.f / vr; gs[5 * nk + k] = sr / (vr * vr); gs[6 * nk + k] = 1.f / vt; gs[7 * nk + k] = st / (vt * vt);
  }
  __syncthreads();
  const int nedge = K * nb;
  const float* Pb = p.pacc + (long long)b * nk * nedge;    // [nk][nedge]
  const int* idxb = p.idx + (long long)b * nedge;
  const float* alb = p.alpha ? p.alpha + (long long)b * nedge : nullptr;
  float* dab = p.dalpha ? p.dalpha + (long long)b * nedge : nullptr;
  for (int kc = 0; kc < nk; kc += 8) {                      // kernels in chunks of 8: 32 sums per thread
    const int kn_ = min(8, nk - kc);
    float* acc = red + tid * 33;                            // [which][u]: mean_rho | sigma_rho | mean_theta | sigma_theta
    bool first = true;
    for (int e = tid; e < nedge; e += blockDim.x) {
      const int i = e / nb, j = idxb[e];
      const float a = alb ? alb[e] : 1.f;
      float P[8];                                           // this chunk's kernels
#pragma unroll
      for (int u = 0; u < 8; ++u) P[u] = u < kn_ ? Pb[(kc + u) * nedge + e] : 0.f;
      const float dx = cen[2 * i] - cen[2 * j], dy = cen[2 * i + 1] - cen[2 * j + 1];
      const float rho = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      const float theta = atan2f(dx, dy);
      float g[8];
      float Ssum = 0.f, da = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {                         // (layers.py:109-117; one ex2 for both exponentials, NaN -> 0 as layers.py:120)
        const int k = kc + u;
        g[u] = 0.f;
        if (u < kn_) {
          g[u] = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
          Ssum += g[u];
          da = fmaf(g[u], P[u], da);
        }
      }
      if (nk > 8) {                                         // more than one chunk: S and dalpha need every kernel
        Ssum = 0.f; da = 0.f;
        for (int k = 0; k < nk; ++k) {
          const float gk = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
          Ssum += gk;
          da = fmaf(gk, Pb[k * nedge + e], da);
        }
      }
      const float invS = __fdiv_rn(1.f, Ssum);               // Ssum == 0 -> inf -> NaN below, as the reference
      da *= invS;                                          // dalpha = sum_k w_k P_k
      if (kc == 0 && dab) dab[e] = da;
      const float T = a * da;                              // sum_k dw_k w_k
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = min(kc + u, nk - 1);
        const float gam = g[u] * invS * fmaf(a, P[u], -T);   // g_k * dL/dg_k ; masked (NaN->0) kernels contribute 0
        const float d = rho - gs[k], df = theta - gs[2 * nk + k];          // (cheaper to recompute than to keep: registers)
        const float phi = fabsf(df), two = GM_TWO_PI_F - phi, psi = fabsf(two), del = fminf(phi, psi);
        const float x = phi < psi ? df : (df > 0.f ? -two : (df < 0.f ? two : 0.f));
        const float c0 = gam * d, c1 = c0 * d, c2 = gam * x, c3 = gam * del * del;
        if (first) { acc[u] = c0; acc[8 + u] = c1; acc[16 + u] = c2; acc[24 + u] = c3; }
        else { acc[u] += c0; acc[8 + u] += c1; acc[16 + u] += c2; acc[24 + u] += c3; }
      }
      first = false;
    }
    if (first) {
#pragma unroll
      for (int u = 0; u < 32; ++u) acc[u] = 0.f;
    }
    __syncthreads();
    {
      float s_ = 0.f;
#pragma unroll 8
      for (int t = 0; t < 32; ++t) s_ += red[(warp * 32 + t) * 33 + lane];
      part[warp][lane] = s_;
    }
    __syncthreads();
    if (tid < 32) {
      float s_ = 0.f;
      for (int w = 0; w < nwarp; ++w) s_ += part[w][tid];
      const int which = tid >> 3, u = tid & 7;              // 0: mean_rho, 1: sigma_rho, 2: mean_theta, 3: sigma_theta
      if (u < kn_) p.partial[(long long)b * 4 * nk + which * nk + kc + u] = s_ * gs[(4 + which) * nk + kc + u];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}
static int make_plane_map(CUtensorMap* tm, const void* ptr, long long ldp, long long rows, int cols, int box_c, int box_r, bool swizzle) {
  auto enc = get_encode();
  if (!enc) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ldp * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r}, estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled(graphconv plane) failed: CUresult %d", (int)r);
  return VQA_OK;
}

template <int MODE>
static int agg_launch(const void* in_hi, const void* in_lo, long long ldin, void* out_hi, void* out_lo, long long ldout, AggParams p,
                      cudaStream_t stream, const char* who) {
  const int K = p.K, B = p.B;
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128, "%s: need 0 < K <= 128 (K=%d)", who, K);
  VQA_CHECK_ARG(p.nb > 0 && p.nb <= K, "%s: neighbourhood size must be in [1,K] (nb=%d, K=%d)", who, p.nb, K);
  VQA_CHECK_ARG(p.nk > 0 && p.nk <= MAX_NK && p.out_dim > 0 && p.out_dim % p.nk == 0, "%s: out_dim (%d) must be divisible by n_kernels (%d <= %d)", who, p.out_dim, p.nk, MAX_NK);
  p.D = p.out_dim / p.nk;
  if (p.D % MT != 0) return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: out_dim / n_kernels (%d) must be a multiple of %d for the tensor-core aggregate", who, p.D, MT);
  VQA_CHECK_ARG(in_hi && aligned16(in_hi) && (!in_lo || aligned16(in_lo)) && (ldin & 7) == 0 && ldin >= p.out_dim, "%s: input planes need 16-byte aligned rows (ld %% 8 == 0)", who);
  p.with_lo = in_lo != nullptr;
  p.KP = (K + 15) & ~15;
  const int KP = p.KP, planes = p.with_lo ? 2 : 1;
  p.ntiles = p.out_dim / MT;
  const int tpk = p.D / MT;
  p.coef_plane = ((KP + 63) / 64) * KP * 128;
  p.stage_bytes = planes * 2 * KP * 128;
  p.out_plane = MODE == AGG_FWD_POOL ? 0 : ((K * MT * 2 + 127) & ~127);
  const int misc = ((K * p.nb + 15) & ~15) + 2 * ((K + 1) & ~1) * 4 + 4 * p.nk * 4 + 64;
  const int budget = 226 * 1024 / 2 - 2048;                 // two CTAs per SM
  // kernels per CTA: enough M-tiles to amortise the prologue, small enough to keep >= 3 input stages at 2 CTAs / SM
  int nkc = (4 + tpk - 1) / tpk;
  if (nkc > p.nk) nkc = p.nk;
  int S = 0;
  for (;; --nkc) {
    const int fixed = nkc * planes * p.coef_plane + 2 * planes * p.out_plane + misc + 256;
    S = (budget - fixed) / p.stage_bytes;
    if (S >= 2 || nkc == 1) break;
  }
  if (S < 1) {   // very large K: one CTA per SM
    const int fixed = nkc * planes * p.coef_plane + 2 * planes * p.out_plane + misc + 256;
    S = (226 * 1024 - 2048 - fixed) / p.stage_bytes;
    if (S < 1) return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: no shared-memory plan for K=%d", who, K);
  }
  if (S > 6) S = 6;
  p.nkc = nkc; p.nstage = S; p.tiles_per_cta = nkc * tpk;
  int off = 0;
  p.off_coef = off; off += nkc * planes * p.coef_plane; off = (off + 1023) & ~1023;
  p.off_stage = off; off += S * p.stage_bytes;
  p.off_out = off; off += 2 * planes * p.out_plane; off = (off + 127) & ~127;
  p.off_misc = off; off += misc; off = (off + 15) & ~15;
  p.off_bars = off; off += (2 * S + 5) * 8;
  const size_t smem = (size_t)off + 1024;
  int tc = 2 * KP; p.tmem_cols = 32; while (p.tmem_cols < tc) p.tmem_cols <<= 1;
  Maps tm;
  memset(&tm, 0, sizeof(tm));
  const long long rows = (long long)B * K;
  int rc = make_plane_map(&tm.in_hi, in_hi, ldin, rows, p.out_dim, 64, KP, true);
  if (!rc && p.with_lo) rc = make_plane_map(&tm.in_lo, in_lo, ldin, rows, p.out_dim, 64, KP, true);
  if (MODE != AGG_FWD_POOL) {
    VQA_CHECK_ARG(out_hi && aligned16(out_hi) && (!out_lo || aligned16(out_lo)) && (ldout & 7) == 0 && ldout >= p.out_dim, "%s: output planes need 16-byte aligned rows", who);
    VQA_CHECK_ARG(!p.with_lo || out_lo, "%s: 3-pass input needs both output planes", who);
    if (!rc) rc = make_plane_map(&tm.out_hi, out_hi, ldout, rows, p.out_dim, MT, K, false);
    if (!rc && p.with_lo) rc = make_plane_map(&tm.out_lo, out_lo, ldout, rows, p.out_dim, MT, K, false);
  }
  if (rc) return rc;
  // ---- preferred: the persistent, fully warp-specialised kernel (one CTA per SM)
  Maps tm2 = tm;                                             // its input boxes hold exactly K rows (no over-read into the next image)
  if (int rc2 = make_plane_map(&tm2.in_hi, in_hi, ldin, rows, p.out_dim, 64, K, true)) return rc2;
  if (p.with_lo) { if (int rc2 = make_plane_map(&tm2.in_lo, in_lo, ldin, rows, p.out_dim, 64, K, true)) return rc2; }
  if (p.coef && p.eoff && (MODE != AGG_FWD || (p.flags & VQA_GC_RELU))) {
    Agg2Params a{};
    a.coef = p.coef; a.eoff = p.eoff;
    a.q = p.q; a.pooled = p.pooled; a.argmax = p.argmax; a.hq = p.hq;
    a.B = B; a.K = K; a.KP = KP; a.nb = p.nb; a.nk = p.nk; a.out_dim = p.out_dim; a.D = p.D; a.flags = p.flags; a.with_lo = p.with_lo;
    a.drop_scale = p.drop_scale; a.drop_thresh16 = p.drop_thresh16; a.seed = p.seed; a.offset = p.offset; a.step_ptr = p.step_ptr;
    a.tpk = tpk;
    int nkc2 = (4 + tpk - 1) / tpk;                           // ~4 M-tiles per item
    if (nkc2 > p.nk) nkc2 = p.nk;
    if (nkc2 > P_MAXKC) nkc2 = P_MAXKC;
    while (p.nk % nkc2) --nkc2;                               // slabs tile the kernel axis exactly
    a.nkc = nkc2; a.slabs = p.nk / nkc2; a.nitems = B * a.slabs;
    a.coef_plane = p.coef_plane; a.coef_buf = nkc2 * planes * p.coef_plane; a.stage_bytes = p.stage_bytes; a.out_plane = p.out_plane;
    const int pool = MODE == AGG_FWD_POOL ? 2 * P_EG * MT * 8 : 0;
    a.rpg = (((K + P_EG - 1) / P_EG) + 1) & ~1;               // node rows per epilogue group (even: dropout hashes cover row pairs)
    a.cshift = 1; a.alias = 0;
    int fixed2 = 2 * a.coef_buf + 2 * planes * p.out_plane + pool + 1024;
    int S2 = (226 * 1024 - 2048 - fixed2) / p.stage_bytes;
    if (S2 < 2) {
      // large node counts (K = 100): one kernel per item, ONE coefficient buffer, output tiles staged inside the input stages
      a.nkc = 1; a.slabs = p.nk; a.nitems = B * a.slabs; a.coef_buf = planes * p.coef_plane;
      a.cshift = 0; a.alias = MODE != AGG_FWD_POOL && planes * p.out_plane <= p.stage_bytes ? 1 : 0;
      fixed2 = a.coef_buf + (a.alias || MODE == AGG_FWD_POOL ? 0 : 2 * planes * p.out_plane) + pool;
      S2 = (227 * 1024 - 1024 /* alignment slack */ - 1024 /* barriers, rounding */ - fixed2) / p.stage_bytes;
    }
    if (S2 >= 2) {
      if (S2 > 8) S2 = 8;
      a.nstage = S2;
      a.ni = S2 < P_NI ? S2 : P_NI;
      // every issuer must own a tile of every item: an issuer without one passes the item without waiting for anything and its
      // "not mine" arrival on cempty could land in an earlier item's phase (tile g belongs to issuer g % ni)
      if (a.ni > a.nkc * tpk) a.ni = a.nkc * tpk;
      int o2 = 0;
      a.off_coef = o2; o2 += (a.cshift ? 2 : 1) * a.coef_buf; o2 = (o2 + 1023) & ~1023;
      a.off_stage = o2; o2 += S2 * p.stage_bytes;
      a.off_out = o2; o2 += a.alias ? 0 : 2 * planes * p.out_plane; o2 = (o2 + 127) & ~127;
      a.off_pool = o2; o2 += pool;
      a.off_bars = o2; o2 += (2 * S2 + 26) * 8;
      a.nacc = (512 - 16) / KP < 4 ? (512 - 16) / KP : 4;     // accumulators in flight between the MMA issuers and the epilogue
      a.tmem_cols = 32; while (a.tmem_cols < a.nacc * KP + 16) a.tmem_cols <<= 1;   // +16: the last 16-column epilogue read may overhang
      const size_t smem2 = (size_t)o2 + 1024;
      int grid2 = g_vqa_sm_budget < a.nitems ? g_vqa_sm_budget : a.nitems;
      const bool dropk = MODE == AGG_FWD && a.drop_thresh16 != 0;
      auto go = [&](auto kern) -> int {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        kern<<<grid2, P_THREADS, smem2, stream>>>(tm2, a);
        VQA_LAUNCH_CHECK("graphconv agg_persistent_kernel");
        return VQA_OK;
      };
      const bool fr12 = a.rpg == 12;                          // K = 36: three groups of exactly 12 rows
      if (MODE == AGG_FWD && dropk) {
        if (a.with_lo) return fr12 ? go(agg_persistent_kernel<MODE, true, MODE == AGG_FWD, 12>) : go(agg_persistent_kernel<MODE, true, MODE == AGG_FWD, 16>);
        return fr12 ? go(agg_persistent_kernel<MODE, false, MODE == AGG_FWD, 12>) : go(agg_persistent_kernel<MODE, false, MODE == AGG_FWD, 16>);
      }
      if (a.with_lo) return fr12 ? go(agg_persistent_kernel<MODE, true, false, 12>) : go(agg_persistent_kernel<MODE, true, false, 16>);
      return fr12 ? go(agg_persistent_kernel<MODE, false, false, 12>) : go(agg_persistent_kernel<MODE, false, false, 16>);
    }
  }
  dim3 grid((p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta, B);
  VQA_CUDA(cudaFuncSetAttribute(agg_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  agg_kernel<MODE><<<grid, THREADS, smem, stream>>>(tm, p);
  VQA_LAUNCH_CHECK("graphconv agg_kernel");
  return VQA_OK;
}

template <bool POOLED>
static int edge_p_launch(const void* d_hi, const void* d_lo, long long ldd, const void* y_hi, const void* y_lo, long long ldy, PParams p,
                         cudaStream_t stream, const char* who) {
  const int K = p.K, B = p.B;
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128, "%s: need 0 < K <= 128 (K=%d)", who, K);
  VQA_CHECK_ARG(p.nb > 0 && p.nb <= K, "%s: neighbourhood size must be in [1,K] (nb=%d, K=%d)", who, p.nb, K);
  VQA_CHECK_ARG(p.nk > 0 && p.nk <= MAX_NK && p.out_dim > 0 && p.out_dim % p.nk == 0, "%s: out_dim (%d) must be divisible by n_kernels (%d <= %d)", who, p.out_dim, p.nk, MAX_NK);
  p.D = p.out_dim / p.nk;
  if (p.D % 64 != 0) return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: out_dim / n_kernels (%d) must be a multiple of 64 for the tensor-core edge products", who, p.D);
  VQA_CHECK_ARG(y_hi && aligned16(y_hi) && (!y_lo || aligned16(y_lo)) && (ldy & 7) == 0 && ldy >= p.out_dim, "%s: Y planes need 16-byte aligned rows", who);
  VQA_CHECK_ARG(p.pacc, "%s: pass a (B, n_kernels, K*nb) float scratch buffer for the selected edge products", who);
  p.with_lo = y_lo != nullptr;
  if (!POOLED) VQA_CHECK_ARG(d_hi && aligned16(d_hi) && (!p.with_lo || (d_lo && aligned16(d_lo))) && (ldd & 7) == 0 && ldd >= p.out_dim, "%s: dO planes need 16-byte aligned rows and the same planes as Y", who);
  const int planes = p.with_lo ? 2 : 1;
  p.G = 128 / K < EP_MAX_G ? 128 / K : EP_MAX_G;           // images stacked in one MMA (M = 128 rows)
  if (p.G > B) p.G = B;
  p.GK = p.G * K;
  p.NP = (p.GK + 15) & ~15;
  p.nblk = p.D / 64;
  // The MMA has M = 128 but only the first GK rows of the A tile matter (rows >= GK only produce rows of Pd nobody reads), so
  // an A plane is allotted round8(GK) rows; the tensor core's reads of the other rows run on into the following planes /
  // stages (always inside the ring + slack below), whose contents are irrelevant for those junk rows.
  p.a_plane = ((p.GK + 7) & ~7) * 128;
  p.b_plane = p.NP * 128;
  p.stage_bytes = planes * (p.a_plane + p.b_plane);
  int slack = (planes - 1) * p.a_plane + 128 * 128 - p.stage_bytes;      // junk-row reads of the last stage's last A plane
  if (slack < 0) slack = 0;
  const int pd = p.GK * (K + 1) * 4;
  const int idxb = (p.GK * p.nb + 15) & ~15;
  const int ring = POOLED ? EP_PF * 256 * 12 : 0;
  const int other = pd + idxb + ring + slack + 512;
  const int per_stage = p.stage_bytes + (POOLED ? 1024 : 0); // + where the pooled A tile's entries are
  int S = (227 * 1024 - 2048 - other) / per_stage;           // one CTA per SM
  if (S < 1) return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: no shared-memory plan for K=%d nb=%d nk=%d", who, K, p.nb, p.nk);
  if (S > 8) S = 8;
  p.nstage = S;
  int off = 0;
  p.off_stage = off; off += (S * p.stage_bytes + slack + 15) & ~15;
  p.off_pd = off; off += (pd + 15) & ~15;
  p.off_idx = off; off += idxb;
  p.off_prev = off; off += POOLED ? S * 1024 : 0;
  p.off_ring = off; off += ring;
  p.off_bars = off; off += (2 * S + 5) * 8;
  const size_t smem = (size_t)off + 1024;
  int tc = 2 * p.NP; p.tmem_cols = 32; while (p.tmem_cols < tc) p.tmem_cols <<= 1;
  PMaps tm;
  memset(&tm, 0, sizeof(tm));
  const long long rows = (long long)B * K;
  int rc = make_plane_map(&tm.y_hi, y_hi, ldy, rows, p.out_dim, 64, p.GK, true);
  if (!rc && p.with_lo) rc = make_plane_map(&tm.y_lo, y_lo, ldy, rows, p.out_dim, 64, p.GK, true);
  if (!POOLED) {
    if (!rc) rc = make_plane_map(&tm.d_hi, d_hi, ldd, rows, p.out_dim, 64, p.GK, true);
    if (!rc && p.with_lo) rc = make_plane_map(&tm.d_lo, d_lo, ldd, rows, p.out_dim, 64, p.GK, true);
  }
  if (rc) return rc;
  const long long units = (long long)((B + p.G - 1) / p.G) * p.nk;
  const int grid = (int)(units < g_vqa_sm_budget ? units : g_vqa_sm_budget);
  VQA_CUDA(cudaFuncSetAttribute(edge_p_kernel<POOLED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edge_p_kernel<POOLED><<<grid, POOLED ? EP_THREADS : 192, smem, stream>>>(tm, p);     // warps 6-9 only exist to build pooled A tiles
  VQA_LAUNCH_CHECK("graphconv edge_p_kernel");
  const int nedge32 = (K * p.nb + 31) & ~31;
  const int fthreads = nedge32 < EF_THREADS ? nedge32 : EF_THREADS;
  VQA_CUDA(cudaFuncSetAttribute(edge_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EF_THREADS * 33 * 4));
  edge_finish_kernel<<<B, fthreads, (size_t)fthreads * 33 * 4, stream>>>(p);
  VQA_LAUNCH_CHECK("graphconv edge_finish_kernel");
  return VQA_OK;
}

}  // namespace gm
}  // namespace vqa
using namespace vqa;

extern "C" int vqa_graphconv_mma_bwd_edges(const void* dO_hi, const void* dO_lo, long long lddo, const float* dpooled,
                                           const long long* argmax, const void* Y_hi, const void* Y_lo, long long ldy,
                                           const int* idx, const float* alpha, const float* boxes, long long ldbox,
                                           const float* gauss, float* dalpha, float* dgauss_partial, float* p_scratch, int B,
                                           int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  const char* who = "vqa_graphconv_mma_bwd_edges";
  VQA_CHECK_ARG(idx && boxes && gauss && dgauss_partial, "%s: null pointer", who);
  const bool pooled = dO_hi == nullptr;
  VQA_CHECK_ARG(!pooled || (dpooled && argmax), "%s: need either dO planes or (dpooled, argmax)", who);
  gm::PParams p{};
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss; p.dpooled = dpooled; p.argmax = argmax;
  p.dalpha = dalpha; p.partial = dgauss_partial; p.pacc = p_scratch;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim;
  return pooled ? gm::edge_p_launch<true>(nullptr, nullptr, 0, Y_hi, Y_lo, ldy, p, stream, who)
                : gm::edge_p_launch<false>(dO_hi, dO_lo, lddo, Y_hi, Y_lo, ldy, p, stream, who);
}

extern "C" int vqa_graphconv_mma_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* alpha,
                                     const float* boxes, long long ldbox, const float* gauss, void* out_hi, void* out_lo,
                                     long long ldo, int B, int K, int nb, int nk, int out_dim, int flags, float dropout_p,
                                     unsigned long long seed, unsigned long long offset, const unsigned long long* step_ptr,
                                     const float* coef, const unsigned* eoff, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss, "vqa_graphconv_mma_fwd: null pointer");
  VQA_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "vqa_graphconv_mma_fwd: dropout p must be in [0,1)");
  gm::AggParams p{};
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim; p.flags = flags;
  p.drop_scale = 1.f / (1.f - dropout_p); p.drop_thresh16 = (unsigned)(dropout_p * 65536.f + 0.5f);
  p.seed = seed; p.offset = offset; p.step_ptr = step_ptr; p.coef = coef; p.eoff = eoff;
  return gm::agg_launch<gm::AGG_FWD>(Y_hi, Y_lo, ldy, out_hi, out_lo, ldo, p, stream, "vqa_graphconv_mma_fwd");
}

extern "C" int vqa_graphconv_mma_pool_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* boxes,
                                          long long ldbox, const float* gauss, const float* q, float* pooled, long long* argmax,
                                          float* hq, int B, int K, int nb, int nk, int out_dim, const float* coef,
                                          const unsigned* eoff, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss && q && pooled && argmax && hq, "vqa_graphconv_mma_pool_fwd: null pointer");
  gm::AggParams p{};
  p.idx = idx; p.alpha = nullptr; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.q = q; p.pooled = pooled; p.argmax = argmax; p.hq = hq;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim; p.flags = VQA_GC_RELU; p.coef = coef; p.eoff = eoff;
  return gm::agg_launch<gm::AGG_FWD_POOL>(Y_hi, Y_lo, ldy, nullptr, nullptr, 0, p, stream, "vqa_graphconv_mma_pool_fwd");
}

extern "C" int vqa_graphconv_mma_bwd_data(const void* dO_hi, const void* dO_lo, long long lddo, const int* idx, const float* alpha,
                                          const float* boxes, long long ldbox, const float* gauss, void* dY_hi, void* dY_lo,
                                          long long lddy, int B, int K, int nb, int nk, int out_dim, const float* coef,
                                          const unsigned* eoff, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss, "vqa_graphconv_mma_bwd_data: null pointer");
  gm::AggParams p{};
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim; p.coef = coef; p.eoff = eoff;
  return gm::agg_launch<gm::AGG_BWD>(dO_hi, dO_lo, lddo, dY_hi, dY_lo, lddy, p, stream, "vqa_graphconv_mma_bwd_data");
}

extern "C" int vqa_graphconv_edge_coef(const int* idx, const float* alpha, const float* boxes, long long ldbox, const float* gauss,
                                       float* coef, unsigned* eoff, int B, int K, int nb, int nk, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss && coef && eoff, "vqa_graphconv_edge_coef: null pointer");
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128 && nb > 0 && nb <= K && nk > 0 && nk <= gm::MAX_NK, "vqa_graphconv_edge_coef: bad sizes (B=%d K=%d nb=%d nk=%d)", B, K, nb, nk);
  const long long n_edges = (long long)B * K * nb;
  const int KP = (K + 15) & ~15;
  gm::edge_coef_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, stream>>>(idx, alpha, boxes, ldbox, gauss, coef, eoff, n_edges, K, nb, nk, KP);
  VQA_LAUNCH_CHECK("graphconv edge_coef_kernel");
  return VQA_OK;
}

extern "C" int vqa_graphconv_pool_bwd_data(const float* dpooled, const long long* argmax, const int* idx, const float* coef, void* dY_hi,
                                           void* dY_lo, long long lddy, int B, int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  VQA_CHECK_ARG(dpooled && argmax && idx && coef && dY_hi, "vqa_graphconv_pool_bwd_data: null pointer");
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128 && nb > 0 && nb <= K && nk > 0 && out_dim > 0 && out_dim % nk == 0, "vqa_graphconv_pool_bwd_data: bad sizes (B=%d K=%d nb=%d nk=%d out=%d)", B, K, nb, nk, out_dim);
  VQA_CHECK_ARG(((out_dim / nk) & 7) == 0 && (lddy & 7) == 0 && lddy >= out_dim && aligned16(dY_hi) && (!dY_lo || aligned16(dY_lo)), "vqa_graphconv_pool_bwd_data: planes need 16-byte aligned rows, (out_dim / nk) %% 8 == 0");
  const size_t base = (size_t)out_dim * 4 + (((size_t)out_dim + 15) & ~(size_t)15) + (((size_t)K * K + 15) & ~(size_t)15);
  const size_t cbytes = (size_t)K * nb * nk * 4;
  const int stage = ((K * nb * nk) % 4 == 0 && aligned16(coef) && base + cbytes <= 56 * 1024) ? 1 : 0;   // 4 CTAs per SM
  const size_t smem = base + (stage ? cbytes : 0);
  if (smem > 200 * 1024) return vqa_fail(VQA_ERR_UNSUPPORTED, "vqa_graphconv_pool_bwd_data: out_dim too large for shared memory (%zu B)", smem);
  VQA_CUDA(cudaFuncSetAttribute(gm::pool_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gm::pool_bwd_data_kernel<<<B, gm::PB_THREADS, smem, stream>>>(dpooled, argmax, idx, coef, reinterpret_cast<__nv_bfloat16*>(dY_hi),
                                                                reinterpret_cast<__nv_bfloat16*>(dY_lo), lddy, K, nb, nk, out_dim, out_dim / nk, stage);
  VQA_LAUNCH_CHECK("graphconv pool_bwd_data_kernel");
  return VQA_OK;
}

#ifdef VQA_EDGE_TRACE
extern "C" int vqa_debug_edge_trace(long long* out, int n_ctas) {
  if (n_ctas > 256) n_ctas = 256;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, gm::g_edge_trace, (size_t)n_ctas * 8 * sizeof(long long)) == cudaSuccess ? VQA_OK : VQA_ERR_CUDA;
}
#endif
