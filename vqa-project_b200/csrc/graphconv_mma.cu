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

// ------------------------------------------------------------------------------------------ backward: edge products
// P[i,m,k] = < dO[i, chunk k], Y[idx[i,m], chunk k] >  and everything that depends on it.
//
// Dense per (image, kernel): Pd = dO_k Y_k^T  (K x K, contraction over the D columns of chunk k) on the tensor cores:
//   A = dO tile [i, c] and B = Y tile [j, c], both K-major straight from TMA (rows past K of either tile only produce
//   rows / columns of Pd that are never read).  The selected entries Pd[i, idx[i,m]] are gathered into shared memory
//   for all nk kernels, then the same CTA finishes the edges: dalpha[i,m] = sum_k w_k P_k and the per-image partial
//   sums of the four Gaussian-parameter gradients.  dO and Y are read exactly once.
// POOLED upstream (layer 2): dO[i, c] = (argmax[c] == i) ? dpooled[c] : 0 is synthesised straight into the A tile.
struct PMaps { CUtensorMap d_hi, d_lo, y_hi, y_lo; };
struct PParams {
  const int* idx; const float* alpha; const float* boxes; long long ldbox; const float* gauss;
  const float* dpooled; const long long* argmax;           // pooled upstream (else NULL)
  float* dalpha; float* partial;                           // (B,K,nb) or NULL ; (B, 4*nk)
  float* pacc_global;                                      // optional (B,K,nb,nk) scratch for shapes whose products exceed shared memory
  int B, K, NP, nb, nk, out_dim, D, nblk, nstage, with_lo;
  int off_stage, a_plane, b_plane, stage_bytes, off_pd, off_pacc, off_red, off_misc, off_bars, tmem_cols;
};

template <bool POOLED>
__global__ void __launch_bounds__(THREADS)
edge_p_kernel(const __grid_constant__ PMaps tm, const PParams p) {
  extern __shared__ uint8_t gsm_raw[];
  uint8_t* sm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x, K = p.K, NP = p.NP, nb = p.nb, nk = p.nk, S = p.nstage;
  const int planes = p.with_lo ? 2 : 1;
  const int nblk = p.nblk;                                 // 64-column blocks per kernel chunk
  const int total = nk * nblk;                             // pipeline items

  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + p.off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  float* Pd = reinterpret_cast<float*>(sm + p.off_pd);     // [K][NP+1]
  float* Pacc = p.off_pacc >= 0 ? reinterpret_cast<float*>(sm + p.off_pacc) : p.pacc_global + (long long)b * K * nb * nk;   // [K*nb][nk]
  uint8_t* idx8 = sm + p.off_misc;
  float* cen = reinterpret_cast<float*>(sm + p.off_misc + ((K * nb + 15) & ~15));
  float* gs = cen + 2 * ((K + 1) & ~1);                    // mean_rho | cr | mean_theta | ct | sigma_rho | sigma_theta | 4 derived

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
  for (int v = tid; v < K * nb; v += THREADS) idx8[v] = (uint8_t)p.idx[(long long)b * K * nb + v];
  for (int i = tid; i < K; i += THREADS) {
    const float* bx = p.boxes + ((long long)b * K + i) * p.ldbox;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    cen[2 * i] = x1 + 0.5f * (x2 - x1);
    cen[2 * i + 1] = y1 + 0.5f * (y2 - y1);
  }
  for (int k = tid; k < nk; k += THREADS) {
    const float sr = p.gauss[nk + k], st = p.gauss[3 * nk + k];
    gs[k] = p.gauss[k];
    gs[nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + sr * sr);
    gs[2 * nk + k] = p.gauss[2 * nk + k];
    gs[3 * nk + k] = -0.5f * 1.4426950408889634f / (GM_EPS_F + st * st);
    gs[4 * nk + k] = sr;
    gs[5 * nk + k] = st;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: item = (kernel k, 64-column block)
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(planes * K * 128) * (POOLED ? 1u : 2u);
      for (int it = 0; it < total; ++it) {
        const int s = it % S;
        mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], tx);
        uint8_t* st = sm + p.off_stage + (size_t)s * p.stage_bytes;
        const int c0 = it * 64;                            // items walk the columns in order: k = it / nblk
        for (int pl = 0; pl < planes; ++pl) {
          if (!POOLED) tma_load_2d(st + pl * p.a_plane, pl ? &tm.d_lo : &tm.d_hi, &full[s], c0, b * K);
          tma_load_2d(st + planes * p.a_plane + pl * p.b_plane, pl ? &tm.y_lo : &tm.y_hi, &full[s], c0, b * K);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: Pd(k) += A_blk B_blk^T
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < total; ++it) {
      const int s = it % S, k = it / nblk, blk = it - k * nblk, acc = k & 1;
      if (blk == 0) mbar_wait(&tempty[acc], ((k >> 1) & 1) ^ 1);
      mbar_wait(&full[s], (it / S) & 1);
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
  } else {
    // ------------------------------------------------------------ warps 2-5: (pooled: build A tiles) + gather Pd -> Pacc
    const int q4 = warp & 3, et = tid - 64;                 // et: 0..127
    const int row = q4 * 32 + lane;                         // TMEM lane = node i
    for (int k = 0; k < nk; ++k) {
      const int acc = k & 1;
      if (POOLED) {
        for (int blk = 0; blk < nblk; ++blk) {
          const int it = k * nblk + blk, s = it % S;
          mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
          uint8_t* a_hi = sm + p.off_stage + (size_t)s * p.stage_bytes;
          // zero rows [0, K) of the A tile(s), then drop dpooled[c] into row argmax[c]
          const int n16 = K * 8;                            // 16-byte chunks per plane (K rows x 128 B)
          for (int v = et; v < planes * n16; v += 128) {
            const int pl = v / n16, w = v - pl * n16;
            *reinterpret_cast<uint4*>(a_hi + pl * p.a_plane + w * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
          epi_bar_sync();
          if (et < 64) {
            const long long o = (long long)b * p.out_dim + (long long)it * 64 + et;
            const int n = (int)p.argmax[o];
            const float v = p.dpooled[o];
            const uint32_t off = coef_off(n, et, 128);      // K-major 128B-swizzled tile with 128-row pitch layout
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16*>(a_hi + off) = h;
            if (p.with_lo) *reinterpret_cast<__nv_bfloat16*>(a_hi + p.a_plane + off) = __float2bfloat16_rn(v - __bfloat162float(h));
          }
          fence_proxy_async();
          mbar_arrive(&full[s]);
        }
      }
      mbar_wait(&tfull[acc], (k >> 1) & 1);
      tc_fence_after();
      if (q4 * 32 < K) {                                    // warp-uniform: the .sync.aligned loads need the whole warp
        for (int j0 = 0; j0 < K; j0 += 16) {
          uint32_t r[16];
          tc_ld_32x16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * NP + j0), r);
          tc_wait_ld();
          if (row < K) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (j0 + e < K) Pd[row * (NP + 1) + j0 + e] = __uint_as_float(r[e]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      epi_bar_sync();
      for (int e = et; e < K * nb; e += 128) Pacc[e * nk + k] = Pd[(e / nb) * (NP + 1) + idx8[e]];
      epi_bar_sync();
    }
  }
  __syncthreads();

  // ------------------------------------------------------------ edge finish (all threads): dalpha and Gaussian-parameter partials
  // per-kernel constants without divisions in the edge loop: gs[6nk..10nk) = 1/vr | sr/vr^2 | 1/vt | st/vt^2
  for (int k = tid; k < nk; k += THREADS) {
    const float sr = gs[4 * nk + k], st = gs[5 * nk + k];
    const float vr = GM_EPS_F + sr * sr, vt = GM_EPS_F + st * st;
    gs[6 * nk + k] = 1.f / vr; gs[7 * nk + k] = sr / (vr * vr); gs[8 * nk + k] = 1.f / vt; gs[9 * nk + k] = st / (vt * vt);
  }
  __syncthreads();
  float* red = reinterpret_cast<float*>(sm + p.off_red);   // [THREADS][33]
  const int nedge = K * nb;
  for (int kc = 0; kc < nk; kc += 8) {                      // kernels in chunks of 8: 32 per-thread accumulators
    const int kn_ = min(8, nk - kc);
    float a_mr[8], a_sr[8], a_mt[8], a_st[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a_mr[u] = a_sr[u] = a_mt[u] = a_st[u] = 0.f;
    for (int e = tid; e < nedge; e += THREADS) {
      const int i = e / nb, j = idx8[e];
      const float dx = cen[2 * i] - cen[2 * j], dy = cen[2 * i + 1] - cen[2 * j + 1];
      const float rho = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      const float theta = atan2f(dx, dy);
      const float* Pe = Pacc + e * nk;
      float Ssum = 0.f, da = 0.f;
      for (int k = 0; k < nk; ++k) {
        const float g = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
        Ssum += g;
        da = fmaf(g, Pe[k], da);
      }
      const float invS = __fdiv_rn(1.f, Ssum);               // Ssum == 0 -> inf -> NaN below, as the reference
      da *= invS;                                          // dalpha = sum_k w_k P_k
      const float a = p.alpha ? p.alpha[(long long)b * nedge + e] : 1.f;
      if (kc == 0 && p.dalpha) p.dalpha[(long long)b * nedge + e] = da;
      const float T = a * da;                              // sum_k dw_k w_k
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (u < kn_) {
          const int k = kc + u;
          const float mr = gs[k], mt = gs[2 * nk + k];
          const float g = gauss_val(rho, theta, mr, gs[nk + k], mt, gs[3 * nk + k]);
          const float gam = g * (a * Pe[k] - T) * invS;    // g_k * dL/dg_k ; masked (NaN->0) kernels contribute 0
          const float dr = rho - mr;
          const float df = theta - mt;
          const float phi = fabsf(df), two = GM_TWO_PI_F - phi, psi = fabsf(two);
          const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
          const float ddel = phi < psi ? -sgn : (two > 0.f ? 1.f : (two < 0.f ? -1.f : 0.f)) * sgn;   // d(delta)/d(mean_theta)
          const float del = fminf(phi, psi);
          float c_mr = gam * dr * gs[6 * nk + k];
          float c_sr = gam * dr * dr * gs[7 * nk + k];
          float c_mt = -gam * del * gs[8 * nk + k] * ddel;
          float c_st = gam * del * del * gs[9 * nk + k];
          if (gam != gam) { c_mr = c_sr = c_mt = c_st = gam; }   // keep NaN visible (S == 0 rows), as autograd would
          a_mr[u] += c_mr; a_sr[u] += c_sr; a_mt[u] += c_mt; a_st[u] += c_st;
        }
      }
    }
    // transposed reduction through shared memory: thread t writes its 32 partials, then 32 threads sum a column each
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      red[tid * 33 + u] = a_mr[u]; red[tid * 33 + 8 + u] = a_sr[u]; red[tid * 33 + 16 + u] = a_mt[u]; red[tid * 33 + 24 + u] = a_st[u];
    }
    __syncthreads();
    if (tid < 32) {
      float s_ = 0.f;
      for (int t = 0; t < THREADS; ++t) s_ += red[t * 33 + tid];
      const int which = tid >> 3, u = tid & 7;              // 0: mean_rho, 1: precision_rho, 2: mean_theta, 3: precision_theta
      if (u < kn_) p.partial[(long long)b * 4 * nk + which * nk + kc + u] = s_;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
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
  p.with_lo = y_lo != nullptr;
  if (!POOLED) VQA_CHECK_ARG(d_hi && aligned16(d_hi) && (!p.with_lo || (d_lo && aligned16(d_lo))) && (ldd & 7) == 0 && ldd >= p.out_dim, "%s: dO planes need 16-byte aligned rows and the same planes as Y", who);
  const int planes = p.with_lo ? 2 : 1;
  p.NP = (K + 15) & ~15;
  p.nblk = p.D / 64;
  // The MMA has M = 128 but only the first K rows of the A tile matter (rows >= K only produce rows of Pd nobody reads), so
  // an A plane is allotted round8(K) rows; the tensor core's reads of the other rows run on into the following planes /
  // stages (always inside the ring + slack below), whose contents are irrelevant for those junk rows.
  p.a_plane = ((K + 7) & ~7) * 128;
  p.b_plane = p.NP * 128;
  p.stage_bytes = planes * (p.a_plane + p.b_plane);
  int slack = (planes - 1) * p.a_plane + 128 * 128 - p.stage_bytes;      // junk-row reads of the last stage's last A plane
  if (slack < 0) slack = 0;
  const int pd = K * (p.NP + 1) * 4;
  int pacc = K * p.nb * p.nk * 4;                            // selected products of all kernels: shared memory when it fits,
  if (pacc > 48 * 1024) pacc = 0;                            // else the caller's global scratch (L2-resident round trip)
  VQA_CHECK_ARG(pacc > 0 || p.pacc_global, "%s: K*nb*nk = %d products do not fit in shared memory: pass a (B,K,nb,nk) scratch buffer", who, K * p.nb * p.nk);
  const int red = THREADS * 33 * 4;                          // edge-finish reduction scratch: aliases the (by then idle) stage ring
  const int misc = ((K * p.nb + 15) & ~15) + 2 * ((K + 1) & ~1) * 4 + 10 * p.nk * 4 + 64;
  const int other = pd + pacc + misc + slack + 256;
  int S = (226 * 1024 / 2 - 2048 - other) / p.stage_bytes;  // two CTAs per SM: one finishes its edges while the other streams tiles
  if (S < 2) S = (226 * 1024 - 2048 - other) / p.stage_bytes;
  if (S < 1) return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: no shared-memory plan for K=%d nb=%d nk=%d", who, K, p.nb, p.nk);
  if (S > 6) S = 6;
  p.nstage = S;
  int ring = S * p.stage_bytes + slack;
  if (ring < red) ring = red;
  int off = 0;
  p.off_stage = off; off += (ring + 15) & ~15;
  p.off_pd = off; off += (pd + 15) & ~15;
  p.off_pacc = pacc ? off : -1; off += (pacc + 15) & ~15;
  p.off_red = p.off_stage;
  p.off_misc = off; off += misc; off = (off + 15) & ~15;
  p.off_bars = off; off += (2 * S + 5) * 8;
  const size_t smem = (size_t)off + 1024;
  int tc = 2 * p.NP; p.tmem_cols = 32; while (p.tmem_cols < tc) p.tmem_cols <<= 1;
  PMaps tm;
  memset(&tm, 0, sizeof(tm));
  const long long rows = (long long)B * K;
  int rc = make_plane_map(&tm.y_hi, y_hi, ldy, rows, p.out_dim, 64, K, true);
  if (!rc && p.with_lo) rc = make_plane_map(&tm.y_lo, y_lo, ldy, rows, p.out_dim, 64, K, true);
  if (!POOLED) {
    if (!rc) rc = make_plane_map(&tm.d_hi, d_hi, ldd, rows, p.out_dim, 64, K, true);
    if (!rc && p.with_lo) rc = make_plane_map(&tm.d_lo, d_lo, ldd, rows, p.out_dim, 64, K, true);
  }
  if (rc) return rc;
  VQA_CUDA(cudaFuncSetAttribute(edge_p_kernel<POOLED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edge_p_kernel<POOLED><<<B, THREADS, smem, stream>>>(tm, p);
  VQA_LAUNCH_CHECK("graphconv edge_p_kernel");
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
  p.dalpha = dalpha; p.partial = dgauss_partial; p.pacc_global = p_scratch;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim;
  return pooled ? gm::edge_p_launch<true>(nullptr, nullptr, 0, Y_hi, Y_lo, ldy, p, stream, who)
                : gm::edge_p_launch<false>(dO_hi, dO_lo, lddo, Y_hi, Y_lo, ldy, p, stream, who);
}

extern "C" int vqa_graphconv_mma_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* alpha,
                                     const float* boxes, long long ldbox, const float* gauss, void* out_hi, void* out_lo,
                                     long long ldo, int B, int K, int nb, int nk, int out_dim, int flags, float dropout_p,
                                     unsigned long long seed, unsigned long long offset, const unsigned long long* step_ptr,
                                     cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss, "vqa_graphconv_mma_fwd: null pointer");
  VQA_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "vqa_graphconv_mma_fwd: dropout p must be in [0,1)");
  gm::AggParams p{};
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim; p.flags = flags;
  p.drop_scale = 1.f / (1.f - dropout_p); p.drop_thresh16 = (unsigned)(dropout_p * 65536.f + 0.5f);
  p.seed = seed; p.offset = offset; p.step_ptr = step_ptr;
  return gm::agg_launch<gm::AGG_FWD>(Y_hi, Y_lo, ldy, out_hi, out_lo, ldo, p, stream, "vqa_graphconv_mma_fwd");
}

extern "C" int vqa_graphconv_mma_pool_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* boxes,
                                          long long ldbox, const float* gauss, const float* q, float* pooled, long long* argmax,
                                          float* hq, int B, int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss && q && pooled && argmax && hq, "vqa_graphconv_mma_pool_fwd: null pointer");
  gm::AggParams p{};
  p.idx = idx; p.alpha = nullptr; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.q = q; p.pooled = pooled; p.argmax = argmax; p.hq = hq;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim; p.flags = VQA_GC_RELU;
  return gm::agg_launch<gm::AGG_FWD_POOL>(Y_hi, Y_lo, ldy, nullptr, nullptr, 0, p, stream, "vqa_graphconv_mma_pool_fwd");
}

extern "C" int vqa_graphconv_mma_bwd_data(const void* dO_hi, const void* dO_lo, long long lddo, const int* idx, const float* alpha,
                                          const float* boxes, long long ldbox, const float* gauss, void* dY_hi, void* dY_lo,
                                          long long lddy, int B, int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  VQA_CHECK_ARG(idx && boxes && gauss, "vqa_graphconv_mma_bwd_data: null pointer");
  gm::AggParams p{};
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.B = B; p.K = K; p.nb = nb; p.nk = nk; p.out_dim = out_dim;
  return gm::agg_launch<gm::AGG_BWD>(dO_hi, dO_lo, lddo, dY_hi, dY_lo, lddy, p, stream, "vqa_graphconv_mma_bwd_data");
}
