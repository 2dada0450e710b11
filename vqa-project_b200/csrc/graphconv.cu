// Fused MoNet-style graph convolution on fixed-size neighbourhoods (sm_100a), project-first formulation.
//
//   out[b,i, chunk k] = act( sum_m  w[b,i,m,k] * alpha[b,i,m] * Y[b, idx[b,i,m], chunk k] ),   Y = X W_all^T
//
// One kernel computes the Gaussian patch weights over polar pseudo-coordinates from the box centres (only for the
// B*K*nb SELECTED edges, never the dense K x K table), normalises them over the kernel axis, gathers the neighbour
// rows out of a TMA-staged shared-memory tile and aggregates them; ReLU / dropout / max-pool+gate are epilogues.
// Replaces sparse_graph_model.py:161-195 (expand + torch.gather materialising (B,K,nb,F)), :239-240 (alpha multiply),
// :244-269 (dense pseudo-coordinates), layers.py:100-125 (Gaussian weights), :136-137 (bmm patch operator) and the
// ReLU/dropout/max/gate at sparse_graph_model.py:137-138,148-151.  HBM-bound: Y is read once, out written once.
//
// Data layout: Y / out are (B*K, out_dim) fp32 row-major; a CTA owns image b and a slab of column tiles
// [K rows x TW cols]; tiles arrive through a 2-D TMA tensor map into a multi-stage mbarrier ring.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstring>
#include "../../include/vqa_b200.h"

namespace vqa {

constexpr int GC_THREADS = 256;
constexpr int GC_WARPS = GC_THREADS / 32;
constexpr int GC_SMEM_MAX = 227 * 1024;
#define TWO_PI_F 6.28318530717958647692f
#define GAUSS_EPS_F 1e-14f

struct GcParams {
  const float* Y; long long ldy;
  const float* dO; long long lddo;          // bwd, dense upstream
  const float* dpooled; const long long* argmax_in;   // bwd, pooled upstream
  const int* idx; const float* alpha; const float* boxes; long long ldbox; const float* gauss;
  float* out; long long ldo;                // fwd: out ; bwd: dY
  float* P;                                 // bwd
  const float* q; float* pooled; long long* argmax; float* hq;   // fwd pooled epilogue
  int K, nb, nbp, nk, out_dim, D, TW, tstride, tiles_per_cta, ntiles, nkc, nstage, flags;   // tstride: floats per staged tile
  float drop_p, drop_scale;
  unsigned long long seed, offset;
  const unsigned long long* step_ptr;
};

struct GcSmem {   // byte offsets into dynamic smem (host-computed, identical on both sides)
  int tiles, coef, pacc, idx8, rev_cnt, rev_e, cen, gs, bars, scratch, total;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline GcSmem gc_smem_layout(int K, int nbp, int nk, int TW, int nkc, int nstage, int tiles_per_stage,
                                                 bool bwd, bool pool) {
  GcSmem s;
  int o = 0;
  s.tiles = o; o += nstage * tiles_per_stage * align_up(K * TW * 4, 128); o = align_up(o, 128);
  s.coef = o; o += nkc * K * nbp * 4;
  s.pacc = o; if (bwd) o += nkc * K * nbp * 4;
  s.idx8 = o; o += align_up(K * nbp, 16);
  s.rev_cnt = o; if (bwd) o += align_up(K * 2, 16);
  s.rev_e = o; if (bwd) o += align_up(K * K * 2, 16);
  s.cen = o; o += align_up(K * 2 * 4, 16);
  s.gs = o; o += align_up(4 * nk * 4, 16);
  s.bars = o; o += align_up(nstage * 8, 16);
  s.scratch = o; if (pool) o += 2 * GC_WARPS * TW * 8;
  s.total = o;
  return s;
}

__device__ __forceinline__ float gauss_val(float rho, float theta, float mr, float sr, float mt, float st) {
  float d = rho - mr;
  d = __fmul_rn(d, d);
  const float wr = expf(__fdiv_rn(__fmul_rn(-0.5f, d), __fadd_rn(GAUSS_EPS_F, __fmul_rn(sr, sr))));
  const float a1 = fabsf(theta - mt);
  const float a2 = fabsf(TWO_PI_F - a1);
  const float mn = fminf(a1, a2);
  const float wt = expf(__fdiv_rn(__fmul_rn(-0.5f, __fmul_rn(mn, mn)), __fadd_rn(GAUSS_EPS_F, __fmul_rn(st, st))));
  const float g = wr * wt;
  return (g != g) ? 0.f : g;      // NaN -> 0 BEFORE the kernel-axis normalisation (layers.py:120)
}
__device__ __forceinline__ void polar(float cxi, float cyi, float cxj, float cyj, float& rho, float& theta) {
  const float dx = cxi - cxj, dy = cyi - cyj;           // centre_i - centre_j (sparse_graph_model.py:258-259)
  rho = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  theta = atan2f(dx, dy);                               // x FIRST (sparse_graph_model.py:264-265)
}

// Shared prologue: neighbour ids (uint8) + box centres + Gaussian params to smem, then the per-edge coefficient
// table coef[kk][i][m] = w[i,m,k_lo+kk] * alpha[i,m] for the kernels this CTA's slab touches.
__device__ void gc_prologue(const GcParams& p, const GcSmem& L, uint8_t* sm, int b, int k_lo, int nkc) {
  const int tid = threadIdx.x, K = p.K, nb = p.nb, nbp = p.nbp, nk = p.nk;
  uint8_t* idx8 = sm + L.idx8;
  float* cen = reinterpret_cast<float*>(sm + L.cen);
  float* gs = reinterpret_cast<float*>(sm + L.gs);
  float* coef = reinterpret_cast<float*>(sm + L.coef);
  for (int v = tid; v < K * nbp; v += GC_THREADS) {
    const int i = v / nbp, m = v - i * nbp;
    idx8[v] = m < nb ? (uint8_t)p.idx[((long long)b * K + i) * nb + m] : (uint8_t)0;
  }
  for (int i = tid; i < K; i += GC_THREADS) {
    const float* bx = p.boxes + ((long long)b * K + i) * p.ldbox;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    cen[2 * i] = x1 + 0.5f * (x2 - x1);                 // sparse_graph_model.py:106-108
    cen[2 * i + 1] = y1 + 0.5f * (y2 - y1);
  }
  for (int v = tid; v < 4 * nk; v += GC_THREADS) gs[v] = p.gauss[v];
  __syncthreads();
  for (int v = tid; v < K * nbp; v += GC_THREADS) {
    const int i = v / nbp, m = v - i * nbp;
    if (m >= nb) {
      for (int kk = 0; kk < nkc; ++kk) coef[(kk * K + i) * nbp + m] = 0.f;
      continue;
    }
    const int j = idx8[v];
    float rho, theta;
    polar(cen[2 * i], cen[2 * i + 1], cen[2 * j], cen[2 * j + 1], rho, theta);
    float S = 0.f;
    for (int k = 0; k < nk; ++k) S += gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
    const float a = p.alpha ? p.alpha[((long long)b * K + i) * nb + m] : 1.f;
    for (int kk = 0; kk < nkc; ++kk) {
      const int k = k_lo + kk;
      const float g = gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
      coef[(kk * K + i) * nbp + m] = __fdiv_rn(g, S) * a;    // S == 0 -> NaN, exactly as the reference
    }
  }
}

// single-exp form of gauss_val with precomputed -0.5/(eps+sigma^2): exp(a)*exp(b) == exp(a+b) up to 1-2 ulp
__device__ __forceinline__ float gauss_fast(float rho, float theta, float mr, float cr, float mt, float ct) {
  const float d = rho - mr;
  const float a1 = fabsf(theta - mt);
  const float mn = fminf(a1, fabsf(TWO_PI_F - a1));
  const float g = expf(fmaf(d * d, cr, mn * mn * ct));
  return (g != g) ? 0.f : g;
}
// packed 2 x fp32 FMA (Blackwell FFMA2): d = a * b + c on both halves, each rounded to nearest like fmaf
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo2(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
// counter-based 32-bit hash (murmur3 finaliser over a seeded 64-bit counter): the fused dropout needs 16 random bits
// per output element and the kernel is issue-bound, so Philox4x32-10 (~90 instructions / call) is too expensive here.
__device__ __forceinline__ uint32_t hash32(unsigned long long ctr, unsigned long long seed, unsigned long long offset) {
  uint32_t x = (uint32_t)ctr * 0x9E3779B1u ^ (uint32_t)(ctr >> 32) * 0x85EBCA77u ^ (uint32_t)seed ^ (uint32_t)(seed >> 32) * 0xC2B2AE3Du ^
               (uint32_t)offset * 0x27D4EB2Fu;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  x ^= (uint32_t)ctr; x *= 0x9E3779B1u; x ^= x >> 15;
  return x;
}

// ------------------------------------------------------------------------------------------ dense-register aggregate
// out[r, c] = sum_{r'} M_k[r][r'] * in[r', c]  for every row r of one image, k = kernel owning column c.
//
// The neighbourhood aggregate written as a tiny dense matrix product per (image, kernel): M_k is K x K with the
// <= nb non-zeros per row w[i,m,k]*alpha[i,m] at column idx[i,m] (forward), or its transpose (backward dY = M^T dO).
// Why dense: a gather out of shared memory needs one 16-byte LDS per 4 FMAs and is bound by the 128 B/clk smem
// crossbar (measured 22 % of the HBM roofline); here every lane keeps its column slice of ALL K input rows in
// registers (K independent, coalesced global loads in flight per thread), coefficients arrive as warp-broadcast
// LDS.128 (4 per load), and the inner loop is >90 % FFMA.  K^2/nb more FLOPs, but FP32-FMA time (K=36: ~38 us for
// layer 1 at B=512) stays under the HBM time (46 us).  One warp owns a [K x 32*VEC] tile; max-pool over the nodes
// is a per-lane running max.  CTA = 4 warps = (image, slab of tiles): the Gaussian/alpha coefficient matrices of the
// slab's kernels are built once per CTA in shared memory.
enum { DM_FWD = 0, DM_FWD_POOL = 1, DM_BWD_DENSE = 2, DM_BWD_POOLED = 3 };
constexpr int DN_WARPS = 3;                 // 3 warps x 2 CTAs / SM: leaves ~250 registers per thread for the resident input rows
constexpr int DN_THREADS = DN_WARPS * 32;
constexpr int DN_CTA = DN_THREADS;

struct DenseParams {
  const float* in; long long ldin;           // Y (fwd) or dO (bwd dense)
  const float* dpooled; const long long* argmax_in;   // bwd pooled upstream
  const int* idx; const float* alpha; const float* boxes; long long ldbox; const float* gauss;
  float* out; long long ldo;
  const float* q; float* pooled; long long* argmax; float* hq;
  int K, nb, nk, out_dim, D, tiles_per_cta, ntiles, nkc, flags;
  int nstage, tile_stride, ring_off, bar_off;   // per-warp TMA landing slot: bytes per staged tile, byte offsets in dynamic smem
  float drop_p, drop_scale;
  unsigned int drop_thresh16;                // keep iff 16-bit uniform >= thresh
  unsigned long long seed, offset;
  const unsigned long long* step_ptr;        // optional device-side step counter added to offset (CUDA-graph replays)
};

template <int KT, int VEC, int MODE>
__global__ void __launch_bounds__(DN_CTA, 2)
graphconv_dense_kernel(const __grid_constant__ CUtensorMap tmIn, const DenseParams p) {
  static_assert(KT % 4 == 0 && (VEC == 1 || VEC == 2 || VEC == 4), "KT must be a multiple of 4, VEC 1, 2 or 4");
  constexpr int KP = KT;                      // row stride of the coefficient matrices (floats)
  constexpr int TWD = 32 * VEC;               // tile width in columns
  constexpr bool BWD = MODE == DM_BWD_DENSE || MODE == DM_BWD_POOLED;
  constexpr bool RING = MODE != DM_BWD_POOLED;   // input tiles arrive through a TMA ring (the pooled upstream is synthesised)
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* dsm = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);   // pointer arithmetic keeps the .shared address space (LDS/STS)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, K = p.K, nb = p.nb, nk = p.nk;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, p.ntiles - t0);
  const unsigned long long rng_offset = p.offset + (p.step_ptr ? *p.step_ptr * 16ull : 0ull);
  const int k_lo = (t0 * TWD) / p.D;
  const int nkc = ((t0 + nt) * TWD - 1) / p.D - k_lo + 1;
  // Input tiles: every warp owns one shared-memory landing slot.  Lane 0 issues the TMA load of the warp's NEXT tile as
  // soon as the current one has been copied into registers, so the load overlaps the whole FMA phase; with the
  // register copy this is a 2-deep pipeline per warp and ~2 x DN_WARPS x 18 KB in flight per SM, no cross-warp sync.
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + p.bar_off) + warp;
  float* slot = reinterpret_cast<float*>(dsm + p.ring_off + (size_t)warp * p.tile_stride);
  if (RING && lane == 0) {
    tma_prefetch_desc(&tmIn);
    mbar_init(full, 1);
    fence_barrier_init();
    if (warp < nt) {
      mbar_arrive_expect_tx(full, (uint32_t)K * TWD * 4);
      tma_load_2d(slot, &tmIn, full, (t0 + warp) * TWD, b * K);
    }
  }
  __syncwarp();
  constexpr int DUP = VEC >= 2 ? 2 : 1;       // packed paths: coefficients stored twice ({c,c}) as ready-made FFMA2 operands
  float* Ms = reinterpret_cast<float*>(dsm);                       // [nkc][K][KP][DUP]
  float* cen = Ms + p.nkc * K * KP * DUP;                          // [K][2]
  float* gs = cen + 2 * ((K + 1) & ~1);                            // [4][nk]: mean_rho, -0.5/(eps+prec_rho^2), mean_theta, -0.5/(eps+prec_theta^2)
  uint8_t* idx8 = reinterpret_cast<uint8_t*>(gs + 4 * nk);         // [K][nb]

  // ---- prologue: zero M, stage neighbour ids / box centres / Gaussian parameters
  for (int v = tid; v < nkc * K * KP * DUP / 4; v += DN_THREADS) reinterpret_cast<float4*>(Ms)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = tid; v < K * nb; v += DN_THREADS) idx8[v] = (uint8_t)p.idx[(long long)b * K * nb + v];
  for (int i = tid; i < K; i += DN_THREADS) {
    const float* bx = p.boxes + ((long long)b * K + i) * p.ldbox;
    const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    cen[2 * i] = x1 + 0.5f * (x2 - x1);                 // sparse_graph_model.py:106-108
    cen[2 * i + 1] = y1 + 0.5f * (y2 - y1);
  }
  for (int k = tid; k < nk; k += DN_THREADS) {
    const float sr = p.gauss[nk + k], st = p.gauss[3 * nk + k];
    gs[k] = p.gauss[k];
    gs[nk + k] = -0.5f / (GAUSS_EPS_F + sr * sr);
    gs[2 * nk + k] = p.gauss[2 * nk + k];
    gs[3 * nk + k] = -0.5f / (GAUSS_EPS_F + st * st);
  }
  __syncthreads();
  // ---- per-edge Gaussian weights -> dense coefficient matrices of the slab's kernels
  for (int e = tid; e < K * nb; e += DN_THREADS) {
    const int i = e / nb, j = idx8[e];
    float rho, theta;
    polar(cen[2 * i], cen[2 * i + 1], cen[2 * j], cen[2 * j + 1], rho, theta);
    float S = 0.f;
    for (int k = 0; k < nk; ++k) S += gauss_fast(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
    const float a_over_S = __fdiv_rn(p.alpha ? p.alpha[(long long)b * K * nb + e] : 1.f, S);   // S == 0 -> inf/NaN as in the reference
    for (int kk = 0; kk < nkc; ++kk) {
      const int k = k_lo + kk;
      const float c = gauss_fast(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]) * a_over_S;
      const int at = BWD ? (kk * K + j) * KP + i      // dY[j] += c * dO[i]
                         : (kk * K + i) * KP + j;     // out[i] += c * Y[j]
      if constexpr (DUP == 2) reinterpret_cast<float2*>(Ms)[at] = make_float2(c, c);
      else Ms[at] = c;
    }
  }
  __syncthreads();

  // ---- main: one [K x TWD] tile per warp iteration, input rows held in registers
  for (int t = warp; t < nt; t += DN_WARPS) {
    const int colg = (t0 + t) * TWD + lane * VEC;
    const float* Mk = Ms + (((t0 + t) * TWD) / p.D - k_lo) * K * KP * DUP;
    float y[KT][VEC];
    if constexpr (MODE == DM_BWD_POOLED) {
      float dp[VEC];
      int ar[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        dp[v] = p.dpooled[(long long)b * p.out_dim + colg + v];
        ar[v] = (int)p.argmax_in[(long long)b * p.out_dim + colg + v];
      }
#pragma unroll
      for (int j = 0; j < KT; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) y[j][v] = (ar[v] == j) ? dp[v] : 0.f;
    } else {
      mbar_wait(full, ((t - warp) / DN_WARPS) & 1);
      const float* src = slot + lane * VEC;
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        if (j < K) {
          if constexpr (VEC == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(src + j * TWD);
            y[j][0] = t4.x; y[j][1] = t4.y; y[j][2] = t4.z; y[j][3] = t4.w;
          } else if constexpr (VEC == 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(src + j * TWD);
            y[j][0] = t2.x; y[j][1] = t2.y;
          } else {
            y[j][0] = src[j * TWD];
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) y[j][v] = 0.f;
        }
      }
      __syncwarp();
      if (lane == 0 && t + DN_WARPS < nt) {        // slot is free again: fetch this warp's next tile under the FMA phase
        fence_proxy_async();
        mbar_arrive_expect_tx(full, (uint32_t)K * TWD * 4);
        tma_load_2d(slot, &tmIn, full, (t0 + t + DN_WARPS) * TWD, b * K);
      }
    }
    float best[VEC];
    int barg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { best[v] = -1.f; barg[v] = 0; }
#pragma unroll 1
    for (int i = 0; i < K; ++i) {
      float acc[VEC];
      if constexpr (VEC == 4) {
        // packed path, 128 columns per warp: four independent FFMA2 chains (2 column pairs x even/odd j)
        const ulonglong2* mrow = reinterpret_cast<const ulonglong2*>(Mk + i * KP * 2);
        unsigned long long a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull;
#pragma unroll
        for (int j2 = 0; j2 < KT / 2; ++j2) {
          const ulonglong2 c = mrow[j2];             // {c_j, c_j}, {c_j+1, c_j+1}: warp-broadcast LDS.128
          a0 = ffma2(c.x, pack2(y[2 * j2][0], y[2 * j2][1]), a0);
          b0 = ffma2(c.x, pack2(y[2 * j2][2], y[2 * j2][3]), b0);
          a1 = ffma2(c.y, pack2(y[2 * j2 + 1][0], y[2 * j2 + 1][1]), a1);
          b1 = ffma2(c.y, pack2(y[2 * j2 + 1][2], y[2 * j2 + 1][3]), b1);
        }
        acc[0] = lo2(a0) + lo2(a1);
        acc[1] = hi2(a0) + hi2(a1);
        acc[2] = lo2(b0) + lo2(b1);
        acc[3] = hi2(b0) + hi2(b1);
      } else if constexpr (VEC == 2) {
        // packed path: {acc0,acc1} += {c,c} * {y0,y1}; two independent chains for ILP
        const ulonglong2* mrow = reinterpret_cast<const ulonglong2*>(Mk + i * KP * 2);
        unsigned long long a0 = 0ull, a1 = 0ull;
#pragma unroll
        for (int j2 = 0; j2 < KT / 2; ++j2) {
          const ulonglong2 c = mrow[j2];             // {c_j, c_j}, {c_j+1, c_j+1}: warp-broadcast LDS.128
          a0 = ffma2(c.x, pack2(y[2 * j2][0], y[2 * j2][1]), a0);
          a1 = ffma2(c.y, pack2(y[2 * j2 + 1][0], y[2 * j2 + 1][1]), a1);
        }
        acc[0] = lo2(a0) + lo2(a1);
        acc[1] = hi2(a0) + hi2(a1);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        const float4* mrow = reinterpret_cast<const float4*>(Mk + i * KP);
#pragma unroll
        for (int j4 = 0; j4 < KT / 4; ++j4) {
          const float4 c = mrow[j4];                 // same address for the whole warp: broadcast
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            acc[v] = fmaf(c.x, y[4 * j4 + 0][v], acc[v]);
            acc[v] = fmaf(c.y, y[4 * j4 + 1][v], acc[v]);
            acc[v] = fmaf(c.z, y[4 * j4 + 2][v], acc[v]);
            acc[v] = fmaf(c.w, y[4 * j4 + 3][v], acc[v]);
          }
        }
      }
      if constexpr (MODE == DM_FWD_POOL) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          acc[v] = fmaxf(acc[v], 0.f);
          if (acc[v] > best[v]) { best[v] = acc[v]; barg[v] = i; }   // strict > : first index on ties
        }
      } else {
        if constexpr (MODE == DM_FWD) {
          if (p.flags & VQA_GC_RELU) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fmaxf(acc[v], 0.f);
          }
          if (p.drop_p > 0.f) {
            // 16 random bits per element from one 32-bit counter hash per (row, VEC-column group)
#pragma unroll
            for (int v0 = 0; v0 < VEC; v0 += 2) {
              const uint32_t r = hash32((unsigned long long)(((long long)b * K + i) * (p.out_dim / 2) + (colg + v0) / 2), p.seed, rng_offset);
              acc[v0] = (r & 0xFFFFu) >= p.drop_thresh16 ? acc[v0] * p.drop_scale : 0.f;
              if (v0 + 1 < VEC) acc[v0 + 1 < VEC ? v0 + 1 : v0] = (r >> 16) >= p.drop_thresh16 ? acc[v0 + 1 < VEC ? v0 + 1 : v0] * p.drop_scale : 0.f;
            }
          }
        }
        float* dst = p.out + ((long long)b * K + i) * p.ldo + colg;
        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else if constexpr (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(acc[0], acc[1]);
        else dst[0] = acc[0];
      }
    }
    if constexpr (MODE == DM_FWD_POOL) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const long long o = (long long)b * p.out_dim + colg + v;
        p.pooled[o] = best[v];
        p.argmax[o] = barg[v];
        p.hq[o] = fmaxf(p.q[o], 0.f) * best[v];
      }
    }
  }
}

// host: pick the compile-time K bucket and launch.  Returns 1 when the shape is not eligible (caller falls back to the
// generic gather kernels): the tile width 32*VEC must divide D = out_dim / nk so that a tile belongs to one kernel.
static int make_tile_map(CUtensorMap* tm, const float* ptr, long long ld, long long rows, int cols, int TW, int K);

template <int MODE>
static int dense_launch(const DenseParams& base, int B, cudaStream_t stream) {
  DenseParams p = base;
  const int K = p.K, D = p.out_dim / p.nk;
  const int KT = K <= 36 ? 36 : (K <= 52 ? 52 : (K <= 64 ? 64 : (K <= 100 ? 100 : 128)));
  const int VEC = (KT == 36 && D % 128 == 0 && (p.ldo % 4) == 0) ? 4 : (KT <= 64 ? 2 : 1);
  const int TWD = 32 * VEC, dup = VEC >= 2 ? 2 : 1;
  if (K > 128 || D % TWD != 0 || (MODE != DM_BWD_POOLED && (p.ldin % 4) != 0) || (p.ldo % VEC) != 0) return 1;
  p.D = D;
  p.ntiles = p.out_dim / TWD;
  const int tpk = D / TWD;                                    // tiles per kernel
  const int mat_bytes = K * KT * 4 * dup;
  p.tile_stride = (K * TWD * 4 + 127) & ~127;
  const int slots = MODE == DM_BWD_POOLED ? 0 : DN_WARPS * p.tile_stride;
  const int misc = (2 * ((K + 1) & ~1) + 4 * p.nk) * 4 + K * p.nb + 512;
  int nkc = ((226 * 1024) / 2 - 1024 - slots - misc) / mat_bytes;   // coefficient matrices per CTA at 2 CTAs / SM
  if (nkc > (p.nk + 1) / 2) nkc = (p.nk + 1) / 2;             // at least 2 slabs per image: finer load balance
  if (nkc < 1) nkc = 1;
  if ((size_t)nkc * mat_bytes + slots + misc > 226 * 1024) return 1;
  p.nkc = nkc;
  p.tiles_per_cta = nkc * tpk;
  const int nslab = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  // dynamic smem: [coefficients | centres | gauss | idx8 | pad128 | per-warp tile slots | barriers]
  int off = (nkc * K * KT * dup + 2 * ((K + 1) & ~1) + 4 * p.nk) * 4 + K * p.nb;
  off = (off + 127) & ~127;
  p.ring_off = off;
  p.nstage = 1;
  p.bar_off = off + slots;
  const size_t smem = (size_t)p.bar_off + DN_WARPS * 8 + 128;
  CUtensorMap tm;
  if (MODE != DM_BWD_POOLED) {
    if (int rc = make_tile_map(&tm, p.in, p.ldin, (long long)B * K, p.out_dim, TWD, K)) return rc;
  } else {
    memset(&tm, 0, sizeof(tm));
  }
  dim3 grid(nslab, B);
#define VQA_DENSE_CASE(KT_, VEC_)                                                                                      \
  {                                                                                                                    \
    VQA_CUDA(cudaFuncSetAttribute(graphconv_dense_kernel<KT_, VEC_, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    graphconv_dense_kernel<KT_, VEC_, MODE><<<grid, DN_CTA, smem, stream>>>(tm, p);                                    \
  }
  if (KT == 36 && VEC == 4) VQA_DENSE_CASE(36, 4)
  else if (KT == 36) VQA_DENSE_CASE(36, 2)
  else if (KT == 52) VQA_DENSE_CASE(52, 2)
  else if (KT == 64) VQA_DENSE_CASE(64, 2)
  else if (KT == 100) VQA_DENSE_CASE(100, 1)
  else VQA_DENSE_CASE(128, 1)
#undef VQA_DENSE_CASE
  VQA_LAUNCH_CHECK("graphconv_dense_kernel");
  return VQA_OK;
}

// ------------------------------------------------------------------------------------------ forward
template <bool POOL>
__global__ void __launch_bounds__(GC_THREADS)
graphconv_fwd_kernel(const __grid_constant__ CUtensorMap tmY, const GcParams p, const GcSmem L) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);   // pointer arithmetic keeps the .shared address space (LDS/STS)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, K = p.K, nbp = p.nbp, TW = p.TW;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, p.ntiles - t0);
  const int k_lo = (t0 * TW) / p.D;
  const int nkc = min(p.nkc, p.nk - k_lo);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  float* tiles = reinterpret_cast<float*>(sm + L.tiles);
  const uint32_t tile_bytes = (uint32_t)K * TW * 4;
  const int NS = p.nstage;

  if (tid == 0) {
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {   // fill the ring before the (long) prologue so the loads overlap it
    for (int t = 0; t < NS - 1 && t < nt; ++t) {
      mbar_arrive_expect_tx(&bars[t], tile_bytes);
      tma_load_2d(tiles + (size_t)t * p.tstride, &tmY, &bars[t], (t0 + t) * TW, b * K);
    }
  }
  gc_prologue(p, L, sm, b, k_lo, nkc);
  __syncthreads();

  const float* coef = reinterpret_cast<const float*>(sm + L.coef);
  const uint8_t* idx8 = sm + L.idx8;
  const int col = lane * 4;
  const bool active = col < TW;
  for (int t = 0; t < nt; ++t) {
    if (tid == 0) {
      const int tn = t + NS - 1;
      if (tn < nt && NS > 1) {
        const int s = tn % NS;
        mbar_arrive_expect_tx(&bars[s], tile_bytes);
        tma_load_2d(tiles + (size_t)s * p.tstride, &tmY, &bars[s], (t0 + tn) * TW, b * K);
      } else if (NS == 1) {
        mbar_arrive_expect_tx(&bars[0], tile_bytes);
        tma_load_2d(tiles, &tmY, &bars[0], (t0 + t) * TW, b * K);
      }
    }
    mbar_wait(&bars[t % NS], (t / NS) & 1);
    const float* tile = tiles + (size_t)(t % NS) * p.tstride;
    const int colg = (t0 + t) * TW + col;                 // global output column of this lane's float4
    const int kk = ((t0 + t) * TW) / p.D - k_lo;          // one kernel per tile (TW divides D)
    float4 best = make_float4(-1.f, -1.f, -1.f, -1.f);
    int4 barg = make_int4(0, 0, 0, 0);
    for (int i = warp; i < K; i += GC_WARPS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
        const float* cf = coef + (kk * K + i) * nbp;
        const uint8_t* ix = idx8 + i * nbp;
        for (int m = 0; m < nbp; m += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cf + m);
          const uchar4 j4 = *reinterpret_cast<const uchar4*>(ix + m);
          const float4 v0 = *reinterpret_cast<const float4*>(tile + j4.x * TW + col);
          const float4 v1 = *reinterpret_cast<const float4*>(tile + j4.y * TW + col);
          const float4 v2 = *reinterpret_cast<const float4*>(tile + j4.z * TW + col);
          const float4 v3 = *reinterpret_cast<const float4*>(tile + j4.w * TW + col);
          acc.x = fmaf(c4.x, v0.x, acc.x); acc.y = fmaf(c4.x, v0.y, acc.y); acc.z = fmaf(c4.x, v0.z, acc.z); acc.w = fmaf(c4.x, v0.w, acc.w);
          acc.x = fmaf(c4.y, v1.x, acc.x); acc.y = fmaf(c4.y, v1.y, acc.y); acc.z = fmaf(c4.y, v1.z, acc.z); acc.w = fmaf(c4.y, v1.w, acc.w);
          acc.x = fmaf(c4.z, v2.x, acc.x); acc.y = fmaf(c4.z, v2.y, acc.y); acc.z = fmaf(c4.z, v2.z, acc.z); acc.w = fmaf(c4.z, v2.w, acc.w);
          acc.x = fmaf(c4.w, v3.x, acc.x); acc.y = fmaf(c4.w, v3.y, acc.y); acc.z = fmaf(c4.w, v3.z, acc.z); acc.w = fmaf(c4.w, v3.w, acc.w);
        }
        if (POOL || (p.flags & VQA_GC_RELU)) {
          acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        }
        if (POOL) {
          if (acc.x > best.x) { best.x = acc.x; barg.x = i; }
          if (acc.y > best.y) { best.y = acc.y; barg.y = i; }
          if (acc.z > best.z) { best.z = acc.z; barg.z = i; }
          if (acc.w > best.w) { best.w = acc.w; barg.w = i; }
        } else {
          const long long row = (long long)b * K + i;
          if (p.drop_p > 0.f) {
            const Philox rng(p.seed);
            const uint4 r = rng((unsigned long long)((row * p.out_dim + colg) >> 2), p.offset + (p.step_ptr ? *p.step_ptr * 16ull : 0ull));
            acc.x = u32_to_unit(r.x) >= p.drop_p ? acc.x * p.drop_scale : 0.f;
            acc.y = u32_to_unit(r.y) >= p.drop_p ? acc.y * p.drop_scale : 0.f;
            acc.z = u32_to_unit(r.z) >= p.drop_p ? acc.z * p.drop_scale : 0.f;
            acc.w = u32_to_unit(r.w) >= p.drop_p ? acc.w * p.drop_scale : 0.f;
          }
          *reinterpret_cast<float4*>(p.out + row * p.ldo + colg) = acc;
        }
      }
    }
    if (POOL) {
      float* pv = reinterpret_cast<float*>(sm + L.scratch) + (size_t)(t & 1) * GC_WARPS * TW * 2;
      int* pa = reinterpret_cast<int*>(pv + GC_WARPS * TW);
      if (active) {
        *reinterpret_cast<float4*>(pv + warp * TW + col) = best;
        *reinterpret_cast<int4*>(pa + warp * TW + col) = barg;
      }
      __syncthreads();
      if (tid < TW) {
        float bv = pv[tid];
        int ba = pa[tid];
        for (int w = 1; w < GC_WARPS; ++w) {
          const float v = pv[w * TW + tid];
          const int a = pa[w * TW + tid];
          if (v > bv || (v == bv && a < ba)) { bv = v; ba = a; }   // ties -> first (lowest) node index
        }
        const long long o = (long long)b * p.out_dim + (t0 + t) * TW + tid;
        p.pooled[o] = bv;
        p.argmax[o] = ba;
        p.hq[o] = fmaxf(p.q[o], 0.f) * bv;
      }
    } else {
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------ backward (data path)
// 16 per-lane partials -> 16 warp totals with 16 shuffles; lane l ends up holding total number
// ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1)  (both lanes of an even/odd pair hold the same value).
__device__ __forceinline__ float reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float send = (lane & 16) ? v[r] : v[r + 8], keep = (lane & 16) ? v[r + 8] : v[r];
    v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float send = (lane & 8) ? v[r] : v[r + 4], keep = (lane & 8) ? v[r + 4] : v[r];
    v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float send = (lane & 4) ? v[r] : v[r + 2], keep = (lane & 4) ? v[r + 2] : v[r];
    v[r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v[0] : v[1], keep = (lane & 2) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool POOLED>
__global__ void __launch_bounds__(GC_THREADS)
graphconv_bwd_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmD, const GcParams p,
                     const GcSmem L) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);   // pointer arithmetic keeps the .shared address space (LDS/STS)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, K = p.K, nb = p.nb, nbp = p.nbp, TW = p.TW;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, p.ntiles - t0);
  const int k_lo = (t0 * TW) / p.D;
  const int nkc = min(p.nkc, p.nk - k_lo);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  float* tiles = reinterpret_cast<float*>(sm + L.tiles);
  constexpr int TPS = POOLED ? 1 : 2;                    // tiles per stage: Y (+ dO)
  const uint32_t tile_bytes = (uint32_t)K * TW * 4;
  const int NS = p.nstage;
  auto issue = [&](int t) {
    const int s = t % NS;
    float* dst = tiles + (size_t)s * TPS * p.tstride;
    mbar_arrive_expect_tx(&bars[s], tile_bytes * TPS);
    tma_load_2d(dst, &tmY, &bars[s], (t0 + t) * TW, b * K);
    if (!POOLED) tma_load_2d(dst + p.tstride, &tmD, &bars[s], (t0 + t) * TW, b * K);
  };
  if (tid == 0) {
    tma_prefetch_desc(&tmY);
    if (!POOLED) tma_prefetch_desc(&tmD);
    for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int t = 0; t < NS - 1 && t < nt; ++t) issue(t);
  gc_prologue(p, L, sm, b, k_lo, nkc);

  const float* coef = reinterpret_cast<const float*>(sm + L.coef);
  float* pacc = reinterpret_cast<float*>(sm + L.pacc);
  const uint8_t* idx8 = sm + L.idx8;
  uint16_t* rev_cnt = reinterpret_cast<uint16_t*>(sm + L.rev_cnt);
  uint16_t* rev_e = reinterpret_cast<uint16_t*>(sm + L.rev_e);
  for (int v = tid; v < nkc * K * nbp; v += GC_THREADS) pacc[v] = 0.f;
  __syncthreads();   // idx8 visible
  // reverse neighbour lists, ordered by source row i (deterministic): rev[j] = { e = i*nbp+m : idx[i,m] == j }
  for (int j = warp; j < K; j += GC_WARPS) {
    int cnt = 0;
    for (int i0 = 0; i0 < K; i0 += 32) {
      const int i = i0 + lane;
      int hit = -1;
      if (i < K)
        for (int m = 0; m < nb; ++m)
          if (idx8[i * nbp + m] == j) hit = m;
      const unsigned mask = __ballot_sync(0xffffffffu, hit >= 0);
      if (hit >= 0) rev_e[j * K + cnt + __popc(mask & ((1u << lane) - 1))] = (uint16_t)(i * nbp + hit);
      cnt += __popc(mask);
    }
    if (lane == 0) rev_cnt[j] = (uint16_t)cnt;
  }
  __syncthreads();

  const int col = lane * 4;
  const bool active = col < TW;
  for (int t = 0; t < nt; ++t) {
    if (tid == 0) {
      if (NS > 1) { if (t + NS - 1 < nt) issue(t + NS - 1); }
      else issue(t);
    }
    mbar_wait(&bars[t % NS], (t / NS) & 1);
    const float* ytile = tiles + (size_t)(t % NS) * TPS * p.tstride;
    const float* dtile = ytile + p.tstride;
    const int colg = (t0 + t) * TW + col;
    const int kk = ((t0 + t) * TW) / p.D - k_lo;
    float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 ar = make_int4(-1, -1, -1, -1);
    if (POOLED && active) {
      const long long o = (long long)b * p.out_dim + colg;
      dp = *reinterpret_cast<const float4*>(p.dpooled + o);
      ar = make_int4((int)p.argmax_in[o], (int)p.argmax_in[o + 1], (int)p.argmax_in[o + 2], (int)p.argmax_in[o + 3]);
    }
    auto load_dO = [&](int i) -> float4 {
      if (POOLED) return make_float4(ar.x == i ? dp.x : 0.f, ar.y == i ? dp.y : 0.f, ar.z == i ? dp.z : 0.f, ar.w == i ? dp.w : 0.f);
      return *reinterpret_cast<const float4*>(dtile + i * TW + col);
    };
    // (a) dY[j] = sum over incoming edges (i,m) of coef * dO[i]   (skipped when the dense kernel already produced dY)
    for (int j = warp; j < K && p.out != nullptr; j += GC_WARPS) {
      if (active) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = rev_cnt[j];
        const uint16_t* re = rev_e + j * K;
        const float* cf = coef + kk * K * nbp;
        for (int r = 0; r < n; ++r) {
          const int e = re[r];
          const float c = cf[e];
          const float4 d = load_dO(e / nbp);
          acc.x = fmaf(c, d.x, acc.x); acc.y = fmaf(c, d.y, acc.y); acc.z = fmaf(c, d.z, acc.z); acc.w = fmaf(c, d.w, acc.w);
        }
        *reinterpret_cast<float4*>(p.out + ((long long)b * K + j) * p.ldo + colg) = acc;
      }
    }
    // (b) P[i,m,k] += <dO[i, tile cols], Y[idx[i,m], tile cols]>
    for (int i = warp; i < K; i += GC_WARPS) {
      const float4 d = active ? load_dO(i) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int m0 = 0; m0 < nbp; m0 += 16) {
        float part[16];
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
          part[mm] = 0.f;
          if (m0 + mm < nbp && active) {
            const float4 y = *reinterpret_cast<const float4*>(ytile + idx8[i * nbp + m0 + mm] * TW + col);
            part[mm] = fmaf(d.x, y.x, fmaf(d.y, y.y, fmaf(d.z, y.z, d.w * y.w)));
          }
        }
        const float tot = reduce16(part, lane);
        const int mm = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        if (!(lane & 1) && m0 + mm < nb) pacc[(kk * K + i) * nbp + m0 + mm] += tot;   // the warp owns row i
      }
    }
    __syncthreads();
  }
  for (int v = tid; v < nkc * K * nb; v += GC_THREADS) {
    const int kk = v / (K * nb), r = v - kk * K * nb, i = r / nb, m = r - i * nb;
    p.P[(((long long)b * K + i) * nb + m) * p.nk + k_lo + kk] = pacc[(kk * K + i) * nbp + m];
  }
}

// ------------------------------------------------------------------------------------------ backward (edge finish)
constexpr int EDGE_THREADS = 256;
constexpr int MAX_NK = 64;
__global__ void __launch_bounds__(EDGE_THREADS)
graphconv_edge_bwd_kernel(const float* __restrict__ P, const int* __restrict__ idx, const float* __restrict__ alpha,
                          const float* __restrict__ boxes, long long ldbox, const float* __restrict__ gauss,
                          float* __restrict__ dalpha, float* __restrict__ partial, long long nedges, int K, int nb, int nk) {
  __shared__ float gs[4 * MAX_NK];
  __shared__ float wsum[EDGE_THREADS / 32][4 * MAX_NK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int v = tid; v < 4 * nk; v += EDGE_THREADS) gs[v] = gauss[v];
  for (int v = tid; v < (EDGE_THREADS / 32) * 4 * MAX_NK; v += EDGE_THREADS) (&wsum[0][0])[v] = 0.f;
  __syncthreads();
  const long long e = (long long)blockIdx.x * EDGE_THREADS + tid;
  const bool ok = e < nedges;
  float rho = 0.f, theta = 0.f, S = 1.f, a = 1.f, T = 0.f;
  const float* Pe = P + (ok ? e : 0) * nk;
  if (ok) {
    const long long node = e / nb;                 // b*K + i
    const long long b = node / K;
    const int j = idx[e];
    const float* bi = boxes + node * ldbox;
    const float* bj = boxes + (b * K + j) * ldbox;
    polar(bi[0] + 0.5f * (bi[2] - bi[0]), bi[1] + 0.5f * (bi[3] - bi[1]), bj[0] + 0.5f * (bj[2] - bj[0]),
          bj[1] + 0.5f * (bj[3] - bj[1]), rho, theta);
    S = 0.f;
    for (int k = 0; k < nk; ++k) S += gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
    a = alpha ? alpha[e] : 1.f;
    float da = 0.f;
    for (int k = 0; k < nk; ++k) {
      const float w = __fdiv_rn(gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]), S);
      da = fmaf(w, Pe[k], da);                    // dalpha = sum_k w_k P_k ;  T = sum_k dw_k w_k = alpha * dalpha
    }
    if (dalpha) dalpha[e] = da;
    T = a * da;
  }
  for (int k = 0; k < nk; ++k) {
    float c_mr = 0.f, c_sr = 0.f, c_mt = 0.f, c_st = 0.f;
    if (ok) {
      const float mr = gs[k], sr = gs[nk + k], mt = gs[2 * nk + k], st = gs[3 * nk + k];
      const float g = gauss_val(rho, theta, mr, sr, mt, st);
      const float dg = (a * Pe[k] - T) / S;
      const float gam = g * dg;                    // masked (NaN->0) kernels contribute 0
      const float vr = GAUSS_EPS_F + sr * sr, vt = GAUSS_EPS_F + st * st;
      const float dr = rho - mr;
      c_mr = gam * dr / vr;
      c_sr = gam * dr * dr * sr / (vr * vr);
      const float df = theta - mt;
      const float phi = fabsf(df), two = TWO_PI_F - phi, psi = fabsf(two);
      const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
      float ddel;                                  // d(delta)/d(mean_theta), delta = min(phi, psi)  
      if (phi < psi) ddel = -sgn;
      else ddel = (two > 0.f ? 1.f : (two < 0.f ? -1.f : 0.f)) * sgn;
      const float del = fminf(phi, psi);
      c_mt = gam * (-del / vt) * ddel;
      c_st = gam * del * del * st / (vt * vt);
      if (gam != gam) { c_mr = c_sr = c_mt = c_st = gam; }   // keep NaN visible (S == 0 rows), as autograd would
    }
    c_mr = warp_sum(c_mr); c_sr = warp_sum(c_sr); c_mt = warp_sum(c_mt); c_st = warp_sum(c_st);
    if (lane == 0) { wsum[warp][k] = c_mr; wsum[warp][nk + k] = c_sr; wsum[warp][2 * nk + k] = c_mt; wsum[warp][3 * nk + k] = c_st; }
  }
  __syncthreads();
  for (int v = tid; v < 4 * nk; v += EDGE_THREADS) {
    float s = 0.f;
    for (int w = 0; w < EDGE_THREADS / 32; ++w) s += wsum[w][v];
    partial[(long long)blockIdx.x * 4 * nk + v] = s;
  }
}

__global__ void __launch_bounds__(256)
gaussian_weights_kernel(const float* __restrict__ pseudo, const float* __restrict__ gauss, float* __restrict__ w, long long n, int nk) {
  __shared__ float gs[4 * MAX_NK];
  for (int v = threadIdx.x; v < 4 * nk; v += 256) gs[v] = gauss[v];
  __syncthreads();
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  const float rho = pseudo[2 * e], theta = pseudo[2 * e + 1];
  float S = 0.f;
  for (int k = 0; k < nk; ++k) S += gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]);
  for (int k = 0; k < nk; ++k)
    w[e * nk + k] = __fdiv_rn(gauss_val(rho, theta, gs[k], gs[nk + k], gs[2 * nk + k], gs[3 * nk + k]), S);
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 gc_get_encode() {
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
static int make_tile_map(CUtensorMap* tm, const float* ptr, long long ld, long long rows, int cols, int TW, int K) {
  auto enc = gc_get_encode();
  if (!enc) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TW, (cuuint32_t)K}, estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled(graphconv tile) failed: CUresult %d", (int)r);
  return VQA_OK;
}

struct GcPlan { int D, TW, tstride, ntiles, tiles_per_cta, nslab, nkc, nstage, nbp, smem; GcSmem L; };

// Pick the tile width, the slab (tiles per CTA), the coefficient-table depth and the TMA ring depth that fit in 227 KB.
static int gc_plan(GcPlan* pl, int B, int K, int nb, int nk, int out_dim, bool bwd, bool pool, bool pooled_bwd, const char* who) {
  VQA_CHECK_ARG(B > 0 && K > 0 && K <= 128, "%s: need 0 < K <= 128 (K=%d)", who, K);
  VQA_CHECK_ARG(nb > 0 && nb <= K, "%s: neighbourhood size must be in [1,K] (nb=%d, K=%d)", who, nb, K);
  VQA_CHECK_ARG(nk > 0 && nk <= MAX_NK, "%s: n_kernels must be in [1,%d] (nk=%d)", who, MAX_NK, nk);
  VQA_CHECK_ARG(out_dim > 0 && out_dim % nk == 0, "%s: out_dim (%d) must be divisible by n_kernels (%d)", who, out_dim, nk);
  const int D = out_dim / nk;
  VQA_CHECK_ARG(D % 4 == 0, "%s: out_dim / n_kernels (%d) must be a multiple of 4", who, D);
  const int nbp = (nb + 3) & ~3;
  const int tps = bwd ? (pooled_bwd ? 1 : 2) : 1;
  for (int TW = 128; TW >= 4; TW >>= 1) {
    if (D % TW) continue;
    const int ntiles = out_dim / TW, tpk = D / TW;
    int want = (4 * kNumSMs + B - 1) / B;          // ~2 waves at 2 CTAs / SM
    if (want < 1) want = 1;
    if (want > ntiles) want = ntiles;
    int tpc0 = (ntiles + want - 1) / want;
    const int step = bwd ? tpk : 1;                // bwd: P[.,.,k] sums over ALL tiles of kernel k -> keep them in one CTA
    if (bwd) tpc0 = (tpc0 + tpk - 1) / tpk * tpk;
    for (int tpc = tpc0; tpc >= step; tpc -= step) {
      int nkc;
      if (tpc % tpk == 0) nkc = tpc / tpk;
      else if (tpk % tpc == 0) nkc = 1;
      else nkc = (tpc + tpk - 1) / tpk + 1;
      if (nkc > nk) nkc = nk;
      const int ns_max = tpc < 4 ? tpc : 4;
      for (int ns = ns_max; ns >= 1; --ns) {
        const GcSmem L = gc_smem_layout(K, nbp, nk, TW, nkc, ns, tps, bwd, pool);
        if (L.total + 128 > GC_SMEM_MAX) continue;
        if (ns < 2 && ns_max >= 2 && tpc > step) break;   // a single-stage ring serialises load and compute: shrink the slab first
        pl->D = D; pl->TW = TW; pl->tstride = align_up(K * TW * 4, 128) / 4; pl->ntiles = ntiles; pl->tiles_per_cta = tpc;
        pl->nslab = (ntiles + tpc - 1) / tpc; pl->nkc = nkc; pl->nstage = ns; pl->nbp = nbp; pl->smem = L.total + 128; pl->L = L;
        return VQA_OK;
      }
    }
  }
  return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: no shared-memory plan for K=%d nb=%d nk=%d out=%d", who, K, nb, nk, out_dim);
}

static int gc_fwd_common(bool pool, const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                         long long ldbox, const float* gauss, float* out, long long ldo, const float* q, float* pooled,
                         long long* argmax, float* hq, int B, int K, int nb, int nk, int out_dim, int flags, float drop_p,
                         unsigned long long seed, unsigned long long offset, const unsigned long long* step_ptr, cudaStream_t stream) {
  const char* who = pool ? "vqa_graphconv_pool_fwd_f32" : "vqa_graphconv_fwd_f32";
  VQA_CHECK_ARG(Y && idx && boxes && gauss, "%s: null pointer", who);
  VQA_CHECK_ARG(aligned16(Y) && (ldy & 3) == 0 && ldy >= out_dim, "%s: Y must be 16-byte aligned with ld %% 4 == 0", who);
  VQA_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "%s: dropout p must be in [0,1)", who);
  if (K <= 128 && nb <= K && nk <= MAX_NK && out_dim % nk == 0) {   // preferred: dense-register kernel
    DenseParams dp{};
    dp.in = Y; dp.ldin = ldy; dp.idx = idx; dp.alpha = alpha; dp.boxes = boxes; dp.ldbox = ldbox; dp.gauss = gauss;
    dp.out = out; dp.ldo = ldo; dp.q = q; dp.pooled = pooled; dp.argmax = argmax; dp.hq = hq;
    dp.K = K; dp.nb = nb; dp.nk = nk; dp.out_dim = out_dim; dp.flags = flags;
    dp.drop_p = drop_p; dp.drop_scale = 1.f / (1.f - drop_p); dp.drop_thresh16 = (unsigned)(drop_p * 65536.f + 0.5f);
    dp.seed = seed; dp.offset = offset; dp.step_ptr = step_ptr;
    if (pool) dp.ldo = 2;
    const int rc = pool ? dense_launch<DM_FWD_POOL>(dp, B, stream) : dense_launch<DM_FWD>(dp, B, stream);
    if (rc <= 0) return rc;
  }
  GcPlan pl;
  if (int rc = gc_plan(&pl, B, K, nb, nk, out_dim, false, pool, false, who)) return rc;
  CUtensorMap tm;
  if (int rc = make_tile_map(&tm, Y, ldy, (long long)B * K, out_dim, pl.TW, K)) return rc;
  GcParams p{};
  p.Y = Y; p.ldy = ldy; p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.out = out; p.ldo = ldo; p.q = q; p.pooled = pooled; p.argmax = argmax; p.hq = hq;
  p.K = K; p.nb = nb; p.nbp = pl.nbp; p.nk = nk; p.out_dim = out_dim; p.D = pl.D; p.TW = pl.TW; p.tstride = pl.tstride;
  p.tiles_per_cta = pl.tiles_per_cta; p.ntiles = pl.ntiles; p.nkc = pl.nkc; p.nstage = pl.nstage; p.flags = flags;
  p.drop_p = drop_p; p.drop_scale = 1.f / (1.f - drop_p); p.seed = seed; p.offset = offset; p.step_ptr = step_ptr;
  dim3 grid(pl.nslab, B);
  if (pool) {
    VQA_CUDA(cudaFuncSetAttribute(graphconv_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    graphconv_fwd_kernel<true><<<grid, GC_THREADS, pl.smem, stream>>>(tm, p, pl.L);
  } else {
    VQA_CUDA(cudaFuncSetAttribute(graphconv_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    graphconv_fwd_kernel<false><<<grid, GC_THREADS, pl.smem, stream>>>(tm, p, pl.L);
  }
  VQA_LAUNCH_CHECK("graphconv_fwd_kernel");
  return VQA_OK;
}

}  // namespace vqa
using namespace vqa;

extern "C" int vqa_graphconv_fwd_f32(const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                                     long long ldbox, const float* gauss, float* out, long long ldo, int B, int K, int nb,
                                     int nk, int out_dim, int flags, float dropout_p, unsigned long long seed,
                                     unsigned long long offset, const unsigned long long* step_ptr, cudaStream_t stream) {
  VQA_CHECK_ARG(out && aligned16(out) && (ldo & 3) == 0 && ldo >= out_dim, "vqa_graphconv_fwd_f32: out must be 16-byte aligned with ld %% 4 == 0");
  return gc_fwd_common(false, Y, ldy, idx, alpha, boxes, ldbox, gauss, out, ldo, nullptr, nullptr, nullptr, nullptr, B, K,
                       nb, nk, out_dim, flags, dropout_p, seed, offset, step_ptr, stream);
}

extern "C" int vqa_graphconv_pool_fwd_f32(const float* Y, long long ldy, const int* idx, const float* boxes, long long ldbox,
                                          const float* gauss, const float* q, float* pooled, long long* argmax, float* hq,
                                          int B, int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  VQA_CHECK_ARG(q && pooled && argmax && hq, "vqa_graphconv_pool_fwd_f32: null pointer");
  return gc_fwd_common(true, Y, ldy, idx, nullptr, boxes, ldbox, gauss, nullptr, 0, q, pooled, argmax, hq, B, K, nb, nk,
                       out_dim, VQA_GC_RELU, 0.f, 0, 0, nullptr, stream);
}

extern "C" int vqa_graphconv_bwd_f32(const float* dO, long long lddo, const float* dpooled, const long long* argmax,
                                     const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                                     long long ldbox, const float* gauss, float* dY, long long lddy, float* P, int B, int K,
                                     int nb, int nk, int out_dim, cudaStream_t stream) {
  const char* who = "vqa_graphconv_bwd_f32";
  const bool pooled = dO == nullptr;
  VQA_CHECK_ARG(idx && boxes && gauss && (P || dY) && (Y || !P), "%s: null pointer", who);
  VQA_CHECK_ARG(pooled ? (dpooled && argmax) : true, "%s: need either dO or (dpooled, argmax)", who);
  VQA_CHECK_ARG((!Y || (aligned16(Y) && (ldy & 3) == 0)) && (!dY || (aligned16(dY) && (lddy & 3) == 0)), "%s: Y/dY alignment", who);
  VQA_CHECK_ARG(pooled ? aligned16(dpooled) : (aligned16(dO) && (lddo & 3) == 0), "%s: upstream gradient alignment", who);
  bool dy_done = dY == nullptr;            // dY == NULL: the caller computes dY elsewhere (tensor-core path); only P is wanted
  if (!dy_done && K <= 128 && nb <= K && nk <= MAX_NK && out_dim % nk == 0) {   // dY = M^T dO on the dense-register kernel
    DenseParams dp{};
    dp.in = dO; dp.ldin = pooled ? 2 : lddo; dp.dpooled = dpooled; dp.argmax_in = argmax;
    dp.idx = idx; dp.alpha = alpha; dp.boxes = boxes; dp.ldbox = ldbox; dp.gauss = gauss; dp.out = dY; dp.ldo = lddy;
    dp.K = K; dp.nb = nb; dp.nk = nk; dp.out_dim = out_dim;
    const int rc = pooled ? dense_launch<DM_BWD_POOLED>(dp, B, stream) : dense_launch<DM_BWD_DENSE>(dp, B, stream);
    if (rc < 0) return rc;
    dy_done = rc == 0;
  }
  if (!P) {   // data path only: the edge products are computed elsewhere (tensor-core path)
    if (dy_done) return VQA_OK;
    return vqa_fail(VQA_ERR_UNSUPPORTED, "%s: dY-only mode needs a shape the dense kernel supports", who);
  }
  GcPlan pl;
  if (int rc = gc_plan(&pl, B, K, nb, nk, out_dim, true, false, pooled, who)) return rc;
  CUtensorMap tmY, tmD;
  if (int rc = make_tile_map(&tmY, Y, ldy, (long long)B * K, out_dim, pl.TW, K)) return rc;
  if (!pooled) { if (int rc = make_tile_map(&tmD, dO, lddo, (long long)B * K, out_dim, pl.TW, K)) return rc; }
  else tmD = tmY;
  GcParams p{};
  p.Y = Y; p.ldy = ldy; p.dO = dO; p.lddo = lddo; p.dpooled = dpooled; p.argmax_in = argmax;
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss; p.out = dy_done ? nullptr : dY; p.ldo = lddy; p.P = P;
  p.K = K; p.nb = nb; p.nbp = pl.nbp; p.nk = nk; p.out_dim = out_dim; p.D = pl.D; p.TW = pl.TW; p.tstride = pl.tstride;
  p.tiles_per_cta = pl.tiles_per_cta; p.ntiles = pl.ntiles; p.nkc = pl.nkc; p.nstage = pl.nstage;
  dim3 grid(pl.nslab, B);
  if (pooled) {
    VQA_CUDA(cudaFuncSetAttribute(graphconv_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    graphconv_bwd_kernel<true><<<grid, GC_THREADS, pl.smem, stream>>>(tmY, tmD, p, pl.L);
  } else {
    VQA_CUDA(cudaFuncSetAttribute(graphconv_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    graphconv_bwd_kernel<false><<<grid, GC_THREADS, pl.smem, stream>>>(tmY, tmD, p, pl.L);
  }
  VQA_LAUNCH_CHECK("graphconv_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_graphconv_edge_blocks(int B, int K, int nb) {
  const long long n = (long long)B * K * nb;
  return (int)((n + EDGE_THREADS - 1) / EDGE_THREADS);
}

extern "C" int vqa_graphconv_edge_bwd_f32(const float* P, const int* idx, const float* alpha, const float* boxes,
                                          long long ldbox, const float* gauss, float* dalpha, float* dgauss_partial, int B,
                                          int K, int nb, int nk, cudaStream_t stream) {
  VQA_CHECK_ARG(P && idx && boxes && gauss && dgauss_partial, "vqa_graphconv_edge_bwd_f32: null pointer");
  VQA_CHECK_ARG(nk > 0 && nk <= MAX_NK && K > 0 && nb > 0 && B > 0, "vqa_graphconv_edge_bwd_f32: bad sizes");
  const long long n = (long long)B * K * nb;
  graphconv_edge_bwd_kernel<<<vqa_graphconv_edge_blocks(B, K, nb), EDGE_THREADS, 0, stream>>>(P, idx, alpha, boxes, ldbox, gauss, dalpha,
                                                                                           dgauss_partial, n, K, nb, nk);
  VQA_LAUNCH_CHECK("graphconv_edge_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_gaussian_weights_f32(const float* pseudo, const float* gauss, float* w, long long n, int nk, cudaStream_t stream) {
  VQA_CHECK_ARG(pseudo && gauss && w && n > 0 && nk > 0 && nk <= MAX_NK, "vqa_gaussian_weights_f32: bad arguments");
  gaussian_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pseudo, gauss, w, n, nk);
  VQA_LAUNCH_CHECK("gaussian_weights_kernel");
  return VQA_OK;
}

// ------------------------------------------------------------------------------------------------ layer-level patch operator
// layers.py:127-137 on MATERIALISED neighbourhoods (the module API `NeighbourhoodGraphConvolution.convolution`): per node n
//   Z[n, k, :] = sum_m w[n, m, k] * X[n, m, :]          (torch.bmm(weights^T, neighbourhood) in the reference)
// HBM-bound: X (n, nb, F) is read once, Z (n, nk, F) written once; one CTA per node, a thread owns float4 columns and keeps up
// to 8 kernels' accumulators in registers.  Backward: dX[n,m,:] = sum_k w[n,m,k] dZ[n,k,:], dw[n,m,k] = <X[n,m,:], dZ[n,k,:]>.
namespace vqa {
constexpr int PO_THREADS = 256;

__global__ void __launch_bounds__(PO_THREADS) patch_operator_fwd_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                                        float* __restrict__ Z, int nb, int nk, int F) {
  extern __shared__ float ws[];                         // [nb][nk]
  const long long n = blockIdx.x;
  for (int v = threadIdx.x; v < nb * nk; v += PO_THREADS) ws[v] = w[n * nb * nk + v];
  __syncthreads();
  const float* Xn = X + n * (long long)nb * F;
  float* Zn = Z + n * (long long)nk * F;
  const bool vec = (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Z)) & 15) == 0;
  if (vec) {
    const int F4 = F >> 2;
    for (int c = threadIdx.x; c < F4; c += PO_THREADS)
      for (int k0 = 0; k0 < nk; k0 += 8) {
        float4 acc[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = 0; m < nb; ++m) {
          const float4 x = reinterpret_cast<const float4*>(Xn + (long long)m * F)[c];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float wk = k0 + kk < nk ? ws[m * nk + k0 + kk] : 0.f;
            acc[kk].x = fmaf(wk, x.x, acc[kk].x); acc[kk].y = fmaf(wk, x.y, acc[kk].y);
            acc[kk].z = fmaf(wk, x.z, acc[kk].z); acc[kk].w = fmaf(wk, x.w, acc[kk].w);
          }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (k0 + kk < nk) reinterpret_cast<float4*>(Zn + (long long)(k0 + kk) * F)[c] = acc[kk];
      }
  } else {
    for (int c = threadIdx.x; c < F; c += PO_THREADS)
      for (int k = 0; k < nk; ++k) {
        float acc = 0.f;
        for (int m = 0; m < nb; ++m) acc = fmaf(ws[m * nk + k], Xn[(long long)m * F + c], acc);
        Zn[(long long)k * F + c] = acc;
      }
  }
}

__global__ void __launch_bounds__(PO_THREADS) patch_operator_bwd_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                                        const float* __restrict__ dZ, float* __restrict__ dX,
                                                                        float* __restrict__ dw, int nb, int nk, int F) {
  extern __shared__ float ws[];                         // [nb][nk]
  const long long n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = threadIdx.x; v < nb * nk; v += PO_THREADS) ws[v] = w[n * nb * nk + v];
  __syncthreads();
  const float* Xn = X + n * (long long)nb * F;
  const float* dZn = dZ + n * (long long)nk * F;
  if (dX) {
    float* dXn = dX + n * (long long)nb * F;
    for (int c = threadIdx.x; c < F; c += PO_THREADS)
      for (int m = 0; m < nb; ++m) {
        float acc = 0.f;
        for (int k = 0; k < nk; ++k) acc = fmaf(ws[m * nk + k], dZn[(long long)k * F + c], acc);   // dZ rows of this node stay in L1
        dXn[(long long)m * F + c] = acc;
      }
  }
  if (dw) {
    for (int pr = warp; pr < nb * nk; pr += PO_THREADS / 32) {       // one (m, k) pair per warp pass
      const int m = pr / nk, k = pr - m * nk;
      float acc = 0.f;
      for (int c = lane; c < F; c += 32) acc = fmaf(Xn[(long long)m * F + c], dZn[(long long)k * F + c], acc);
      acc = warp_sum(acc);
      if (lane == 0) dw[n * nb * nk + pr] = acc;
    }
  }
}
}  // namespace vqa

extern "C" int vqa_patch_operator_fwd_f32(const float* X, const float* w, float* Z, long long n, int nb, int nk, int F, cudaStream_t stream) {
  VQA_CHECK_ARG(X && w && Z && n > 0 && nb > 0 && nk > 0 && F > 0 && nb * nk <= 8192, "vqa_patch_operator_fwd_f32: bad arguments");
  vqa::patch_operator_fwd_kernel<<<(unsigned)n, vqa::PO_THREADS, (size_t)nb * nk * sizeof(float), stream>>>(X, w, Z, nb, nk, F);
  VQA_LAUNCH_CHECK("patch_operator_fwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_patch_operator_bwd_f32(const float* X, const float* w, const float* dZ, float* dX, float* dw, long long n, int nb,
                                          int nk, int F, cudaStream_t stream) {
  VQA_CHECK_ARG(X && w && dZ && (dX || dw) && n > 0 && nb > 0 && nk > 0 && F > 0 && nb * nk <= 8192, "vqa_patch_operator_bwd_f32: bad arguments");
  vqa::patch_operator_bwd_kernel<<<(unsigned)n, vqa::PO_THREADS, (size_t)nb * nk * sizeof(float), stream>>>(X, w, dZ, dX, dw, nb, nk, F);
  VQA_LAUNCH_CHECK("patch_operator_bwd_kernel");
  return VQA_OK;
}
