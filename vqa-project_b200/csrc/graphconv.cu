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

// ------------------------------------------------------------------------------------------ forward
template <bool POOL>
__global__ void __launch_bounds__(GC_THREADS)
graphconv_fwd_kernel(const __grid_constant__ CUtensorMap tmY, const GcParams p, const GcSmem L) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 127) & ~uintptr_t(127));
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
            const uint4 r = rng((unsigned long long)((row * p.out_dim + colg) >> 2), p.offset);
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
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 127) & ~uintptr_t(127));
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
    // (a) dY[j] = sum over incoming edges (i,m) of coef * dO[i]
    for (int j = warp; j < K; j += GC_WARPS) {
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
                         unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  const char* who = pool ? "vqa_graphconv_pool_fwd_f32" : "vqa_graphconv_fwd_f32";
  VQA_CHECK_ARG(Y && idx && boxes && gauss, "%s: null pointer", who);
  VQA_CHECK_ARG(aligned16(Y) && (ldy & 3) == 0 && ldy >= out_dim, "%s: Y must be 16-byte aligned with ld %% 4 == 0", who);
  VQA_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "%s: dropout p must be in [0,1)", who);
  GcPlan pl;
  if (int rc = gc_plan(&pl, B, K, nb, nk, out_dim, false, pool, false, who)) return rc;
  CUtensorMap tm;
  if (int rc = make_tile_map(&tm, Y, ldy, (long long)B * K, out_dim, pl.TW, K)) return rc;
  GcParams p{};
  p.Y = Y; p.ldy = ldy; p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss;
  p.out = out; p.ldo = ldo; p.q = q; p.pooled = pooled; p.argmax = argmax; p.hq = hq;
  p.K = K; p.nb = nb; p.nbp = pl.nbp; p.nk = nk; p.out_dim = out_dim; p.D = pl.D; p.TW = pl.TW; p.tstride = pl.tstride;
  p.tiles_per_cta = pl.tiles_per_cta; p.ntiles = pl.ntiles; p.nkc = pl.nkc; p.nstage = pl.nstage; p.flags = flags;
  p.drop_p = drop_p; p.drop_scale = 1.f / (1.f - drop_p); p.seed = seed; p.offset = offset;
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
                                     unsigned long long offset, cudaStream_t stream) {
  VQA_CHECK_ARG(out && aligned16(out) && (ldo & 3) == 0 && ldo >= out_dim, "vqa_graphconv_fwd_f32: out must be 16-byte aligned with ld %% 4 == 0");
  return gc_fwd_common(false, Y, ldy, idx, alpha, boxes, ldbox, gauss, out, ldo, nullptr, nullptr, nullptr, nullptr, B, K,
                       nb, nk, out_dim, flags, dropout_p, seed, offset, stream);
}

extern "C" int vqa_graphconv_pool_fwd_f32(const float* Y, long long ldy, const int* idx, const float* boxes, long long ldbox,
                                          const float* gauss, const float* q, float* pooled, long long* argmax, float* hq,
                                          int B, int K, int nb, int nk, int out_dim, cudaStream_t stream) {
  VQA_CHECK_ARG(q && pooled && argmax && hq, "vqa_graphconv_pool_fwd_f32: null pointer");
  return gc_fwd_common(true, Y, ldy, idx, nullptr, boxes, ldbox, gauss, nullptr, 0, q, pooled, argmax, hq, B, K, nb, nk,
                       out_dim, VQA_GC_RELU, 0.f, 0, 0, stream);
}

extern "C" int vqa_graphconv_bwd_f32(const float* dO, long long lddo, const float* dpooled, const long long* argmax,
                                     const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                                     long long ldbox, const float* gauss, float* dY, long long lddy, float* P, int B, int K,
                                     int nb, int nk, int out_dim, cudaStream_t stream) {
  const char* who = "vqa_graphconv_bwd_f32";
  const bool pooled = dO == nullptr;
  VQA_CHECK_ARG(Y && idx && boxes && gauss && dY && P, "%s: null pointer", who);
  VQA_CHECK_ARG(pooled ? (dpooled && argmax) : true, "%s: need either dO or (dpooled, argmax)", who);
  VQA_CHECK_ARG(aligned16(Y) && (ldy & 3) == 0 && aligned16(dY) && (lddy & 3) == 0, "%s: Y/dY alignment", who);
  VQA_CHECK_ARG(pooled ? aligned16(dpooled) : (aligned16(dO) && (lddo & 3) == 0), "%s: upstream gradient alignment", who);
  GcPlan pl;
  if (int rc = gc_plan(&pl, B, K, nb, nk, out_dim, true, false, pooled, who)) return rc;
  CUtensorMap tmY, tmD;
  if (int rc = make_tile_map(&tmY, Y, ldy, (long long)B * K, out_dim, pl.TW, K)) return rc;
  if (!pooled) { if (int rc = make_tile_map(&tmD, dO, lddo, (long long)B * K, out_dim, pl.TW, K)) return rc; }
  else tmD = tmY;
  GcParams p{};
  p.Y = Y; p.ldy = ldy; p.dO = dO; p.lddo = lddo; p.dpooled = dpooled; p.argmax_in = argmax;
  p.idx = idx; p.alpha = alpha; p.boxes = boxes; p.ldbox = ldbox; p.gauss = gauss; p.out = dY; p.ldo = lddy; p.P = P;
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
