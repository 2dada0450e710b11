// Dense projections of the VQA hot path as tcgen05 / TMEM tensor-core GEMMs fed by TMA (sm_100a).
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] )        fp32 in HBM on both sides.
//
// Replaces the reference's cuBLAS call sites: GraphLearner linears (layers.py:185-190), the per-kernel
// conv Linears (layers.py:140-142, run here as ONE projection), the classifier
// (sparse_graph_model.py:154-157) and every dX / dW product autograd derives from them.
//
// Precision modes (fp32 tensors stay fp32 in HBM; the split/convert happens on the smem tile):
//   TF32X3 : each fp32 operand tile is split in shared memory into hi = rna_tf32(x), lo = rna_tf32(x - hi) by
//            4 "transform" warps; the MMA warp issues lo*hi + hi*lo + hi*hi into the same TMEM accumulator
//            (error ~2^-21 per product: fp32-grade, passes the 1e-3 parity budget; single-pass TF32 does not).
//   TF32   : one kind::tf32 MMA on the raw tile (debug / speed reference).
//   TF32X3_HP : TF32X3 plus "chunked promotion": the tensor core accumulates with truncation, so the error of a long
//            contraction grows linearly with K (measured 1.5e-5 at K=2052).  In HP mode the MMA warp alternates between
//            two TMEM accumulators every 4 k-blocks (K=128); 4 dedicated drain warps add each finished chunk into fp32
//            REGISTER accumulators (round-to-nearest) while the next chunk is being multiplied.  Measured error at
//            K=2052: 1.2e-6 = cuBLAS-SIMT-fp32 grade.  Used for the graph-learner forward, whose output feeds exp().
//
// Structure: one 128 x BN output tile per CTA, 6 warps: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer
// (single thread), warps 2-5 = operand transform during the main loop, then the TMEM -> register -> global epilogue.
// Three mbarrier rings (raw-full, transformed-full, empty) + one accumulator-ready barrier.  Operands may be
// K-major (contraction contiguous) or MN-major (contraction strided: the dW = dY^T X products) - both are loaded
// by TMA with the 128-byte swizzle and described to the tensor core through the UMMA shared-memory descriptor.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>

#include "../../include/vqa_b200.h"

namespace vqa {

constexpr int BM = 128;
constexpr int BK = 32;                  // fp32 elements per k-block: one 128-byte swizzle row
constexpr int A_TILE = BM * 128;        // bytes
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_THREADS_HP = 320;      // + 4 drain warps
constexpr int HP_CHUNK = 4;               // k-blocks per promoted chunk (K = 128)
constexpr int SMEM_BUDGET = 232448 - 1024 - 256;

struct GemmParams {
  float* C;
  long long ldc;
  int M, N, Kc;
  int a_mn, b_mn;
  const float* bias;
  const float* rowb;
  long long ldrb;
  int group;
  const float* aux;
  long long ldaux;
  float aux_scale;
  int flags;
  int kb_per_split, num_kb;
};

template <int BN, int PREC>
struct Cfg {
  static constexpr bool HP = PREC == VQA_PREC_TF32X3_HP;
  static constexpr bool X3 = PREC == VQA_PREC_TF32X3 || HP;
  static constexpr int THREADS = HP ? GEMM_THREADS_HP : GEMM_THREADS;
  static constexpr int TMEM_COLS = HP ? 2 * BN : BN;
  static constexpr int B_TILE = BN * 128;
  static constexpr int RAW = A_TILE + B_TILE;
  static constexpr int STAGE = RAW * (X3 ? 2 : 1);
  static constexpr int S_ = SMEM_BUDGET / STAGE;
  static constexpr int S = S_ > 8 ? 8 : S_;
  static constexpr int SMEM = S * STAGE + 1024 + 256;   // barriers: 3S+1 ring/acc + tmem slot + 4 HP = <= 30 x 8 B
};

// UMMA shared-memory descriptor.
//   K-major  (contraction contiguous): rows of 128 B, 16-byte chunks XOR-swizzled per 8-row group (SWIZZLE_128B),
//                                      SBO = 1024 B between 8-row groups, LBO unused.
//   MN-major (contraction strided)   : 32-bit operands only support the "128B swizzle with 32-byte atoms" layout
//                                      (SWIZZLE_128B_BASE32B <-> TMA SWIZZLE_128B_ATOM_32B): rows of 128 B = 32 MN
//                                      elements, 4 contraction rows per swizzle atom -> SBO = 512 B between 4-row
//                                      groups, LBO = 4096 B between 32-element MN chunks (one TMA box each).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int mn_major) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);               // start address  [0,14)
  d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;              // leading byte offset [16,30)
  d |= (uint64_t)(mn_major ? (512 >> 4) : (1024 >> 4)) << 32;     // stride byte offset  [32,46)
  d |= 1ull << 46;                                                // descriptor version (Blackwell)
  d |= (mn_major ? 1ull : 2ull) << 61;                            // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// epilogue for 32 consecutive columns [cb, cb+32) of one output row
struct Epi {
  const GemmParams& p;
  float* crow; const float* rb; const float* ax; bool vec_ok, relu, atomic;
  __device__ __forceinline__ Epi(const GemmParams& p_, int row) : p(p_) {
    relu = p.flags & VQA_GEMM_RELU; atomic = p.flags & VQA_GEMM_ATOMIC_ADD;
    crow = p.C + (long long)row * p.ldc;
    rb = p.rowb ? p.rowb + (long long)(row / p.group) * p.ldrb : nullptr;
    ax = p.aux ? p.aux + (long long)row * p.ldaux : nullptr;
    vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0) &&
             (!ax || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0))) &&
             (!rb || (((p.ldrb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.rowb) & 15) == 0))) &&
             (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
  }
  __device__ __forceinline__ void store32(int cb, const float* r) const {
    if (vec_ok && !atomic) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = cb + j;
        if (col >= p.N) break;
        float4 v = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        if (rb) { const float4 a = *reinterpret_cast<const float4*>(rb + col); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (p.bias) { const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + col)); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (ax) {
          const float4 a = *reinterpret_cast<const float4*>(ax + col);
          v.x = a.x > 0.f ? v.x * p.aux_scale : 0.f; v.y = a.y > 0.f ? v.y * p.aux_scale : 0.f;
          v.z = a.z > 0.f ? v.z * p.aux_scale : 0.f; v.w = a.w > 0.f ? v.w * p.aux_scale : 0.f;
        }
        *reinterpret_cast<float4*>(crow + col) = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = cb + j;
        if (col >= p.N) break;
        float v = r[j];
        if (rb) v += rb[col];
        if (p.bias) v += p.bias[col];
        if (relu) v = fmaxf(v, 0.f);
        if (ax) v = ax[col] > 0.f ? v * p.aux_scale : 0.f;
        if (atomic) atomicAdd(crow + col, v); else crow[col] = v;
      }
    }
  }
};

template <int BN, int PREC>
__global__ void __launch_bounds__(Cfg<BN, PREC>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = Cfg<BN, PREC>;
  constexpr int S = C::S;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the .shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * C::STAGE);
  uint64_t* full_raw = bars;
  uint64_t* full_xf = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* acc_full = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);
  uint64_t* hp_full = bars + 3 * S + 2;      // HP: accumulator a holds a finished chunk
  uint64_t* hp_empty = bars + 3 * S + 4;     // HP: accumulator a has been drained

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_stop = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_stop - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_raw[s], 1);
        mbar_init(&full_xf[s], 128);
        mbar_init(&empty[s], 1);
      }
      mbar_init(acc_full, 1);
      if (C::HP) {
        mbar_init(&hp_full[0], 1); mbar_init(&hp_full[1], 1);
        mbar_init(&hp_empty[0], 128); mbar_init(&hp_empty[1], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S, ph = (i / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_raw[s], C::RAW);
        const int k0 = (kb_begin + i) * BK;
        uint8_t* a_dst = smem + s * C::STAGE;
        uint8_t* b_dst = a_dst + A_TILE;
        if (!p.a_mn) {
          tma_load_2d(a_dst, &tmA, &full_raw[s], k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_dst + c * 4096, &tmA, &full_raw[s], m0 + c * 32, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_dst, &tmB, &full_raw[s], k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(b_dst + c * 4096, &tmB, &full_raw[s], n0 + c * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t a_step = p.a_mn ? 1024 : 32, b_step = p.b_mn ? 1024 : 32;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S, ph = (i / S) & 1;
      const int chunk = C::HP ? i / HP_CHUNK : 0, in_chunk = C::HP ? i % HP_CHUNK : i;
      const int accsel = chunk & 1;
      if (C::HP && in_chunk == 0) {               // wait until the drain warps emptied this accumulator
        mbar_wait(&hp_empty[accsel], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(C::X3 ? &full_xf[s] : &full_raw[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_tmem = tmem_base + (C::HP ? accsel * BN : 0);
        const uint32_t a_hi = smem_u32(smem + s * C::STAGE), b_hi = a_hi + A_TILE;
        const uint32_t a_lo = a_hi + C::RAW, b_lo = a_lo + A_TILE;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t dah = umma_desc(a_hi + ks * a_step, p.a_mn), dbh = umma_desc(b_hi + ks * b_step, p.b_mn);
          const uint32_t acc = (in_chunk > 0 || ks > 0) ? 1u : 0u;
          if (C::X3) {
            const uint64_t dal = umma_desc(a_lo + ks * a_step, p.a_mn), dbl = umma_desc(b_lo + ks * b_step, p.b_mn);
            tc_mma<0>(d_tmem, dal, dbh, idesc, acc);   // small terms first
            tc_mma<0>(d_tmem, dah, dbl, idesc, 1u);
            tc_mma<0>(d_tmem, dah, dbh, idesc, 1u);
          } else {
            tc_mma<0>(d_tmem, dah, dbh, idesc, acc);
          }
        }
        tc_commit(&empty[s]);                       // smem slot reusable once these MMAs retire
        if (C::HP) { if (in_chunk == HP_CHUNK - 1 || i == nkb - 1) tc_commit(&hp_full[accsel]); }
        else if (i == nkb - 1) tc_commit(acc_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ warps 2-5: operand transform, then (non-HP) epilogue
    const int t = threadIdx.x - 64;
    if (C::X3) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S, ph = (i / S) & 1;
        mbar_wait(&full_raw[s], ph);
        float4* raw = reinterpret_cast<float4*>(smem + s * C::STAGE);
        float4* lo = reinterpret_cast<float4*>(smem + s * C::STAGE + C::RAW);
#pragma unroll 4
        for (int v = t; v < C::RAW / 16; v += 128) {
          const float4 x = raw[v];
          float4 h, l;
          h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
          l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y); l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
          raw[v] = h;
          lo[v] = l;
        }
        fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&full_xf[s]);
      }
    }
    if (!C::HP) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int q = warp & 3;                 // TMEM lane quarter this warp may access
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const Epi epi(p, row);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;            // warp-uniform
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tc_wait_ld();
        if (row_ok) epi.store32(n0 + c0, reinterpret_cast<const float*>(r));
      }
    }
  } else {
    // ------------------------------------------------------------ HP only, warps 6-9: chunk drain + epilogue
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    float racc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) racc[j] = 0.f;
    const int nchunks = (nkb + HP_CHUNK - 1) / HP_CHUNK;
    for (int c = 0; c < nchunks; ++c) {
      const int accsel = c & 1;
      mbar_wait(&hp_full[accsel], (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(accsel * BN + c0), r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) racc[c0 + j] += __uint_as_float(r[j]);   // fp32 round-to-nearest promotion
      }
      tc_fence_before();
      mbar_arrive(&hp_empty[accsel]);
    }
    if (row < p.M) {
      const Epi epi(p, row);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32)
        if (n0 + c0 < p.N) epi.store32(n0 + c0, racc + c0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
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

// operand stored either (rows = MN, cols = K contiguous) [k-major] or (rows = K, cols = MN contiguous) [mn-major]
static int make_operand_map(CUtensorMap* tm, const float* ptr, long long ld, int mn_extent, int k_extent, int mn_major, int tile_mn) {
  auto enc = get_encode();
  if (!enc) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) { dims[0] = k_extent; dims[1] = mn_extent; box[0] = BK; box[1] = tile_mn; }
  else           { dims[0] = mn_extent; dims[1] = k_extent; box[0] = 32; box[1] = BK; }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return vqa_fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (ld=%lld mn=%d k=%d)", (int)r, ld, mn_extent, k_extent);
  return VQA_OK;
}

template <int BN, int PREC>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits, cudaStream_t st) {
  using C = Cfg<BN, PREC>;
  static bool attr_set = false;
  if (!attr_set) {
    VQA_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  gemm_tcgen05_kernel<BN, PREC><<<grid, C::THREADS, C::SMEM, st>>>(ta, tb, p);
  VQA_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return VQA_OK;
}

}  // namespace vqa

using namespace vqa;

extern "C" int vqa_gemm_f32(const float* A, long long lda, int a_mn_major, const float* B, long long ldb, int b_mn_major,
                            float* C, long long ldc, int M, int N, int Kc, const float* bias, const float* rowbcast,
                            long long ldrb, int group, const float* aux, long long ldaux, float aux_scale, int flags,
                            int precision, int split_k, int tile_n, cudaStream_t stream) {
  VQA_CHECK_ARG(A && B && C, "vqa_gemm_f32: null operand");
  VQA_CHECK_ARG(M > 0 && N > 0 && Kc > 0, "vqa_gemm_f32: empty problem M=%d N=%d K=%d", M, N, Kc);
  VQA_CHECK_ARG(aligned16(A) && aligned16(B), "vqa_gemm_f32: operands must be 16-byte aligned for TMA");
  VQA_CHECK_ARG((lda & 3) == 0 && (ldb & 3) == 0, "vqa_gemm_f32: leading dimensions must be multiples of 4 floats (TMA 16-byte strides), got lda=%lld ldb=%lld", lda, ldb);
  VQA_CHECK_ARG(lda >= (a_mn_major ? M : Kc) && ldb >= (b_mn_major ? N : Kc) && ldc >= N, "vqa_gemm_f32: leading dimension smaller than the row length");
  VQA_CHECK_ARG(precision == VQA_PREC_TF32X3 || precision == VQA_PREC_TF32 || precision == VQA_PREC_TF32X3_HP, "vqa_gemm_f32: unknown precision %d", precision);
  VQA_CHECK_ARG(!rowbcast || group > 0, "vqa_gemm_f32: rowbcast needs group > 0");
  const int num_kb = (Kc + BK - 1) / BK;
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > num_kb) splits = num_kb;
  int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;   // no empty split
  if (splits > 1) {
    VQA_CHECK_ARG(!bias && !rowbcast && !aux && !(flags & VQA_GEMM_RELU), "vqa_gemm_f32: split-K supports the plain epilogue only");
    flags |= VQA_GEMM_ATOMIC_ADD;      // caller zero-fills C
  }
  int bn = tile_n;
  if (bn == 0) bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  if (precision == VQA_PREC_TF32X3_HP && bn > 128) bn = 128;   // register accumulators: 128 columns per drain thread
  VQA_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "vqa_gemm_f32: tile_n must be 64, 128 or 256");

  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, A, lda, M, Kc, a_mn_major, BM);
  if (rc) return rc;
  rc = make_operand_map(&tb, B, ldb, N, Kc, b_mn_major, bn);
  if (rc) return rc;
  GemmParams p{C, ldc, M, N, Kc, a_mn_major ? 1 : 0, b_mn_major ? 1 : 0, bias, rowbcast, ldrb, group, aux, ldaux, aux_scale, flags, per, num_kb};
#define VQA_DISPATCH(BN_)                                                                     \
  (precision == VQA_PREC_TF32X3 ? launch<BN_, VQA_PREC_TF32X3>(ta, tb, p, splits, stream)     \
                                : launch<BN_, VQA_PREC_TF32>(ta, tb, p, splits, stream))
  if (precision == VQA_PREC_TF32X3_HP)
    return bn == 128 ? launch<128, VQA_PREC_TF32X3_HP>(ta, tb, p, splits, stream) : launch<64, VQA_PREC_TF32X3_HP>(ta, tb, p, splits, stream);
  if (bn == 256) return VQA_DISPATCH(256);
  if (bn == 128) return VQA_DISPATCH(128);
  return VQA_DISPATCH(64);
#undef VQA_DISPATCH
}
