// Tail of the training step: the criterion and the optimiser of the reference drivers as
// single HBM-streaming kernels.
//   mlsm_loss_fwd_kernel / mlsm_loss_bwd_kernel : nn.MultiLabelSoftMarginLoss (run.py:382,431) - one pass each instead of
//                                                 torch's ~20 pointwise/reduce launches;
//   adam_flat_kernel                            : torch.optim.Adam (run.py:392,435) over the flat gradient buffer of
//                                                 vqa_b200.ddp.GradReducer - every parameter tensor in ONE launch driven by a
//                                                 device-side chunk table, the 1/world gradient average folded in.
// All three are bandwidth bound: 8 / 12 bytes per logit, 28 bytes per parameter.
#include "common.cuh"
#include "../../include/vqa_b200.h"

namespace vqa {

// log(sigmoid(x)) the way ATen evaluates it: min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (wid == 0) {
    t = lane < 8 ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// loss = scale * sum_i -( y_i logsig(x_i) + (1 - y_i) logsig(-x_i) ); scale = 1/(B*A) for reduction='mean'.
// Each block reduces a contiguous slice and stores its partial sum; the last block to finish adds the partials in block order
// (deterministic) and leaves the counter zero for the next launch.
__global__ void __launch_bounds__(256) mlsm_loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            long long n, float scale, float* __restrict__ partial,
                                                            int* __restrict__ counter, float* __restrict__ loss, bool vec) {
  __shared__ float red[8];
  __shared__ bool last;
  const long long per = ((n + gridDim.x - 1) / gridDim.x + 3) & ~3LL;
  const long long lo = per * blockIdx.x, hi = min(n, lo + per);
  float acc = 0.f;
  if (vec) {
    for (long long i = lo + 4LL * threadIdx.x; i < hi; i += 1024) {
      if (i + 4 <= hi) {
        const float4 a = *reinterpret_cast<const float4*>(x + i);
        const float4 t = *reinterpret_cast<const float4*>(y + i);
        acc -= t.x * log_sigmoid(a.x) + (1.f - t.x) * log_sigmoid(-a.x);
        acc -= t.y * log_sigmoid(a.y) + (1.f - t.y) * log_sigmoid(-a.y);
        acc -= t.z * log_sigmoid(a.z) + (1.f - t.z) * log_sigmoid(-a.z);
        acc -= t.w * log_sigmoid(a.w) + (1.f - t.w) * log_sigmoid(-a.w);
      } else {
        for (long long j = i; j < hi; ++j) acc -= y[j] * log_sigmoid(x[j]) + (1.f - y[j]) * log_sigmoid(-x[j]);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) acc -= y[i] * log_sigmoid(x[i]) + (1.f - y[i]) * log_sigmoid(-x[i]);
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += __ldcg(partial + i);
  __syncthreads();
  t = block_sum_256(t, red);
  if (threadIdx.x == 0) {
    *loss = t * scale;
    *counter = 0;
  }
}

// dx_i = (sigmoid(x_i) - y_i) * scale * (*gout)
__global__ void __launch_bounds__(256) mlsm_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gout, float* __restrict__ dx, long long n,
                                                            float scale, bool vec) {
  const float g = scale * (gout ? __ldg(gout) : 1.f);
  const long long stride = (long long)gridDim.x * 256;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
      const float4 a = reinterpret_cast<const float4*>(x)[i];
      const float4 t = reinterpret_cast<const float4*>(y)[i];
      float4 o;
      o.x = (1.f / (1.f + expf(-a.x)) - t.x) * g;
      o.y = (1.f / (1.f + expf(-a.y)) - t.y) * g;
      o.z = (1.f / (1.f + expf(-a.z)) - t.z) * g;
      o.w = (1.f / (1.f + expf(-a.w)) - t.w) * g;
      reinterpret_cast<float4*>(dx)[i] = o;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
      dx[i] = (1.f / (1.f + expf(-x[i])) - y[i]) * g;
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
      dx[i] = (1.f / (1.f + expf(-x[i])) - y[i]) * g;
  }
}

// ------------------------------------------------------------------------------------------------ Adam
// chunk table: 3 x int64 per chunk = {address of the parameter elements, offset into the flat grad/m/v buffers, count}.
// state[0] = number of steps taken so far, state[1] = block counter (zero between launches).  hyper[0] = learning rate.
struct AdamChunk {
  long long param;
  long long off;
  long long count;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float gs, float wd, float b1, float b2,
                                         float step_size, float inv_sqrt_bc2, float eps) {
  g *= gs;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(g - m, 1.f - b1, m);
  v = fmaf(v, b2, (1.f - b2) * g * g);
  const float denom = fmaf(sqrtf(v), inv_sqrt_bc2, eps);
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(const AdamChunk* __restrict__ chunks, int nchunks,
                                                        const float* __restrict__ grad, float* __restrict__ exp_avg,
                                                        float* __restrict__ exp_avg_sq, const float* __restrict__ hyper,
                                                        float b1, float b2, float eps, float wd, float grad_scale,
                                                        int* __restrict__ state) {
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  __shared__ int s_step;
  if (threadIdx.x == 0) {
    const int step = state[0] + 1;                      // every block reads it before it adds itself to state[1]
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    s_step = step;
    s_step_size = (float)((double)__ldg(hyper) / bc1);
    s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, isb2 = s_inv_sqrt_bc2;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const AdamChunk ch = chunks[c];
    float* p = reinterpret_cast<float*>(ch.param);
    const float* g = grad + ch.off;
    float* m = exp_avg + ch.off;
    float* v = exp_avg_sq + ch.off;
    const int cnt = (int)ch.count;
    const bool vec = ((ch.param | (ch.off << 2)) & 15) == 0;
    if (vec) {
      const int n4 = cnt >> 2;
      for (int i = threadIdx.x; i < n4; i += 256) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, grad_scale, wd, b1, b2, step_size, isb2, eps);
        adam_one(pp.y, gg.y, mm.y, vv.y, grad_scale, wd, b1, b2, step_size, isb2, eps);
        adam_one(pp.z, gg.z, mm.z, vv.z, grad_scale, wd, b1, b2, step_size, isb2, eps);
        adam_one(pp.w, gg.w, mm.w, vv.w, grad_scale, wd, b1, b2, step_size, isb2, eps);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += 256)
        adam_one(p[i], g[i], m[i], v[i], grad_scale, wd, b1, b2, step_size, isb2, eps);
    } else {
      for (int i = threadIdx.x; i < cnt; i += 256)
        adam_one(p[i], g[i], m[i], v[i], grad_scale, wd, b1, b2, step_size, isb2, eps);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[0] = s_step;
      state[1] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ data-parallel Adam over NVLink peer memory
// One process per GPU; symmetric memory maps every rank's buffers into every other rank's address space over NVLink / NVSwitch.
// The reference has no distributed code; the stock recipe is an all-reduce of the gradients (measured at N = 8: 370 us for the
// 121 MB buffer, with NCCL's CTAs competing with backward for SMs) followed by N identical 848 MB Adam passes (155 us).  Here:
//   reduce-scatter, pushed: every rank owns a 1/N slice of the flat index space.  While backward runs, each finished gradient bucket
//       is COPIED (cudaMemcpyAsync over NVLink: copy engines, no SMs) into the owners' receive buffers R_r[q] (vqa_b200.ddp).
//       Pushing, not pulling: measured with this kernel's first version, remote LOADS sustain ~330 GB/s per GPU, remote STORES and
//       copy-engine pushes ~600 GB/s (profiles/r02_p2p_adam.md).
//   this kernel, per rank: g[i] = sum_q (q == rank ? G[i] : R[q][i - lo]) in rank order (all local reads), Adam on the slice (moments
//       are sharded: each rank touches 1/N of exp_avg / exp_avg_sq), and the all-gather: the updated parameters are stored into every
//       rank's flat parameter buffer P_q[i] (remote stores).
// Two flag barriers (p2p_barrier_kernel) bracket it: all pushes landed before anyone sums, all parameters written before anyone's
// next forward.
constexpr int P2P_MAX_WORLD = 16;
struct P2PPtrs { float* p[P2P_MAX_WORLD]; float* mc; };   // mc: multicast address of the parameter buffers (NVLS) or null

// Ownership is interleaved: the flat index space is cut into chunks of 2^ch_log2 floats, chunk c belongs to rank c % world ("row" k =
// chunks k * world .. k * world + world - 1), so every gradient bucket spreads evenly over the owners and its push can start the
// moment the bucket is complete.  The rank's elements in its own order: n4 float4s, element e = (row e >> chl4, offset e & mask).
__global__ void __launch_bounds__(256) adam_p2p_kernel(const __grid_constant__ P2PPtrs pp, const float* __restrict__ grad,
                                                       const float* __restrict__ recv, long long n4, int chl4, float* __restrict__ exp_avg,
                                                       float* __restrict__ exp_avg_sq, int rank, int world,
                                                       const float* __restrict__ hyper, float b1, float b2, float eps, float wd,
                                                       float grad_scale, int* __restrict__ state) {
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  __shared__ int s_step;
  if (threadIdx.x == 0) {
    const int step = state[0] + 1;
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    s_step = step;
    s_step_size = (float)((double)__ldg(hyper) / bc1);
    s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, isb2 = s_inv_sqrt_bc2;
  float* pl = pp.p[rank];
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  const float4* r4 = reinterpret_cast<const float4*>(recv);
  const long long mask = (1LL << chl4) - 1;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n4; e += (long long)gridDim.x * 256) {
    const long long i = (((e >> chl4) * world + rank) << chl4) + (e & mask);     // flat float4 index of my e-th element
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int q = 0; q < world; ++q) {                      // rank order: the sum does not depend on who owns the element
      const float4 t = q == rank ? __ldcs(g4 + i) : __ldcs(r4 + q * n4 + e);
      gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
    }
    float4 w = reinterpret_cast<const float4*>(pl)[i];
    float4 mm = reinterpret_cast<float4*>(exp_avg)[i];
    float4 vv = reinterpret_cast<float4*>(exp_avg_sq)[i];
    adam_one(w.x, gg.x, mm.x, vv.x, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.y, gg.y, mm.y, vv.y, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.z, gg.z, mm.z, vv.z, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.w, gg.w, mm.w, vv.w, grad_scale, wd, b1, b2, step_size, isb2, eps);
    reinterpret_cast<float4*>(exp_avg)[i] = mm;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = vv;
    if (pp.mc) {                                            // one multicast store: the switch delivers it to every rank
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                   ::"l"(pp.mc + 4 * i), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
    } else {
#pragma unroll 4
      for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(pp.p[q])[i] = w;
    }
  }
  __threadfence_system();                                   // my parameter stores are performed before the barrier kernel announces them
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[0] = s_step;
      state[1] = 0;
    }
  }
}

// The same step through the NVSwitch's multicast objects (NVLS): with the gradient and parameter buffers bound to a multicast address,
// ONE multimem.ld_reduce returns the sum of an element over all ranks (added inside the switch: the rank receives its 1/N slice once
// instead of N-1 copies of it) and ONE multimem.st delivers the updated parameter to every rank.  No pushes, no receive buffers, one
// store instruction instead of N.  The rank owns the contiguous flat slice [lo4, hi4) (float4 units).
__global__ void __launch_bounds__(256) adam_mc_kernel(const float* __restrict__ mc_grad, float* __restrict__ mc_param,
                                                      const float* __restrict__ param, float* __restrict__ exp_avg,
                                                      float* __restrict__ exp_avg_sq, long long lo4, long long hi4,
                                                      const float* __restrict__ hyper, float b1, float b2, float eps, float wd,
                                                      float grad_scale, int* __restrict__ state) {
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  __shared__ int s_step;
  if (threadIdx.x == 0) {
    const int step = state[0] + 1;
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    s_step = step;
    s_step_size = (float)((double)__ldg(hyper) / bc1);
    s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, isb2 = s_inv_sqrt_bc2;
  for (long long i = lo4 + (long long)blockIdx.x * 256 + threadIdx.x; i < hi4; i += (long long)gridDim.x * 256) {
    float4 gg;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(gg.x), "=f"(gg.y), "=f"(gg.z), "=f"(gg.w) : "l"(mc_grad + 4 * i) : "memory");
    float4 w = reinterpret_cast<const float4*>(param)[i];
    float4 mm = reinterpret_cast<float4*>(exp_avg)[i];
    float4 vv = reinterpret_cast<float4*>(exp_avg_sq)[i];
    adam_one(w.x, gg.x, mm.x, vv.x, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.y, gg.y, mm.y, vv.y, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.z, gg.z, mm.z, vv.z, grad_scale, wd, b1, b2, step_size, isb2, eps);
    adam_one(w.w, gg.w, mm.w, vv.w, grad_scale, wd, b1, b2, step_size, isb2, eps);
    reinterpret_cast<float4*>(exp_avg)[i] = mm;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = vv;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(mc_param + 4 * i), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
      state[0] = s_step;
      state[1] = 0;
    }
  }
}

// Flag barrier across the ranks of one node: flags[q] = address of rank q's flag array (world ints, zero at start) as mapped HERE;
// *epoch counts the barriers this rank has passed.  Thread q announces the new epoch in rank q's slot [rank] (release, system scope)
// and waits for rank q's announcement in the local slot [q] (acquire).  Bounded: a rank that never arrives traps the launch after ~4 s
// of SM clock instead of hanging the GPU.
struct P2PFlags { int* f[P2P_MAX_WORLD]; };
__global__ void p2p_barrier_kernel(const __grid_constant__ P2PFlags fl, int rank, int world, int* __restrict__ epoch) {
  __shared__ int s_e;
  if (threadIdx.x == 0) { s_e = *epoch + 1; }
  __syncthreads();
  const int e = s_e, q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    asm volatile("st.global.release.sys.b32 [%0], %1;" ::"l"(fl.f[q] + rank), "r"(e) : "memory");
    const int* mine = fl.f[rank] + q;
    const long long t0 = clock64();
    for (;;) {
      int seen;
      asm volatile("ld.global.acquire.sys.b32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if (seen - e >= 0) break;
      if (clock64() - t0 > 8000000000LL) {
        printf("vqa_b200: p2p barrier timed out (rank %d waiting for rank %d, epoch %d, saw %d)\n", rank, q, e, seen);
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}

}  // namespace vqa
using namespace vqa;

static inline int loss_blocks(long long n) { return (int)max(1LL, min((long long)kNumSMs * 4, (n + 2047) / 2048)); }

extern "C" int vqa_mlsm_loss_blocks(long long n) { return n > 0 ? loss_blocks(n) : 0; }

extern "C" int vqa_mlsm_loss_fwd_f32(const float* logits, const float* target, long long n, float scale, float* partial,
                                     int* counter, float* loss, cudaStream_t stream) {
  VQA_CHECK_ARG(logits && target && partial && counter && loss && n > 0, "vqa_mlsm_loss_fwd_f32: bad arguments");
  const bool vec = aligned16(logits) && aligned16(target);
  mlsm_loss_fwd_kernel<<<loss_blocks(n), 256, 0, stream>>>(logits, target, n, scale, partial, counter, loss, vec);
  VQA_LAUNCH_CHECK("mlsm_loss_fwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_mlsm_loss_bwd_f32(const float* logits, const float* target, const float* grad_out, float* dlogits,
                                     long long n, float scale, cudaStream_t stream) {
  VQA_CHECK_ARG(logits && target && dlogits && n > 0, "vqa_mlsm_loss_bwd_f32: bad arguments");
  const bool vec = aligned16(logits) && aligned16(target) && aligned16(dlogits);
  const int blocks = (int)min((long long)kNumSMs * 8, (n / 4 + 255) / 256 + 1);
  mlsm_loss_bwd_kernel<<<blocks, 256, 0, stream>>>(logits, target, grad_out, dlogits, n, scale, vec);
  VQA_LAUNCH_CHECK("mlsm_loss_bwd_kernel");
  return VQA_OK;
}

extern "C" int vqa_adam_flat_f32(const long long* chunks, int nchunks, const float* grad, float* exp_avg, float* exp_avg_sq,
                                 const float* lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                 int* state, cudaStream_t stream) {
  VQA_CHECK_ARG(chunks && grad && exp_avg && exp_avg_sq && lr && state && nchunks > 0, "vqa_adam_flat_f32: bad arguments");
  VQA_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                "vqa_adam_flat_f32: betas must be in [0,1) and eps >= 0 (got %f, %f, %g)", beta1, beta2, eps);
  static_assert(sizeof(AdamChunk) == 24, "chunk table layout");
  const int blocks = min(nchunks, kNumSMs * 4);      // 56 registers: four 256-thread blocks per SM
  adam_flat_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const AdamChunk*>(chunks), nchunks, grad, exp_avg, exp_avg_sq,
                                               lr, beta1, beta2, eps, weight_decay, grad_scale, state);
  VQA_LAUNCH_CHECK("adam_flat_kernel");
  return VQA_OK;
}

extern "C" int vqa_p2p_barrier(const long long* flag_addrs, int rank, int world, int* epoch, cudaStream_t stream) {
  VQA_CHECK_ARG(flag_addrs && epoch && world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world,
                "vqa_p2p_barrier: bad arguments (rank %d of %d, at most %d ranks)", rank, world, P2P_MAX_WORLD);
  P2PFlags fl{};
  for (int q = 0; q < world; ++q) {
    VQA_CHECK_ARG(flag_addrs[q] != 0, "vqa_p2p_barrier: null flag array for rank %d", q);
    fl.f[q] = reinterpret_cast<int*>(flag_addrs[q]);
  }
  p2p_barrier_kernel<<<1, 32, 0, stream>>>(fl, rank, world, epoch);
  VQA_LAUNCH_CHECK("p2p_barrier_kernel");
  return VQA_OK;
}

extern "C" int vqa_adam_flat_p2p(const float* grad, const float* recv, long long n_own, int chunk_log2, const long long* param_addrs,
                                 float* mc_param, float* exp_avg, float* exp_avg_sq, int rank, int world, const float* lr, float beta1, float beta2,
                                 float eps, float weight_decay, float grad_scale, int* state, cudaStream_t stream) {
  VQA_CHECK_ARG(grad && recv && param_addrs && exp_avg && exp_avg_sq && lr && state, "vqa_adam_flat_p2p: null pointer");
  VQA_CHECK_ARG(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "vqa_adam_flat_p2p: rank %d of %d (at most %d ranks)", rank, world, P2P_MAX_WORLD);
  VQA_CHECK_ARG(chunk_log2 >= 2 && chunk_log2 <= 30 && n_own >= 0 && (n_own & ((1LL << chunk_log2) - 1)) == 0,
                "vqa_adam_flat_p2p: the rank's %lld elements must be whole chunks of 2^%d floats", n_own, chunk_log2);
  VQA_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "vqa_adam_flat_p2p: betas must be in [0,1) and eps >= 0");
  VQA_CHECK_ARG(aligned16(grad) && aligned16(recv) && aligned16(exp_avg) && aligned16(exp_avg_sq), "vqa_adam_flat_p2p: buffers must be 16-byte aligned");
  P2PPtrs pp{};
  for (int q = 0; q < world; ++q) {
    VQA_CHECK_ARG(param_addrs[q] && (param_addrs[q] & 15) == 0, "vqa_adam_flat_p2p: rank %d's parameter buffer must be mapped and 16-byte aligned", q);
    pp.p[q] = reinterpret_cast<float*>(param_addrs[q]);
  }
  pp.mc = mc_param;
  const long long n4 = n_own >> 2;
  if (n4 == 0) return VQA_OK;
  long long blocks = (n4 + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  adam_p2p_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pp, grad, recv, n4, chunk_log2 - 2, exp_avg, exp_avg_sq, rank, world, lr, beta1, beta2, eps,
                                                        weight_decay, grad_scale, state);
  VQA_LAUNCH_CHECK("adam_p2p_kernel");
  return VQA_OK;
}

extern "C" int vqa_memcpy2d_async(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height,
                                  cudaStream_t stream) {
  VQA_CHECK_ARG(dst && src && width > 0 && height > 0 && dpitch >= width && spitch >= width, "vqa_memcpy2d_async: bad arguments");
  VQA_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, cudaMemcpyDeviceToDevice, stream));
  return VQA_OK;
}

extern "C" int vqa_adam_flat_mc(const float* mc_grad, float* mc_param, const float* param, float* exp_avg, float* exp_avg_sq, long long lo,
                                long long hi, const float* lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                int* state, cudaStream_t stream) {
  VQA_CHECK_ARG(mc_grad && mc_param && param && exp_avg && exp_avg_sq && lr && state, "vqa_adam_flat_mc: null pointer");
  VQA_CHECK_ARG(lo >= 0 && hi >= lo && (lo & 3) == 0 && (hi & 3) == 0, "vqa_adam_flat_mc: the slice [%lld, %lld) must be float4 aligned", lo, hi);
  VQA_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "vqa_adam_flat_mc: betas must be in [0,1) and eps >= 0");
  VQA_CHECK_ARG(aligned16(mc_grad) && aligned16(mc_param) && aligned16(param) && aligned16(exp_avg) && aligned16(exp_avg_sq), "vqa_adam_flat_mc: buffers must be 16-byte aligned");
  const long long n4 = (hi - lo) >> 2;
  if (n4 == 0) return VQA_OK;
  long long blocks = (n4 + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  adam_mc_kernel<<<(unsigned)blocks, 256, 0, stream>>>(mc_grad, mc_param, param, exp_avg, exp_avg_sq, lo >> 2, hi >> 2, lr, beta1, beta2, eps,
                                                       weight_decay, grad_scale, state);
  VQA_LAUNCH_CHECK("adam_mc_kernel");
  return VQA_OK;
}
