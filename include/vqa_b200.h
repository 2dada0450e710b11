/* libvqa_sm100.so -- C ABI of the B200-native conditioned-graph VQA hot path.
 *
 * The reference (Originofamonia/vqa-project) has no FFI layer: its hot path is a chain of torch library calls
 * inside sparse_graph_model.py / layers.py.  This header is the boundary a maintainer binds instead (ctypes stub
 * in INTEGRATION.md); each entry point names the reference lines it replaces (paths relative to the reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch owns all buffers, incl. workspaces);
 *   - kernels are enqueued on `stream`, never allocate, never synchronise;
 *   - return 0 on success, negative on error; vqa_last_error() gives the (thread-local) message;
 *   - row-major fp32 everywhere; index tensors are int32 (neighbour ids) or int64 (argmax, as the reference
 *     returns it); B = images, K = nodes/image (<= 128), nb = neighbourhood size (<= K), nk = Gaussian kernels.
 */
#ifndef VQA_B200_H_
#define VQA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vqa_stream_t; /* cudaStream_t */

#define VQA_ABI_VERSION 11

/* GEMM precision modes */
#define VQA_PREC_TF32X3 0 /* 3-pass split TF32 on tcgen05: fp32-grade results (parity mode)      */
#define VQA_PREC_TF32 1   /* single-pass TF32 on tcgen05                                           */
#define VQA_PREC_TF32X3_HP 2 /* TF32X3 + fp32 promotion of the accumulator every K=128: cuBLAS-fp32-grade */

/* GEMM epilogue flags */
#define VQA_GEMM_RELU 1
#define VQA_GEMM_ATOMIC_ADD 4 /* set internally for split-K: C must be zero-filled by the caller */
#define VQA_GEMM_NO_CLUSTER 16 /* vqa_gemm_bf16s: never pair CTAs (B-tile multicast); for A/B measurements */
#define VQA_GEMM_ACCUMULATE 8 /* C += A.B^T with fp32 atomics into a caller-initialised C (plain epilogue, any split_k) */

/* graph-conv flags */
#define VQA_GC_RELU 1

const char* vqa_last_error(void);
/* SMs the persistent (one CTA per SM, statically partitioned) kernels may occupy: 8 .. 148, default 148; returns the previous
 * value.  Read at launch time (a captured graph keeps the grids it was captured with).  Used while a collective runs under
 * backward: a persistent grid that finds some SMs taken would otherwise finish a whole wave late. */
int vqa_set_sm_budget(int n_sms);
int vqa_abi_version(void);

/* C[M,N] = epi( sum_k A[m,k] * B[n,k] ) on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulator, TMA operands).
 * Operand storage: x_mn_major = 0 -> (MN rows, K contiguous) e.g. activations / nn.Linear weight (out,in);
 *                  x_mn_major = 1 -> (K rows, MN contiguous) e.g. dY and X in dW = dY^T X.
 * epi(v) : v += rowbcast[(m / group), n] ; v += bias[n] ; relu ; aux-mask: v = aux[m,n] > 0 ? v*aux_scale : 0.
 * Replaces: F.linear / weight-norm Linear of layers.py:185-190, the nk conv Linears layers.py:140-142 (one GEMM),
 * sparse_graph_model.py:154-157, and autograd's mm/addmm backward nodes for all of them.
 * split_k > 1 accumulates partial tiles with fp32 atomics into a caller-zeroed C (plain epilogue only).
 * tile_n in {0 (auto), 64, 128, 256}. */
int vqa_gemm_f32(const float* A, long long lda, int a_mn_major, const float* B, long long ldb, int b_mn_major,
                 float* C, long long ldc, int M, int N, int Kc, const float* bias, const float* rowbcast,
                 long long ldrb, int group, const float* aux, long long ldaux, float aux_scale, int flags,
                 int precision, int split_k, int tile_n, vqa_stream_t stream);

/* Split-bf16 operand planes: hi = bf16(x), lo = bf16(x - hi), each (rows, ldp) bf16 with ldp % 8 == 0.  lo may be
 * NULL (plain bf16).  Producers of large activations write the planes from their own epilogue instead. */
int vqa_split_bf16_f32(const float* x, long long ldx, void* hi, void* lo, long long ldp, long long rows, int cols,
                       vqa_stream_t stream);

/* Inverted dropout fused with the split: planes of dropout(x); same RNG family and step_ptr convention as
 * vqa_dropout_f32 (counter = (row * ceil(cols/4) + col/4, offset)).  sparse_graph_model.py:111 feeding the projections. */
int vqa_dropout_split_f32(const float* x, long long ldx, void* hi, void* lo, long long ldp, long long rows, int cols, float p,
                          unsigned long long seed, unsigned long long offset, const unsigned long long* step_ptr,
                          vqa_stream_t stream);

/* C[M,N] = epi( sum_k A[m,k] * B[n,k] ) with A = A_hi + A_lo, B = B_hi + B_lo given as bf16 planes (kind::f16
 * tcgen05.mma, fp32 TMEM accumulator).  passes = 3: lo*hi + hi*lo + hi*hi (fp32-grade, ~2^-17 per product);
 * passes = 1: hi planes only (bf16 mode).  Same operand-major convention, epilogue and call sites as vqa_gemm_f32;
 * the mask may also be given as the hi plane of a split tensor (aux_hi), and the result can be written as fp32 (C),
 * as split planes (C_hi, C_lo; C_lo may be NULL), or both.
 * tile_gate (optional, device, one int per 128-row tile of C): the CTAs of row tile m do nothing when
 * tile_gate[m] <= gate_t - used by the padded GRU recurrence (tile_gate[m] = longest question in the tile, gate_t = time
 * step), where those rows are never read (forward) or would only accumulate zeros (backward). */
int vqa_gemm_bf16s(const void* A_hi, const void* A_lo, long long lda, int a_mn_major, const void* B_hi,
                   const void* B_lo, long long ldb, int b_mn_major, float* C, long long ldc, void* C_hi, void* C_lo,
                   long long ldcs, int M, int N, int Kc, const float* bias, const float* rowbcast, long long ldrb,
                   int group, const float* aux, long long ldaux, const void* aux_hi, long long ldauxh,
                   float aux_scale, int flags, int passes, int split_k, int tile_n, const int* tile_gate, int gate_t,
                   vqa_stream_t stream);

/* y = x * keep / (1-p), keep ~ Bernoulli(1-p) from Philox4x32-10(seed; counter = (element/4, offset)).
 * step_ptr (optional, device): *step_ptr * 16 is added to offset at run time, so a CUDA-graph replay that bumps the
 * counter draws fresh masks.  Replaces nn.Dropout on the image tensor and the classifier hidden
 * (sparse_graph_model.py:111,156). */
int vqa_dropout_f32(const float* x, float* y, long long n, float p, unsigned long long seed,
                    unsigned long long offset, const unsigned long long* step_ptr, vqa_stream_t stream);

/* Old-style weight norm (dim=0): w[r,:] = v[r,:] * (g[r] / ||v[r,:]||).  torch._weight_norm at layers.py:171-172,
 * sparse_graph_model.py:88-89.  bwd: given dw returns dv, dg. */
int vqa_weight_norm_fwd_f32(const float* v, const float* g, float* w, int rows, int cols, vqa_stream_t stream);
/* The same, fused with the operand split: columns [c0, c1) of the effective weight written as (hi, lo) bf16 planes
 * (lo may be NULL), no fp32 intermediate.  cols % 4 == 0, c0 % 4 == 0, ldp % 8 == 0. */
int vqa_weight_norm_split_f32(const float* v, const float* g, int rows, int cols, int c0, int c1, void* hi, void* lo,
                              long long ldp, vqa_stream_t stream);
int vqa_weight_norm_bwd_f32(const float* dw, const float* v, const float* g, float* dv, float* dg, int rows,
                            int cols, vqa_stream_t stream);

/* out[c] = sum_r x[r,c]  (bias gradients; deterministic two-stage reduction, `scratch` >= 256*cols floats).
 * counters (optional): >= ceil(cols/256) ints that are ZERO on entry and are left zero - the last row block of every column
 * block then adds the partial sums itself (one launch instead of two, same summation order). */
int vqa_colsum_f32(const float* x, long long ldx, float* out, float* scratch, long long rows, int cols, int* counters,
                   vqa_stream_t stream);
/* out[s,c] = sum_{i<seg_len} x[s*seg_len+i, c]  (gradient of the per-image broadcast question term) */
int vqa_segment_sum_f32(const float* x, float* out, int segments, int seg_len, int cols, vqa_stream_t stream);

/* Fused graph-learner tail: per image A = h h^T (fp32 FMA), per-row top-nb selection and softmax.
 * h (B,K,C) -> adjacency (B,K,K), idx (B,K,nb) int32 in DESCENDING value order, alpha (B,K,nb).
 * Replaces layers.py:193-195 and sparse_graph_model.py:225-227 (topk sorted=False + K softmax launches). */
int vqa_adjacency_topk_fwd_f32(const float* h, float* adjacency, int* idx, float* alpha, int B, int K, int C,
                               int nb, vqa_stream_t stream);
/* Same selection/softmax from a GIVEN adjacency (used for the bit-exact index parity test and the layer API). */
int vqa_topk_softmax_f32(const float* adjacency, int* idx, float* alpha, int B, int K, int nb, vqa_stream_t stream);
/* Backward of the above: dalpha (B,K,nb) [+ optional dadj (B,K,K)] -> dh (B,K,C), already multiplied by the
 * ReLU mask (h > 0) of the layer that produced h. */
int vqa_adjacency_topk_bwd_f32(const float* h, const int* idx, const float* alpha, const float* dalpha,
                               const float* dadj, float* dh, int B, int K, int C, int nb, vqa_stream_t stream);

/* Fused graph convolution, project-first: Y = X W_all^T comes from vqa_gemm_f32; this kernel computes the
 * Gaussian patch weights over polar pseudo-coordinates of the box centres (boxes = xyxy, row stride ldbox floats),
 * normalises them over the kernel axis, gathers neighbours and aggregates:
 *   out[b,i, chunk k] = act( sum_m w[b,i,m,k] * alpha[b,i,m] * Y[b, idx[b,i,m], chunk k] ),  chunk = out_dim/nk.
 * gauss = {mean_rho[nk], precision_rho[nk], mean_theta[nk], precision_theta[nk]}.  alpha may be NULL (== 1).
 * Optional fused dropout (p > 0) after the ReLU.  Replaces sparse_graph_model.py:161-195,239-240,244-269 and
 * layers.py:100-144 (+ relu/dropout :137-138,148). */
int vqa_graphconv_fwd_f32(const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                          long long ldbox, const float* gauss, float* out, long long ldo, int B, int K, int nb,
                          int nk, int out_dim, int flags, float dropout_p, unsigned long long seed,
                          unsigned long long offset, const unsigned long long* step_ptr, vqa_stream_t stream);
/* Second layer with the pooling tail fused: relu, max over the K nodes (first index on ties), gate with relu(q):
 * pooled (B,out), argmax (B,out) int64, hq = relu(q) * pooled.  sparse_graph_model.py:146-151. */
int vqa_graphconv_pool_fwd_f32(const float* Y, long long ldy, const int* idx, const float* boxes, long long ldbox,
                               const float* gauss, const float* q, float* pooled, long long* argmax, float* hq,
                               int B, int K, int nb, int nk, int out_dim, vqa_stream_t stream);
/* Backward data path of the aggregate.  Upstream gradient is either dense dO (B,K,out) (already ReLU/dropout
 * masked) or, for the pooled layer, dpooled (B,out) + argmax (scatter by argmax is done on the fly).
 * Writes dY (B,K,out) (skipped when dY is NULL) and the per-edge, per-kernel dot products
 * P (B,K,nb,nk) = <dO[i,chunk k], Y[idx,chunk k]> (skipped when P is NULL; Y may then be NULL too). */
int vqa_graphconv_bwd_f32(const float* dO, long long lddo, const float* dpooled, const long long* argmax,
                          const float* Y, long long ldy, const int* idx, const float* alpha, const float* boxes,
                          long long ldbox, const float* gauss, float* dY, long long lddy, float* P, int B, int K,
                          int nb, int nk, int out_dim, vqa_stream_t stream);
/* Edge-level finish of the backward: from P computes dalpha (B,K,nb) (NULL when alpha is NULL) and the partial
 * sums of the four Gaussian parameter gradients, dgauss_partial (nblocks, 4*nk) with nblocks returned by
 * vqa_graphconv_edge_blocks(); reduce them with vqa_colsum_f32.  SURVEY.md 9.2. */
int vqa_graphconv_edge_blocks(int B, int K, int nb);
int vqa_graphconv_edge_bwd_f32(const float* P, const int* idx, const float* alpha, const float* boxes,
                               long long ldbox, const float* gauss, float* dalpha, float* dgauss_partial, int B,
                               int K, int nb, int nk, vqa_stream_t stream);

/* ---- tensor-core graph convolution on split-bf16 planes (graphconv_mma.cu).  Same maths and reference call sites as
 * vqa_graphconv_fwd_f32 / _pool_fwd_f32 / the dY part of _bwd_f32; Y and the results are (hi, lo) bf16 planes (lo may
 * be NULL: bf16 mode), so the projections before and after exchange planes with no fp32 round trip.  Requires
 * (out_dim / nk) % 128 == 0 and K <= 128; returns VQA_ERR_UNSUPPORTED otherwise (callers use the fp32 kernels).
 *
 * vqa_graphconv_edge_coef evaluates, once per layer and step, everything that depends only on the selected edges:
 * coef (B,K,nb,nk) = Gaussian weight (normalised over the kernel axis, layers.py:109-123) * alpha (1 if alpha is NULL)
 * and eoff (B,K,nb) = the packed positions of each edge inside the shared-memory coefficient matrices (forward |
 * transposed << 16).  Passing coef/eoff to the three aggregates selects the persistent streaming kernels; with
 * coef == NULL they evaluate the weights themselves (slower, same results). */
int vqa_graphconv_edge_coef(const int* idx, const float* alpha, const float* boxes, long long ldbox, const float* gauss,
                            float* coef, unsigned* eoff, int B, int K, int nb, int nk, vqa_stream_t stream);
int vqa_graphconv_mma_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* alpha,
                          const float* boxes, long long ldbox, const float* gauss, void* out_hi, void* out_lo, long long ldo,
                          int B, int K, int nb, int nk, int out_dim, int flags, float dropout_p, unsigned long long seed,
                          unsigned long long offset, const unsigned long long* step_ptr, const float* coef,
                          const unsigned* eoff, vqa_stream_t stream);
int vqa_graphconv_mma_pool_fwd(const void* Y_hi, const void* Y_lo, long long ldy, const int* idx, const float* boxes,
                               long long ldbox, const float* gauss, const float* q, float* pooled, long long* argmax,
                               float* hq, int B, int K, int nb, int nk, int out_dim, const float* coef,
                               const unsigned* eoff, vqa_stream_t stream);
/* dY[b,j, chunk k] = sum_{(i,m): idx[i,m]=j} w[i,m,k] alpha[i,m] dO[b,i, chunk k]  (the transposed aggregate) */
int vqa_graphconv_mma_bwd_data(const void* dO_hi, const void* dO_lo, long long lddo, const int* idx, const float* alpha,
                               const float* boxes, long long ldbox, const float* gauss, void* dY_hi, void* dY_lo,
                               long long lddy, int B, int K, int nb, int nk, int out_dim, const float* coef,
                               const unsigned* eoff, vqa_stream_t stream);

/* Data path of the backward of the POOLED layer (max over nodes, sparse_graph_model.py:150, then the transposed
 * aggregate): dY[b, idx[b,a,m], c] = coef[b,a,m,k(c)] * dpooled[b,c] with a = argmax[b,c], zero elsewhere, written as
 * (hi, lo) planes (lo may be NULL).  coef from vqa_graphconv_edge_coef of that layer (alpha = NULL). */
int vqa_graphconv_pool_bwd_data(const float* dpooled, const long long* argmax, const int* idx, const float* coef, void* dY_hi,
                                void* dY_lo, long long lddy, int B, int K, int nb, int nk, int out_dim, vqa_stream_t stream);

/* Edge part of the backward on the tensor cores: P[i,m,k] = <dO[i,chunk k], Y[idx[i,m],chunk k]> per image as
 * dO_k Y_k^T, then dalpha (B,K,nb) (NULL when alpha is NULL) and the per-image partial sums of the Gaussian-parameter
 * gradients dgauss_partial (B, 4*nk) (reduce over B with vqa_colsum_f32).  Upstream: dO planes, or (dpooled, argmax)
 * with dO_hi == NULL for the pooled layer.  p_scratch: (B, nk, K*nb) floats, required: the selected products, written by the
 * streaming kernel (persistent, one CTA per SM over B*nk units) and read by the edge-finish kernel of the same call.
 * Requires (out_dim / nk) % 64 == 0.  SURVEY.md 9.2. */
int vqa_graphconv_mma_bwd_edges(const void* dO_hi, const void* dO_lo, long long lddo, const float* dpooled,
                                const long long* argmax, const void* Y_hi, const void* Y_lo, long long ldy, const int* idx,
                                const float* alpha, const float* boxes, long long ldbox, const float* gauss, float* dalpha,
                                float* dgauss_partial, float* p_scratch, int B, int K, int nb, int nk, int out_dim,
                                vqa_stream_t stream);

/* Gaussian patch weights for explicit pseudo-coordinates (n,2) -> (n,nk): NeighbourhoodGraphConvolution.
 * get_gaussian_weights, layers.py:100-125 (layer-level API). */
/* The patch operator of the layer-level API on MATERIALISED neighbourhoods, layers.py:127-137 (torch.bmm(weights^T, neighbourhood)):
 * Z[n,k,:] = sum_m w[n,m,k] X[n,m,:]  with X (n, nb, F), w (n, nb, nk) -> Z (n, nk, F), all contiguous fp32.
 * Backward: dX[n,m,:] = sum_k w[n,m,k] dZ[n,k,:] (optional), dw[n,m,k] = <X[n,m,:], dZ[n,k,:]> (optional). */
int vqa_patch_operator_fwd_f32(const float* X, const float* w, float* Z, long long n, int nb, int nk, int F, vqa_stream_t stream);
int vqa_patch_operator_bwd_f32(const float* X, const float* w, const float* dZ, float* dX, float* dw, long long n, int nb, int nk,
                               int F, vqa_stream_t stream);
int vqa_gaussian_weights_f32(const float* pseudo, const float* gauss, float* w, long long n, int nk,
                             vqa_stream_t stream);

/* ---- question encoder (sparse_graph_model.py:117-121: nn.Embedding + pack_padded_sequence + nn.GRU, final state) ----
 * The recurrence runs over padded time-major steps; a sequence past its length keeps its state, which equals the
 * packed-sequence result.  The matrix products (x W_ih^T for all steps, h W_hh^T per step, and the backward
 * products) are vqa_gemm_bf16s calls; these entry points are the gather/scatter and the pointwise cell. */
/* E[(t*B + b), :] = W[question[b,t], :] written as split planes (rows time-major), t < T.  A token outside [0, vocab) - for
 * which nn.Embedding raises - sets bit 0 of *err (device int, optional) and reads row 0; the host raises at its next sync point. */
int vqa_embed_gather_split(const long long* question, long long ldq, const float* W, long long vocab, int emb, void* hi,
                           void* lo, long long ldp, int B, int T, int* err, vqa_stream_t stream);
/* dW[question[b,t], :] += dE[(t*B + b), :] for t < len[b]  (fp32 atomics; dW pre-zeroed / accumulated by the caller). */
int vqa_embed_scatter_add_f32(const float* dE, long long ldd, const long long* question, long long ldq, const int* len,
                              float* dW, long long vocab, int emb, int B, int T, vqa_stream_t stream);
/* One GRU step, torch gate order (r,z,n).  gi = x_t W_ih^T + b_ih (B,3H); gh = h_{t-1} W_hh^T (B,3H) WITHOUT bias
 * (b_hh is added by this kernel) or NULL at t = 0; h_t = t < len[b] ? cell : h_{t-1}, written as fp32 and split planes; gates (B,4H)
 * = r | z | n | gh_n saved for backward. */
int vqa_gru_cell_fwd_f32(const float* gi, long long ldgi, const float* gh, const float* b_hh, const float* h_prev,
                         const int* len, int t, float* h_out, void* h_hi, void* h_lo, long long ldp, float* gates, int B,
                         int H, vqa_stream_t stream);
/* The same step as ONE kernel: gh = h_{t-1} W_hh^T on the tensor cores (3-pass split-bf16) with the cell evaluated straight
 * out of TMEM - no (B,3H) gh round trip, no separate cell launch.  Whh planes (3H,H), gi (B,3H; includes b_ih) and b_hh (3H)
 * are given in UNIT-BLOCK order: block u = rows/columns [r | z | n] of hidden units 32u..32u+31 (row u*96 + g*32 + i <-
 * original row g*H + u*32 + i); h_t, its planes and the gates are written in the original layouts.  tile_gate (optional, one
 * int per 128 batch rows): row tiles with tile_gate[m] <= t are skipped, their outputs are NOT written.  H % 32 == 0. */
int vqa_gru_step_fused(const void* hprev_hi, const void* hprev_lo, long long ldh, const void* Whh_hi, const void* Whh_lo,
                       long long ldw, const float* gi, long long ldgi, const float* b_hh, const float* h_prev, const int* len,
                       int t, float* h_out, void* hout_hi, void* hout_lo, long long ldp, float* gates, const int* tile_gate,
                       int B, int H, vqa_stream_t stream);
/* All T steps in ONE cooperative launch (grid barrier between steps instead of kernel boundaries; needs H/32 * ceil(B/128)
 * <= 148 CTAs, else VQA_ERR_UNSUPPORTED).  H planes: ((T+1)*B, H) with rows [0,B) = h_{-1} = 0 supplied by the caller, rows
 * (t+1)*B.. receive h_t; Hall: (T+1, B, H) fp32 likewise; gi: (T*B, 3H) and Whh / b_hh in unit-block order; gates: (T, B, 4H);
 * counter: one device word used by the grid barrier (zeroed by this call). */
int vqa_gru_seq_fused(const void* H_hi, const void* H_lo, long long ldh, const void* Whh_hi, const void* Whh_lo, long long ldw,
                      const float* gi, long long ldgi, const float* b_hh, float* Hall, const int* len, float* gates,
                      const int* tile_gate, unsigned* counter, int T, int B, int H, vqa_stream_t stream);
/* Backward of one step: dh (B,H) -> dgi, dgh (B,3H; fp32 and split planes, ld = ldp) and dh_part (B,H), the direct
 * part of dL/dh_{t-1}; the caller adds dgh W_hh. */
int vqa_gru_cell_bwd_f32(const float* dh, const float* gates, const float* h_prev, const int* len, int t, float* dgi,
                         float* dgh, void* dgi_hi, void* dgi_lo, void* dgh_hi, void* dgh_lo, long long ldp, float* dh_part,
                         int B, int H, vqa_stream_t stream);

/* Gate/pool backward: dpooled = pooled > 0 ? dhq * relu(q) : 0 ;  dq = q > 0 ? dhq * pooled : 0.
 * sparse_graph_model.py:150-151 (autograd of max + relu*mul). */
int vqa_gate_bwd_f32(const float* dhq, const float* q, const float* pooled, float* dpooled, float* dq, long long n,
                     vqa_stream_t stream);

/* ---- tail of the training step (the reference drivers' criterion and optimiser) ----
 * nn.MultiLabelSoftMarginLoss(logits, target) of run.py:382,431 (run_imageclef.py / run_mimic.py alike) over n = B*A contiguous
 * elements: *loss = scale * sum_i -( y_i logsigmoid(x_i) + (1 - y_i) logsigmoid(-x_i) ), scale = 1/(B*A) for reduction='mean'.
 * partial: >= vqa_mlsm_loss_blocks(n) floats of scratch; counter: one int that is ZERO on entry and is left zero (the last
 * block adds the partial sums in block order, so the result does not depend on scheduling). */
int vqa_mlsm_loss_blocks(long long n);
int vqa_mlsm_loss_fwd_f32(const float* logits, const float* target, long long n, float scale, float* partial, int* counter,
                          float* loss, vqa_stream_t stream);
/* dlogits_i = (sigmoid(x_i) - y_i) * scale * (*grad_out)   (grad_out: device scalar, NULL == 1). */
int vqa_mlsm_loss_bwd_f32(const float* logits, const float* target, const float* grad_out, float* dlogits, long long n,
                          float scale, vqa_stream_t stream);

/* torch.optim.Adam (run.py:392,435; amsgrad=False, maximize=False) for every parameter tensor in ONE launch.
 * chunks: device table of nchunks x {int64 address of the parameter elements, int64 element offset into grad / exp_avg /
 * exp_avg_sq, int64 count}; grad is the flat gradient buffer (vqa_b200.ddp.GradReducer), exp_avg / exp_avg_sq have its layout.
 * g = grad * grad_scale (+ weight_decay * p); grad_scale = 1/world folds the data-parallel average into this pass.
 * lr: device float (so a CUDA-graph replay sees a scheduler's new rate).  state: two device ints, {steps taken, 0}; the
 * launch applies step state[0]+1 (bias corrections in double) and its last block stores the new count. */
/* Data-parallel Adam over NVLink peer memory (one process per GPU on one node; no counterpart in the reference, which is single
 * process - SURVEY.md 8e): the tail of a pushed reduce-scatter, Adam on the rank's own elements and the all-gather of the updated
 * parameters in ONE kernel.  Ownership is interleaved: chunk c of 2^chunk_log2 floats of the flat index space belongs to rank
 * c % world; the rank owns n_own elements (whole chunks), in its own order e -> flat index ((e / chunk) * world + rank) * chunk +
 * e % chunk.  grad: its own flat gradient buffer; recv: its receive buffer, `world` strides of n_own floats, stride q holding rank
 * q's gradients for the rank's elements in that order (pushed there by rank q before the barrier, e.g. with vqa_memcpy2d_async;
 * stride `rank` is unused); param_addrs: HOST array of `world` addresses - every rank's flat parameter buffer (layout of the
 * gradients) as mapped into THIS process; exp_avg / exp_avg_sq local, same layout.  g = grad_scale * sum over ranks in rank order,
 * Adam, P_q[i] = p for every q.  lr / state as in vqa_adam_flat_f32.  Bracket with vqa_p2p_barrier on every rank.
 * mc_param (optional): the MULTICAST address of the parameter buffers (NVLS); when given, the all-gather is one multimem.st per
 * element instead of `world` stores. */
int vqa_adam_flat_p2p(const float* grad, const float* recv, long long n_own, int chunk_log2, const long long* param_addrs, float* mc_param,
                      float* exp_avg, float* exp_avg_sq, int rank, int world, const float* lr, float beta1, float beta2, float eps, float weight_decay,
                      float grad_scale, int* state, vqa_stream_t stream);
/* The same exchange through NVSwitch multicast objects (NVLS): mc_grad / mc_param are the MULTICAST addresses of the ranks' flat gradient
 * / parameter buffers (symmetric allocations bound to one multicast object), param / exp_avg / exp_avg_sq the local ones.  The rank
 * owns the contiguous flat slice [lo, hi): one multimem.ld_reduce per element returns the sum over all ranks (added in the switch),
 * one multimem.st delivers the updated parameter to every rank.  Bracket with vqa_p2p_barrier on every rank. */
int vqa_adam_flat_mc(const float* mc_grad, float* mc_param, const float* param, float* exp_avg, float* exp_avg_sq, long long lo,
                     long long hi, const float* lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, int* state,
                     vqa_stream_t stream);
/* cudaMemcpy2DAsync, device to device (pitches / width in bytes): the strided pushes of the interleaved layout above as ONE copy per
 * (bucket, peer) on the copy engines; capturable into a CUDA graph as a memcpy node. */
int vqa_memcpy2d_async(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height,
                       vqa_stream_t stream);
/* Flag barrier between the ranks of one node, stream ordered: flag_addrs = HOST array of `world` addresses of the ranks' flag arrays
 * (`world` ints each, zero-initialised, peer mapped); epoch = local device int, zero-initialised, incremented by every barrier.
 * A rank that does not arrive within ~4 s of SM clock makes the launch trap (error to the caller) instead of hanging the GPU. */
int vqa_p2p_barrier(const long long* flag_addrs, int rank, int world, int* epoch, vqa_stream_t stream);
int vqa_adam_flat_f32(const long long* chunks, int nchunks, const float* grad, float* exp_avg, float* exp_avg_sq,
                      const float* lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, int* state,
                      vqa_stream_t stream);

/* ---- batch assembly on the device (input side of the path; replaces the per-question host work of torch_dataset.py:105-164
 * and most of the H2D copy of utils.py:22-31).  The feature table lives in HBM (or in pinned host memory for the rows of one
 * batch): features (n_rows, K, D) fp32 or bf16 (features_bf16 != 0), boxes (n_rows, K, 4) fp32 xyxy already divided by the
 * image size (torch_dataset.py:148-154).  image[b] = [ features[rows[b]] | boxes[rows[b]] ] as (B, K, D+4) fp32, the layout
 * Model.forward takes.  A row index outside [0, n_rows) writes zeros and sets *err_flag = 1 (never reads outside the table). */
int vqa_gather_image_f32(const void* features, int features_bf16, const float* boxes, const long long* rows, long long n_rows,
                         float* image, int B, int K, int D, int* err_flag, vqa_stream_t stream);
/* Dense (B, A) rows from CSR triplets: out[b, ids[e]] = vals[e] for e in [ptr[b], ptr[b+1]), zero elsewhere; entries are
 * applied in order (a repeated id keeps the last value, as the assignment loops of torch_dataset.py:117-130 do).  An id
 * outside [0, A) is skipped and sets *err_flag = 2. */
int vqa_scatter_targets_f32(const long long* ptr, const int* ids, const float* vals, float* out, int B, int A, int* err_flag,
                            vqa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_B200_H_ */
