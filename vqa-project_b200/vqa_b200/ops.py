"""Autograd operators of the conditioned-graph VQA hot path, built on the sm_100a kernels.

``ConditionedGraphFn`` is the fused entry ``Model.forward`` uses: everything between the question
encoding and the logits (reference ``sparse_graph_model.py:106-157`` minus the embedding/GRU) as one
autograd node whose forward and backward are sequences of hand-written CUDA kernels -- tcgen05 GEMMs
for the dense projections, the fused adjacency/top-k/softmax kernel, the fused graph-convolution
kernels and their backward counterparts (SURVEY.md section 9 is the maths).  Nothing is materialised that
the reference materialises only because of its operator granularity: no (B,K,F+H) concat, no
(B,K,nb,F) gathers, no dense (B,K,K,2) pseudo-coordinates, no per-kernel slices.

The smaller Functions (``LinearFn``, ``AdjacencyFn``, ``GaussianWeightsFn``) back the layer-level
module API (``layers.GraphLearner`` / ``layers.NeighbourhoodGraphConvolution``).
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import torch

from . import kernels as kn
from ._cabi import PREC_TF32, PREC_TF32X3, PREC_TF32X3_HP

# GEMM precision policy (fp32 tensors stay fp32 at the module boundary; see csrc/gemm_bf16s.cu, csrc/gemm_tcgen05.cu).
#   "fp32" (default): every dense product runs on the split-bf16 tcgen05 GEMM with 3 passes (operands carried as
#            hi/lo bf16 planes, ~2^-17 per product, measured 8e-6 at K=2052).
#   "fp32_strict": as "fp32", but the graph-learner FORWARD chain -- whose output is exponentiated by the neighbourhood
#            softmax (an absolute error e in the adjacency is a relative error e in alpha) -- uses the chunk-promoted
#            3xTF32 kernel (1.2e-6, cuBLAS-fp32 grade).  Measured on the parity workloads the two modes agree to within
#            2e-5 on every gradient; "fp32" is ~0.3 ms/step faster at B=512.
#   "bf16":  one pass on the hi planes (plain bf16 tensor-core GEMM, fp32 accumulate); the graph-learner forward chain
#            stays 3-pass so the neighbourhood selection does not drift.  Stated tolerance (max-norm relative): logits 2e-2,
#            weight gradients 1e-1, Gaussian-kernel parameter gradients 2e-1.
_PRECISION_NAME = "fp32"
_PASSES = 3
_GL_STRICT = False


def set_precision(name: str) -> None:
    global _PRECISION_NAME, _PASSES, _GL_STRICT
    if name in ("fp32", "fp32_strict"):
        _PRECISION_NAME, _PASSES, _GL_STRICT = name, 3, name == "fp32_strict"
    elif name == "bf16":
        _PRECISION_NAME, _PASSES, _GL_STRICT = "bf16", 1, False
    else:
        raise ValueError(f"unknown precision {name!r}: use 'fp32', 'fp32_strict' or 'bf16'")


def get_precision() -> str:
    return _PRECISION_NAME


def _gemm(a, b, **kw):
    """fp32-in/fp32-out product on the 3xTF32 kernel (layer-level API)."""
    return kn.gemm(a, b, precision=PREC_TF32X3, **kw)


def _gemm_gl(a, b, **kw):
    """GEMMs on the path image/question -> adjacency (feeds exp): chunk-promoted accumulation."""
    return kn.gemm(a, b, precision=PREC_TF32X3_HP, **kw)


def _split(x):
    return kn.split(x, with_lo=_PASSES == 3)


def _gemm_s(a, b, **kw):
    return kn.gemm_s(a, b, passes=_PASSES, **kw)


def _plan(M: int, N: int, K: int, plain: bool = True) -> dict:
    """Tile width / split-K for the products the library's own heuristic serves badly (measured sweep, tools/gemm_small_sweep.py):
    few row tiles (M = batch) or a long contraction with a narrow output.  128-wide tiles keep ~100 CTAs busy without the L2
    re-read cost of 64-wide ones; plain-epilogue products additionally split the contraction to fill / balance the 148 SMs."""
    mt, kb = (M + 127) // 128, (K + 63) // 64
    if mt * ((N + 255) // 256) >= 120 and not (plain and kb >= 24 and mt * ((N + 127) // 128) < 296):
        return {}                                             # plenty of 256-wide tiles: leave it to the library
    tiles = mt * ((N + 127) // 128)
    out = {"tile_n": 128} if tiles >= 48 or N > 128 else {}
    if plain and kb >= 16:
        if tiles <= 148:
            sk = max(1, min(148 // tiles, kb // 8))
        else:
            sk = 3 if tiles < 296 and kb >= 24 else 1         # 1-2 waves of long tiles: cut them into thirds to balance
        if sk > 1:
            out["split_k"] = sk
    return out


def _split_for(out_rows: int, out_cols: int, contraction: int) -> int:
    """Split-K factor for dW = dY^T X products whose output has too few tiles to fill 148 SMs."""
    tiles = ((out_rows + 127) // 128) * ((out_cols + 255) // 256)
    kblocks = (contraction + 63) // 64
    s = max(1, min(148 // max(tiles, 1), kblocks // 8))
    return s


_GRAD_SINK = None   # callable(parameter tensor) -> fresh view of a flat gradient buffer, or None (ddp.GradReducer.sink)


_GRAD_READY = None  # callable(parameter tensor): "its gradient is final in the sink view" (ddp.GradReducer.mark_ready)


def set_grad_sink(fn=None, ready=None) -> None:
    """Install (or remove) the gradient sink: backward then lets the LAST kernel of every parameter gradient write straight
    into the view the sink returns and hands that view to autograd, which adopts it as ``.grad`` without an add or a copy
    (the reference's autograd allocates, zero-fills and accumulates: SURVEY.md 2a k20)."""
    global _GRAD_SINK, _GRAD_READY
    _GRAD_SINK = fn
    _GRAD_READY = ready if fn is not None else None


def _ready(*params):
    """Tell the reducer that these parameters' gradients are complete in their sink views (their bucket may start its
    all-reduce now, under the rest of backward)."""
    # (whether the reducer wants these calls is its decision: ddp.GradReducer(early=...) installs the callback or not)
    if _GRAD_READY is not None:
        for t in params:
            if t is not None:
                _GRAD_READY(t)


def _sink(t):
    return None if _GRAD_SINK is None or t is None else _GRAD_SINK(t)


def _into(sink_view, value):
    """Tiny gradients assembled elsewhere (Gaussian parameters): copy into the sink view when there is one."""
    if sink_view is None:
        return value
    sink_view.copy_(value.view(sink_view.shape))
    return sink_view


_GRAPH_RNG = None   # (seed, device int64 step counter) while a CUDA-graph-safe RNG is installed (engine.TrainStep)
_GRAPH_RNG_SITE = 0


def set_graph_rng(seed=None, step: "torch.Tensor | None" = None) -> None:
    """Install (or remove, with no arguments) a graph-safe dropout RNG: masks are a function of (seed, call site,
    *step) with ``step`` read on the device at run time, so replays of a captured step differ once ``step`` is bumped."""
    global _GRAPH_RNG, _GRAPH_RNG_SITE
    _GRAPH_RNG = None if seed is None else (int(seed) & 0xFFFFFFFFFFFFFFFF, step)
    _GRAPH_RNG_SITE = 0


def next_philox(device: torch.device):
    """(seed, offset, step tensor or None) for one fused-dropout call.  Eager mode: drawn from (and advancing) torch's
    CUDA generator state, so ``torch.manual_seed`` controls the masks exactly as it does for nn.Dropout.  Graph mode
    (``set_graph_rng``): per-call-site offsets plus the device-side step counter."""
    global _GRAPH_RNG_SITE
    if _GRAPH_RNG is not None:
        _GRAPH_RNG_SITE = _GRAPH_RNG_SITE % 15 + 1
        return _GRAPH_RNG[0], _GRAPH_RNG_SITE, _GRAPH_RNG[1]
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    off = gen.get_offset()
    gen.set_offset(off + 4)
    return gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, off // 4 + 1, None


def pack_gauss(mean_rho, precision_rho, mean_theta, precision_theta) -> torch.Tensor:
    return torch.cat((mean_rho.reshape(-1), precision_rho.reshape(-1), mean_theta.reshape(-1), precision_theta.reshape(-1))).contiguous()


def flat_weight(ws: Sequence[torch.Tensor]) -> torch.Tensor:
    """The nk per-kernel conv weights (D,in) as ONE (nk*D, in) matrix; zero-copy when they already are
    consecutive views of one buffer (layers.NeighbourhoodGraphConvolution arranges that)."""
    w0 = ws[0]
    d, fin = w0.shape
    step = d * fin * w0.element_size()
    if all(w.is_contiguous() and w.data_ptr() == w0.data_ptr() + i * step for i, w in enumerate(ws)):
        try:
            return w0.detach().as_strided((len(ws) * d, fin), (fin, 1))
        except RuntimeError:
            pass
    return torch.cat([w.detach() for w in ws], dim=0)


def _conv_sink(ws, zero: bool):
    """One (nk*D, in) destination covering the gradient views of the nk per-kernel conv weights, when the sink lays them out
    consecutively (it does: registration order); None otherwise.  ``zero``: the product accumulates split-K partials."""
    vs = [_sink(w) for w in ws]
    if any(v is None for v in vs):
        return None
    v0 = vs[0]
    d, fin = v0.shape
    step = d * fin * v0.element_size()
    if not all(v.is_contiguous() and v.data_ptr() == v0.data_ptr() + i * step for i, v in enumerate(vs)):
        return None
    out = v0.as_strided((len(vs) * d, fin), (fin, 1))
    if zero:
        out.zero_()
    return out


class ConditionedGraphFn(torch.autograd.Function):
    """(image, qenc, parameters) -> (logits, adjacency, h_max_indices)."""

    @staticmethod
    def forward(ctx, cfg, image, qenc, v1, g1, b1, v2, g2, b2, mr1, pr1, mt1, pt1, mr2, pr2, mt2, pt2,
                vo1, go1, bo1, vo2, go2, bo2, *conv_ws):
        nk, nb, p_drop, training = cfg["n_kernels"], cfg["neighbourhood_size"], float(cfg["dropout"]), bool(cfg["training"])
        B, K, F = image.shape
        H = qenc.shape[1]
        dev = image.device
        if image.requires_grad:
            raise RuntimeError("ConditionedGraphFn: image.requires_grad is set, but no gradient with respect to the image features is "
                               "computed (the reference's data layer never asks for one: torch_dataset.py:157-164, utils.py:22-31)")
        image = image.contiguous()
        qenc = qenc.contiguous()
        drop = training and p_drop > 0.0
        scale = 1.0 / (1.0 - p_drop) if drop else 1.0

        # dropout on the WHOLE image tensor incl. box columns (sparse_graph_model.py:111), written directly as split planes;
        # box centres are taken from the un-dropped image inside the graph-conv kernels (:106-108 precede :111)
        img2 = image.view(B * K, F)
        if drop:
            seed, off, step = next_philox(dev)
            Xs = kn.dropout_split(img2, p_drop, seed, off, step)          # lo plane always: the graph-learner chain is 3-pass
        else:
            Xs = kn.split(img2)

        # weight-norm effective weights (layers.py:171-172, sparse_graph_model.py:88-89), written directly as operand planes
        Wc1 = flat_weight(conv_ws[:nk])
        Wc2 = flat_weight(conv_ws[nk:])
        lo3 = _PASSES == 3
        if _GL_STRICT:
            W1 = kn.weight_norm_fwd(v1, g1)
            W2 = kn.weight_norm_fwd(v2, g2)
            W1xs, W1qs, W2s = kn.split(W1[:, :F]), kn.split(W1[:, F:].contiguous()), kn.split(W2)
        else:
            W1xs, W1qs = kn.weight_norm_split(v1, g1, 0, F), kn.weight_norm_split(v1, g1, F, F + H)
            W2s = kn.weight_norm_split(v2, g2)
        Wo1s, Wo2s = kn.weight_norm_split(vo1, go1, with_lo=lo3), kn.weight_norm_split(vo2, go2, with_lo=lo3)
        Wc1s, Wc2s = _split(Wc1), _split(Wc2)
        qs = kn.split(qenc)

        # graph learner: [X || q] W1^T = X W1[:, :F]^T + (q W1[:, F:]^T) broadcast over the K nodes  (no concat/repeat)
        h1s = kn.empty_split(B * K, v1.shape[0], dev, True)
        if _GL_STRICT:
            X2 = Xs.float()
            qt = _gemm_gl(qenc, W1[:, F:])
            h1 = _gemm_gl(X2, W1[:, :F], bias=b1, rowbcast=qt, group=K, relu=True)
            h2 = _gemm_gl(h1, W2, bias=b2, relu=True)
            h1s = kn.split(h1)
            del X2, h1
        else:
            qt = kn.gemm_s(qs, W1qs, passes=3)
            kn.gemm_s(Xs, W1xs, bias=b1, rowbcast=qt, group=K, relu=True, out_split=h1s, want_f32=False, passes=3)
            h2 = kn.gemm_s(h1s, W2s, bias=b2, relu=True, passes=3)
        C = h2.shape[1]
        adj, idx, alpha = kn.adjacency_topk_fwd(h2.view(B, K, C), nb)

        # graph convolution 1 (project first, then fused Gaussian-weight/gather/aggregate + ReLU + dropout)
        gs1 = pack_gauss(mr1, pr1, mt1, pt1)
        gs2 = pack_gauss(mr2, pr2, mt2, pt2)
        mma1 = kn.mma_eligible(K, Wc1s.rows, nk)      # tensor-core aggregate needs (out/nk) % 128 == 0; else CUDA-core kernels
        mma2 = kn.mma_eligible(K, Wc2s.rows, nk)
        with_lo = _PASSES == 3
        gseed, goff, gstep = next_philox(dev) if drop else (0, 0, None)
        ec1 = ec2 = None
        if mma1:
            Y1 = _gemm_s(Xs, Wc1s, out_split=kn.empty_split(B * K, Wc1s.rows, dev, with_lo), want_f32=False)   # planes only
            ec1 = kn.graphconv_edge_coef(idx, alpha, image, gs1, B, K)      # edge coefficients: once per layer, reused by backward
            G1s = kn.graphconv_fwd_s(Y1, idx, alpha, image, gs1, B, K, relu=True, dropout_p=p_drop if drop else 0.0,
                                     seed=gseed, offset=goff, step=gstep, ec=ec1)
        else:
            Y1 = _gemm_s(Xs, Wc1s)
            G1 = kn.graphconv_fwd(Y1, idx, alpha, image, gs1, B, K, relu=True, dropout_p=p_drop if drop else 0.0,
                                  seed=gseed, offset=goff, step=gstep)
            G1s = _split(G1)
            del G1
        # graph convolution 2 with max-pool over nodes + question gate fused (sparse_graph_model.py:146-151)
        if mma2:
            Y2 = _gemm_s(G1s, Wc2s, out_split=kn.empty_split(B * K, Wc2s.rows, dev, with_lo), want_f32=False)
            ec2 = kn.graphconv_edge_coef(idx, None, image, gs2, B, K)
            pooled, argmax, hq = kn.graphconv_pool_fwd_s(Y2, idx, image, gs2, qenc, B, K, ec=ec2)
        else:
            Y2 = _gemm_s(G1s, Wc2s)
            pooled, argmax, hq = kn.graphconv_pool_fwd(Y2, idx, image, gs2, qenc, B, K)

        # classifier
        hqs = _split(hq)
        o1 = _gemm_s(hqs, Wo1s, bias=bo1, relu=True, **_plan(B, Wo1s.rows, H, plain=False))
        if drop:
            seed, off, step = next_philox(dev)
            o1 = kn.dropout(o1, p_drop, seed, off, step)
        o1s = _split(o1)
        logits = _gemm_s(o1s, Wo2s, bias=bo2, **_plan(B, Wo2s.rows, Wo2s.cols, plain=False))

        ctx.cfg = dict(B=B, K=K, F=F, H=H, nk=nk, nb=nb, scale=scale, mma1=mma1, mma2=mma2)
        ctx.prm = (b1, b2, bo1, bo2, (mr1, pr1, mt1, pt1), (mr2, pr2, mt2, pt2), conv_ws)   # identities for the gradient sink
        ctx.ec1, ctx.ec2 = ec1, ec2
        ctx.splits = (Xs, qs, h1s, G1s, hqs, o1s, W1qs, W2s, Wo1s, Wo2s, Wc1s, Wc2s, Y1, Y2)   # Y1/Y2: SplitT on the tensor-core path, fp32 else
        ctx.save_for_backward(image, qenc, v1, g1, v2, g2, vo1, go1, vo2, go2, gs1, gs2, h2, idx, alpha, pooled, argmax)
        ctx.mark_non_differentiable(argmax)
        return logits, adj, argmax

    @staticmethod
    def backward(ctx, dlogits, dadj, _dargmax):
        (image, qenc, v1, g1, v2, g2, vo1, go1, vo2, go2, gs1, gs2, h2, idx, alpha, pooled, argmax) = ctx.saved_tensors
        Xs, qs, h1s, G1s, hqs, o1s, W1qs, W2s, Wo1s, Wo2s, Wc1s, Wc2s, Y1, Y2 = ctx.splits
        c = ctx.cfg
        B, K, F, H, nk, scale = c["B"], c["K"], c["F"], c["H"], c["nk"], c["scale"]
        M = B * K
        dev = image.device
        with_lo = _PASSES == 3
        dlogits = dlogits.contiguous()

        # classifier
        dls = _split(dlogits)
        b1_, b2_, bo1_, bo2_, g1p, g2p, conv_ws = ctx.prm
        dbo2 = kn.colsum(dlogits, out=_sink(bo2_))
        dWo2 = _gemm_s(dls, o1s, a_mn=True, b_mn=True)
        dvo2, dgo2 = kn.weight_norm_bwd(dWo2, vo2, go2, out=(_sink(vo2), _sink(go2)))
        _ready(bo2_, vo2, go2)                              # the 36 MB out_2 bucket starts its all-reduce under the rest of backward
        do1s = kn.empty_split(o1s.rows, o1s.cols, dev, with_lo)
        do1 = _gemm_s(dls, Wo2s, b_mn=True, aux=o1s, aux_scale=scale, out_split=do1s,   # ReLU + dropout mask from the stored output
                      **_plan(dls.rows, Wo2s.cols, Wo2s.rows, plain=False))
        dbo1 = kn.colsum(do1, out=_sink(bo1_))
        dWo1 = _gemm_s(do1s, hqs, a_mn=True, b_mn=True)
        dvo1, dgo1 = kn.weight_norm_bwd(dWo1, vo1, go1, out=(_sink(vo1), _sink(go1)))
        _ready(bo1_, vo1, go1)
        dhq = _gemm_s(do1s, Wo1s, b_mn=True, **_plan(do1s.rows, Wo1s.cols, Wo1s.rows))
        dpooled, dq = kn.gate_bwd(dhq, qenc, pooled)

        # graph convolution 2: max-pool scatter by argmax is done inside the kernel
        if c["mma2"]:
            _, dgs2 = kn.graphconv_bwd_edges_s(Y2, idx, None, image, gs2, B, K, dpooled=dpooled, argmax=argmax)
            dY2s = kn.graphconv_pool_bwd_data_s(dpooled, argmax, idx, ctx.ec2, B, K, Y2.cols, with_lo)   # scatter, no contraction
        else:
            dY2, _, dgs2 = kn.graphconv_bwd(Y2, idx, None, image, gs2, B, K, dpooled=dpooled, argmax=argmax)
            dY2s = _split(dY2)
        sk2 = _split_for(Wc2s.rows, Wc2s.cols, M)
        out_c2 = _conv_sink(conv_ws[nk:], sk2 > 1)
        dWc2 = _gemm_s(dY2s, G1s, a_mn=True, b_mn=True, split_k=sk2, out=out_c2)
        gsh = (nk, 1)
        gauss2 = [_into(_sink(prm), dgs2[i * nk:(i + 1) * nk].view(gsh)) for i, prm in enumerate(g2p)]
        _ready(*(conv_ws[nk:] if out_c2 is not None else ()), *g2p)
        # graph convolution 1
        if c["mma1"]:
            dG1s = _gemm_s(dY2s, Wc2s, b_mn=True, aux=G1s, aux_scale=scale, out_split=kn.empty_split(M, G1s.cols, dev, with_lo), want_f32=False)
            dY1s = kn.graphconv_bwd_data_s(dG1s, idx, alpha, image, gs1, B, K, ec=ctx.ec1)   # dY = M^T dO on the tensor cores
            dalpha, dgs1 = kn.graphconv_bwd_edges_s(Y1, idx, alpha, image, gs1, B, K, dOs=dG1s)
        else:
            dG1 = _gemm_s(dY2s, Wc2s, b_mn=True, aux=G1s, aux_scale=scale)
            dY1, dalpha, dgs1 = kn.graphconv_bwd(Y1, idx, alpha, image, gs1, B, K, dO=dG1)
            dY1s = _split(dY1)
        sk1 = _split_for(Wc1s.rows, Wc1s.cols, M)
        out_c1 = _conv_sink(conv_ws[:nk], sk1 > 1)
        dWc1 = _gemm_s(dY1s, Xs, a_mn=True, b_mn=True, split_k=sk1, out=out_c1)
        gauss1 = [_into(_sink(prm), dgs1[i * nk:(i + 1) * nk].view(gsh)) for i, prm in enumerate(g1p)]
        _ready(*(conv_ws[:nk] if out_c1 is not None else ()), *g1p)

        # graph learner
        Cdim = h2.shape[1]
        dh2 = kn.adjacency_topk_bwd(h2.view(B, K, Cdim), idx, alpha, dalpha, dadj).view(M, Cdim)
        dh2s = _split(dh2)
        db2 = kn.colsum(dh2, out=_sink(b2_))
        dW2 = _gemm_s(dh2s, h1s, a_mn=True, b_mn=True, split_k=_split_for(Cdim, Cdim, M))
        dh1s = kn.empty_split(M, Cdim, dev, with_lo)
        dh1 = _gemm_s(dh2s, W2s, b_mn=True, aux=h1s, aux_scale=1.0, out_split=dh1s)
        db1 = kn.colsum(dh1, out=_sink(b1_))
        s1 = _split_for(Cdim, F, M)
        dW1 = torch.zeros((Cdim, F + H), device=dev, dtype=torch.float32) if s1 > 1 else torch.empty((Cdim, F + H), device=dev, dtype=torch.float32)
        _gemm_s(dh1s, Xs, a_mn=True, b_mn=True, out=dW1[:, :F], split_k=s1)
        dqt = kn.segment_sum(dh1, K)
        dqts = _split(dqt)
        _gemm_s(dqts, qs, a_mn=True, b_mn=True, out=dW1[:, F:])
        dq_gl = _gemm_s(dqts, W1qs, b_mn=True)
        dq = dq + dq_gl

        dv1, dg1 = kn.weight_norm_bwd(dW1, v1, g1, out=(_sink(v1), _sink(g1)))
        dv2, dg2 = kn.weight_norm_bwd(dW2, v2, g2, out=(_sink(v2), _sink(g2)))

        d1 = Wc1s.rows // nk
        d2 = Wc2s.rows // nk
        conv_grads = [dWc1[i * d1:(i + 1) * d1] for i in range(nk)] + [dWc2[i * d2:(i + 1) * d2] for i in range(nk)]
        _ready(b1_, b2_, v1, g1, v2, g2)
        return (None, None, dq, dv1, dg1, db1, dv2, dg2, db2, *gauss1, *gauss2,
                dvo1, dgo1, dbo1, dvo2, dgo2, dbo2, *conv_grads)


# One kernel per forward step (product + cell out of TMEM, kernels.gru_step_fused) instead of product and cell as two launches.
# Correct and tested both ways; measured at B=512, H=1024: 3.92 ms/step fused vs 3.86 ms unfused, so it is OFF.  Either way every
# CTA of a step streams ~64 KB of operand planes per k-block into its SM (~1000 cycles per k-block against ~770 of tensor
# time, 16 k-blocks, plus launch / prologue / epilogue): the fusion removes the 6 MB gh round trip and a launch, not that.
# Multicasting the shared A tile between horizontally adjacent CTAs was tried too (same bytes arrive in every SM): no gain.
GRU_FUSED = False
# The same fused step for ALL steps in one cooperative launch (grid barrier between steps, kernels.gru_seq_fused; needs
# H/32 * ceil(B/128) <= 148 CTAs).  Also correct and tested, also slower inside the captured step: 3.93 vs 3.89 ms at B=512
# and 1.33 vs 1.22 ms at B=8 - a graph replay has no launch overhead left to save, and inside one kernel the cell of step t
# cannot overlap the product of step t+1.  "1" enables it, "0" (default) keeps product and cell as two launches per step.
GRU_FUSED_SEQ = os.environ.get("VQA_GRU_SEQ", "0")
_UB_PERM = {}


def _unit_block_perm(H: int, device) -> torch.Tensor:
    """Row permutation of the (3H, .) GRU weights into unit-block order: new row u*96 + g*32 + i <- old row g*H + u*32 + i."""
    key = (H, str(device))
    if key not in _UB_PERM:
        _UB_PERM[key] = torch.arange(3 * H, device=device).view(3, H // 32, 32).permute(1, 0, 2).reshape(-1).contiguous()
    return _UB_PERM[key]


_STEP_INDEX = {}


def _all_steps_gate(tile_len: torch.Tensor, T: int, B: int):
    """Row-tile gate of a product over all T*B rows (time-major): tile (t, j) is live while tile_len[j] - t > 0.  Only when a step
    is a whole number of 128-row tiles; otherwise None (no gating)."""
    if B % 128 != 0:
        return None
    key = (T, str(tile_len.device))
    if key not in _STEP_INDEX:
        _STEP_INDEX[key] = torch.arange(T, device=tile_len.device, dtype=torch.int32).view(T, 1)
    return ((tile_len.view(1, -1) - _STEP_INDEX[key]).reshape(-1).contiguous(), 0, "all_steps")


class QuestionEncoderFn(torch.autograd.Function):
    """(question tokens, lengths, embedding + GRU parameters) -> final GRU state per question (B, H).

    Restates nn.Embedding + pack_padded_sequence + nn.GRU + ``hid[0]`` (reference sparse_graph_model.py:117-121) as a
    padded, length-masked recurrence whose matrix products run on the split-bf16 tcgen05 GEMM (always 3 passes: the
    recurrence compounds errors over T steps).  No host-side packing, no data-dependent shapes."""

    @staticmethod
    def forward(ctx, question, qlen, T, wemb, w_ih, w_hh, b_ih, b_hh):
        B = question.shape[0]
        H = w_hh.shape[1]
        dev = wemb.device
        Es = kn.embed_gather_split(question, wemb, T)
        Wihs, Whhs = kn.split(w_ih), kn.split(w_hh)
        b_ih = b_ih.contiguous()
        b_hh = b_hh.contiguous()
        Hall = torch.empty((T + 1, B, H), device=dev, dtype=torch.float32)     # Hall[t+1] = h_t, Hall[0] = 0
        gates = torch.empty((T, B, 4 * H), device=dev, dtype=torch.float32)
        # longest question per 128-row tile: the per-step products skip row tiles whose sequences have all ended (the reference's
        # collate_fn sorts a batch by descending length, so the active rows are a shrinking prefix; any order stays correct)
        pad = (-B) % 128
        tile_len = torch.nn.functional.pad(qlen.to(torch.int32), (0, pad)).view(-1, 128).amax(dim=1).to(torch.int32).contiguous()
        fused_seq = kn.gru_seq_supported(B, H) and GRU_FUSED_SEQ in ("1", True)
        if fused_seq or (GRU_FUSED and H % 32 == 0):
            # one kernel per step: product + cell out of TMEM.  Weights / input projections in unit-block order
            # (block u = [r | z | n] of units 32u..32u+31) so that one 128 x 96 accumulator holds all gates of its units
            perm = _unit_block_perm(H, dev)
            Wih_ub, Whh_ub = kn.split(w_ih.index_select(0, perm)), kn.split(w_hh.index_select(0, perm))
            GI = kn.gemm_s(Es, Wih_ub, bias=b_ih.index_select(0, perm))        # (T*B, 3H) in unit-block column order
            b_hh_ub = b_hh.index_select(0, perm)
            Hs = kn.zeros_split((T + 1) * B, H, dev)                           # skipped row tiles stay zero (dW_hh reads all rows)
            Hall[0].zero_()
            if fused_seq:
                kn.gru_seq_fused(Hs, Whh_ub, GI, b_hh_ub, Hall, qlen, gates, tile_len, T)
            else:
                for t in range(T):
                    kn.gru_step_fused(Hs.rows_slice(t * B, (t + 1) * B), Whh_ub, GI[t * B:(t + 1) * B], b_hh_ub, Hall[t], qlen, t, Hall[t + 1],
                                      Hs.rows_slice((t + 1) * B, (t + 2) * B), gates[t], tile_len)
            # every question's state at its OWN last step (rows of skipped tiles are not carried forward)
            out = Hall[qlen.to(torch.int64).clamp(min=0, max=T), torch.arange(B, device=dev)]
        else:
            # (T*B, 3H), all steps at once; like the per-step products it skips the row tiles of (step t, tile j) whose questions have all ended
            GI = kn.gemm_s(Es, Wihs, bias=b_ih, row_gate=_all_steps_gate(tile_len, T, B))
            Hs = kn.empty_split((T + 1) * B, H, dev)
            Hall[0].zero_(); Hs.hi[:B].zero_(); Hs.lo[:B].zero_()
            # per-step product h W_hh^T: L2-bandwidth bound at M = B rows (measured sweep, tools/gru_gemm_sweep.py): 128-wide tiles, no split
            tile = 128 if B <= 1024 else 0
            for t in range(T):
                GH = kn.gemm_s(Hs.rows_slice(t * B, (t + 1) * B), Whhs, tile_n=tile, row_gate=(tile_len, t)) if t > 0 else None
                kn.gru_cell_fwd(GI[t * B:(t + 1) * B], GH, b_hh, Hall[t] if t > 0 else None, qlen, t, Hall[t + 1],
                                Hs.rows_slice((t + 1) * B, (t + 2) * B), gates[t])
            out = Hall[T]
        ctx.T = T
        ctx.prm = (w_ih, w_hh, b_ih, b_hh)
        ctx.tile_len = tile_len
        ctx.splits = (Es, Wihs, Whhs, Hs)
        ctx.save_for_backward(question, qlen, wemb, Hall, gates)
        return out

    @staticmethod
    def backward(ctx, dq):
        question, qlen, wemb, Hall, gates = ctx.saved_tensors
        Es, Wihs, Whhs, Hs = ctx.splits
        T = ctx.T
        _, B, H = Hall.shape
        dev = wemb.device
        dh = dq.contiguous()
        dGI = torch.empty((T * B, 3 * H), device=dev, dtype=torch.float32)
        dGH = torch.empty((T * B, 3 * H), device=dev, dtype=torch.float32)
        dGIs = kn.empty_split(T * B, 3 * H, dev)
        dGHs = kn.empty_split(T * B, 3 * H, dev)
        tile = 128 if B <= 1024 else 0
        ksplit = max(1, min(8, 128 // max(1, ((B + 127) // 128) * ((H + 127) // 128))))   # ~128 CTAs for the per-step product (sweep)
        for t in reversed(range(T)):
            r0, r1 = t * B, (t + 1) * B
            dh_part = torch.empty((B, H), device=dev, dtype=torch.float32)
            kn.gru_cell_bwd(dh, gates[t], Hall[t] if t > 0 else None, qlen, t, dGI[r0:r1], dGH[r0:r1],
                            dGIs.rows_slice(r0, r1), dGHs.rows_slice(r0, r1), dh_part)
            if t > 0:   # dL/dh_{t-1} = direct part + dgh . W_hh  (split-K accumulating into the direct part)
                kn.gemm_s(dGHs.rows_slice(r0, r1), Whhs, b_mn=True, out=dh_part, accumulate=True, split_k=ksplit, tile_n=tile,
                          row_gate=(ctx.tile_len, t))
            dh = dh_part
        w_ih_, w_hh_, b_ih_, b_hh_ = ctx.prm
        TB = T * B
        # the embedding gradient first: it is the largest tensor of the last all-reduce bucket, which then travels under the two
        # weight-gradient products below instead of after them
        dwemb = None
        if ctx.needs_input_grad[3]:
            # (T*B, E); rows of ended questions are never read by the scatter below: their row tiles are skipped
            dE = kn.gemm_s(dGIs, Wihs, b_mn=True, row_gate=_all_steps_gate(ctx.tile_len, T, B), **_plan(TB, Wihs.cols, Wihs.rows))
            dwemb = _sink(wemb)
            sunk = dwemb is not None
            if dwemb is None:
                dwemb = torch.zeros_like(wemb)
            else:
                dwemb.zero_()
            kn.embed_scatter_add(dE, question, qlen, dwemb, T)
            if sunk:
                _ready(wemb)
        db_ih = kn.colsum(dGI, out=_sink(b_ih_))
        # dgh equals dgi in the r and z thirds (the reset gate only scales the candidate's hidden term, gru.cu): only the last third
        # of b_hh's gradient needs a column sum of its own
        db_hh = _sink(b_hh_)
        if db_hh is None:
            db_hh = torch.empty_like(db_ih)
        db_hh[:2 * H].copy_(db_ih[:2 * H])
        kn.colsum(dGH[:, 2 * H:], out=db_hh[2 * H:])

        def wgrad(a, b, prm):
            plan = _plan(a.cols, b.cols, TB)
            plan.setdefault("split_k", _split_for(a.cols, b.cols, TB))
            out = _sink(prm)
            if out is not None and plan["split_k"] > 1:
                out.zero_()
            return kn.gemm_s(a, b, a_mn=True, b_mn=True, out=out, **plan)
        dW_hh = wgrad(dGHs, Hs.rows_slice(0, TB), w_hh_)
        _ready(b_ih_, b_hh_, w_hh_)
        dW_ih = wgrad(dGIs, Es, w_ih_)                      # the smallest weight last: what is left exposed after backward is its bucket
        _ready(w_ih_)
        return None, None, None, dwemb, dW_ih, dW_hh, db_ih, db_hh


# ------------------------------------------------------------------------------------------- layer-level operators
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on the tcgen05 GEMM (x: (M,in), W: (out,in))."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        x = x.contiguous()
        w = w.contiguous()
        y = _gemm(x, w, bias=b, relu=relu)
        ctx.relu = relu
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = torch.where(y > 0, dy, torch.zeros_like(dy))
        dx = _gemm(dy, w, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = _gemm(dy, x, a_mn=True, b_mn=True, split_k=_split_for(w.shape[0], w.shape[1], x.shape[0]))
        db = kn.colsum(dy) if ctx.has_bias else None
        return dx, dw, db, None


class WeightNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, g):
        ctx.save_for_backward(v, g)
        return kn.weight_norm_fwd(v, g)

    @staticmethod
    def backward(ctx, dw):
        v, g = ctx.saved_tensors
        return kn.weight_norm_bwd(dw, v, g)


class AdjacencyFn(torch.autograd.Function):
    """A = h h^T per image via the fused adjacency kernel (h must be a ReLU output, as in GraphLearner)."""

    @staticmethod
    def forward(ctx, h):
        h = h.contiguous()
        adj, idx, alpha = kn.adjacency_topk_fwd(h, 1)
        ctx.save_for_backward(h, idx, alpha)
        return adj

    @staticmethod
    def backward(ctx, dadj):
        h, idx, alpha = ctx.saved_tensors
        return kn.adjacency_topk_bwd(h, idx, alpha, torch.zeros_like(alpha), dadj.contiguous())


class PatchOperatorFn(torch.autograd.Function):
    """Z[n,k,:] = sum_m w[n,m,k] X[n,m,:] on materialised neighbourhoods (reference layers.py:136-137), kernels both ways."""

    @staticmethod
    def forward(ctx, X, w):
        X = X.contiguous(); w = w.contiguous()
        ctx.save_for_backward(X, w)
        return kn.patch_operator_fwd(X, w)

    @staticmethod
    def backward(ctx, dZ):
        X, w = ctx.saved_tensors
        return kn.patch_operator_bwd(X, w, dZ.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1])


class PerKernelLinearFn(torch.autograd.Function):
    """out[:, k*D:(k+1)*D] = Z[:, k, :] W_k^T (+ b_k) for the nk bias-free (by default) linear maps of layers.py:139-142, every product
    a tcgen05 GEMM that reads its slice of Z / writes its slice of the output in place.  The backward writes dZ slice by slice too:
    the reference's autograd zero-fills a Z-sized tensor per kernel here (``select_backward``)."""

    @staticmethod
    def forward(ctx, z, nk, has_bias, *wb):
        z = z.contiguous()
        n, _, fin = z.shape
        ws, bs = wb[:nk], (wb[nk:] if has_bias else (None,) * nk)
        d = ws[0].shape[0]
        if fin % 4 or d % 4:
            raise RuntimeError(f"NeighbourhoodGraphConvolution.convolution: in_feat_dim ({fin}) and out_feat_dim / n_kernels ({d}) must be "
                               "multiples of 4 (16-byte aligned rows for the TMA loads)")
        z2 = z.view(n, nk * fin)
        out = torch.empty((n, nk * d), device=z.device, dtype=torch.float32)
        for k in range(nk):
            _gemm(z2[:, k * fin:(k + 1) * fin], ws[k].contiguous(), out=out[:, k * d:(k + 1) * d], bias=bs[k])
        ctx.nk, ctx.has_bias = nk, has_bias
        ctx.save_for_backward(z, *ws)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, *ws = ctx.saved_tensors
        nk = ctx.nk
        n, _, fin = z.shape
        d = ws[0].shape[0]
        dout = dout.contiguous()
        z2 = z.view(n, nk * fin)
        dz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        dws, dbs = [], []
        for k in range(nk):
            dk = dout[:, k * d:(k + 1) * d]
            if dz is not None:
                _gemm(dk, ws[k].contiguous(), b_mn=True, out=dz.view(n, nk * fin)[:, k * fin:(k + 1) * fin])
            dws.append(_gemm(dk, z2[:, k * fin:(k + 1) * fin], a_mn=True, b_mn=True, split_k=_split_for(d, fin, n)))
            if ctx.has_bias:
                dbs.append(kn.colsum(dk))
        return (dz, None, None, *dws, *dbs)


def _gaussian_weights_torch(pseudo, mr, pr, mt, pt):
    """Differentiable re-evaluation (torch ops) used only for the BACKWARD of the layer-level API."""
    import math
    rho = pseudo[..., 0].reshape(-1, 1)
    theta = pseudo[..., 1].reshape(-1, 1)
    wr = torch.exp(-0.5 * (rho - mr.view(1, -1)) ** 2 / (1e-14 + pr.view(1, -1) ** 2))
    a1 = torch.abs(theta - mt.view(1, -1))
    a2 = torch.abs(2 * math.pi - a1)
    wt = torch.exp(-0.5 * torch.minimum(a1, a2) ** 2 / (1e-14 + pt.view(1, -1) ** 2))
    w = wr * wt
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
    return w / w.sum(dim=1, keepdim=True)


class GaussianWeightsFn(torch.autograd.Function):
    """get_gaussian_weights of the layer API: forward = CUDA kernel, backward = autograd through the torch formula."""

    @staticmethod
    def forward(ctx, pseudo, mr, pr, mt, pt):
        ctx.save_for_backward(pseudo, mr, pr, mt, pt)
        return kn.gaussian_weights(pseudo, pack_gauss(mr, pr, mt, pt))

    @staticmethod
    def backward(ctx, dw):
        pseudo, mr, pr, mt, pt = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in (mr, pr, mt, pt)]
            w = _gaussian_weights_torch(pseudo.detach(), *leaves)
            grads = torch.autograd.grad(w, leaves, dw)
        return (None, *grads)
