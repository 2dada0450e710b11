"""Tensor-level wrappers over the C ABI: argument checking, output allocation, stream plumbing.

PyTorch is used here only for device memory and streams.  Every function enqueues hand-written
sm_100a kernels from ``libvqa_sm100.so`` on the current CUDA stream and returns immediately.
``LAUNCHES`` counts kernel launches issued through this module (bench.py reports it).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _cabi
from ._cabi import PREC_TF32, PREC_TF32X3, PREC_TF32X3_HP, GEMM_RELU, GEMM_ACCUMULATE, GEMM_NO_CLUSTER, GC_RELU

LAUNCHES = 0
_LAUNCH_COST = {}          # every entry point is one launch (the column sum became a single kernel)
GEMM_LOG = None            # list of (M, N, K, passes, row-gate step or None) per split-bf16 product while a caller (bench.py) collects it


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the vqa_b200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def set_sm_budget(n_sms: int) -> int:
    """SMs the persistent one-CTA-per-SM kernels may occupy from now on (8..148); returns the previous value.  Host-side state read at
    launch time; see include/vqa_b200.h."""
    return int(_cabi.load().vqa_set_sm_budget(int(n_sms)))


# name -> list of (start_event, end_event): filled when a caller (bench.py) asks for in-stream kernel timing
TIMERS = {}


def enable_timing(*names: str) -> None:
    """Record CUDA events on the launching stream around every launch of the named entry points."""
    TIMERS.clear()
    for n in names:
        TIMERS[n] = []


def _call(name, *args):
    global LAUNCHES
    rec = TIMERS.get(name)
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.call(name, *args)
        e1.record()
        rec.append((e0, e1))
    else:
        _cabi.call(name, *args)
    LAUNCHES += _LAUNCH_COST.get(name, 1)


def _rows_view(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """2-D tensor whose rows are contiguous -> (tensor, leading dimension in elements)."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name}: need a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t, ld


# ------------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None, rowbcast: Optional[torch.Tensor] = None, group: int = 1,
         aux: Optional[torch.Tensor] = None, aux_scale: float = 1.0, relu: bool = False,
         precision: int = PREC_TF32X3, split_k: int = 1, tile_n: int = 0) -> torch.Tensor:
    """C[M,N] = epi(A . B^T) on tcgen05.  ``a`` is (M,K) [a_mn=False] or (K,M) [a_mn=True]; same for ``b`` with N."""
    _chk(a, "gemm A"); _chk(b, "gemm B")
    a, lda = _rows_view(a, "gemm A")
    b, ldb = _rows_view(b, "gemm B")
    M, Ka = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if Ka != Kb:
        raise RuntimeError(f"gemm: contraction mismatch {Ka} vs {Kb}")
    if lda % 4 or ldb % 4 or a.data_ptr() % 16 or b.data_ptr() % 16:
        raise RuntimeError("gemm: TMA needs 16-byte aligned operands with leading dimensions that are multiples of 4 floats "
                           f"(lda={lda}, ldb={ldb}); pad the feature dimension")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
        if split_k > 1:
            out.zero_()
    else:
        _chk(out, "gemm out")
        if out.shape != (M, N):
            raise RuntimeError(f"gemm: out has shape {tuple(out.shape)}, expected {(M, N)}")
    out, ldc = _rows_view(out, "gemm out")
    ldrb = ldaux = 0
    if rowbcast is not None:
        rowbcast, ldrb = _rows_view(_chk(rowbcast, "gemm rowbcast"), "gemm rowbcast")
    if aux is not None:
        aux, ldaux = _rows_view(_chk(aux, "gemm aux"), "gemm aux")
    _call("vqa_gemm_f32", a.data_ptr(), lda, int(a_mn), b.data_ptr(), ldb, int(b_mn), out.data_ptr(), ldc, M, N, Ka,
          _ptr(bias), _ptr(rowbcast), ldrb, group, _ptr(aux), ldaux, float(aux_scale), GEMM_RELU if relu else 0,
          precision, split_k, tile_n, _stream())
    return out


# ------------------------------------------------------------------------------------------- split-bf16 GEMM
class SplitT:
    """An fp32 matrix stored as two bf16 planes (hi, lo) of shape (rows, ld), ld % 8 == 0; lo is None in bf16 mode.
    ``shape`` is the logical (rows, cols)."""
    __slots__ = ("hi", "lo", "rows", "cols", "ld")

    def __init__(self, hi, lo, rows, cols, ld):
        self.hi, self.lo, self.rows, self.cols, self.ld = hi, lo, rows, cols, ld

    @property
    def shape(self):
        return (self.rows, self.cols)

    def float(self) -> torch.Tensor:
        x = self.hi[:, :self.cols].float()
        return x + self.lo[:, :self.cols].float() if self.lo is not None else x

    def rows_slice(self, r0: int, r1: int) -> "SplitT":
        return SplitT(self.hi[r0:r1], None if self.lo is None else self.lo[r0:r1], r1 - r0, self.cols, self.ld)

    def cols_slice(self, c0: int, c1: int) -> "SplitT":
        if c0 % 8:
            raise RuntimeError("SplitT.cols_slice: start column must be a multiple of 8 (16-byte TMA alignment)")
        return SplitT(self.hi[:, c0:], None if self.lo is None else self.lo[:, c0:], self.rows, c1 - c0, self.ld)


def empty_split(rows: int, cols: int, device, with_lo: bool = True) -> SplitT:
    ld = (cols + 7) // 8 * 8
    buf = torch.empty((2 if with_lo else 1, rows, ld), device=device, dtype=torch.bfloat16)
    return SplitT(buf[0], buf[1] if with_lo else None, rows, cols, ld)


def zeros_split(rows: int, cols: int, device, with_lo: bool = True) -> SplitT:
    ld = (cols + 7) // 8 * 8
    buf = torch.zeros((2 if with_lo else 1, rows, ld), device=device, dtype=torch.bfloat16)
    return SplitT(buf[0], buf[1] if with_lo else None, rows, cols, ld)


def split(x: torch.Tensor, with_lo: bool = True) -> SplitT:
    """fp32 (rows, cols) -> bf16 planes hi = bf16(x), lo = bf16(x - hi)."""
    x, ldx = _rows_view(_chk(x, "split x"), "split x")
    rows, cols = x.shape
    out = empty_split(rows, cols, x.device, with_lo)
    _call("vqa_split_bf16_f32", x.data_ptr(), ldx, out.hi.data_ptr(), _ptr(out.lo), out.ld, rows, cols, _stream())
    return out


def dropout_split(x: torch.Tensor, p: float, seed: int, offset: int, step: Optional[torch.Tensor] = None, with_lo: bool = True) -> SplitT:
    """Planes of dropout(x) for a 2-D fp32 x (rows, cols) in one pass."""
    x, ldx = _rows_view(_chk(x, "dropout_split x"), "dropout_split x")
    rows, cols = x.shape
    out = empty_split(rows, cols, x.device, with_lo)
    _call("vqa_dropout_split_f32", x.data_ptr(), ldx, out.hi.data_ptr(), _ptr(out.lo), out.ld, rows, cols, float(p), seed, offset,
          _ptr(step), _stream())
    return out


def gemm_s(a: SplitT, b: SplitT, *, a_mn: bool = False, b_mn: bool = False, out: Optional[torch.Tensor] = None,
           out_split: Optional[SplitT] = None, want_f32: bool = True, bias: Optional[torch.Tensor] = None,
           rowbcast: Optional[torch.Tensor] = None, group: int = 1, aux=None, aux_scale: float = 1.0, relu: bool = False,
           passes: int = 3, split_k: int = 1, tile_n: int = 0, accumulate: bool = False, cluster: bool = True,
           row_gate=None):
    """C[M,N] = epi(A . B^T) on the split-bf16 tcgen05 GEMM (``accumulate``: ``out += A . B^T`` with fp32 atomics).  ``a`` is (M,K) [a_mn=False] or (K,M) [a_mn=True]; same
    for ``b`` with N.  ``aux`` (mask source) may be an fp32 tensor or a SplitT.  Returns ``out`` (fp32) or, when
    ``want_f32`` is False, ``out_split``.  ``row_gate`` = (int32 device tensor with one entry per 128-row tile of the output,
    t): row tiles with entry <= t are skipped entirely (padded recurrences; a third element "all_steps" marks a gate tensor that
    covers the row tiles of every time step of a (T*B, .) product - only bench.py's FLOP count reads it)."""
    M, Ka = (a.cols, a.rows) if a_mn else (a.rows, a.cols)
    N, Kb = (b.cols, b.rows) if b_mn else (b.rows, b.cols)
    if Ka != Kb:
        raise RuntimeError(f"gemm_s: contraction mismatch {Ka} vs {Kb}")
    if passes == 3 and (a.lo is None or b.lo is None):
        raise RuntimeError("gemm_s: 3-pass mode needs operands split with with_lo=True")
    dev = a.hi.device
    ldc = 0
    if want_f32:
        if out is None:
            if accumulate:
                raise RuntimeError("gemm_s: accumulate=True needs an initialised out")
            out = torch.empty((M, N), device=dev, dtype=torch.float32)
            if split_k > 1:
                out.zero_()
        elif out.shape != (M, N):
            raise RuntimeError(f"gemm_s: out has shape {tuple(out.shape)}, expected {(M, N)}")
        out, ldc = _rows_view(_chk(out, "gemm_s out"), "gemm_s out")
    else:
        out = None
        if out_split is None:
            raise RuntimeError("gemm_s: want_f32=False needs out_split")
    if out_split is not None and out_split.shape != (M, N):
        raise RuntimeError(f"gemm_s: out_split has shape {out_split.shape}, expected {(M, N)}")
    ldrb = ldaux = ldauxh = 0
    aux_f = aux_h = None
    if GEMM_LOG is not None:
        GEMM_LOG.append((M, N, Ka, passes, None if row_gate is None else (-1 if len(row_gate) > 2 else int(row_gate[1]))))
    if rowbcast is not None:
        rowbcast, ldrb = _rows_view(_chk(rowbcast, "gemm_s rowbcast"), "gemm_s rowbcast")
    if isinstance(aux, SplitT):
        aux_h, ldauxh = aux.hi, aux.ld
    elif aux is not None:
        aux_f, ldaux = _rows_view(_chk(aux, "gemm_s aux"), "gemm_s aux")
    _call("vqa_gemm_bf16s", a.hi.data_ptr(), _ptr(a.lo) if passes == 3 else None, a.ld, int(a_mn),
          b.hi.data_ptr(), _ptr(b.lo) if passes == 3 else None, b.ld, int(b_mn), _ptr(out), ldc,
          None if out_split is None else out_split.hi.data_ptr(), None if out_split is None else _ptr(out_split.lo),
          0 if out_split is None else out_split.ld, M, N, Ka, _ptr(bias), _ptr(rowbcast), ldrb, group,
          _ptr(aux_f), ldaux, _ptr(aux_h), ldauxh, float(aux_scale), (GEMM_RELU if relu else 0) | (GEMM_ACCUMULATE if accumulate else 0) | (0 if cluster else GEMM_NO_CLUSTER),
          passes, split_k, tile_n, None if row_gate is None else row_gate[0].data_ptr(), 0 if row_gate is None else int(row_gate[1]), _stream())
    return out if want_f32 else out_split


# ------------------------------------------------------------------------------------------- elementwise
def dropout(x: torch.Tensor, p: float, seed: int, offset: int, step: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``step``: optional CUDA int64 scalar tensor; 16 * its run-time value is added to ``offset`` (graph replays)."""
    x = _chk(x, "dropout x").contiguous()
    y = torch.empty_like(x)
    _call("vqa_dropout_f32", x.data_ptr(), y.data_ptr(), x.numel(), float(p), seed, offset, _ptr(step), _stream())
    return y


def weight_norm_fwd(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    v = _chk(v, "weight_norm v").contiguous(); g = _chk(g, "weight_norm g").contiguous()
    w = torch.empty_like(v)
    _call("vqa_weight_norm_fwd_f32", v.data_ptr(), g.data_ptr(), w.data_ptr(), v.shape[0], v.shape[1], _stream())
    return w


def weight_norm_split(v: torch.Tensor, g: torch.Tensor, c0: int = 0, c1: Optional[int] = None, with_lo: bool = True) -> "SplitT":
    """Planes of columns [c0, c1) of the weight-normed matrix v * g / ||v|| (no fp32 intermediate)."""
    v = _chk(v, "weight_norm v").contiguous(); g = _chk(g, "weight_norm g").contiguous()
    rows, cols = v.shape
    c1 = cols if c1 is None else c1
    if cols % 4 or c0 % 4 or v.data_ptr() % 16:            # rows not 16-byte aligned: two kernels instead of one
        w = weight_norm_fwd(v, g)
        return split(w[:, c0:c1] if (c0, c1) == (0, cols) else w[:, c0:c1].contiguous(), with_lo)
    out = empty_split(rows, c1 - c0, v.device, with_lo)
    _call("vqa_weight_norm_split_f32", v.data_ptr(), g.data_ptr(), rows, cols, c0, c1, out.hi.data_ptr(), _ptr(out.lo), out.ld, _stream())
    return out


def weight_norm_bwd(dw: torch.Tensor, v: torch.Tensor, g: torch.Tensor, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``out``: optional (dv, dg) destinations (contiguous, e.g. views of a flat gradient buffer)."""
    dw = _chk(dw, "weight_norm dw").contiguous(); v = v.contiguous(); g = g.contiguous()
    dv, dg = out if out is not None and out[0] is not None and out[1] is not None else (None, None)
    if dv is None:
        dv = torch.empty_like(v)
        dg = torch.empty_like(g)
    elif dv.shape != v.shape or dg.shape != g.shape or not dv.is_contiguous() or not dg.is_contiguous():
        raise RuntimeError("weight_norm_bwd: out tensors must be contiguous and shaped like (v, g)")
    _call("vqa_weight_norm_bwd_f32", dw.data_ptr(), v.data_ptr(), g.data_ptr(), dv.data_ptr(), dg.data_ptr(),
          v.shape[0], v.shape[1], _stream())
    return dv, dg


_COLSUM_COUNTERS = {}


def _colsum_counters(device: torch.device) -> torch.Tensor:
    """Persistent zero-initialised ticket counters of the single-launch column sum (the kernel leaves them zero).  One set per
    (device, stream): launches on one stream are ordered, launches on different streams may overlap and must not share tickets."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    buf = _COLSUM_COUNTERS.get(key)
    if buf is None:
        buf = _COLSUM_COUNTERS[key] = torch.zeros(4096, device=device, dtype=torch.int32)
    return buf


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    x, ldx = _rows_view(_chk(x, "colsum x"), "colsum x")
    if out is None:
        out = torch.empty(x.shape[1], device=x.device, dtype=torch.float32)
    elif out.numel() != x.shape[1] or not out.is_contiguous():
        raise RuntimeError("colsum: out must be a contiguous tensor with one element per column")
    scratch = torch.empty(256 * x.shape[1], device=x.device, dtype=torch.float32)
    cnt = _colsum_counters(x.device) if x.shape[1] <= 4096 * 256 else None
    _call("vqa_colsum_f32", x.data_ptr(), ldx, out.data_ptr(), scratch.data_ptr(), x.shape[0], x.shape[1], _ptr(cnt), _stream())
    return out


def segment_sum(x: torch.Tensor, seg_len: int) -> torch.Tensor:
    x = _chk(x, "segment_sum x").contiguous()
    rows, cols = x.shape
    if rows % seg_len:
        raise RuntimeError("segment_sum: rows not divisible by segment length")
    out = torch.empty((rows // seg_len, cols), device=x.device, dtype=torch.float32)
    _call("vqa_segment_sum_f32", x.data_ptr(), out.data_ptr(), rows // seg_len, seg_len, cols, _stream())
    return out


def gate_bwd(dhq, q, pooled):
    dhq = _chk(dhq, "gate dhq").contiguous(); q = q.contiguous(); pooled = pooled.contiguous()
    dpooled = torch.empty_like(dhq)
    dq = torch.empty_like(dhq)
    _call("vqa_gate_bwd_f32", dhq.data_ptr(), q.data_ptr(), pooled.data_ptr(), dpooled.data_ptr(), dq.data_ptr(),
          dhq.numel(), _stream())
    return dpooled, dq


# ------------------------------------------------------------------------------------------- graph learner tail
def adjacency_topk_fwd(h: torch.Tensor, nb: int):
    """h (B,K,C) -> adjacency (B,K,K), idx (B,K,nb) int32 (descending value), alpha (B,K,nb)."""
    h = _chk(h, "adjacency h").contiguous()
    B, K, Cdim = h.shape
    adj = torch.empty((B, K, K), device=h.device, dtype=torch.float32)
    idx = torch.empty((B, K, nb), device=h.device, dtype=torch.int32)
    alpha = torch.empty((B, K, nb), device=h.device, dtype=torch.float32)
    _call("vqa_adjacency_topk_fwd_f32", h.data_ptr(), adj.data_ptr(), idx.data_ptr(), alpha.data_ptr(), B, K, Cdim, nb, _stream())
    return adj, idx, alpha


def topk_softmax(adj: torch.Tensor, nb: int):
    adj = _chk(adj, "topk adjacency").contiguous()
    B, K, _ = adj.shape
    idx = torch.empty((B, K, nb), device=adj.device, dtype=torch.int32)
    alpha = torch.empty((B, K, nb), device=adj.device, dtype=torch.float32)
    _call("vqa_topk_softmax_f32", adj.data_ptr(), idx.data_ptr(), alpha.data_ptr(), B, K, nb, _stream())
    return idx, alpha


def adjacency_topk_bwd(h, idx, alpha, dalpha, dadj=None):
    h = _chk(h, "adjacency h").contiguous()
    B, K, Cdim = h.shape
    nb = idx.shape[-1]
    dalpha = _chk(dalpha, "dalpha").contiguous()
    if dadj is not None:
        dadj = _chk(dadj, "dadj").contiguous()
    dh = torch.empty_like(h)
    _call("vqa_adjacency_topk_bwd_f32", h.data_ptr(), idx.data_ptr(), alpha.data_ptr(), dalpha.data_ptr(), _ptr(dadj),
          dh.data_ptr(), B, K, Cdim, nb, _stream())
    return dh


# ------------------------------------------------------------------------------------------- graph convolution
def _boxes_view(image: torch.Tensor):
    """(B,K,F) image -> pointer to the first box column + row stride; no copy."""
    _chk(image, "image")
    if image.dim() != 3 or image.stride(2) != 1 or image.stride(0) != image.shape[1] * image.stride(1):
        raise RuntimeError("image must be (B,K,F) with contiguous rows")
    return image.data_ptr() + (image.shape[2] - 4) * 4, image.stride(1)


def graphconv_fwd(Y, idx, alpha, image, gauss, B, K, relu=True, dropout_p=0.0, seed=0, offset=0, step=None):
    Y, ldy = _rows_view(_chk(Y, "graphconv Y"), "graphconv Y")
    nb = idx.shape[-1]
    nk = gauss.numel() // 4
    out_dim = Y.shape[1]
    bptr, ldbox = _boxes_view(image)
    out = torch.empty((B * K, out_dim), device=Y.device, dtype=torch.float32)
    _call("vqa_graphconv_fwd_f32", Y.data_ptr(), ldy, idx.data_ptr(), _ptr(alpha), bptr, ldbox, gauss.data_ptr(),
          out.data_ptr(), out_dim, B, K, nb, nk, out_dim, GC_RELU if relu else 0, float(dropout_p), seed, offset, _ptr(step), _stream())
    return out


def graphconv_pool_fwd(Y, idx, image, gauss, q, B, K):
    Y, ldy = _rows_view(_chk(Y, "graphconv Y"), "graphconv Y")
    nb = idx.shape[-1]
    nk = gauss.numel() // 4
    out_dim = Y.shape[1]
    bptr, ldbox = _boxes_view(image)
    q = _chk(q, "q").contiguous()
    pooled = torch.empty((B, out_dim), device=Y.device, dtype=torch.float32)
    argmax = torch.empty((B, out_dim), device=Y.device, dtype=torch.int64)
    hq = torch.empty((B, out_dim), device=Y.device, dtype=torch.float32)
    _call("vqa_graphconv_pool_fwd_f32", Y.data_ptr(), ldy, idx.data_ptr(), bptr, ldbox, gauss.data_ptr(), q.data_ptr(),
          pooled.data_ptr(), argmax.data_ptr(), hq.data_ptr(), B, K, nb, nk, out_dim, _stream())
    return pooled, argmax, hq


def graphconv_bwd(Y, idx, alpha, image, gauss, B, K, dO=None, dpooled=None, argmax=None, want_dY=True, want_edges=True, out_dim=None):
    """-> dY (B*K,out) (None when ``want_dY`` is False), dalpha (B,K,nb) or None, dgauss (4*nk,) (None when ``want_edges``
    is False; ``Y`` may then be None and ``out_dim`` must be given)."""
    nb = idx.shape[-1]
    nk = gauss.numel() // 4
    ldy = 0
    if Y is not None:
        Y, ldy = _rows_view(_chk(Y, "graphconv Y"), "graphconv Y")
        out_dim = Y.shape[1]
    dev = idx.device
    bptr, ldbox = _boxes_view(image)
    dY = torch.empty((B * K, out_dim), device=dev, dtype=torch.float32) if want_dY else None
    P = torch.empty((B, K, nb, nk), device=dev, dtype=torch.float32) if want_edges else None
    lddo = 0
    if dO is not None:
        dO, lddo = _rows_view(_chk(dO, "graphconv dO"), "graphconv dO")
    else:
        dpooled = _chk(dpooled, "dpooled").contiguous()
    _call("vqa_graphconv_bwd_f32", _ptr(dO), lddo, _ptr(dpooled), _ptr(argmax), _ptr(Y), ldy, idx.data_ptr(),
          _ptr(alpha), bptr, ldbox, gauss.data_ptr(), _ptr(dY), out_dim, _ptr(P), B, K, nb, nk, out_dim, _stream())
    if not want_edges:
        return dY, None, None
    nblk = _cabi.load().vqa_graphconv_edge_blocks(B, K, nb)
    partial = torch.empty((nblk, 4 * nk), device=dev, dtype=torch.float32)
    dalpha = torch.empty((B, K, nb), device=dev, dtype=torch.float32) if alpha is not None else None
    _call("vqa_graphconv_edge_bwd_f32", P.data_ptr(), idx.data_ptr(), _ptr(alpha), bptr, ldbox, gauss.data_ptr(),
          _ptr(dalpha), partial.data_ptr(), B, K, nb, nk, _stream())
    dgauss = colsum(partial)
    return dY, dalpha, dgauss


# ---- tensor-core variants on split planes -------------------------------------------------------------------------
def mma_eligible(K: int, out_dim: int, nk: int) -> bool:
    return K <= 128 and out_dim % nk == 0 and (out_dim // nk) % 128 == 0


def graphconv_edge_coef(idx, alpha, image, gauss, B, K):
    """Per-edge coefficients of one layer, evaluated once per step: (coef (B,K,nb,nk) f32 = normalised Gaussian weight * alpha,
    eoff (B,K,nb) i32 = packed positions inside the shared-memory coefficient matrices).  Feeds the three aggregates below."""
    nb, nk = idx.shape[-1], gauss.numel() // 4
    bptr, ldbox = _boxes_view(image)
    coef = torch.empty((B, K, nb, nk), device=idx.device, dtype=torch.float32)
    eoff = torch.empty((B, K, nb), device=idx.device, dtype=torch.int32)
    _call("vqa_graphconv_edge_coef", idx.data_ptr(), _ptr(alpha), bptr, ldbox, gauss.data_ptr(), coef.data_ptr(), eoff.data_ptr(),
          B, K, nb, nk, _stream())
    return coef, eoff


def graphconv_fwd_s(Ys: SplitT, idx, alpha, image, gauss, B, K, relu=True, dropout_p=0.0, seed=0, offset=0, step=None, ec=None) -> SplitT:
    """Layer-1 style aggregate on planes: Ys (B*K, out) -> relu/dropout(aggregate) as planes.  ec: graphconv_edge_coef() result."""
    nb, nk, out_dim = idx.shape[-1], gauss.numel() // 4, Ys.cols
    bptr, ldbox = _boxes_view(image)
    if ec is None:
        ec = graphconv_edge_coef(idx, alpha, image, gauss, B, K)
    out = empty_split(B * K, out_dim, Ys.hi.device, Ys.lo is not None)
    _call("vqa_graphconv_mma_fwd", Ys.hi.data_ptr(), _ptr(Ys.lo), Ys.ld, idx.data_ptr(), _ptr(alpha), bptr, ldbox, gauss.data_ptr(),
          out.hi.data_ptr(), _ptr(out.lo), out.ld, B, K, nb, nk, out_dim, GC_RELU if relu else 0, float(dropout_p), seed, offset,
          _ptr(step), ec[0].data_ptr(), ec[1].data_ptr(), _stream())
    return out


def graphconv_pool_fwd_s(Ys: SplitT, idx, image, gauss, q, B, K, ec=None):
    nb, nk, out_dim = idx.shape[-1], gauss.numel() // 4, Ys.cols
    bptr, ldbox = _boxes_view(image)
    if ec is None:
        ec = graphconv_edge_coef(idx, None, image, gauss, B, K)
    q = _chk(q, "q").contiguous()
    dev = Ys.hi.device
    pooled = torch.empty((B, out_dim), device=dev, dtype=torch.float32)
    argmax = torch.empty((B, out_dim), device=dev, dtype=torch.int64)
    hq = torch.empty((B, out_dim), device=dev, dtype=torch.float32)
    _call("vqa_graphconv_mma_pool_fwd", Ys.hi.data_ptr(), _ptr(Ys.lo), Ys.ld, idx.data_ptr(), bptr, ldbox, gauss.data_ptr(), q.data_ptr(),
          pooled.data_ptr(), argmax.data_ptr(), hq.data_ptr(), B, K, nb, nk, out_dim, ec[0].data_ptr(), ec[1].data_ptr(), _stream())
    return pooled, argmax, hq


def graphconv_bwd_data_s(dOs: SplitT, idx, alpha, image, gauss, B, K, ec=None) -> SplitT:
    """dY = M^T dO on planes."""
    nb, nk, out_dim = idx.shape[-1], gauss.numel() // 4, dOs.cols
    bptr, ldbox = _boxes_view(image)
    if ec is None:
        ec = graphconv_edge_coef(idx, alpha, image, gauss, B, K)
    out = empty_split(B * K, out_dim, dOs.hi.device, dOs.lo is not None)
    _call("vqa_graphconv_mma_bwd_data", dOs.hi.data_ptr(), _ptr(dOs.lo), dOs.ld, idx.data_ptr(), _ptr(alpha), bptr, ldbox, gauss.data_ptr(),
          out.hi.data_ptr(), _ptr(out.lo), out.ld, B, K, nb, nk, out_dim, ec[0].data_ptr(), ec[1].data_ptr(), _stream())
    return out


def graphconv_pool_bwd_data_s(dpooled, argmax, idx, ec, B, K, out_dim, with_lo=True) -> SplitT:
    """dY of the pooled layer as planes: the nb products coef * dpooled of every column land in the rows of its arg-max node's neighbours."""
    nb, nk = idx.shape[-1], ec[0].shape[-1]
    dpooled = _chk(dpooled, "dpooled").contiguous()
    out = empty_split(B * K, out_dim, dpooled.device, with_lo)
    _call("vqa_graphconv_pool_bwd_data", dpooled.data_ptr(), argmax.data_ptr(), idx.data_ptr(), ec[0].data_ptr(), out.hi.data_ptr(), _ptr(out.lo),
          out.ld, B, K, nb, nk, out_dim, _stream())
    return out


def graphconv_bwd_edges_s(Ys: SplitT, idx, alpha, image, gauss, B, K, dOs: Optional[SplitT] = None, dpooled=None, argmax=None):
    """Edge part of the backward on planes -> (dalpha (B,K,nb) or None, dgauss (4*nk,))."""
    nb, nk, out_dim = idx.shape[-1], gauss.numel() // 4, Ys.cols
    bptr, ldbox = _boxes_view(image)
    dev = Ys.hi.device
    dalpha = torch.empty((B, K, nb), device=dev, dtype=torch.float32) if alpha is not None else None
    partial = torch.empty((B, 4 * nk), device=dev, dtype=torch.float32)
    if dOs is None:
        dpooled = _chk(dpooled, "dpooled").contiguous()
    scratch = torch.empty((B, nk, K * nb), device=dev, dtype=torch.float32)       # selected edge products, written and read within the call
    _call("vqa_graphconv_mma_bwd_edges", None if dOs is None else dOs.hi.data_ptr(), None if dOs is None else _ptr(dOs.lo),
          0 if dOs is None else dOs.ld, _ptr(dpooled), _ptr(argmax), Ys.hi.data_ptr(), _ptr(Ys.lo), Ys.ld, idx.data_ptr(), _ptr(alpha),
          bptr, ldbox, gauss.data_ptr(), _ptr(dalpha), partial.data_ptr(), _ptr(scratch), B, K, nb, nk, out_dim, _stream())
    return dalpha, colsum(partial)


def patch_operator_fwd(X: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """X (n, nb, F), w (n, nb, nk) -> Z (n, nk, F): Z[n,k] = sum_m w[n,m,k] X[n,m]  (layers.py:136-137)."""
    X = _chk(X, "patch operator X").contiguous(); w = _chk(w, "patch operator w").contiguous()
    n, nb, F = X.shape
    if w.shape[:2] != (n, nb):
        raise RuntimeError(f"patch_operator: weights {tuple(w.shape)} do not match neighbourhoods {tuple(X.shape)}")
    nk = w.shape[2]
    Z = torch.empty((n, nk, F), device=X.device, dtype=torch.float32)
    _call("vqa_patch_operator_fwd_f32", X.data_ptr(), w.data_ptr(), Z.data_ptr(), n, nb, nk, F, _stream())
    return Z


def patch_operator_bwd(X: torch.Tensor, w: torch.Tensor, dZ: torch.Tensor, want_dX: bool = True, want_dw: bool = True):
    X = _chk(X, "patch operator X").contiguous(); w = _chk(w, "patch operator w").contiguous(); dZ = _chk(dZ, "patch operator dZ").contiguous()
    n, nb, F = X.shape
    nk = w.shape[2]
    dX = torch.empty_like(X) if want_dX else None
    dw = torch.empty_like(w) if want_dw else None
    _call("vqa_patch_operator_bwd_f32", X.data_ptr(), w.data_ptr(), dZ.data_ptr(), _ptr(dX), _ptr(dw), n, nb, nk, F, _stream())
    return dX, dw


def gaussian_weights(pseudo: torch.Tensor, gauss: torch.Tensor) -> torch.Tensor:
    pseudo = _chk(pseudo, "pseudo").contiguous().view(-1, 2)
    nk = gauss.numel() // 4
    w = torch.empty((pseudo.shape[0], nk), device=pseudo.device, dtype=torch.float32)
    _call("vqa_gaussian_weights_f32", pseudo.data_ptr(), gauss.data_ptr(), w.data_ptr(), pseudo.shape[0], nk, _stream())
    return w


# ------------------------------------------------------------------------------------------- question encoder
_ERR_WORDS = {}


def device_error_word(device: torch.device) -> torch.Tensor:
    """One int32 per device that kernels OR error bits into (bit 0: a token id outside the embedding table)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    w = _ERR_WORDS.get(key)
    if w is None:
        w = _ERR_WORDS[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return w


def check_device_errors(device=None) -> None:
    """Read the error words back (a sync point: call it where the host waits anyway, e.g. with the loss / score read-back) and raise
    what the reference's modules would have raised at the offending call."""
    for key, w in list(_ERR_WORDS.items()):
        if device is not None and torch.device(device).index not in (None, key):
            continue
        e = int(w.item())
        if e:
            w.zero_()
            if e & 1:
                raise IndexError("vqa_b200: a question token id lies outside the embedding table (nn.Embedding raises 'index out of range' here)")
            raise RuntimeError(f"vqa_b200: device error word {e:#x}")


def embed_gather_split(question: torch.Tensor, wemb: torch.Tensor, T: int, with_lo: bool = True) -> SplitT:
    """question (B, >=T) int64, wemb (V, E) fp32 -> split planes of the time-major embeddings (T*B, E)."""
    if question.dtype != torch.int64 or not question.is_cuda or question.stride(1) != 1:
        raise RuntimeError("embed_gather_split: question must be a CUDA int64 tensor with contiguous rows (no CPU fallback)")
    wemb = _chk(wemb, "embedding weight").contiguous()
    B = question.shape[0]
    if T > question.shape[1]:
        raise RuntimeError(f"embed_gather_split: T={T} exceeds the question width {question.shape[1]}")
    out = empty_split(T * B, wemb.shape[1], wemb.device, with_lo)
    _call("vqa_embed_gather_split", question.data_ptr(), question.stride(0), wemb.data_ptr(), wemb.shape[0], wemb.shape[1],
          out.hi.data_ptr(), _ptr(out.lo), out.ld, B, T, device_error_word(wemb.device).data_ptr(), _stream())
    return out


def embed_scatter_add(dE: torch.Tensor, question: torch.Tensor, qlen: torch.Tensor, dW: torch.Tensor, T: int) -> torch.Tensor:
    dE, ldd = _rows_view(_chk(dE, "dE"), "dE")
    _chk(dW, "dW")
    _chk(qlen, "qlen", torch.int32)
    _call("vqa_embed_scatter_add_f32", dE.data_ptr(), ldd, question.data_ptr(), question.stride(0), qlen.data_ptr(), dW.data_ptr(),
          dW.shape[0], dW.shape[1], question.shape[0], T, _stream())
    return dW


def gru_cell_fwd(gi, gh, b_hh, h_prev, qlen, t, h_out, h_split: SplitT, gates):
    gi, ldgi = _rows_view(_chk(gi, "gi"), "gi")
    B, H = h_out.shape
    _call("vqa_gru_cell_fwd_f32", gi.data_ptr(), ldgi, _ptr(gh), b_hh.data_ptr(), _ptr(h_prev), qlen.data_ptr(), t, h_out.data_ptr(),
          h_split.hi.data_ptr(), _ptr(h_split.lo), h_split.ld, gates.data_ptr(), B, H, _stream())


def gru_step_fused(hprev: "SplitT", Whh_ub: "SplitT", gi_ub: torch.Tensor, b_hh_ub: torch.Tensor, h_prev: torch.Tensor, qlen: torch.Tensor,
                   t: int, h_out: torch.Tensor, hout: "SplitT", gates: torch.Tensor, tile_len: Optional[torch.Tensor] = None) -> None:
    """One GRU step as one kernel (product + cell).  ``Whh_ub`` / ``gi_ub`` / ``b_hh_ub`` in unit-block order (see the header)."""
    B, H = h_prev.shape
    gi_ub, ldgi = _rows_view(_chk(gi_ub, "gru gi"), "gru gi")
    _call("vqa_gru_step_fused", hprev.hi.data_ptr(), hprev.lo.data_ptr(), hprev.ld, Whh_ub.hi.data_ptr(), Whh_ub.lo.data_ptr(), Whh_ub.ld,
          gi_ub.data_ptr(), ldgi, b_hh_ub.data_ptr(), h_prev.data_ptr(), qlen.data_ptr(), int(t), h_out.data_ptr(), hout.hi.data_ptr(),
          hout.lo.data_ptr(), hout.ld, gates.data_ptr(), _ptr(tile_len), B, H, _stream())


_GRID_COUNTERS = {}


def gru_seq_supported(B: int, H: int) -> bool:
    """The whole-sequence kernel synchronises its CTAs with a grid barrier: all of them must fit on the 148 SMs at once."""
    return H % 32 == 0 and (H // 32) * ((B + 127) // 128) <= 148


def gru_seq_fused(Hs: "SplitT", Whh_ub: "SplitT", gi_ub: torch.Tensor, b_hh_ub: torch.Tensor, Hall: torch.Tensor, qlen: torch.Tensor,
                  gates: torch.Tensor, tile_len: Optional[torch.Tensor], T: int) -> None:
    """All T GRU steps in one cooperative launch.  Hs rows [0,B) and Hall[0] must hold h_{-1} = 0."""
    _, B, H = Hall.shape
    gi_ub, ldgi = _rows_view(_chk(gi_ub, "gru gi"), "gru gi")
    key = Hall.device.index if Hall.device.index is not None else torch.cuda.current_device()
    cnt = _GRID_COUNTERS.get(key)
    if cnt is None:
        cnt = _GRID_COUNTERS[key] = torch.zeros(4, device=Hall.device, dtype=torch.int32)
    _call("vqa_gru_seq_fused", Hs.hi.data_ptr(), Hs.lo.data_ptr(), Hs.ld, Whh_ub.hi.data_ptr(), Whh_ub.lo.data_ptr(), Whh_ub.ld,
          gi_ub.data_ptr(), ldgi, b_hh_ub.data_ptr(), Hall.data_ptr(), qlen.data_ptr(), gates.data_ptr(), _ptr(tile_len), cnt.data_ptr(),
          int(T), B, H, _stream())


def gru_cell_bwd(dh, gates, h_prev, qlen, t, dgi, dgh, dgi_s: SplitT, dgh_s: SplitT, dh_part):
    B, H = dh.shape
    _call("vqa_gru_cell_bwd_f32", dh.data_ptr(), gates.data_ptr(), _ptr(h_prev), qlen.data_ptr(), t, dgi.data_ptr(), dgh.data_ptr(),
          dgi_s.hi.data_ptr(), _ptr(dgi_s.lo), dgh_s.hi.data_ptr(), _ptr(dgh_s.lo), dgi_s.ld, dh_part.data_ptr(), B, H, _stream())


# ------------------------------------------------------------------------------------------- criterion / optimiser (train_step.cu)
_LOSS_SCRATCH = {}


_LOSS_SCRATCH_RETIRED = []      # outgrown buffers stay alive: a captured graph may still hold their addresses


def _loss_scratch(device: torch.device, blocks: int):
    """Partial sums + ticket of the loss kernel, one set per (device, stream) (see _colsum_counters)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ent = _LOSS_SCRATCH.get(key)
    if ent is None or ent[0].numel() < blocks:
        if ent is not None:
            _LOSS_SCRATCH_RETIRED.append(ent)
        ent = _LOSS_SCRATCH[key] = (torch.empty(max(blocks, 4096), device=device, dtype=torch.float32),
                                    torch.zeros(1, device=device, dtype=torch.int32))
    return ent


def mlsm_loss_fwd(logits: torch.Tensor, target: torch.Tensor, scale: float) -> torch.Tensor:
    """0-d loss = scale * sum -(y logsigmoid(x) + (1-y) logsigmoid(-x)) over contiguous (B, A) tensors, one launch."""
    _chk(logits, "mlsm_loss logits"); _chk(target, "mlsm_loss target")
    if logits.shape != target.shape or not logits.is_contiguous() or not target.is_contiguous():
        raise RuntimeError(f"mlsm_loss: logits {tuple(logits.shape)} and target {tuple(target.shape)} must be contiguous and of one shape")
    n = logits.numel()
    partial, counter = _loss_scratch(logits.device, _cabi.load().vqa_mlsm_loss_blocks(n))
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    _call("vqa_mlsm_loss_fwd_f32", logits.data_ptr(), target.data_ptr(), n, float(scale), partial.data_ptr(), counter.data_ptr(),
          loss.data_ptr(), _stream())
    return loss


def mlsm_loss_bwd(logits: torch.Tensor, target: torch.Tensor, grad_out: Optional[torch.Tensor], scale: float) -> torch.Tensor:
    """dlogits = (sigmoid(x) - y) * scale * grad_out (a device scalar; None == 1), one launch."""
    if grad_out is not None:
        _chk(grad_out, "mlsm_loss grad_out")
        if grad_out.numel() != 1:
            raise RuntimeError("mlsm_loss: grad_out must be a scalar")
    d = torch.empty_like(logits)
    _call("vqa_mlsm_loss_bwd_f32", logits.data_ptr(), target.data_ptr(), _ptr(grad_out), d.data_ptr(), logits.numel(), float(scale),
          _stream())
    return d


def adam_flat(chunks: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: torch.Tensor,
              beta1: float, beta2: float, eps: float, weight_decay: float, grad_scale: float, state: torch.Tensor) -> None:
    """One Adam step of every parameter named by the (nchunks, 3) int64 chunk table, one launch (see include/vqa_b200.h)."""
    _chk(chunks, "adam chunks", torch.int64); _chk(grad, "adam grad"); _chk(exp_avg, "adam exp_avg"); _chk(exp_avg_sq, "adam exp_avg_sq")
    _chk(lr, "adam lr"); _chk(state, "adam state", torch.int32)
    if chunks.dim() != 2 or chunks.shape[1] != 3 or not chunks.is_contiguous():
        raise RuntimeError("adam_flat: chunk table must be a contiguous (nchunks, 3) int64 tensor")
    if exp_avg.numel() != grad.numel() or exp_avg_sq.numel() != grad.numel() or state.numel() < 2:
        raise RuntimeError("adam_flat: exp_avg / exp_avg_sq must have the flat gradient buffer's size, state two ints")
    _call("vqa_adam_flat_f32", chunks.data_ptr(), chunks.shape[0], grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
          lr.data_ptr(), float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale), state.data_ptr(), _stream())


def _addr_array(addrs):
    import ctypes
    return (ctypes.c_longlong * len(addrs))(*[int(a) for a in addrs])


def p2p_barrier(flag_addrs, rank: int, world: int, epoch: torch.Tensor) -> None:
    """Stream-ordered flag barrier between the ranks of one node (see include/vqa_b200.h)."""
    _chk(epoch, "p2p epoch", torch.int32)
    _call("vqa_p2p_barrier", _addr_array(flag_addrs), int(rank), int(world), epoch.data_ptr(), _stream())


def adam_flat_p2p(grad: torch.Tensor, recv: torch.Tensor, n_own: int, chunk_log2: int, param_addrs, exp_avg: torch.Tensor,
                  exp_avg_sq: torch.Tensor, rank: int, world: int, lr: torch.Tensor, beta1: float, beta2: float, eps: float,
                  weight_decay: float, grad_scale: float, state: torch.Tensor, mc_param_addr: int = 0) -> None:
    """Sum of the pushed gradient chunks + Adam on the rank's elements + all-gather of the new parameters over NVLink peer memory, one
    launch (include/vqa_b200.h)."""
    _chk(grad, "adam grad"); _chk(recv, "adam recv"); _chk(exp_avg, "adam exp_avg"); _chk(exp_avg_sq, "adam exp_avg_sq")
    _chk(lr, "adam lr"); _chk(state, "adam state", torch.int32)
    if recv.numel() < world * n_own:
        raise RuntimeError("adam_flat_p2p: the receive buffer must hold world strides of n_own floats")
    _call("vqa_adam_flat_p2p", grad.data_ptr(), recv.data_ptr(), int(n_own), int(chunk_log2), _addr_array(param_addrs),
          int(mc_param_addr) or None, exp_avg.data_ptr(), exp_avg_sq.data_ptr(), int(rank), int(world), lr.data_ptr(), float(beta1), float(beta2), float(eps), float(weight_decay),
          float(grad_scale), state.data_ptr(), _stream())


def adam_flat_mc(mc_grad_addr: int, mc_param_addr: int, param: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lo: int, hi: int,
                 lr: torch.Tensor, beta1: float, beta2: float, eps: float, weight_decay: float, grad_scale: float, state: torch.Tensor) -> None:
    """In-switch sum of the ranks' gradients (multimem.ld_reduce) + Adam on flat elements [lo, hi) + multicast store of the new
    parameters (multimem.st), one launch (include/vqa_b200.h)."""
    _chk(param, "adam param"); _chk(exp_avg, "adam exp_avg"); _chk(exp_avg_sq, "adam exp_avg_sq"); _chk(lr, "adam lr"); _chk(state, "adam state", torch.int32)
    _call("vqa_adam_flat_mc", int(mc_grad_addr), int(mc_param_addr), param.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), int(lo), int(hi),
          lr.data_ptr(), float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale), state.data_ptr(), _stream())


def memcpy2d_async(dst_addr: int, dpitch: int, src_addr: int, spitch: int, width: int, height: int) -> None:
    """cudaMemcpy2DAsync device to device on the current stream (addresses / pitches / width in bytes)."""
    _call("vqa_memcpy2d_async", int(dst_addr), int(dpitch), int(src_addr), int(spitch), int(width), int(height), _stream())


# ------------------------------------------------------------------------------------------- batch assembly (loader.cu)
def gather_image(features: torch.Tensor, boxes: torch.Tensor, rows: torch.Tensor, err: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """image (B, K, D+4) fp32 = [features[rows] | boxes[rows]] from a (n, K, D) fp32/bf16 table and (n, K, 4) fp32 boxes.  ``out``: write
    into this contiguous (B, K, D+4) fp32 tensor (a training step's input slot) instead of a fresh one."""
    if features.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"gather_image: features must be fp32 or bf16, got {features.dtype}")
    _chk(features, "gather_image features", features.dtype); _chk(boxes, "gather_image boxes"); _chk(rows, "gather_image rows", torch.int64)
    _chk(err, "gather_image err", torch.int32)
    if features.dim() != 3 or boxes.shape != (features.shape[0], features.shape[1], 4) or not features.is_contiguous() or not boxes.is_contiguous():
        raise RuntimeError(f"gather_image: need contiguous features (n,K,D) and boxes (n,K,4), got {tuple(features.shape)} / {tuple(boxes.shape)}")
    n, K, D = features.shape
    rows = rows.contiguous().reshape(-1)
    B = rows.numel()
    if out is None:
        out = torch.empty((B, K, D + 4), device=features.device, dtype=torch.float32)
    elif tuple(out.shape) != (B, K, D + 4) or not out.is_contiguous():
        raise RuntimeError(f"gather_image: out must be a contiguous {(B, K, D + 4)} tensor, got {tuple(out.shape)}")
    else:
        _chk(out, "gather_image out")
    _call("vqa_gather_image_f32", features.data_ptr(), int(features.dtype == torch.bfloat16), boxes.data_ptr(), rows.data_ptr(), n,
          out.data_ptr(), B, K, D, err.data_ptr(), _stream())
    return out


def scatter_targets(ptr: torch.Tensor, ids: torch.Tensor, vals: torch.Tensor, B: int, A: int, err: torch.Tensor) -> torch.Tensor:
    """Dense (B, A) fp32 rows from CSR triplets (ptr int64 (B+1,), ids int32, vals fp32)."""
    _chk(ptr, "scatter_targets ptr", torch.int64); _chk(ids, "scatter_targets ids", torch.int32); _chk(vals, "scatter_targets vals")
    _chk(err, "scatter_targets err", torch.int32)
    if ptr.numel() != B + 1 or ids.numel() != vals.numel():
        raise RuntimeError("scatter_targets: ptr must have B+1 entries and ids / vals one entry per triplet")
    out = torch.empty((B, A), device=ptr.device, dtype=torch.float32)
    _call("vqa_scatter_targets_f32", ptr.data_ptr(), ids.data_ptr(), vals.data_ptr(), out.data_ptr(), B, A, err.data_ptr(), _stream())
    return out
