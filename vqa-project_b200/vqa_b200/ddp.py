"""Data-parallel training plumbing: one process per GPU, gradients all-reduced over NCCL (NVLink/NVSwitch),
bucketed and overlapped with the rest of backward.

The reference has no distributed code (single process, ``nn.DataParallel`` only as a comment).
The model path shards over the batch with no forward exchange (every question/image is independent), so
the only collective is the gradient all-reduce.

``GradReducer`` owns one flat fp32 gradient buffer; every ``param.grad`` is a view into it.  The buffer is laid out
in registration order (so the nk per-kernel conv weights stay consecutive and one GEMM can write all their gradients);
buckets are contiguous ranges taken from its END backwards (= the order autograd finishes them: classifier first,
GRU and embedding last).  Gradients are WRITTEN, not accumulated: ``zero_grad`` sets every ``.grad`` to None (no memset),
the operators of ``vqa_b200.ops`` ask ``sink(param)`` for the parameter's view and let their last kernel write straight
into it, and autograd adopts that view as ``.grad`` without a copy; a gradient that arrives any other way is copied
into its view by the hook.  A post-accumulate-grad hook counts finished parameters per bucket and launches
``all_reduce(bucket, async_op=True)`` the moment a bucket is complete, so the transfer of the 36 MB ``out_2``
bucket runs under the graph-convolution backward kernels.  ``finish()`` waits and leaves averaged gradients.
Works with any backend (``gloo`` in the CPU tests, ``nccl`` on the B200s).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def symmetric_empty(numel: int, dtype: torch.dtype, device: torch.device, group=None):
    """A zero-filled buffer every rank of the node can address: (local tensor, [address of rank q's buffer in THIS process], handle).
    torch's symmetric-memory allocator (CUDA VMM handles exchanged at rendezvous) maps the peers' allocations over NVLink."""
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(numel, dtype=dtype, device=device)
    t.zero_()
    hdl = symm.rendezvous(t, dist.group.WORLD if group is None else group)
    return t, [int(a) for a in hdl.buffer_ptrs], hdl


class GradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, early: Optional[bool] = None, p2p: Optional[bool] = None):
        """``early``: let the fused operators start a bucket's all-reduce from INSIDE their backward, the moment the bucket's last
        gradient kernel is enqueued (``mark_ready``), instead of when the operator's autograd node returns.  Default: the
        ``VQA_EARLY_READY`` environment variable ("1" unless set to "0")."""
        import os
        self.early = (os.environ.get("VQA_EARLY_READY", "1") != "0") if early is None else bool(early)
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradReducer: no trainable parameters")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        import os
        if os.environ.get("VQA_NO_EXCHANGE") == "1":      # measurement aid: N independent replicas (what the box does without any exchange)
            self.world = 1
        dev, dt = self.params[0].device, self.params[0].dtype
        # pass 1: offsets in registration order (64-element alignment keeps views 256-byte aligned)
        off = 0
        for p in self.params:
            off = (off + 63) // 64 * 64
            p._vqa_flat_off = off
            off += p.numel()
        total = off
        # pass 2: buckets = contiguous ranges from the end backwards (completion order of backward)
        self.bucket_of = [0] * len(self.params)
        self.bucket_slices: List[slice] = []
        self.bucket_size: List[int] = []
        end, nbytes, count = total, 0, 0
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            pbytes = p.numel() * p.element_size()
            if count and pbytes >= bucket_bytes // 2:     # a large tensor never drags small, later-finishing neighbours into its bucket
                self.bucket_slices.append(slice(self.params[i + 1]._vqa_flat_off, end))
                self.bucket_size.append(count)
                end, nbytes, count = self.params[i + 1]._vqa_flat_off, 0, 0
            self.bucket_of[i] = len(self.bucket_slices)
            nbytes += pbytes
            count += 1
            if nbytes >= bucket_bytes or i == 0:
                self.bucket_slices.append(slice(p._vqa_flat_off, end))
                self.bucket_size.append(count)
                end, nbytes, count = p._vqa_flat_off, 0, 0
        # pass 3: one buffer, every .grad a view into it.  p2p (world > 1 on one NVLink node, CUDA): the buffer is symmetric memory, the
        # peers read it directly (optim.FlatAdam's fused reduce-scatter + Adam + all-gather kernel) and NO all-reduce is issued here.
        import os
        want = (os.environ.get("VQA_P2P", "1") != "0") if p2p is None else bool(p2p)
        self.p2p, self.peer_grad_addrs, self._symm = False, None, []
        self.total = total = (total + 3) // 4 * 4
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        if want and self.world > 1 and dev.type == "cuda":
            try:
                # every rank owns a contiguous 1/world slice of the flat index space; `recv` holds, per peer, that peer's gradients
                # for MY slice (pushed there by the peer, see _launch)
                # interleaved ownership: chunk c of CH floats belongs to rank c % world, a "row" = world consecutive chunks, the buffers
                # are padded to whole rows.  `recv` holds, per peer, that peer's gradients for MY chunks (pushed there, see _push).
                # 4 MB chunks: one plain cudaMemcpyAsync per chunk (copy engines; strided cudaMemcpy2DAsync pushes of finer chunks were
                # executed by SM copy kernels - 136 us of device time per step at N = 2, profiles/r02_p2p_adam.md)
                self.chunk_log2 = 20                                  # 2^20 floats = 4 MB per chunk
                CH = 1 << self.chunk_log2
                row = CH * self.world
                self.nrows = (total + row - 1) // row
                self.total = total = self.nrows * row
                self.n_own = self.nrows * CH
                self.flat = torch.zeros(total, device=dev, dtype=dt)
                self.recv, self.peer_recv_addrs, h = symmetric_empty(self.world * self.n_own, dt, dev, process_group)
                self._peer_recv = [None if q == self.rank else h.get_buffer(q, (self.world * self.n_own,), dt, 0) for q in range(self.world)]
                self._symm.append(h)
                # a chunk may be pushed once every bucket it touches is complete
                starts = [p._vqa_flat_off for p in self.params]
                self._chunk_buckets = []
                for c in range(self.nrows * self.world):
                    a, e = c * CH, (c + 1) * CH
                    self._chunk_buckets.append({self.bucket_of[i] for i in range(len(self.params)) if starts[i] < e and starts[i] + self.params[i].numel() > a})
                self._chunk_pushed = [False] * (self.nrows * self.world)
                self._bucket_ready = [False] * len(self.bucket_size)
                self.comm_stream = torch.cuda.Stream(device=dev)
                self.p2p = True
            except Exception as e:                       # no NVLink peer access / allocator unavailable: the NCCL path below
                import warnings
                warnings.warn(f"vqa_b200.ddp: symmetric memory unavailable ({e!r}); gradients will be all-reduced with NCCL")
        if not self.p2p:
            self.flat = torch.zeros(total, device=dev, dtype=dt)
        self._by_ptr = {p.data_ptr(): p for p in self.params}
        for p in self.params:
            p.grad = self._view(p)
        self._ready = [0] * len(self.bucket_size)
        self._done = [False] * len(self.params)           # counted towards its bucket in this step (by mark_ready or by the hook)
        self._issued = [False] * len(self.params)         # sink() handed this parameter's view out in this step (at most once)
        self._accumulating = False                        # zero_grad(set_to_none=False): several backwards per step, reduce in finish()
        self._index = {p.data_ptr(): i for i, p in enumerate(self.params)}
        self._handles = []
        self._launched_now = [False] * len(self.bucket_size)
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.launched = 0
        self.average = True                               # finish() scales by 1/world; optim.FlatAdam folds the factor into its pass instead
        try:                                              # the fused operators write their parameter gradients through the sink
            from . import ops
            ops.set_grad_sink(self.sink, self.mark_ready if self.early else None)
        except Exception:                                 # pragma: no cover - CPU-only use of the reducer (tests)
            pass

    def _view(self, p: torch.nn.Parameter) -> torch.Tensor:
        """A fresh view of the flat buffer shaped like ``p`` (a new tensor object each call, so autograd can adopt it)."""
        return self.flat[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)

    def sink(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        """Destination for the gradient of the parameter whose storage ``t`` starts at (None if it is not one of ours).
        Handed out AT MOST ONCE per parameter and step: when a parameter feeds two autograd nodes (the model called twice
        before one backward), the second node gets None, computes into a tensor of its own, and autograd adds the two -
        two aliasing views of the same memory would otherwise be overwritten by the later kernel and summed to 2 * g_last."""
        i = self._index.get(t.data_ptr())
        if i is None:
            return None
        p = self.params[i]
        if p.shape != t.shape or p.grad is not None or self._issued[i]:      # existing .grad: accumulation is autograd's job
            return None
        self._issued[i] = True
        return self._view(p)

    def _count(self, i: int) -> None:
        if self._done[i]:
            return
        self._done[i] = True
        b = self.bucket_of[i]
        self._ready[b] += 1
        if self._ready[b] == self.bucket_size[b] and not self._accumulating:
            self._launch(b)

    def mark_ready(self, t: torch.Tensor) -> None:
        """The operator that wrote this parameter's gradient into its sink view says so the moment its last kernel is enqueued.
        The fused operators are single autograd nodes: without this, the hooks of ALL their parameters would fire together when
        the node returns, and the classifier's 36 MB bucket could not start its all-reduce under the rest of backward."""
        i = self._index.get(t.data_ptr())
        if i is not None and self.params[i].grad is None and not self._accumulating:  # (.grad exists: the hook does the counting)
            self._count(i)

    def _make_hook(self, i: int):
        def hook(param):
            g = param.grad
            if g is not None and g.data_ptr() != self.flat.data_ptr() + param._vqa_flat_off * self.flat.element_size():
                if self._done[i]:
                    raise RuntimeError("GradReducer: a gradient marked ready through the sink was replaced by another tensor")
                v = self._view(param)                     # produced outside the sink: move it into the flat buffer
                v.copy_(g)
                param.grad = v
            self._count(i)
        return hook

    def _push(self, b: int) -> None:
        """p2p: bucket b is complete - copy every chunk that has now all its buckets complete (and that another rank owns) into the
        owner's receive buffer: one cudaMemcpyAsync over NVLink per chunk (copy engines, no SMs taken from backward) on a side stream,
        ordered after everything enqueued so far on the current stream."""
        import os
        self._bucket_ready[b] = True
        W, CH = self.world, 1 << self.chunk_log2
        todo = [c for c in range(self.nrows * W) if not self._chunk_pushed[c] and all(self._bucket_ready[x] for x in self._chunk_buckets[c])]
        for c in todo:
            self._chunk_pushed[c] = True
        todo = [c for c in todo if c % W != self.rank]
        if not todo or "nopush" in os.environ.get("VQA_P2P_DIAG", ""):       # (timing diagnosis only: wrong results)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            for c in sorted(todo, key=lambda c: ((c % W - self.rank) % W, c)):     # staggered: the ranks start with different peers
                r, k = c % W, c // W
                o = self.rank * self.n_own + k * CH
                self._peer_recv[r][o:o + CH].copy_(self.flat[c * CH:(c + 1) * CH], non_blocking=True)

    def _launch(self, b: int) -> None:
        self._launched_now[b] = True
        if self.p2p:
            self._push(b)
            self.launched += 1
            return
        if self.world > 1 and not self.p2p:
            h = dist.all_reduce(self.flat[self.bucket_slices[b]], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append(h)
        self.launched += 1

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Start a step.  Default: drop every ``.grad`` (the step WRITES gradients into the flat buffer, nothing to zero).
        ``set_to_none=False``: zero the flat buffer in one memset and (re)attach the views, for accumulation over micro-batches:
        any number of backwards may follow, autograd adds into the views, and NO bucket is reduced before ``finish()`` (the
        equivalent of DDP's ``no_sync`` for all but the last micro-batch - reducing after the first backward would race the
        later additions and leave them un-reduced)."""
        self._accumulating = not set_to_none
        if set_to_none:
            for p in self.params:
                p.grad = None
        else:
            self.flat.zero_()
            for p in self.params:
                if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + p._vqa_flat_off * self.flat.element_size():
                    p.grad = self._view(p)
        self._ready = [0] * len(self.bucket_size)
        self._done = [False] * len(self.params)
        self._issued = [False] * len(self.params)
        self._launched_now = [False] * len(self.bucket_size)
        if self.p2p:
            self._chunk_pushed = [False] * (self.nrows * self.world)
            self._bucket_ready = [False] * len(self.bucket_size)

    def finish(self) -> None:
        """Wait for the in-flight buckets (launching any that has not started: parameters without a gradient, or accumulation
        mode) and average (unless ``average`` was cleared by an optimiser that applies 1/world itself: the buffer then holds the
        SUM over ranks)."""
        for b in range(len(self.bucket_size)):
            if not self._launched_now[b]:
                self._ready[b] = self.bucket_size[b]
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        import os
        if self.p2p and "nopush" not in os.environ.get("VQA_P2P_DIAG", ""):
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)   # the pushes precede the optimiser's barrier
        if self.world > 1 and self.average and not self.p2p:
            self.flat.mul_(1.0 / self.world)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
        try:
            from . import ops
            if ops._GRAD_SINK == self.sink:
                ops.set_grad_sink(None, None)
        except Exception:                                 # pragma: no cover
            pass


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
