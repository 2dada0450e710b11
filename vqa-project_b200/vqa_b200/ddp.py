"""Data-parallel training plumbing: one process per GPU, gradients all-reduced over NCCL (NVLink/NVSwitch),
bucketed and overlapped with the rest of backward.

The reference has no distributed code (single process, ``nn.DataParallel`` only as a comment).
The model path shards over the batch with no forward exchange (every question/image is independent), so
the only collective is the gradient all-reduce.

``GradReducer`` owns one flat fp32 gradient buffer; every ``param.grad`` is a view into it.  Parameters are
assigned to buckets in REVERSE registration order (= the order autograd finishes them: classifier first, GRU
and embedding last).  A post-accumulate-grad hook counts finished parameters per bucket and launches
``all_reduce(bucket, async_op=True)`` the moment a bucket is complete, so the transfer of the 36 MB ``out_2``
bucket runs under the graph-convolution backward kernels.  ``finish()`` waits and leaves averaged gradients.
Works with any backend (``gloo`` in the CPU tests, ``nccl`` on the B200s).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20,
                 process_group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradReducer: no trainable parameters")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        dev, dt = self.params[0].device, self.params[0].dtype
        # pass 1: offsets.  Reverse order = completion order of backward; 64-element alignment keeps views 256-byte aligned
        self.bucket_of = [0] * len(self.params)
        self.bucket_slices: List[slice] = []
        self.bucket_size: List[int] = []
        off = start = nbytes = count = 0
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            off = (off + 63) // 64 * 64
            p._vqa_flat_off = off
            self.bucket_of[i] = len(self.bucket_slices)
            off += p.numel()
            nbytes += p.numel() * p.element_size()
            count += 1
            if nbytes >= bucket_bytes:
                self.bucket_slices.append(slice(start, off))
                self.bucket_size.append(count)
                start, nbytes, count = off, 0, 0
        if count:
            self.bucket_slices.append(slice(start, off))
            self.bucket_size.append(count)
        # pass 2: one buffer, every .grad a view into it
        self.flat = torch.zeros(off, device=dev, dtype=dt)
        for p in self.params:
            p.grad = self.flat[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)
        self._ready = [0] * len(self.bucket_size)
        self._handles = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.launched = 0

    def _make_hook(self, i: int):
        b = self.bucket_of[i]

        def hook(param):
            self._ready[b] += 1
            if self._ready[b] == self.bucket_size[b]:
                self._launch(b)
        return hook

    def _launch(self, b: int) -> None:
        if self.world > 1:
            h = dist.all_reduce(self.flat[self.bucket_slices[b]], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append(h)
        self.launched += 1

    def zero_grad(self) -> None:
        """Zero the flat buffer in one memset and (re)attach the views (``optimizer.zero_grad()`` would detach them)."""
        self.flat.zero_()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + p._vqa_flat_off * self.flat.element_size():
                p.grad = self.flat[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)
        self._ready = [0] * len(self.bucket_size)

    def finish(self) -> None:
        """Wait for the in-flight buckets (launching any whose parameters received no gradient) and average."""
        for b in range(len(self.bucket_size)):
            if self._ready[b] != self.bucket_size[b]:
                self._ready[b] = self.bucket_size[b]
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        if self.world > 1:
            self.flat.mul_(1.0 / self.world)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
