"""Training-step engine: the reference's inner loop body as one replayable CUDA graph.

The reference drivers do, per batch (``run.py:425-439``)::

    optimizer.zero_grad(); output, _, _ = model(q, v, K, qlen); loss = criterion(output, a); loss.backward(); optimizer.step()

Once the model's kernels take 4-5 ms per B=512 step, ~300 kernel launches issued from Python cost more host time than
the GPU needs.  ``TrainStep`` keeps the same sequence but captures it ONCE (forward, loss, backward, bucketed NCCL
all-reduce, Adam) into a CUDA graph over static input buffers and replays it per batch:

    step = TrainStep(model, optimizer, criterion, reducer=None)          # model.max_question_len should be set
    loss = step(question, image, K, qlen, target)                        # host (pinned) or device tensors; returns a 0-d device tensor

Nothing inside the captured region depends on host data: question lengths are a device tensor, dropout masks come
from a device-side step counter (``ops.set_graph_rng``), shapes are fixed per captured graph: a batch with other shapes
(the short last batch of an epoch) captures its own graph ONCE and every later batch of that shape replays it.  Two
input-buffer sets are captured (sharing one memory pool) so that the H2D copy of batch i+1, issued on a side stream by
``prefetch``, overlaps the replay of batch i.

The trajectory is the reference loop's: the eager warm-up iterations that precede a capture (they set kernel attributes
and size the allocator pool) run on a snapshot - parameters, optimiser state and the dropout step counter are restored
in place afterwards - so the first replay is the first update.  ``model.max_question_len`` (the number of recurrence
steps, fixed at capture time) must be set by the caller: inferring it from one batch would silently truncate longer
questions in later batches.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import kernels as kn
from . import ops


def _as_len_tensor(qlen, device) -> torch.Tensor:
    if torch.is_tensor(qlen):
        return qlen.reshape(-1).to(torch.int32)
    return torch.tensor([int(x) for x in qlen], dtype=torch.int32)


class TrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, criterion, reducer=None,
                 use_graph: bool = True, warmup: int = 3, seed: Optional[int] = None, slots: int = 2):
        self.model, self.opt, self.criterion, self.reducer = model, optimizer, criterion, reducer
        self.use_graph, self.warmup, self.nslots = use_graph, warmup, slots
        import os
        self.sm_reserve = int(os.environ.get("VQA_SM_RESERVE", "0"))          # SMs left to NCCL during backward (N > 1)
        self.device = next(model.parameters()).device
        self.seed = torch.initial_seed() if seed is None else seed
        self.rng_step = torch.zeros((), dtype=torch.int64, device=self.device)
        self.slots: List[Dict[str, torch.Tensor]] = []
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.losses: List[torch.Tensor] = []
        self.sig = None
        self._cache: Dict[tuple, tuple] = {}    # signature -> (slots, graphs, losses, launches): one capture per batch shape
        self._pool = None
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.filled = [False] * slots
        self.cur = 0
        self.launches_per_step = 0          # our kernels launched by one step (counted while capturing / running it)

    # ------------------------------------------------------------------------------------------------ step body
    def _body(self, b: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=False)
        logits, _, _ = self.model(b["question"], b["image"], b["K"], b["qlen"])
        loss = self.criterion(logits, b["target"])
        # while buckets travel under backward, the persistent one-CTA-per-SM kernels leave `sm_reserve` SMs to the collective: a
        # statically partitioned persistent grid that finds SMs taken finishes a whole wave late (vqa_set_sm_budget)
        reserve = self.sm_reserve if self.reducer is not None and self.reducer.world > 1 else 0
        if reserve:
            old = kn.set_sm_budget(148 - reserve)
        try:
            loss.backward()
            if self.reducer is not None:
                self.reducer.finish()
        finally:
            if reserve:
                kn.set_sm_budget(old)
        self.opt.step()
        self.rng_step += 1
        return loss

    def _signature(self, question, image, target):
        return (tuple(question.shape), tuple(image.shape), tuple(target.shape))

    # ------------------------------------------------------------------------------------------------ warm-up on a snapshot
    def _snapshot(self):
        params = [p.detach().clone() for p in self.model.parameters()]
        state = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in self.opt.state.items()}
        return params, state, self.rng_step.clone()

    def _restore(self, snap) -> None:
        """In place (captured graphs and FlatAdam's chunk table hold these addresses).  Optimiser state created by the warm-up
        itself (torch.optim.Adam allocates lazily) is reset to its initial value, zero."""
        params, state, rng = snap
        with torch.no_grad():
            for p, old in zip(self.model.parameters(), params):
                p.copy_(old)
            for p, st in self.opt.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        old = state.get(p, {}).get(k)
                        v.copy_(old) if old is not None else v.zero_()
            self.rng_step.copy_(rng)

    def _check_lengths(self, qlen) -> None:
        """Host-side lengths (the drivers' list / a CPU tensor) are checked for free; device-side lengths cannot be without a sync."""
        if torch.is_tensor(qlen) and qlen.is_cuda:
            return
        longest = int(qlen.max()) if torch.is_tensor(qlen) else max(int(x) for x in qlen)
        if longest > int(self.model.max_question_len):
            raise ValueError(f"TrainStep: a question of {longest} tokens exceeds model.max_question_len={self.model.max_question_len} "
                             "(the captured recurrence has that many steps)")

    def _build(self, question, image, K, qlen, target) -> None:
        dev = self.device
        qlen_t = _as_len_tensor(qlen, dev)
        if not getattr(self.model, "max_question_len", 0):
            raise ValueError("TrainStep: set model.max_question_len to the dataset's longest question (14 for VQA2, torch_dataset.py:425) "
                             "before the first step: the number of recurrence steps is fixed when the step is captured")
        self.sig = self._signature(question, image, target)
        if self.sig in self._cache:                                  # this batch shape has its graphs already
            self.slots, self.graphs, self.losses, self.launches_per_step = self._cache[self.sig]
            self.filled = [False] * self.nslots
            for s in range(self.nslots):
                self.consumed[s].record()
            return
        self.slots = [dict(question=torch.zeros(question.shape, dtype=torch.int64, device=dev),
                           image=torch.zeros(image.shape, dtype=torch.float32, device=dev),
                           K=torch.zeros(K.shape, dtype=K.dtype, device=dev),
                           qlen=torch.ones(qlen_t.shape, dtype=torch.int32, device=dev),
                           target=torch.zeros(target.shape, dtype=torch.float32, device=dev)) for _ in range(self.nslots)]
        self.graphs, self.losses = [], []
        self.filled = [False] * self.nslots
        for s in range(self.nslots):
            self.consumed[s].record()
        if not self.use_graph:
            self._cache[self.sig] = (self.slots, self.graphs, self.losses, 0)
            return
        ops.set_graph_rng(self.seed, self.rng_step)
        self._fill(0, question, image, K, qlen_t, target, torch.cuda.current_stream())
        # eager warm-up on a side stream (sets kernel attributes, sizes the allocator pool) before capture - on a snapshot: these
        # iterations are not part of the training trajectory
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._body(self.slots[0])
        torch.cuda.current_stream().wait_stream(side)
        self._restore(snap)
        torch.cuda.synchronize()
        pool = self._pool
        for s in range(self.nslots):
            g = torch.cuda.CUDAGraph()
            n0 = kn.LAUNCHES
            with torch.cuda.graph(g, pool=pool):
                loss = self._body(self.slots[s])
            self.launches_per_step = kn.LAUNCHES - n0
            pool = self._pool = g.pool()
            self.graphs.append(g)
            self.losses.append(loss)
        self._cache[self.sig] = (self.slots, self.graphs, self.losses, self.launches_per_step)

    def _fill(self, s, question, image, K, qlen_t, target, stream) -> None:
        sl = self.slots[s]
        with torch.cuda.stream(stream):
            stream.wait_event(self.consumed[s])
            sl["question"].copy_(question, non_blocking=True)
            if image.data_ptr() != sl["image"].data_ptr():               # a producer that wrote into input_slot("image") has left it in place
                sl["image"].copy_(image, non_blocking=True)
            sl["K"].copy_(K, non_blocking=True)
            sl["qlen"].copy_(qlen_t, non_blocking=True)
            sl["target"].copy_(target, non_blocking=True)
            self.ready[s].record(stream)
        self.filled[s] = True

    # ------------------------------------------------------------------------------------------------ public API
    def input_slot(self, name: str = "image") -> Optional[torch.Tensor]:
        """The idle buffer set's tensor ``name`` - what the NEXT step will read - or None before the first step.  The copy stream is
        made to wait for the replay that last read it, so a producer launched on ``copy_stream`` (``shards.ShardLoader.assemble(...,
        image_out=...)``) can write the next batch straight into it; ``prefetch`` / ``__call__`` then find the data in place and skip
        that copy (151 MB device-to-device per step at the VQA2 shapes)."""
        if not self.slots:
            return None
        s = (self.cur + 1) % self.nslots
        self.copy_stream.wait_event(self.consumed[s])
        return self.slots[s][name]

    def prefetch(self, question, image, K, qlen, target) -> None:
        """Start copying the NEXT batch into the idle buffer set on the copy stream (overlaps the running step)."""
        if self.sig != self._signature(question, image, target):
            return                                                   # first batch / new shapes: handled by __call__
        self._check_lengths(qlen)
        s = (self.cur + 1) % self.nslots
        self._fill(s, question, image, K, _as_len_tensor(qlen, self.device), target, self.copy_stream)

    def __call__(self, question, image, K, qlen, target) -> torch.Tensor:
        if self.sig != self._signature(question, image, target):
            self._build(question, image, K, qlen, target)
            self.cur = self.nslots - 1
        s = (self.cur + 1) % self.nslots
        if not self.filled[s]:
            self._check_lengths(qlen)
            self._fill(s, question, image, K, _as_len_tensor(qlen, self.device), target, self.copy_stream)
        torch.cuda.current_stream().wait_event(self.ready[s])
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()                                        # a scheduler's new rate reaches the captured Adam launch
        if self.use_graph:
            self.graphs[s].replay()
            loss = self.losses[s]
        else:
            n0 = kn.LAUNCHES
            loss = self._body(self.slots[s])
            self.launches_per_step = kn.LAUNCHES - n0
        self.consumed[s].record()
        self.filled[s] = False
        self.cur = s
        return loss

    def close(self) -> None:
        ops.set_graph_rng()
