"""Adam over the flat gradient buffer: every parameter tensor updated by ONE kernel launch.

The reference drivers use ``torch.optim.Adam(model.parameters(), lr=args.lr)`` (``run.py:392``) stepped once per batch
(``run.py:435``) and ``MultiStepLR`` (``run.py:393``).  ``FlatAdam`` keeps that interface (``param_groups`` with ``lr`` / ``betas`` /
``eps`` / ``weight_decay``, ``step()``, ``zero_grad()``, ``state_dict()`` in ``torch.optim.Adam``'s layout) and the same update rule
(``amsgrad=False``), but reads the gradients where ``vqa_b200.ddp.GradReducer`` keeps them - one flat fp32 buffer - and holds
``exp_avg`` / ``exp_avg_sq`` in two buffers of the same layout, so one launch of ``adam_flat_kernel`` (``csrc/train_step.cu``),
driven by a device-side chunk table, replaces the multi-tensor launches.  With ``world > 1`` the reducer is told not to
average (``average=False``): the 1/world factor is folded into this pass and the separate sweep over the gradient buffer
disappears.  The step count and the learning rate live on the device, so the step is CUDA-graph capturable; ``sync_lr()``
(called by ``TrainStep`` before every replay) pushes a scheduler's new rate.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as kn
from .ddp import GradReducer

CHUNK = 4096          # elements per table entry: 16 KB of each of p / g / m / v, four float4 per thread


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, reducer: GradReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError(f"FlatAdam: invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(reducer.params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.reducer = reducer
        flat = reducer.flat
        if not flat.is_cuda:
            raise RuntimeError("FlatAdam: the gradient buffer is not on a CUDA device (the vqa_b200 path has no CPU fallback)")
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self._state = torch.zeros(2, device=flat.device, dtype=torch.int32)          # {steps taken, block ticket}
        self._lr = torch.full((1,), float(lr), device=flat.device, dtype=torch.float32)
        self._lr_host = float(lr)
        self._build_chunks()
        reducer.average = False                                                       # 1/world is applied by the Adam pass
        self.p2p = bool(getattr(reducer, "p2p", False))
        if self.p2p:
            self._setup_p2p()
        for p in reducer.params:                                                      # torch.optim.Adam-style per-parameter views
            st = self.state[p]
            st["exp_avg"] = self.exp_avg[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)
            st["exp_avg_sq"] = self.exp_avg_sq[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)
            st["step"] = self._state[0]

    def _build_chunks(self) -> None:
        """The device table {parameter address, flat offset, count} in CHUNK-element pieces, from the parameters' CURRENT storage."""
        rows = []
        for p in self.reducer.params:
            if not p.is_contiguous() or p.dtype != torch.float32:
                raise RuntimeError("FlatAdam: parameters must be contiguous fp32 tensors")
            n, off, addr = p.numel(), p._vqa_flat_off, p.data_ptr()
            for s in range(0, n, CHUNK):
                rows.append((addr + 4 * s, off + s, min(CHUNK, n - s)))
        self._ptrs = [p.data_ptr() for p in self.reducer.params]
        self.chunks = torch.tensor(rows, dtype=torch.int64).to(self.reducer.flat.device)

    def _setup_p2p(self) -> None:
        """Data-parallel mode over NVLink peer memory (ddp.GradReducer(p2p=True)): the parameters move into ONE symmetric flat buffer
        with the gradient buffer's layout (every ``p.data`` becomes a view of it, values kept), so that the fused kernel
        (``kernels.adam_flat_p2p``) can write every rank's parameters.  Each rank owns every world-th 256 KB chunk of the flat index
        space: the peers push their gradients for those chunks into its receive buffer while backward runs (``ddp.GradReducer._push``),
        it sums them, updates its chunks (its moments are the only ones that ever change: optimizer state is sharded) and stores the
        new values into all ranks' parameter buffers."""
        import torch.distributed as dist
        from .ddp import symmetric_empty
        r = self.reducer
        dev = r.flat.device
        import os
        self.pflat, self.peer_param_addrs, h = symmetric_empty(r.total, torch.float32, dev, r.group)
        # NVLS (opt-in, VQA_P2P_MULTICAST=1): with the switch's multicast address of the buffer the all-gather is one multimem.st per
        # element - measured SLOWER than the plain peer stores at N = 2 (199 vs 116 us for the kernel, 4.04 vs 3.95 ms/step)
        self.mc_param_addr = int(getattr(h, "multicast_ptr", 0) or 0) if os.environ.get("VQA_P2P_MULTICAST", "0") == "1" else 0
        self.flags, self.peer_flag_addrs, h2 = symmetric_empty(64, torch.int32, dev, r.group)
        r._symm += [h, h2]
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        with torch.no_grad():
            for p in r.params:
                v = self.pflat[p._vqa_flat_off:p._vqa_flat_off + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
        torch.cuda.synchronize(dev)
        dist.barrier(group=r.group)                   # every rank's flags are zero and its parameters in place before the first kernel barrier

    def _check_p2p_views(self) -> None:
        base = self.pflat.data_ptr()
        for p in self.reducer.params:
            if p.data_ptr() != base + 4 * p._vqa_flat_off:
                raise RuntimeError("FlatAdam (p2p): a parameter no longer lives in the shared flat buffer (was the module moved with .to() / "
                                   "re-flattened after the optimiser was built?)")

    def gather_state(self) -> None:
        """Make every rank's exp_avg / exp_avg_sq complete (each rank only ever updates its own chunks): call before ``state_dict()``."""
        if not self.p2p:
            return
        import torch.distributed as dist
        r = self.reducer
        CH = 1 << r.chunk_log2
        for buf in (self.exp_avg, self.exp_avg_sq):
            rows = buf.view(r.nrows, r.world, CH)
            for q in range(r.world):
                mine = rows[:, q, :].contiguous()
                dist.broadcast(mine, src=dist.get_global_rank(r.group, q) if r.group is not None else q, group=r.group)
                rows[:, q, :].copy_(mine)

    @property
    def steps_taken(self) -> int:
        return int(self._state[0].item())

    def sync_lr(self) -> None:
        """Copy ``param_groups[0]['lr']`` to the device if a scheduler changed it (call outside graph capture)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if len(self.param_groups) != 1:
            raise RuntimeError("FlatAdam: one parameter group (the reducer's parameters) is supported")
        g = self.param_groups[0]
        if self.p2p:
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                self.sync_lr()
                self._check_p2p_views()
            missing = [i for i, p in enumerate(self.reducer.params) if p.grad is None]
            if missing:
                raise RuntimeError(f"FlatAdam.step: parameters {missing} (reducer order) received no gradient in this step")
            r = self.reducer
            import os
            diag = os.environ.get("VQA_P2P_DIAG", "")             # timing diagnosis only (wrong results)
            if "nobarrier" in diag or "nokernel" in diag:
                if "nokernel" not in diag:
                    kn.adam_flat_p2p(r.flat, r.recv, r.n_own, r.chunk_log2, self.peer_param_addrs, self.exp_avg, self.exp_avg_sq, r.rank, r.world,
                                     self._lr, g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], 1.0 / r.world, self._state)
                return loss
            kn.p2p_barrier(self.peer_flag_addrs, r.rank, r.world, self.epoch)      # every rank's gradients are written
            kn.adam_flat_p2p(r.flat, r.recv, r.n_own, r.chunk_log2, self.peer_param_addrs, self.exp_avg, self.exp_avg_sq, r.rank, r.world,
                             self._lr, g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], 1.0 / r.world, self._state, self.mc_param_addr)
            kn.p2p_barrier(self.peer_flag_addrs, r.rank, r.world, self.epoch)      # every rank's parameters are complete
            return loss
        moved = [p.data_ptr() for p in self.reducer.params] != self._ptrs
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
            if moved:                                 # e.g. a module re-joined its weights into one buffer on its first forward
                self._build_chunks()
        elif moved:
            raise RuntimeError("FlatAdam: parameter storage moved since the last eager step; run one step before capturing")
        missing = [i for i, p in enumerate(self.reducer.params) if p.grad is None]
        if missing:                                   # the flat buffer would hold a previous step's values for them
            raise RuntimeError(f"FlatAdam.step: parameters {missing} (reducer order) received no gradient in this step")
        kn.adam_flat(self.chunks, self.reducer.flat, self.exp_avg, self.exp_avg_sq, self._lr, g["betas"][0], g["betas"][1], g["eps"],
                     g["weight_decay"], 1.0 / self.reducer.world, self._state)
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.reducer.zero_grad(set_to_none=set_to_none)

    def load_state_dict(self, state_dict) -> None:
        """Accepts ``torch.optim.Adam.state_dict()`` (and this class's own): moments are copied INTO the flat buffers, the step
        count (one value for all parameters here) is taken from the first entry."""
        groups, state = state_dict["param_groups"], state_dict["state"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.reducer.params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group over the reducer's parameters")
        g = self.param_groups[0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                g[k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        if groups[0].get("amsgrad") or groups[0].get("maximize"):
            raise ValueError("FlatAdam.load_state_dict: amsgrad / maximize are not supported")
        steps = None
        with torch.no_grad():
            for idx, p in zip(groups[0]["params"], self.reducer.params):
                st = state.get(idx)
                if st is None:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                s = int(st["step"])
                if steps is not None and s != steps:
                    raise ValueError("FlatAdam.load_state_dict: parameters with different step counts")
                steps = s
            if steps is not None:
                self._state[0] = steps
        self.sync_lr()
