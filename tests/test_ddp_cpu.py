"""world_size-2 gloo test of the bucketed gradient reducer (host logic of the multi-GPU path, runs on CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer, broadcast_parameters
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # deliberately different init per rank
    net = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.ReLU(), torch.nn.Linear(33, 7), torch.nn.Linear(7, 3))
    broadcast_parameters(net)
    red = GradReducer(net.parameters(), bucket_bytes=1024)      # tiny buckets -> several collectives
    assert len(red.bucket_size) > 1
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    torch.manual_seed(7)
    x_all = torch.randn(2 * world, 20)
    y_all = torch.randn(2 * world, 3)
    for step in range(3):
        red.zero_grad()
        xs, ys = x_all[rank * 2:(rank + 1) * 2], y_all[rank * 2:(rank + 1) * 2]
        ((net(xs) - ys) ** 2).mean().backward()
        red.finish()
        opt.step()
    # single-process reference: same init (rank 0's), full batch
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.ReLU(), torch.nn.Linear(33, 7), torch.nn.Linear(7, 3))
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    for step in range(3):
        ropt.zero_grad()
        ((ref(x_all) - y_all) ** 2).mean().backward()
        ropt.step()
    err = max((a - b).abs().max().item() for a, b in zip(net.parameters(), ref.parameters()))
    q.put((rank, err, red.launched))
    dist.destroy_process_group()


def test_two_rank_training_equals_single_process_on_the_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, launched in res:
        assert err < 1e-6, (rank, err)
        assert launched >= 6          # >= 2 buckets x 3 steps, launched from the backward hooks


def test_reducer_single_process_keeps_views_and_zeroes():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    red = GradReducer(net.parameters())
    w = net[0].weight

    def in_flat(p):
        return p.grad.data_ptr() == red.flat.data_ptr() + p._vqa_flat_off * 4

    # registration order inside the flat buffer (consecutive conv weights stay consecutive), buckets cover it exactly once
    offs = [p._vqa_flat_off for p in net.parameters()]
    assert offs == sorted(offs)
    covered = sorted((s.start, s.stop) for s in red.bucket_slices)
    assert covered[0][0] == 0 and 0 <= red.flat.numel() - covered[-1][1] < 4      # (the buffer is padded to whole float4s)
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    net(torch.randn(3, 5)).sum().backward()               # .grad defined: autograd accumulates in place
    red.finish()
    assert all(in_flat(p) for p in net.parameters()) and red.flat.abs().sum() > 0
    # default step start: gradients dropped, nothing zeroed; a gradient produced outside the sink is moved into its view
    red.zero_grad()
    assert all(p.grad is None for p in net.parameters())
    v = red.sink(w)
    assert v is not None and v.data_ptr() == red.flat.data_ptr() + w._vqa_flat_off * 4
    red.zero_grad()                                       # (a view is handed out once per step; start the step again)
    x = torch.randn(3, 5)
    net(x).sum().backward()
    red.finish()
    assert all(in_flat(p) for p in net.parameters())
    ref = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    ref.load_state_dict(net.state_dict())
    ref(x).sum().backward()
    assert all(torch.allclose(a.grad, b.grad) for a, b in zip(net.parameters(), ref.parameters()))
    assert red.sink(w) is None                            # .grad exists: accumulation is autograd's job
    # accumulation mode: one memset, views kept
    red.zero_grad(set_to_none=False)
    assert red.flat.abs().sum() == 0 and w.grad.abs().sum() == 0 and in_flat(w)
    red.remove()


def test_sink_views_are_adopted_by_autograd_without_a_copy():
    """The mechanism vqa_b200.ops relies on: a custom Function whose backward writes into the sink's view and returns it."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer
    lin = torch.nn.Linear(4, 5, bias=False)
    red = GradReducer(lin.parameters())

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            return x @ w.t()

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            out = red.sink(w)
            assert out is not None
            torch.mm(g.t(), x, out=out)
            return None, out

    x = torch.randn(3, 4)
    for _ in range(2):
        red.zero_grad()
        red.flat.fill_(7.0)                               # stale content must be overwritten, not accumulated
        F.apply(x, lin.weight).sum().backward()
        red.finish()
        assert lin.weight.grad.data_ptr() == red.flat.data_ptr() + lin.weight._vqa_flat_off * 4
        assert torch.allclose(lin.weight.grad, torch.ones(3, 5).t() @ x)
    red.remove()


def test_mark_ready_launches_a_bucket_before_the_node_returns_and_is_not_counted_twice():
    """Fused operators are single autograd nodes: they report finished parameter gradients themselves (mark_ready), so a bucket's
    all-reduce can start inside their backward; the post-accumulate hook that fires later for the same parameter must not count again."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer
    a = torch.nn.Parameter(torch.randn(300, 10)); b = torch.nn.Parameter(torch.randn(10))
    red = GradReducer([a, b], bucket_bytes=4096)           # two buckets: {b} is not enough alone -> buckets from the end: {b, a}? sizes decide
    seen = []

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, wa, wb):
            ctx.save_for_backward(x, wa, wb)
            return x @ wa + wb

        @staticmethod
        def backward(ctx, g):
            x, wa, wb = ctx.saved_tensors
            ga, gb = red.sink(wa), red.sink(wb)
            torch.mm(x.t(), g, out=ga)
            red.mark_ready(wa)
            seen.append(("after a", red.launched, list(red._ready)))
            torch.sum(g, dim=0, out=gb)
            red.mark_ready(wb)
            red.mark_ready(wb)                               # idempotent
            seen.append(("after b", red.launched, list(red._ready)))
            return None, ga, gb

    x = torch.randn(7, 300)
    for _ in range(2):
        red.zero_grad()
        launched0 = red.launched
        seen.clear()
        F.apply(x, a, b).sum().backward()
        red.finish()
        nb = len(red.bucket_size)
        assert red.launched - launched0 == nb               # every bucket exactly once, hooks did not recount
        assert seen[-1][1] - launched0 == nb                 # ... and all of them already inside backward
        assert torch.allclose(a.grad, x.t() @ torch.ones(7, 10)) and torch.allclose(b.grad, torch.full((10,), 7.0))
        assert a.grad.data_ptr() == red.flat.data_ptr() + a._vqa_flat_off * 4
    red.remove()


def test_reducer_average_flag_leaves_the_sum_for_an_optimiser_that_scales_itself():
    """optim.FlatAdam clears ``average`` and folds 1/world into its own pass; with one process finish() never scales."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer
    net = torch.nn.Linear(4, 2)
    red = GradReducer(net.parameters())
    assert red.average is True
    red.world, red.average = 2, False                     # pretend: the buffer holds a two-rank SUM
    red._launch = lambda b: None                          # no process group in this test
    red.zero_grad()
    net(torch.ones(1, 4)).sum().backward()
    before = red.flat.clone()
    red.finish()
    assert torch.equal(red.flat, before)
    red.average = True
    red.finish()
    assert torch.allclose(red.flat, before * 0.5)
    red.remove()


def _accum_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer, broadcast_parameters
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.ReLU(), torch.nn.Linear(33, 3))
    broadcast_parameters(net)
    red = GradReducer(net.parameters(), bucket_bytes=1024)
    torch.manual_seed(7)
    x_all, y_all = torch.randn(4 * world, 20), torch.randn(4 * world, 3)
    xs, ys = x_all[rank * 4:(rank + 1) * 4], y_all[rank * 4:(rank + 1) * 4]
    red.zero_grad(set_to_none=False)                   # accumulation over two micro-batches of 2
    launched0 = red.launched
    for m in range(2):
        (((net(xs[2 * m:2 * m + 2]) - ys[2 * m:2 * m + 2]) ** 2).sum() / (4 * 3)).backward()
        assert red.launched == launched0               # nothing is reduced before finish()
    red.finish()
    assert red.launched - launched0 == len(red.bucket_size)
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.ReLU(), torch.nn.Linear(33, 3))
    (((ref(x_all) - y_all) ** 2).sum() / (4 * world * 3)).backward()
    err = max((a.grad - b.grad).abs().max().item() for a, b in zip(net.parameters(), ref.parameters()))
    q.put((rank, err))
    dist.destroy_process_group()


def test_two_rank_gradient_accumulation_reduces_once_in_finish():
    """zero_grad(set_to_none=False) + two backwards per step at world 2: the averaged gradient equals the single-process gradient of
    the concatenated batch (a bucket reduced after the FIRST backward would miss the second micro-batch and race it)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-6, (rank, err)


def test_sink_hands_a_view_out_once_per_step_so_shared_parameters_sum_correctly():
    """A parameter feeding two autograd nodes (the model called twice before one backward): the second node must not get an aliasing
    view of the same memory - autograd would add two views of one buffer and produce 2 * g_last instead of g_a + g_b."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
    from vqa_b200.ddp import GradReducer
    lin = torch.nn.Linear(4, 5, bias=False)
    red = GradReducer(lin.parameters())

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            return x @ w.t()

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            out = red.sink(w)
            if out is None:
                out = torch.empty_like(w)
            torch.mm(g.t(), x, out=out)
            return None, out

    xa, xb = torch.randn(3, 4), torch.randn(6, 4)
    red.zero_grad()
    (F.apply(xa, lin.weight).sum() + 2.0 * F.apply(xb, lin.weight).sum()).backward()
    red.finish()
    want = torch.ones(3, 5).t() @ xa + 2.0 * torch.ones(6, 5).t() @ xb
    assert torch.allclose(lin.weight.grad, want, atol=1e-6)
    assert lin.weight.grad.data_ptr() == red.flat.data_ptr() + lin.weight._vqa_flat_off * 4
    red.zero_grad()
    assert red.sink(lin.weight) is not None and red.sink(lin.weight) is None       # once per step
    red.remove()
