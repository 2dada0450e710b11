"""The tail of the training step: criterion (run.py:382,431) and Adam (run.py:392,435) as single kernels of libvqa_sm100.so,
checked against torch's own modules (fp64 for the loss, torch.optim.Adam for the update rule) and, through the captured step,
against a step that uses torch's criterion and fused Adam."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, r):
    a, r = a.double().cpu(), r.double().cpu()
    return ((a - r).abs().max() / r.abs().max().clamp(min=1e-30)).item()


@pytest.mark.parametrize("B,A", [(512, 3000), (7, 13), (1, 1), (64, 3001), (3, 4096)])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_multilabel_soft_margin_loss_matches_torch(B, A, reduction):
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    g = torch.Generator().manual_seed(B * 131 + A)
    x = (4 * torch.randn(B, A, generator=g)).to(DEV).requires_grad_()
    y = torch.rand(B, A, generator=g).to(DEV)
    y[y < 0.7] = 0                                        # soft labels: mostly zero, as the VQA targets
    loss = MultiLabelSoftMarginLoss(reduction=reduction)(x, y)
    (3.0 * loss).backward()
    xr = x.detach().double().requires_grad_()
    ref = torch.nn.MultiLabelSoftMarginLoss(reduction=reduction)(xr, y.double())
    (3.0 * ref).backward()
    assert loss.shape == () and loss.dtype == torch.float32
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert _rel(x.grad, xr.grad) < 2e-6
    # same inputs, second launch: the ticket counter was left zero and the sum does not depend on block scheduling
    again = MultiLabelSoftMarginLoss(reduction=reduction)(x.detach(), y)
    assert again.item() == loss.item()


def test_loss_argument_errors():
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    with pytest.raises(NotImplementedError):
        MultiLabelSoftMarginLoss(reduction="none")
    with pytest.raises(NotImplementedError):
        MultiLabelSoftMarginLoss(weight=torch.ones(3))
    with pytest.raises(RuntimeError):
        MultiLabelSoftMarginLoss()(torch.zeros(2, 3, device=DEV), torch.zeros(2, 4, device=DEV))
    with pytest.raises(RuntimeError):                     # no CPU fallback
        MultiLabelSoftMarginLoss()(torch.zeros(2, 3), torch.zeros(2, 3))


def _make_params(seed):
    """Odd sizes, plus consecutive views of one buffer (like conv_weights.{k}.weight) whose addresses are not 16-byte aligned."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(300, 17), (5,), (4096,), (129, 65), (1,), (3, 2052)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    buf = torch.randn(3 * 7 * 9 + 1, generator=g).to(DEV)
    ps += [torch.nn.Parameter(buf[1 + k * 63:1 + (k + 1) * 63].view(7, 9)) for k in range(3)]
    return ps


@pytest.mark.parametrize("wd,world", [(0.0, 1), (0.01, 1), (0.0, 2)])
def test_flat_adam_matches_torch_adam(wd, world):
    from vqa_b200.ddp import GradReducer
    from vqa_b200.optim import FlatAdam
    ps = _make_params(3)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    red = GradReducer(ps)
    red.world = world                                     # the buffer then holds the SUM over ranks; Adam applies 1/world
    opt = FlatAdam(red, lr=3e-3, betas=(0.8, 0.95), eps=1e-8, weight_decay=wd)
    assert red.average is False
    ropt = torch.optim.Adam(ref, lr=3e-3, betas=(0.8, 0.95), eps=1e-8, weight_decay=wd)
    g = torch.Generator().manual_seed(11)
    for it in range(6):
        if it == 4:                                       # a scheduler lowers the rate
            opt.param_groups[0]["lr"] = ropt.param_groups[0]["lr"] = 1e-3
        opt.zero_grad(set_to_none=False)
        for p, r in zip(ps, ref):
            gr = torch.randn(p.shape, generator=g).to(DEV) * (10.0 ** (it - 3))
            p.grad.copy_(gr)
            r.grad = gr / world
        opt.step()
        ropt.step()
    assert opt.steps_taken == 6
    for i, (p, r) in enumerate(zip(ps, ref)):
        assert _rel(p.detach(), r.detach()) < 2e-6, i
        assert _rel(opt.state[p]["exp_avg"], ropt.state[r]["exp_avg"]) < 2e-6, i
        assert _rel(opt.state[p]["exp_avg_sq"], ropt.state[r]["exp_avg_sq"]) < 2e-6, i
    # a torch.optim.Adam checkpoint continues in FlatAdam (and FlatAdam's own state_dict has the same layout)
    ps2 = [torch.nn.Parameter(r.detach().clone()) for r in ref]
    red2 = GradReducer(ps2)
    red2.world = world
    opt2 = FlatAdam(red2, lr=1.0)
    opt2.load_state_dict(ropt.state_dict())
    assert opt2.steps_taken == 6 and opt2.param_groups[0]["lr"] == 1e-3 and opt2.param_groups[0]["betas"] == (0.8, 0.95)
    opt2.zero_grad(set_to_none=False)
    for p, r in zip(ps2, ref):
        gr = torch.randn(p.shape, generator=g).to(DEV)
        p.grad.copy_(gr)
        r.grad = gr / world
    opt2.step()
    ropt.step()
    for i, (p, r) in enumerate(zip(ps2, ref)):
        assert _rel(p.detach(), r.detach()) < 2e-6, i
    sd = opt2.state_dict()
    assert set(sd["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"} and int(sd["state"][0]["step"]) == 7
    red.remove()
    red2.remove()


def test_flat_adam_refuses_a_step_with_missing_gradients():
    from vqa_b200.ddp import GradReducer
    from vqa_b200.optim import FlatAdam
    ps = _make_params(5)
    red = GradReducer(ps)
    opt = FlatAdam(red)
    opt.zero_grad()                                       # default: every .grad dropped, nothing written since
    with pytest.raises(RuntimeError, match="received no gradient"):
        opt.step()
    red.remove()


def test_captured_step_with_fused_tail_matches_torch_tail():
    """Same model, batches and seeds through TrainStep (CUDA graph) with our criterion + FlatAdam and with torch's modules."""
    from vqa_b200.ddp import GradReducer
    from vqa_b200.engine import TrainStep
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    from vqa_b200.optim import FlatAdam
    from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
    import sparse_graph_model as M
    w = WORKLOADS["small"]
    batches = [make_batch(w, seed=40 + i) for i in range(3)]
    results = []
    for tail in ("fused", "torch"):
        torch.manual_seed(1000)
        kw = w.model_kwargs()
        kw["dropout"] = 0.0
        model = M.Model(pretrained_wemb=make_wemb(w), **kw).to(DEV).train()
        model.max_question_len = w.max_qlen
        red = GradReducer(model.parameters())
        if tail == "fused":
            crit, opt = MultiLabelSoftMarginLoss(), FlatAdam(red, lr=1e-3)
        else:
            crit, opt = torch.nn.MultiLabelSoftMarginLoss(), torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
        step = TrainStep(model, opt, crit, reducer=red, use_graph=True, seed=77)
        losses = []
        for it in range(5):
            b = batches[it % 3]
            losses.append(step(b["question"], b["image"], b["K"], b["qlen"], b["target"]).item())
        torch.cuda.synchronize()
        results.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
        step.close()
        red.remove()
    (lf, pf), (lt, pt) = results
    for a, b in zip(lf, lt):
        assert abs(a - b) <= 1e-3 * abs(b), (lf, lt)
    # Adam's first steps move an element by ~lr * sign(g) whatever |g| is, so an element whose gradient is a rounding-level residue may
    # legitimately step the other way under the other tail's (equally valid) rounding: bound such elements by count and by the total
    # distance five steps can cover, and require everything else to agree far below one step's size
    lr, steps = 1e-3, 5
    total = off = 0
    for k in pt:
        d = (pf[k] - pt[k]).abs()
        assert d.max().item() <= 2 * lr * steps, k
        total += d.numel()
        off += int((d > 1e-4).sum())
    assert off <= 1e-3 * total, (off, total)


def test_captured_step_follows_the_plain_loop_trajectory():
    """TrainStep == the reference loop body run eagerly (run.py:425-460), update for update: the warm-up iterations before a capture
    run on a snapshot (no hidden parameter / Adam-moment / step-count updates), a second batch shape captures once and is then
    replayed from the cache, and a question longer than model.max_question_len is rejected instead of truncated."""
    from vqa_b200.ddp import GradReducer
    from vqa_b200.engine import TrainStep
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    from vqa_b200.optim import FlatAdam
    from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
    import sparse_graph_model as M
    w = WORKLOADS["small"]
    full = [make_batch(w, seed=60 + i) for i in range(2)]
    short = make_batch(w, seed=70, batch=5)                       # the ragged last batch of an epoch
    order = [full[0], full[1], short, full[0], short, full[1]]
    runs = []
    for mode in ("graph", "plain"):
        torch.manual_seed(1000)
        kw = w.model_kwargs()
        kw["dropout"] = 0.0
        model = M.Model(pretrained_wemb=make_wemb(w), **kw).to(DEV).train()
        model.max_question_len = w.max_qlen
        red = GradReducer(model.parameters())
        crit, opt = MultiLabelSoftMarginLoss(), FlatAdam(red, lr=1e-3)
        losses = []
        if mode == "graph":
            step = TrainStep(model, opt, crit, reducer=red, use_graph=True, seed=77)
            for b in order:
                losses.append(step(b["question"], b["image"], b["K"], b["qlen"], b["target"]).item())
            assert len(step._cache) == 2                          # one capture per batch shape, however often the shape comes back
            with pytest.raises(ValueError, match="max_question_len"):
                b = dict(full[0]); b["qlen"] = [torch.tensor(w.max_qlen + 1)] + list(b["qlen"][1:])
                step(b["question"], b["image"], b["K"], b["qlen"], b["target"])
            step.close()
        else:
            qlen_t = lambda b: torch.tensor([int(x) for x in b["qlen"]], dtype=torch.int32, device=DEV)
            for b in order:
                red.zero_grad()
                logits, _, _ = model(b["question"].to(DEV), b["image"].to(DEV), b["K"].to(DEV), qlen_t(b))
                loss = crit(logits, b["target"].to(DEV))
                loss.backward()
                red.finish()
                opt.step()
                losses.append(loss.item())
        torch.cuda.synchronize()
        assert opt.steps_taken == len(order)                      # Adam's bias correction saw exactly the real steps
        runs.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
        red.remove()
    (lg, pg), (lp, pp) = runs
    assert lg == pytest.approx(lp, rel=1e-6), (lg, lp)            # same kernels, same order: the first loss would already differ after a hidden update
    # Parameters: the split-K weight-gradient products and the embedding scatter add with atomics, so two runs of the SAME loop differ
    # in the last bits of the question-path gradients (measured: 1e-11 at step 0), Adam turns that into up to lr * dg / eps = 1e5 * dg
    # for elements whose gradient is itself rounding noise, and a ReLU mask that flips on such an element changes later gradients
    # discretely (measured between two plain runs as well as between graph and plain).  What a captured step must reproduce is
    # the movement: per tensor, the two runs' distance is a small fraction of the distance travelled from the initial weights
    # (one hidden or missing update would be ~1/6 of it), and no element moved further than Adam can move it in six steps.
    torch.manual_seed(1000)
    kw = w.model_kwargs(); kw["dropout"] = 0.0
    init = {k: v.detach().clone() for k, v in M.Model(pretrained_wemb=make_wemb(w), **kw).to(DEV).state_dict().items()}
    for k in pp:
        d = (pg[k].float() - pp[k].float())
        moved = (pp[k].float() - init[k].float()).norm().item()
        assert float(d.abs().max()) <= 2 * 6 * 1e-3, k
        assert d.norm().item() <= 2e-2 * moved + 1e-7, (k, d.norm().item(), moved)


def test_train_step_requires_max_question_len():
    from vqa_b200.engine import TrainStep
    from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
    import sparse_graph_model as M
    w = WORKLOADS["tiny"]
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(DEV).train()
    step = TrainStep(model, torch.optim.Adam(model.parameters()), torch.nn.MultiLabelSoftMarginLoss())
    b = make_batch(w, seed=1)
    with pytest.raises(ValueError, match="max_question_len"):
        step(b["question"], b["image"], b["K"], b["qlen"], b["target"])
