"""Error study for the split-bf16 tensor-core products (CPU, emulation): how do logits and gradients of one training step move
when the Linear-type products (graph-learner, per-kernel projections, classifier, GRU: forward, dX and dW) are evaluated from
bf16 operand planes with 3, 2 or 1 passes?  The arbiter is the same step in fp64.

    python tools/split_error_study.py [--workloads small medium wide] > profiles/rNN_split_error_study.md

Emulation: x = hi + lo with hi = bf16(x), lo = bf16(x - hi); a product is the sum of the chosen plane pairs, each evaluated by an
fp32 matmul of bf16-valued operands (the tensor core multiplies bf16 exactly and accumulates in fp32).  The step is the oracle's
(``oracle/vqa_oracle.py``, reference operation order) with ``Tensor.__matmul__`` patched for the ``x @ W.t()`` call sites; the
adjacency ``h h^T`` and the neighbourhood aggregation stay exact, as in the CUDA path (fp32 FMA / fp32-grade planes).

Modes:  3   = hi.hi + lo.hi + hi.lo everywhere (the shipped parity mode)
        2w  = weights as ONE plane in forward and dX ((x_hi + x_lo) . W_hi), dW still 3 passes
        2   = additionally dW = (dY_hi + dY_lo)^T . X_hi
        1   = hi.hi everywhere
        "+GL3" = the same, but the FORWARD products of the graph learner stay 3-pass (what the shipped bf16 mode does): the top-k
                 neighbour selection is discrete, so a perturbed adjacency flips near-tied neighbours and every gradient with them.
        "+GRU3" = the question encoder's forward products stay 3-pass too (its output feeds the graph learner: the softmax over the
                 selected adjacency values turns an absolute error of the adjacency into a relative error of the edge weights).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
from oracle import vqa_oracle as O  # noqa: E402  (a study tool, not the product: tools/ is outside the package)
from vqa_b200.synthetic import WORKLOADS, Workload, make_batch, make_wemb  # noqa: E402
import sparse_graph_model as M  # noqa: E402

MODE = {"fwd": (2, 2), "dx": (2, 2), "dw": (2, 2)}          # planes used of (left, right) operand per product kind
STATE = {"in_gl": False, "gl3": False, "in_gru": False, "gru3": False}


def planes(x, n):
    hi = x.bfloat16().float()
    return [hi] if n == 1 else [hi, (x - hi).bfloat16().float()]


def emu(a, b, na, nb):
    pa, pb = planes(a, na), planes(b, nb)
    out = None
    for i, x in enumerate(pa):
        for j, y in enumerate(pb):
            if i == 1 and j == 1:
                continue                                     # lo.lo is never issued
            t = torch.mm(x, y)
            out = t if out is None else out + t
    return out


class SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wt):                                 # x (M, in), wt (in, out) = W^T
        ctx.save_for_backward(x, wt)
        ctx.mode = dict(MODE)
        fwd = (2, 2) if (STATE["in_gl"] and STATE["gl3"]) or (STATE["in_gru"] and STATE["gru3"]) else MODE["fwd"]
        return emu(x, wt, *fwd)

    @staticmethod
    def backward(ctx, dy):
        x, wt = ctx.saved_tensors
        m = ctx.mode
        dx = emu(dy, wt.t(), *m["dx"]) if ctx.needs_input_grad[0] else None
        dwt = emu(x.t(), dy, m["dw"][1], m["dw"][0]) if ctx.needs_input_grad[1] else None   # dW^T = X^T dY: (right, left) roles
        return dx, dwt


_orig = torch.Tensor.__matmul__


def patched(self, other):
    if self.dtype == torch.float32 and other.dim() == 2 and self.dim() in (2, 3) and not (self.dim() == 3 and other.dim() == 3):
        lead = self.shape[:-1]
        return SplitLinear.apply(self.reshape(-1, self.shape[-1]), other).reshape(*lead, other.shape[1])
    return _orig(self, other)


_gl = O.graph_learner


def graph_learner_marked(*a, **k):
    STATE["in_gl"] = True
    try:
        return _gl(*a, **k)
    finally:
        STATE["in_gl"] = False


O.graph_learner = graph_learner_marked
_gru = O.gru_last_hidden


def gru_marked(*a, **k):
    STATE["in_gru"] = True
    try:
        return _gru(*a, **k)
    finally:
        STATE["in_gru"] = False


O.gru_last_hidden = gru_marked


def set_mode(name):
    STATE["gru3"] = "+GRU3" in name
    STATE["gl3"] = "+GL3" in name
    name = name.replace("+GL3", "").replace("+GRU3", "")
    full, one = (2, 2), (1, 1)
    table = {"3": dict(fwd=full, dx=full, dw=full), "2w": dict(fwd=(2, 1), dx=(2, 1), dw=full),
             "2": dict(fwd=(2, 1), dx=(2, 1), dw=(2, 1)), "1": dict(fwd=one, dx=one, dw=one)}
    MODE.update(table[name])


def rel(a, r):
    return ((a.double() - r).abs().max() / r.abs().max().clamp(min=1e-300)).item()


GROUPS = [("logits", None), ("graph learner W", "adjacency_1."), ("GC1 conv W", "graph_convolution_1.conv_weights"),
          ("GC2 conv W", "graph_convolution_2.conv_weights"), ("Gaussian params", ("mean_", "precision_")), ("classifier", "out_"),
          ("GRU", "q_gru."), ("embedding", "wembed.")]


def study(w):
    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs())
    with torch.no_grad():
        for gc in (model.graph_convolution_1, model.graph_convolution_2):
            gc.precision_rho.clamp_(min=0.05)
            gc.precision_theta.clamp_(min=0.05)
    p32 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    b = make_batch(w, seed=5)
    qlen = [int(x) for x in b["qlen"]]
    p64 = {k: v.double() for k, v in p32.items()}
    _, g64, (l64, a64, m64) = O.train_step_grads(p64, b["question"], b["image"].double(), qlen, b["target"].double(), w.neighbourhood, w.n_kernels)
    rows = []
    runs = [("fp32 (torch CPU)", None)] + [(f"split-bf16, mode {m}", m) for m in ("3", "2w", "2", "1", "2w+GL3", "2+GL3", "1+GL3", "2+GL3+GRU3", "1+GL3+GRU3")]
    for label, mode in runs:
        if mode is not None:
            set_mode(mode)
            torch.Tensor.__matmul__ = patched
        try:
            _, g, (lg, adj, amax) = O.train_step_grads(p32, b["question"], b["image"], qlen, b["target"], w.neighbourhood, w.n_kernels)
        finally:
            torch.Tensor.__matmul__ = _orig
        errs = []
        for _, key in GROUPS:
            if key is None:
                errs.append(rel(lg, l64))
            else:
                keys = [k for k in g if (any(s in k for s in key) if isinstance(key, tuple) else k.startswith(key))]
                errs.append(max(rel(g[k], g64[k]) for k in keys))
        nb_sets = lambda a: torch.topk(a, w.neighbourhood, dim=-1).indices.sort(dim=-1).values      # noqa: E731
        flips = (int((nb_sets(adj.double()) != nb_sets(a64)).any(dim=-1).sum()), int((amax != m64).sum()))
        rows.append((label, errs, flips))
    print(f"\n### {w.name}: B={w.batch}, K={w.n_obj}, F={w.feat_dim}, H={w.hid_dim}, nk={w.n_kernels}, nb={w.neighbourhood}, A={w.out_dim}\n")
    print("| products | " + " | ".join(n for n, _ in GROUPS) + f" | nodes with another top-{w.neighbourhood} set (of {w.batch * w.n_obj}) | other max-pool winners (of {w.batch * w.hid_dim}) |")
    print("|---|" + "---:|" * (len(GROUPS) + 2))
    for label, errs, flips in rows:
        print(f"| {label} | " + " | ".join(f"{e:.1e}" for e in errs) + f" | {flips[0]} | {flips[1]} |")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="+", default=["small", "medium", "wide"])
    args = ap.parse_args()
    wl = dict(WORKLOADS)
    wl["wide"] = Workload("wide", 4, 36, 2052, vocab=2000, out_dim=3000)       # VQA2 widths at a CPU-sized batch
    print("# Max-norm relative error of one training step against the fp64 step (CPU emulation, tools/split_error_study.py)")
    print("\nColumns: logits, then the worst gradient of each parameter group (parity budget of BASELINE.json north_star: 1e-3), then how many of the"
          "\nstep's DISCRETE choices differ from the fp64 step's: per-node top-k neighbour sets (sparse_graph_model.py:225) and per-column winners of the"
          "\nmax over nodes (:150). One flipped choice re-routes a gradient row; that, not the products' own rounding, is what the gradient columns show.")
    for name in args.workloads:
        study(wl[name])


if __name__ == "__main__":
    main()
