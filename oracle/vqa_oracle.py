"""CPU oracle for the conditioned-graph VQA hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (torch CPU / numpy, functional style) of the
algorithm implemented by the reference model, used as the parity checker for the CUDA
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``vqa-project_b200/``) never imports anything from ``oracle/``.

Parity status: **pinned**.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the pin is the reference code itself: ``tests/golden/make_golden.py``
imports the unmodified reference modules from ``/root/reference`` in the build container,
runs them on seeded synthetic inputs and commits inputs + outputs + gradients as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against
those files (forward and autograd backward).

Every function cites the reference lines it restates (paths relative to /root/reference).
The operation ORDER deliberately follows the reference (concat -> GraphLearner,
aggregate-first patch operator, per-kernel linears) rather than the re-associated
order the CUDA path uses, so that the comparison is between two genuinely different
evaluation orders of the same maths.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]

TWO_PI = 2.0 * math.pi
GAUSS_EPS = 1e-14  # layers.py:111,117


# --------------------------------------------------------------------------------------
# parameter helpers
# --------------------------------------------------------------------------------------
def weight_norm_effective(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Old-style ``nn.utils.weight_norm`` (dim=0): W[r,:] = g[r] * v[r,:] / ||v[r,:]||_2.

    layers.py:171-172 and sparse_graph_model.py:88-89 wrap their Linears with it, which
    is what produces the ``weight_g`` (out,1) / ``weight_v`` (out,in) state_dict keys.
    """
    norm = v.pow(2).sum(dim=1, keepdim=True).sqrt()
    return v * (g / norm)


def wn_linear(x: torch.Tensor, p: Params, prefix: str) -> torch.Tensor:
    w = weight_norm_effective(p[prefix + ".weight_v"], p[prefix + ".weight_g"])
    return x @ w.t() + p[prefix + ".bias"]


def init_params(
    vocab_size: int,
    emb_dim: int,
    feat_dim: int,
    hid_dim: int,
    out_dim: int,
    n_kernels: int,
    combined_dim: int = 512,
    seed: int = 1000,
    dtype: torch.dtype = torch.float32,
) -> Params:
    """Random parameters with the reference's state_dict keys and shapes.

    Distributions are sensible stand-ins (Gaussian params follow layers.py:65-70); parity
    tests always copy one set of numbers into both implementations, so the exact init
    stream of the reference does not matter here.
    """
    gen = torch.Generator().manual_seed(seed)

    def uni(shape, lo, hi):
        return (torch.rand(shape, generator=gen, dtype=torch.float64) * (hi - lo) + lo).to(dtype)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return uni((out_f, in_f), -bound, bound), uni((out_f,), -bound, bound)

    p: Params = {}
    p["wembed.weight"] = (0.4 * torch.randn(vocab_size, emb_dim, generator=gen, dtype=torch.float64)).to(dtype)
    k = 1.0 / math.sqrt(hid_dim)
    p["q_gru.weight_ih_l0"] = uni((3 * hid_dim, emb_dim), -k, k)
    p["q_gru.weight_hh_l0"] = uni((3 * hid_dim, hid_dim), -k, k)
    p["q_gru.bias_ih_l0"] = uni((3 * hid_dim,), -k, k)
    p["q_gru.bias_hh_l0"] = uni((3 * hid_dim,), -k, k)

    def wn(prefix, out_f, in_f):
        v, b = lin(out_f, in_f)
        p[prefix + ".bias"] = b
        p[prefix + ".weight_g"] = v.pow(2).sum(1, keepdim=True).sqrt() * uni((out_f, 1), 0.8, 1.2)
        p[prefix + ".weight_v"] = v

    wn("adjacency_1.edge_layer_1", combined_dim, feat_dim + hid_dim)
    wn("adjacency_1.edge_layer_2", combined_dim, combined_dim)
    for name, fin, fout in (
        ("graph_convolution_1", feat_dim, 2 * hid_dim),
        ("graph_convolution_2", 2 * hid_dim, hid_dim),
    ):
        p[name + ".mean_rho"] = uni((n_kernels, 1), 0.0, 1.0)
        p[name + ".mean_theta"] = uni((n_kernels, 1), -math.pi, math.pi)
        p[name + ".precision_rho"] = uni((n_kernels, 1), 0.05, 1.0)
        p[name + ".precision_theta"] = uni((n_kernels, 1), 0.05, 1.0)
        for i in range(n_kernels):
            p[f"{name}.conv_weights.{i}.weight"] = lin(fout // n_kernels, fin)[0]
    wn("out_1", out_dim, hid_dim)
    wn("out_2", out_dim, out_dim)
    return p


# --------------------------------------------------------------------------------------
# question encoder
# --------------------------------------------------------------------------------------
def gru_last_hidden(emb: torch.Tensor, qlen: Sequence[int], p: Params, prefix: str = "q_gru") -> torch.Tensor:
    """Final hidden state of a 1-layer GRU run over each sequence's first qlen[b] tokens.

    Restates ``pack_padded_sequence`` + ``nn.GRU`` + ``hid[0]`` (sparse_graph_model.py:117-121):
    a packed sequence simply stops updating sample b after qlen[b] steps.  Gate order in the
    stacked weights is (reset, update, new) as in torch.nn.GRU.
    """
    w_ih, w_hh = p[prefix + ".weight_ih_l0"], p[prefix + ".weight_hh_l0"]
    b_ih, b_hh = p[prefix + ".bias_ih_l0"], p[prefix + ".bias_hh_l0"]
    hid = w_hh.shape[1]
    bsz = emb.shape[0]
    lens = torch.as_tensor([int(x) for x in qlen], dtype=torch.long, device=emb.device)
    h = emb.new_zeros(bsz, hid)
    for t in range(int(lens.max())):
        gi = emb[:, t] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h_new = (1.0 - z) * n + z * h
        alive = (lens > t).unsqueeze(1)
        h = torch.where(alive, h_new, h)
    return h


# --------------------------------------------------------------------------------------
# graph pieces
# --------------------------------------------------------------------------------------
def box_centres(image: torch.Tensor) -> torch.Tensor:
    """Centres of the xyxy boxes stored in the last 4 feature columns (sparse_graph_model.py:106-108)."""
    bb = image[..., -4:]
    return bb[..., :2] + 0.5 * (bb[..., 2:] - bb[..., :2])


def polar_pseudo_coordinates(centres: torch.Tensor) -> torch.Tensor:
    """Dense (B,K,K,2) tensor of (rho, theta) for centre_i - centre_j (sparse_graph_model.py:244-269).

    theta = atan2(dx, dy): x is the FIRST argument (line 264-265).
    """
    d = centres.unsqueeze(2) - centres.unsqueeze(1)
    dx, dy = d[..., 0], d[..., 1]
    rho = torch.sqrt(dx * dx + dy * dy)
    theta = torch.atan2(dx, dy)
    return torch.stack((rho, theta), dim=-1)


def graph_learner(nodes: torch.Tensor, p: Params, prefix: str = "adjacency_1") -> torch.Tensor:
    """A = h h^T with h = relu(WN2(relu(WN1(nodes))))  (layers.py:174-197).

    The reference constructs a Dropout in GraphLearner.__init__ (layers.py:170) but never
    applies it in forward; neither do we.
    """
    h = torch.relu(wn_linear(nodes, p, prefix + ".edge_layer_1"))
    h = torch.relu(wn_linear(h, p, prefix + ".edge_layer_2"))
    return h @ h.transpose(1, 2)


def select_neighbourhood(adjacency: torch.Tensor, nb: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-row top-nb (unordered) and softmax over the selected values.

    sparse_graph_model.py:225-227.  The reference loops over the K rows calling F.softmax on
    each (B,nb) slice; that is bitwise one softmax over the last axis.
    """
    vals, idx = torch.topk(adjacency, k=nb, dim=-1, sorted=False)
    return torch.softmax(vals, dim=-1), idx


def gather_neighbours(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(B,K,D),(B,K,nb) -> (B,K,nb,D): row idx[b,i,m] of sample b (sparse_graph_model.py:161-178)."""
    bsz, k, nb = idx.shape
    flat = idx.reshape(bsz, k * nb)
    out = torch.gather(x, 1, flat.unsqueeze(-1).expand(bsz, k * nb, x.shape[-1]))
    return out.view(bsz, k, nb, x.shape[-1])


def gather_pseudo(pseudo: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(B,K,K,2),(B,K,nb) -> (B,K,nb,2) (sparse_graph_model.py:180-195)."""
    return torch.gather(pseudo, 2, idx.unsqueeze(-1).expand(*idx.shape, pseudo.shape[-1]))


def gaussian_kernel_weights(pseudo: torch.Tensor, p: Params, prefix: str) -> torch.Tensor:
    """(B,K,nb,2) -> (B*K*nb, nk) patch weights, normalised over the KERNEL axis (layers.py:100-125)."""
    rho = pseudo[..., 0].reshape(-1, 1)
    theta = pseudo[..., 1].reshape(-1, 1)
    mu_r, mu_t = p[prefix + ".mean_rho"].view(1, -1), p[prefix + ".mean_theta"].view(1, -1)
    s_r, s_t = p[prefix + ".precision_rho"].view(1, -1), p[prefix + ".precision_theta"].view(1, -1)
    w_rho = torch.exp(-0.5 * (rho - mu_r) ** 2 / (GAUSS_EPS + s_r ** 2))
    a1 = torch.abs(theta - mu_t)
    a2 = torch.abs(TWO_PI - a1)
    w_theta = torch.exp(-0.5 * torch.minimum(a1, a2) ** 2 / (GAUSS_EPS + s_t ** 2))
    w = w_rho * w_theta
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)  # NaN -> 0 BEFORE normalising (layers.py:120)
    return w / w.sum(dim=1, keepdim=True)


def graph_convolution(nbr: torch.Tensor, nbr_pseudo: torch.Tensor, p: Params, prefix: str, n_kernels: int) -> torch.Tensor:
    """MoNet-style convolution on fixed-size neighbourhoods (layers.py:72-98, 127-144).

    Aggregate-first, like the reference: Z[i,k,:] = sum_m w[i,m,k] nbr[i,m,:], then the k-th
    bias-free Linear maps Z[:,k,:] to the k-th output chunk.
    """
    bsz, k, nb, fin = nbr.shape
    w = gaussian_kernel_weights(nbr_pseudo, p, prefix).view(bsz * k, nb, n_kernels)
    z = torch.bmm(w.transpose(1, 2), nbr.reshape(bsz * k, nb, fin))  # (B*K, nk, fin)
    chunks = [z[:, i] @ p[f"{prefix}.conv_weights.{i}.weight"].t() for i in range(n_kernels)]
    return torch.cat(chunks, dim=1).view(bsz, k, -1)


# --------------------------------------------------------------------------------------
# full model
# --------------------------------------------------------------------------------------
def forward(
    p: Params,
    question: torch.Tensor,
    image: torch.Tensor,
    qlen: Sequence[int],
    neighbourhood_size: int,
    n_kernels: int,
    dropout_p: float = 0.0,
    training: bool = False,
    generator: Optional[torch.Generator] = None,
    return_intermediates: bool = False,
):
    """Restatement of ``Model.forward`` (sparse_graph_model.py:91-159).

    Returns (logits (B,A), adjacency (B,K,K), h_max_indices (B,H) int64) and, optionally, a
    dict of intermediates used by the kernel-level parity tests.
    """
    def drop(x):
        if not training or dropout_p == 0.0:
            return x
        keep = (torch.rand(x.shape, generator=generator, dtype=x.dtype, device=x.device) >= dropout_p).to(x.dtype)
        return x * keep / (1.0 - dropout_p)

    centres = box_centres(image)  # from the UN-dropped image (lines 106-108 precede 111)
    x = drop(image)
    pseudo = polar_pseudo_coordinates(centres)

    emb = p["wembed.weight"][question]
    qenc = gru_last_hidden(emb, qlen, p)  # (B,H)
    k_nodes = image.shape[1]
    nodes = torch.cat((x, qenc.unsqueeze(1).expand(-1, k_nodes, -1)), dim=-1)
    adjacency = graph_learner(nodes, p)

    alpha, idx = select_neighbourhood(adjacency, neighbourhood_size)
    nbr1 = alpha.unsqueeze(-1) * gather_neighbours(x, idx)  # weight=True (line 239-240)
    nbr_pseudo = gather_pseudo(pseudo, idx)
    g1 = drop(torch.relu(graph_convolution(nbr1, nbr_pseudo, p, "graph_convolution_1", n_kernels)))

    nbr2 = gather_neighbours(g1, idx)  # weight=False (line 145); same top-k as above
    g2 = torch.relu(graph_convolution(nbr2, nbr_pseudo, p, "graph_convolution_2", n_kernels))

    pooled, arg = g2.max(dim=1)
    hq = torch.relu(qenc) * pooled
    hidden = drop(torch.relu(wn_linear(hq, p, "out_1")))
    logits = wn_linear(hidden, p, "out_2")
    if return_intermediates:
        inter = dict(centres=centres, qenc=qenc, alpha=alpha, idx=idx, g1=g1, g2=g2, pooled=pooled, hq=hq)
        return logits, adjacency, arg, inter
    return logits, adjacency, arg


def multilabel_soft_margin_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.MultiLabelSoftMarginLoss with mean reduction (run.py:382,431)."""
    ls = torch.nn.functional.logsigmoid
    return (-(target * ls(logits) + (1.0 - target) * ls(-logits))).mean(dim=1).mean()


def train_step_grads(p: Params, question, image, qlen, target, neighbourhood_size, n_kernels) -> Tuple[torch.Tensor, Params, tuple]:
    """loss, {name: grad}, forward outputs -- dropout off, via torch autograd through this file."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = forward(leaves, question, image, qlen, neighbourhood_size, n_kernels)
    loss = multilabel_soft_margin_loss(out[0], target)
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    gd = {n: (g if g is not None else torch.zeros_like(leaves[n])) for n, g in zip(names, grads)}
    return loss.detach(), gd, tuple(o.detach() for o in out)


# --------------------------------------------------------------------------------------
# tiny numpy restatements (hand-checkable known-answer tests for the graph kernels)
# --------------------------------------------------------------------------------------
def np_topk_softmax(adj_row: np.ndarray, nb: int) -> Tuple[np.ndarray, np.ndarray]:
    """One adjacency row -> (sorted neighbour index set, softmax weights in that index order)."""
    order = np.argsort(-adj_row, kind="stable")[:nb]
    order = np.sort(order)
    v = adj_row[order].astype(np.float64)
    e = np.exp(v - v.max())
    return order, (e / e.sum()).astype(adj_row.dtype)


def np_edge_kernel_weights(ci, cj, mu_r, s_r, mu_t, s_t) -> np.ndarray:
    """Gaussian weights of ONE edge (query node centre ci, neighbour centre cj) over nk kernels."""
    dx, dy = float(ci[0] - cj[0]), float(ci[1] - cj[1])
    rho = math.sqrt(dx * dx + dy * dy)
    theta = math.atan2(dx, dy)
    a = np.exp(-0.5 * (rho - mu_r) ** 2 / (GAUSS_EPS + s_r ** 2))
    f1 = np.abs(theta - mu_t)
    f2 = np.abs(TWO_PI - f1)
    b = np.exp(-0.5 * np.minimum(f1, f2) ** 2 / (GAUSS_EPS + s_t ** 2))
    g = a * b
    g = np.where(np.isnan(g), 0.0, g)
    return g / g.sum()
