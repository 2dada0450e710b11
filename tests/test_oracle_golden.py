"""The oracle restatement must reproduce the reference's golden vectors (CPU, fp32 and fp64)."""
import math

import numpy as np
import pytest
import torch

from conftest import golden_params, rel_err
from oracle import vqa_oracle as O
from vqa_b200.synthetic import WORKLOADS

TOL32 = 2e-5   # fp32 oracle vs fp32 reference: different evaluation order only


def _inputs(g):
    q = torch.from_numpy(g["in.question"])
    img = torch.from_numpy(g["in.image"])
    qlen = [int(x) for x in g["in.qlen"]]
    tgt = torch.from_numpy(g["in.target"])
    return q, img, qlen, tgt


def test_forward_matches_reference(golden):
    name, g = golden
    w = WORKLOADS[name]
    q, img, qlen, _ = _inputs(g)
    logits, adj, arg, inter = O.forward(golden_params(g), q, img, qlen, w.neighbourhood, w.n_kernels,
                                        return_intermediates=True)
    assert rel_err(logits, g["out.logits"]) < TOL32
    assert rel_err(adj, g["out.adjacency"]) < TOL32
    assert rel_err(inter["qenc"], g["layer.qenc"]) < TOL32
    # argmax over K: only compare where the top-2 gap is not a rounding artefact
    g2 = inter["g2"]
    top2 = g2.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5 * top2[:, 0].abs().clamp(min=1e-6)
    assert torch.equal(arg[safe], torch.from_numpy(g["out.h_max_indices"])[safe])
    assert safe.float().mean() > 0.5


def test_neighbourhood_sets_exact_given_reference_adjacency(golden):
    name, g = golden
    w = WORKLOADS[name]
    adj = torch.from_numpy(g["out.adjacency"])
    alpha, idx = O.select_neighbourhood(adj, w.neighbourhood)
    order = idx.argsort(-1)
    assert np.array_equal(torch.gather(idx, -1, order).numpy(), g["nbr.idx_sorted"])
    assert rel_err(torch.gather(alpha, -1, order), g["nbr.alpha_sorted"]) < 1e-6
    # numpy single-row restatement agrees too
    for b, i in ((0, 0), (adj.shape[0] - 1, adj.shape[1] - 1)):
        s, a = O.np_topk_softmax(g["out.adjacency"][b, i], w.neighbourhood)
        assert np.array_equal(s, g["nbr.idx_sorted"][b, i])
        assert np.allclose(a, g["nbr.alpha_sorted"][b, i], rtol=1e-5, atol=1e-7)


def test_layer_level_pieces(golden):
    name, g = golden
    w = WORKLOADS[name]
    p = golden_params(g)
    pseudo = torch.from_numpy(g["layer.nbr_pseudo"])
    gw = O.gaussian_kernel_weights(pseudo, p, "graph_convolution_1")
    assert rel_err(gw, g["layer.gauss_w"]) < 1e-5
    assert torch.allclose(gw.sum(1), torch.ones(gw.shape[0]), atol=1e-5)  # normalised over KERNELS
    img = torch.from_numpy(g["in.image"])
    adj = torch.from_numpy(g["out.adjacency"])
    alpha, idx = O.select_neighbourhood(adj, w.neighbourhood)
    nbr = alpha.unsqueeze(-1) * O.gather_neighbours(img, idx)
    if "layer.nbr_feat" in g:
        assert rel_err(nbr, g["layer.nbr_feat"]) < 1e-6
    cen = O.box_centres(img)
    nbr_pseudo = O.gather_pseudo(O.polar_pseudo_coordinates(cen), idx)
    assert rel_err(nbr_pseudo, g["layer.nbr_pseudo"]) < 1e-6
    out = O.graph_convolution(nbr, nbr_pseudo, p, "graph_convolution_1", w.n_kernels)
    assert rel_err(out, g["layer.gc1_out"]) < TOL32
    if "layer.graph_nodes" in g:
        a = O.graph_learner(torch.from_numpy(g["layer.graph_nodes"]), p)
        assert rel_err(a, g["layer.adjacency"]) < TOL32
    # single-edge numpy KAT
    b, i, m = 0, 1, 2
    j = int(idx[b, i, m])
    gp = {k: p["graph_convolution_1." + k].view(-1).double().numpy() for k in
          ("mean_rho", "precision_rho", "mean_theta", "precision_theta")}
    ref = O.np_edge_kernel_weights(cen[b, i].double().numpy(), cen[b, j].double().numpy(),
                                   gp["mean_rho"], gp["precision_rho"], gp["mean_theta"], gp["precision_theta"])
    nb = w.neighbourhood
    got = gw.view(img.shape[0], img.shape[1], nb, -1)[b, i, m].double().numpy()
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-6)


def test_gradients_match_reference(golden):
    name, g = golden
    w = WORKLOADS[name]
    q, img, qlen, tgt = _inputs(g)
    loss, grads, _ = O.train_step_grads(golden_params(g), q, img, qlen, tgt, w.neighbourhood, w.n_kernels)
    assert abs(loss.item() - float(g["out.loss"])) < 1e-6
    worst = 0.0
    for k, v in g.items():
        if not k.startswith("grad."):
            continue
        e = rel_err(grads[k[5:]], v)
        worst = max(worst, e)
        assert e < 2e-4, (k, e)
    assert worst > 0.0  # different evaluation order: not the same code


def test_fp64_oracle_is_the_arbiter(golden):
    """fp64 oracle vs fp32 reference differ only by fp32 rounding (~1e-6)."""
    name, g = golden
    w = WORKLOADS[name]
    q, img, qlen, _ = _inputs(g)
    logits, adj, _ = O.forward(golden_params(g, torch.float64), q, img.double(), qlen, w.neighbourhood, w.n_kernels)
    assert rel_err(logits, g["out.logits"]) < 2e-5
    assert rel_err(adj, g["out.adjacency"]) < 2e-5


def test_gaussian_nan_rule_is_replicated_not_fixed():
    """layers.py:120-123: NaN -> 0 happens before the kernel-axis sum; an all-underflow row stays 0/0 = NaN."""
    p = {"gc.mean_rho": torch.tensor([[0.5], [0.6]]), "gc.mean_theta": torch.tensor([[0.0], [1.0]]),
         "gc.precision_rho": torch.tensor([[1e-4], [1e-4]]), "gc.precision_theta": torch.tensor([[0.5], [0.5]])}
    pseudo = torch.tensor([[[[10.0, 0.3]]]])  # far from every mean_rho -> exp underflows to 0 for all kernels
    w = O.gaussian_kernel_weights(pseudo, p, "gc")
    assert torch.isnan(w).all()
    # atan2(0,0) = 0 for self edges, theta uses x first
    c = torch.tensor([[[0.2, 0.3], [0.5, 0.3]]])
    ps = O.polar_pseudo_coordinates(c)
    assert ps[0, 0, 0, 1] == 0 and ps[0, 0, 0, 0] == 0
    assert math.isclose(ps[0, 0, 1, 1].item(), math.atan2(-0.3, 0.0), rel_tol=1e-6)


def test_packed_gru_matches_torch_gru():
    torch.manual_seed(0)
    gru = torch.nn.GRU(8, 16)
    emb = torch.randn(5, 7, 8)
    qlen = [7, 5, 5, 3, 1]
    packed = torch.nn.utils.rnn.pack_padded_sequence(emb, qlen, batch_first=True, enforce_sorted=False)
    _, hid = gru(packed)
    p = {"q_gru." + k: v.detach() for k, v in gru.state_dict().items()}
    assert rel_err(O.gru_last_hidden(emb, qlen, p), hid[0].detach()) < 1e-6
