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


def test_hand_checkable_known_answers():
    """tests/golden/kat_tiny.json (produced by the unmodified reference, tests/golden/make_kat.py) against closed forms derived with
    ``math`` alone, and the oracle against both: 4 unit-square nodes, 3 Gaussian kernels, top-2 neighbourhoods."""
    import json
    import math
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_tiny.json")) as f:
        kat = json.load(f)
    c = kat["centres"]
    pi = math.pi
    # ---- pseudo-coordinates: (rho, theta) of c_i - c_j with theta = atan2(dx, dy)  [x first: sparse_graph_model.py:264-265]
    hand_theta = [[0.0, -pi / 2, pi, -3 * pi / 4],          # from node 0 (0,0): to (1,0) d=(-1,0); to (0,1) d=(0,-1); to (1,1) d=(-1,-1)
                  [pi / 2, 0.0, 3 * pi / 4, pi],            # from node 1 (1,0): d=(1,0); -; (1,-1); (0,-1)
                  [0.0, -pi / 4, 0.0, -pi / 2],             # from node 2 (0,1): d=(0,1); (-1,1); -; (-1,0)
                  [pi / 4, 0.0, pi / 2, 0.0]]               # from node 3 (1,1): d=(1,1); (0,1); (1,0); -
    r2 = math.sqrt(2.0)
    hand_rho = [[0, 1, 1, r2], [1, 0, r2, 1], [1, r2, 0, 1], [r2, 1, 1, 0]]
    assert np.allclose(kat["rho"], hand_rho, atol=1e-15) and np.allclose(kat["theta"], hand_theta, atol=1e-15)
    pseudo = O.polar_pseudo_coordinates(torch.tensor([c], dtype=torch.float64))
    assert np.allclose(pseudo[0, :, :, 0].numpy(), hand_rho, atol=1e-15) and np.allclose(pseudo[0, :, :, 1].numpy(), hand_theta, atol=1e-15)

    # ---- Gaussian patch weights, unit precisions: w_k ~ exp(-(rho - mu_rho_k)^2 / 2) * exp(-ang(theta, mu_theta_k)^2 / 2), normalised over k
    g = kat["gauss"]

    def hand_w(rho, theta):
        raw = []
        for mr, mt in zip(g["mean_rho"], g["mean_theta"]):
            a = abs(theta - mt)
            a = min(a, abs(2 * pi - a))                     # wrap-around (layers.py:114-116)
            raw.append(math.exp(-0.5 * (rho - mr) ** 2) * math.exp(-0.5 * a * a))
        s = sum(raw)
        return [x / s for x in raw]

    hand = [[hand_w(hand_rho[i][j], hand_theta[i][j]) for j in range(4)] for i in range(4)]
    assert np.allclose(kat["gaussian_weights"], hand, rtol=1e-12)
    # one value spelled out: edge 1 <- 0 has rho = 1, theta = pi/2 -> raw = (e^{-1/2 - pi^2/8}, 1, e^{-pi^2/8})
    e1, e2 = math.exp(-0.5 - pi * pi / 8), math.exp(-pi * pi / 8)
    assert np.allclose(kat["gaussian_weights"][1][0], [e1 / (e1 + 1 + e2), 1 / (e1 + 1 + e2), e2 / (e1 + 1 + e2)], rtol=1e-12)
    # wrap-around: edge 0 <- 3 has theta = -3pi/4; against mu_theta = pi the angle is pi/4, not 7pi/4
    # (against mu_theta = pi/2 it is 3pi/4, not 5pi/4); both kernels have mu_rho = 1, so w_2 / w_1 = e^{((3pi/4)^2 - (pi/4)^2) / 2} = e^{pi^2/4}
    assert math.isclose(kat["gaussian_weights"][0][3][2] / kat["gaussian_weights"][0][3][1], math.exp(pi * pi / 4), rel_tol=1e-12)
    p = {"gc." + k: torch.tensor(v, dtype=torch.float64).view(3, 1) for k, v in g.items()}
    w = O.gaussian_kernel_weights(pseudo, p, "gc").view(4, 4, 3)
    assert np.allclose(w.numpy(), hand, rtol=1e-12)
    for i in range(4):
        for j in range(4):
            one = O.np_edge_kernel_weights(np.array(c[i]), np.array(c[j]), np.array(g["mean_rho"]), np.array(g["precision_rho"]),
                                           np.array(g["mean_theta"]), np.array(g["precision_theta"]))
            assert np.allclose(one, hand[i][j], rtol=1e-12)

    # ---- top-2 + softmax: rows whose two largest entries differ by ln 2, ln 3, ln 4, ln 1.5 -> (1/3, 2/3), (1/4, 3/4), (1/5, 4/5), (2/5, 3/5)
    assert kat["topk_index_sets"] == [[1, 3], [0, 2], [1, 3], [0, 1]]
    hand_alpha = [[1 / 3, 2 / 3], [1 / 4, 3 / 4], [1 / 5, 4 / 5], [3 / 5, 2 / 5]]
    assert np.allclose(kat["topk_softmax_by_index"], hand_alpha, rtol=1e-12)
    adj = torch.tensor([kat["adjacency"]], dtype=torch.float64)
    alpha, idx = O.select_neighbourhood(adj, 2)
    for i in range(4):
        pairs = sorted(zip(idx[0, i].tolist(), alpha[0, i].tolist()))
        assert [a for a, _ in pairs] == kat["topk_index_sets"][i] and np.allclose([v for _, v in pairs], hand_alpha[i], rtol=1e-12)
        order, sm = O.np_topk_softmax(np.array(kat["adjacency"][i]), 2)
        assert order.tolist() == kat["topk_index_sets"][i] and np.allclose(sm, hand_alpha[i], rtol=1e-12)
