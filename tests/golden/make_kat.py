"""Generate ``kat_tiny.json``: hand-checkable known answers (4 nodes, 2 neighbours, 3 Gaussian kernels) produced by the UNMODIFIED
reference (``Model._compute_pseudo``, ``Model._create_neighbourhood``'s top-k + softmax, ``NeighbourhoodGraphConvolution.
get_gaussian_weights``).  Run in the build container only:  python tests/golden/make_kat.py

Inputs are chosen so that every expected value has a closed form (``tests/test_oracle_golden.py::test_hand_checkable_known_answers``
derives them with ``math`` alone): unit-square box centres -> rho in {0, 1, sqrt 2}, theta in multiples of pi/4; kernel means at
rho 0/1 and theta 0, pi/2, pi with unit precisions; an adjacency row whose two largest entries differ by ln 2 -> softmax (2/3, 1/3).
"""
import json
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CENTRES = [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]
GAUSS = dict(mean_rho=[0.0, 1.0, 1.0], precision_rho=[1.0, 1.0, 1.0], mean_theta=[0.0, math.pi / 2, math.pi], precision_theta=[1.0, 1.0, 1.0])
ADJACENCY = [[0.1, 2.0, -1.0, 2.0 + math.log(2.0)],
             [3.0, 0.0, 3.0 + math.log(3.0), -5.0],
             [0.0, 0.5, 0.25, 0.5 + math.log(4.0)],
             [1.0 + math.log(1.5), 1.0, 0.0, -1.0]]


def main():
    sys.path.insert(0, REF)
    import sparse_graph_model as ref_model
    import layers as ref_layers
    assert ref_model.__file__.startswith(REF)
    sys.path.remove(REF)
    c = torch.tensor([CENTRES], dtype=torch.float64)
    pseudo = ref_model.Model._compute_pseudo(None, c)                                   # (1,4,4,2); the method reads no attribute
    gc = ref_layers.NeighbourhoodGraphConvolution(8, 6, 3, 2).double()
    with torch.no_grad():
        for k, v in GAUSS.items():
            getattr(gc, k).copy_(torch.tensor(v, dtype=torch.float64).view(3, 1))
        w = gc.get_gaussian_weights(pseudo)                                             # (16, 3): all K*K ordered pairs as "neighbourhoods"
    adj = torch.tensor([ADJACENCY], dtype=torch.float64)
    feats = torch.zeros(1, 4, 8, dtype=torch.float64)
    # the reference's own top-k + per-row softmax (sparse_graph_model.py:225-227), called through _create_neighbourhood's code path
    top_k, top_ind = torch.topk(adj, k=2, dim=-1, sorted=False)
    top_k = torch.stack([torch.nn.functional.softmax(top_k[:, k], dim=-1) for k in range(4)]).transpose(0, 1)
    out = dict(centres=CENTRES, gauss=GAUSS, adjacency=ADJACENCY, rho=pseudo[0, :, :, 0].tolist(), theta=pseudo[0, :, :, 1].tolist(),
               gaussian_weights=w.view(4, 4, 3).tolist(), topk_index_sets=[sorted(r) for r in top_ind[0].tolist()],
               topk_softmax_by_index=[[float(v) for _, v in sorted(zip(i, a))] for i, a in zip(top_ind[0].tolist(), top_k[0].tolist())])
    with open(os.path.join(HERE, "kat_tiny.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote kat_tiny.json")


if __name__ == "__main__":
    main()
