"""Drop-in ``layers`` module: GraphLearner and NeighbourhoodGraphConvolution on sm_100a kernels.

Same class names, constructor arguments, public methods and ``state_dict`` keys as the reference's
``layers.py`` (GraphLearner: reference layers.py:147-197, NeighbourhoodGraphConvolution: :24-144), so
``from layers import NeighbourhoodGraphConvolution as GraphConvolution, GraphLearner`` keeps working and
reference checkpoints load unchanged.  The arithmetic runs in ``libvqa_sm100.so`` (tcgen05 GEMMs, fused
adjacency kernel, Gaussian-weight kernel); there is no CPU path.

``Model.forward`` does not go through these ``forward`` methods: it hands the parameters held here to the
fused operator ``vqa_b200.ops.ConditionedGraphFn``.  The methods below serve callers that use the layers
on materialised neighbourhoods, exactly like the reference API.
"""
import math

import torch
import torch.nn as nn

from vqa_b200 import ops


class WeightNormLinear(nn.Module):
    """Linear layer in old-style weight-norm parametrisation: parameters ``bias``, ``weight_g`` (out,1),
    ``weight_v`` (out,in) -- the keys ``nn.utils.weight_norm(nn.Linear(...))`` produces in the reference."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        seed_layer = nn.Linear(in_features, out_features)          # same init stream as the reference
        self.bias = nn.Parameter(seed_layer.bias.detach().clone())
        v = seed_layer.weight.detach().clone()
        self.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(v)

    def effective_weight(self):
        return ops.WeightNormFn.apply(self.weight_v, self.weight_g)

    def forward(self, x, relu=False):
        lead = x.shape[:-1]
        y = ops.LinearFn.apply(x.reshape(-1, self.in_features), self.effective_weight(), self.bias, relu)
        return y.view(*lead, self.out_features)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, weight_norm=True"


class GraphLearner(nn.Module):
    """Question-conditioned adjacency: A = h h^T, h = relu(L2(relu(L1(nodes))))."""

    def __init__(self, in_feature_dim, combined_feature_dim, n_obj, dropout=0.0):
        super().__init__()
        self.in_dim = in_feature_dim
        self.combined_dim = combined_feature_dim
        self.n_obj = n_obj
        self.edge_layer_1 = WeightNormLinear(in_feature_dim, combined_feature_dim)
        self.edge_layer_2 = WeightNormLinear(combined_feature_dim, combined_feature_dim)
        self.dropout = nn.Dropout(p=dropout)     # constructed but never applied, as in the reference (layers.py:170)

    def forward(self, graph_nodes):
        """graph_nodes (B, K, in) -> adjacency (B, K, K)"""
        if graph_nodes.size(-2) != self.n_obj:
            raise ValueError(f"GraphLearner was built for n_obj={self.n_obj} nodes, got {graph_nodes.size(-2)}")
        h = self.edge_layer_1(graph_nodes, relu=True)
        h = self.edge_layer_2(h, relu=True)
        return ops.AdjacencyFn.apply(h.view(-1, self.n_obj, self.combined_dim))


class NeighbourhoodGraphConvolution(nn.Module):
    """MoNet-style convolution over fixed-size neighbourhoods with n_kernels Gaussian kernels in polar
    pseudo-coordinates; out_feat_dim is split evenly over the kernels."""

    def __init__(self, in_feat_dim, out_feat_dim, n_kernels, coordinate_dim, bias=False):
        super().__init__()
        self.n_kernels = n_kernels
        self.coordinate_dim = coordinate_dim
        self.in_feat_dim = in_feat_dim
        self.out_feat_dim = out_feat_dim
        self.bias = bias
        self.conv_weights = nn.ModuleList(
            [nn.Linear(in_feat_dim, out_feat_dim // n_kernels, bias=bias) for _ in range(n_kernels)])
        self.mean_rho = nn.Parameter(torch.empty(n_kernels, 1))
        self.mean_theta = nn.Parameter(torch.empty(n_kernels, 1))
        self.precision_rho = nn.Parameter(torch.empty(n_kernels, 1))
        self.precision_theta = nn.Parameter(torch.empty(n_kernels, 1))
        self.init_parameters()
        self.flatten_parameters()

    def init_parameters(self):
        with torch.no_grad():
            self.mean_theta.uniform_(-math.pi, math.pi)
            self.mean_rho.uniform_(0.0, 1.0)
            self.precision_theta.uniform_(0.0, 1.0)
            self.precision_rho.uniform_(0.0, 1.0)

    # --- one contiguous (out, in) weight matrix behind the per-kernel Parameters -------------------------
    def flatten_parameters(self):
        """Re-point the nk ``conv_weights[k].weight`` at consecutive slices of one buffer so the projection is a
        single GEMM with no gather copy.  Idempotent; called again after ``.to()/.cuda()`` split them."""
        ws = [lin.weight for lin in self.conv_weights]
        d, fin = ws[0].shape
        step = d * fin * ws[0].element_size()
        if all(w.is_contiguous() and w.data_ptr() == ws[0].data_ptr() + i * step for i, w in enumerate(ws)):
            return
        with torch.no_grad():
            flat = torch.cat([w.detach() for w in ws], dim=0).contiguous()
            for i, w in enumerate(ws):
                w.data = flat[i * d:(i + 1) * d]

    def _apply(self, fn, *args, **kwargs):
        """``.to()/.cuda()/.float()`` give every Parameter storage of its own: re-join them at once, so their addresses are final
        before a gradient reducer / optimiser records them."""
        out = super()._apply(fn, *args, **kwargs)
        self.flatten_parameters()
        return out

    def conv_weight_list(self):
        self.flatten_parameters()
        return [lin.weight for lin in self.conv_weights]

    def gaussian_parameters(self):
        return self.mean_rho, self.precision_rho, self.mean_theta, self.precision_theta

    # --- reference API -----------------------------------------------------------------------------------
    def forward(self, neighbourhood_features, neighbourhood_pseudo_coord):
        """(B,K,nb,in), (B,K,nb,2) -> (B,K,out)"""
        bsz, k, nb = neighbourhood_features.shape[:3]
        weights = self.get_gaussian_weights(neighbourhood_pseudo_coord).view(bsz * k, nb, self.n_kernels)
        out = self.convolution(neighbourhood_features.reshape(bsz * k, nb, -1), weights)
        return out.view(-1, k, self.out_feat_dim)

    def get_gaussian_weights(self, pseudo_coord):
        """(B,K,nb,2) -> (B*K*nb, n_kernels), normalised over the kernel axis."""
        return ops.GaussianWeightsFn.apply(pseudo_coord.contiguous(), *self.gaussian_parameters())

    def convolution(self, neighbourhood, weights):
        """(B*K, nb, in), (B*K, nb, nk) -> (B*K, out), in the reference's own order (layers.py:127-144): the patch operator
        Z[:, k] = sum_m weights[:, m, k] * neighbourhood[:, m] as one HBM-bound kernel (the reference's ``torch.bmm``), then the
        k-th bias-free linear map on Z[:, k] as a tcgen05 GEMM reading its slice of Z in place."""
        n, nb, fin = neighbourhood.shape
        nk, d = self.n_kernels, self.out_feat_dim // self.n_kernels
        z = ops.PatchOperatorFn.apply(neighbourhood, weights)                  # (n, nk, in)
        wb = [lin.weight for lin in self.conv_weights] + ([lin.bias for lin in self.conv_weights] if self.bias else [])
        return ops.PerKernelLinearFn.apply(z, nk, bool(self.bias), *wb)
