"""Drop-in ``sparse_graph_model`` module: the conditioned-graph VQA ``Model`` on B200 kernels.

Keeps the reference's constructor signature, attribute names, ``state_dict`` keys and
``forward(question, image, K, qlen) -> (logits, adjacency_matrix, h_max_indices)`` contract
(reference sparse_graph_model.py:28-159), so ``run.py`` / ``run_imageclef.py`` / ``run_mimic.py`` /
``plot.py`` import it unchanged when ``vqa-project_b200/`` precedes the reference on ``sys.path``.

What differs is underneath: the question encoder (embedding + GRU) is one autograd node (``ops.QuestionEncoderFn``) and everything
after it is ONE more (``vqa_b200.ops.ConditionedGraphFn``) made of hand-written sm_100a kernels.
"""
import torch
import torch.nn as nn

from layers import NeighbourhoodGraphConvolution as GraphConvolution
from layers import GraphLearner, WeightNormLinear
from vqa_b200 import ops


class Model(nn.Module):

    def __init__(self, vocab_size, emb_dim, feat_dim, hid_dim, out_dim, pretrained_wemb, dropout,
                 n_kernels=8, neighbourhood_size=9, n_obj=30):
        super().__init__()
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.feat_dim = feat_dim
        self.hid_dim = hid_dim
        self.out_dim = out_dim
        self.neighbourhood_size = neighbourhood_size
        self.n_kernels = n_kernels

        # question encoder (unchanged components)
        self.wembed = nn.Embedding(vocab_size, emb_dim)
        with torch.no_grad():
            self.wembed.weight.copy_(torch.as_tensor(pretrained_wemb))
        self.q_gru = nn.GRU(input_size=emb_dim, hidden_size=hid_dim)

        # graph learner over [image features || question encoding]
        self.adjacency_1 = GraphLearner(in_feature_dim=feat_dim + hid_dim, combined_feature_dim=512,
                                        n_obj=n_obj, dropout=dropout)
        self.dropout = nn.Dropout(p=dropout)
        self.dropout_q = nn.Dropout(p=dropout / 2)   # defined but unused, as in the reference

        self.graph_convolution_1 = GraphConvolution(feat_dim, hid_dim * 2, n_kernels, 2)
        self.graph_convolution_2 = GraphConvolution(hid_dim * 2, hid_dim, n_kernels, 2)

        self.out_1 = WeightNormLinear(hid_dim, out_dim)
        self.out_2 = WeightNormLinear(out_dim, out_dim)

    def encode_question(self, question, qlen):
        """Embedding + GRU final state (reference :117-121) through ``ops.QuestionEncoderFn``.  ``qlen`` is the
        reference's list of lengths (ints or 0-d tensors), or -- for CUDA-graph capture, where nothing may depend on
        host data -- a CUDA int32 tensor (B,), in which case the number of steps is ``self.max_question_len`` (set it
        to the dataset's maximum, 14 for VQA2) or the question width."""
        if torch.is_tensor(qlen) and qlen.is_cuda:
            qlen_dev = qlen.to(torch.int32)
            steps = int(getattr(self, "max_question_len", 0) or question.shape[1])
        else:
            lens = [int(x) for x in qlen]
            steps = max(lens)
            if min(lens) < 1:
                raise ValueError("question lengths must be >= 1 (pack_padded_sequence rejects empty sequences)")
            qlen_dev = torch.tensor(lens, dtype=torch.int32).to(question.device, non_blocking=True)
        g = self.q_gru
        return ops.QuestionEncoderFn.apply(question, qlen_dev, steps, self.wembed.weight, g.weight_ih_l0, g.weight_hh_l0,
                                           g.bias_ih_l0, g.bias_hh_l0)

    def forward(self, question, image, K, qlen):
        """question (B,T) int64, image (B,K,F) float32 whose last 4 columns are xyxy boxes, K (B,1) (all equal),
        qlen list of lengths -> logits (B,out_dim), adjacency (B,K,K), h_max_indices (B,hid_dim) int64."""
        n_nodes = image.size(1)   # the reference syncs on K[0]; the shape carries the same number without a D2H copy
        if n_nodes != self.adjacency_1.n_obj:
            raise ValueError(f"Model was built with n_obj={self.adjacency_1.n_obj} but the batch has {n_nodes} nodes per image")
        if self.neighbourhood_size > n_nodes:
            raise ValueError(f"neighbourhood_size={self.neighbourhood_size} exceeds the {n_nodes} nodes per image")
        if image.size(2) != self.feat_dim:
            raise ValueError(f"expected feat_dim={self.feat_dim}, got {image.size(2)}")
        qenc = self.encode_question(question, qlen)

        gl, gc1, gc2 = self.adjacency_1, self.graph_convolution_1, self.graph_convolution_2
        cfg = dict(n_kernels=self.n_kernels, neighbourhood_size=self.neighbourhood_size,
                   dropout=self.dropout.p, training=self.training)
        logits, adjacency_matrix, h_max_indices = ops.ConditionedGraphFn.apply(
            cfg, image, qenc,
            gl.edge_layer_1.weight_v, gl.edge_layer_1.weight_g, gl.edge_layer_1.bias,
            gl.edge_layer_2.weight_v, gl.edge_layer_2.weight_g, gl.edge_layer_2.bias,
            *gc1.gaussian_parameters(), *gc2.gaussian_parameters(),
            self.out_1.weight_v, self.out_1.weight_g, self.out_1.bias,
            self.out_2.weight_v, self.out_2.weight_g, self.out_2.bias,
            *gc1.conv_weight_list(), *gc2.conv_weight_list())
        return logits, adjacency_matrix, h_max_indices
