"""Launch the layer-1 / layer-2 graph-conv kernels a few times at VQA2 B=512 shapes (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
B, K, F, H, C, nb, nk = 512, 36, 2052, 1024, 512, 16, 8
M = B * K
torch.manual_seed(0)
img = torch.rand(B, K, F, device=dev)
gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
h = torch.randn(B, K, C, device=dev).clamp_(min=0)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
Y1 = torch.randn(M, 2048, device=dev); Y2 = torch.randn(M, 1024, device=dev); q = torch.randn(B, 1024, device=dev)
dO1 = torch.randn(M, 2048, device=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
for _ in range(3):
    if which == "fwd":
        kn.graphconv_fwd(Y1, idx, alpha, img, gauss, B, K, dropout_p=0.5, seed=1, offset=1)
        kn.graphconv_pool_fwd(Y2, idx, img, gauss, q, B, K)
    elif which == "bwd":
        kn.graphconv_bwd(Y1, idx, alpha, img, gauss, B, K, dO=dO1)
    elif which == "adj":
        kn.adjacency_topk_fwd(h, nb); kn.adjacency_topk_bwd(h, idx, alpha, torch.randn_like(alpha))
torch.cuda.synchronize()
print("ok")
