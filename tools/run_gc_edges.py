import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
B, K, F, nb, nk = 512, 36, 2052, 16, 8
M = B * K
torch.manual_seed(0)
img = torch.rand(B, K, F, device=dev); gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
h = torch.randn(B, K, 512, device=dev).clamp_(min=0)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
Y1s = kn.split(torch.randn(M, 2048, device=dev)); dO1s = kn.split(torch.randn(M, 2048, device=dev))
for _ in range(3):
    kn.graphconv_bwd_edges_s(Y1s, idx, alpha, img, gauss, B, K, dOs=dO1s)
torch.cuda.synchronize(); print("ok")
