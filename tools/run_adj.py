"""A few launches of the adjacency / top-k kernels at VQA2 B=512 shapes (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
B, K, C, nb = 512, 36, 512, 16
torch.manual_seed(0)
h = torch.randn(B, K, C, device=dev).clamp_(min=0)
for _ in range(3):
    adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
    dalpha = torch.randn_like(alpha)
    dh = kn.adjacency_topk_bwd(h, idx, alpha, dalpha, None)
torch.cuda.synchronize()
print("ok")
