"""One launch set of the persistent aggregate (ncu target): python tools/run_agg_one.py [fwd|pool|bwd]"""
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
Y1s = kn.split(torch.randn(M, 2048, device=dev)); Y2s = kn.split(torch.randn(M, 1024, device=dev)); q = torch.randn(B, 1024, device=dev)
ec1 = kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K); ec2 = kn.graphconv_edge_coef(idx, None, img, gauss, B, K)
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
for _ in range(2):
    if mode == "fwd": kn.graphconv_fwd_s(Y1s, idx, alpha, img, gauss, B, K, dropout_p=0.5, seed=1, offset=1, ec=ec1)
    elif mode == "pool": kn.graphconv_pool_fwd_s(Y2s, idx, img, gauss, q, B, K, ec=ec2)
    else: kn.graphconv_bwd_data_s(Y1s, idx, alpha, img, gauss, B, K, ec=ec1)
torch.cuda.synchronize()
print("ok")
