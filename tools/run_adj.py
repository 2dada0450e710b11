"""One launch each of the adjacency / top-k forward and backward kernels at a BASELINE shape (for ncu captures).
    python tools/run_adj.py [vqa2|med|k100]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
B, K, nb = {"vqa2": (512, 36, 16), "med": (512, 51, 19), "k100": (1024, 100, 32)}[sys.argv[1] if len(sys.argv) > 1 else "vqa2"]
torch.manual_seed(0)
h = torch.randn(B, K, 512, device="cuda").clamp_(min=0)
for _ in range(2):
    adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
    dh = kn.adjacency_topk_bwd(h, idx, alpha, torch.randn_like(alpha))
torch.cuda.synchronize()
print("ok", float(adj.sum()), float(dh.abs().sum()))
