"""One launch set of the split-bf16 GEMM on the layer-1 projection shape (ncu target): Y1 = X . Wc1^T, M=18432, N=2048, K=2052."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
torch.manual_seed(0)
M, N, K = 18432, 2048, 2052
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Xs = kn.split(torch.randn(M, K, device=dev).clamp_(min=0)); Ws = kn.split(torch.randn(N, K, device=dev) * 0.02)
out = kn.empty_split(M, N, dev, passes == 3)
for _ in range(3):
    kn.gemm_s(Xs, Ws, out_split=out, want_f32=False, passes=passes)
torch.cuda.synchronize()
print("ok")
