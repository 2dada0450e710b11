"""Tensor-core graph-conv kernels at VQA2 B=512 shapes: time + HBM fraction (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(name, fn, nbytes, iters=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize(); torch.cuda._sleep(400000); flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)   # host runs ahead: no launch gap inside the bracket
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {t:8.1f} us  {nbytes / t / 1e3:7.1f} GB/s  {nbytes / t / 1e3 / 6544:5.3f} of HBM peak", flush=True)
B, K, F, nb, nk = 512, 36, 2052, 16, 8
M = B * K
torch.manual_seed(0)
img = torch.rand(B, K, F, device=dev); gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
h = torch.randn(B, K, 512, device=dev).clamp_(min=0)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
Y1s = kn.split(torch.randn(M, 2048, device=dev)); Y2s = kn.split(torch.randn(M, 1024, device=dev)); q = torch.randn(B, 1024, device=dev)
b1 = 2 * M * 2048 * 4 + 2 * M * nb * 4 + M * 16
ec1 = kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K); ec2 = kn.graphconv_edge_coef(idx, None, img, gauss, B, K)
timeit("edge coefficients (one layer)", lambda: kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K), M * nb * (4 + 4 + 4 * nk + 4))
timeit("mma fwd L1 relu", lambda: kn.graphconv_fwd_s(Y1s, idx, alpha, img, gauss, B, K, ec=ec1), b1)
timeit("mma fwd L1 relu+dropout", lambda: kn.graphconv_fwd_s(Y1s, idx, alpha, img, gauss, B, K, dropout_p=0.5, seed=1, offset=1, ec=ec1), b1)
timeit("mma fwd L1 relu (bf16 planes)", lambda: kn.graphconv_fwd_s(kn.SplitT(Y1s.hi, None, Y1s.rows, Y1s.cols, Y1s.ld), idx, alpha, img, gauss, B, K, ec=ec1), b1 // 2)
timeit("mma pool fwd L2", lambda: kn.graphconv_pool_fwd_s(Y2s, idx, img, gauss, q, B, K, ec=ec2), M * 1024 * 4 + M * nb * 4 + M * 16 + 4 * B * 1024 * 4)
timeit("mma bwd data L1", lambda: kn.graphconv_bwd_data_s(Y1s, idx, alpha, img, gauss, B, K, ec=ec1), b1)
timeit("mma bwd data L2", lambda: kn.graphconv_bwd_data_s(Y2s, idx, None, img, gauss, B, K, ec=ec2), 2 * M * 1024 * 4 + M * nb * 4 + M * 16)
dO1s = kn.split(torch.randn(M, 2048, device=dev))
timeit("mma bwd edges L1 (P + edge finish)", lambda: kn.graphconv_bwd_edges_s(Y1s, idx, alpha, img, gauss, B, K, dOs=dO1s), 2 * M * 2048 * 4 + 3 * M * nb * 4)
pooled, arg, hq = kn.graphconv_pool_fwd_s(Y2s, idx, img, gauss, q, B, K)
dp = torch.randn(B, 1024, device=dev)
timeit("mma bwd edges L2 (pooled upstream)", lambda: kn.graphconv_bwd_edges_s(Y2s, idx, None, img, gauss, B, K, dpooled=dp, argmax=arg), M * 1024 * 4 + M * nb * 4)
timeit("pool bwd data L2 (scatter coef * dpooled)", lambda: kn.graphconv_pool_bwd_data_s(dp, arg, idx, ec2, B, K, 1024), M * 1024 * 4 + 2 * B * 1024 * 4)
