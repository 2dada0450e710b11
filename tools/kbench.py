"""Per-kernel timing at the VQA2 B=512 shapes (CUDA events, L2 flushed between iterations).
Usage: python tools/kbench.py [--quick]   -> prints one line per kernel and writes gpurun_out/kbench.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
PEAK_HBM = 6544.0
res = []


def timeit(name, fn, nbytes=None, flops=None, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    r = dict(name=name, us=round(t, 2))
    if nbytes: r["GBs"] = round(nbytes / t / 1e3, 1); r["hbm_frac"] = round(nbytes / t / 1e3 / PEAK_HBM, 3)
    if flops: r["TFs"] = round(flops / t / 1e6, 1)
    print(r, flush=True)
    res.append(r)


B, K, F, H, C, nb, nk, A = 512, 36, 2052, 1024, 512, 16, 8, 3000
M = B * K
torch.manual_seed(0)
X = torch.randn(M, F, device=dev).clamp_(min=0)
img = X.view(B, K, F)
img[..., -4:] = torch.rand(B, K, 4, device=dev)
W1 = torch.randn(2048, F, device=dev) * 0.02
W2 = torch.randn(1024, 2048, device=dev) * 0.02
Wg = torch.randn(512, F + H, device=dev) * 0.02
gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1

for prec, pn in ((0, "tf32x3"), (1, "tf32")):
    for bn in (128, 256):
        timeit(f"gemm Y1=X.W1^T {pn} bn{bn}", lambda: kn.gemm(X, W1, precision=prec, tile_n=bn), flops=2 * M * 2048 * F)
    timeit(f"gemm GL1 {pn}", lambda: kn.gemm(X, Wg[:, :F], precision=prec, relu=True), flops=2 * M * 512 * F)
G1 = torch.randn(M, 2048, device=dev)
timeit("gemm Y2=G1.W2^T tf32x3", lambda: kn.gemm(G1, W2), flops=2 * M * 1024 * 2048)
dY1 = torch.randn(M, 2048, device=dev)
timeit("gemm dW1=dY1^T.X tf32x3", lambda: kn.gemm(dY1, X, a_mn=True, b_mn=True), flops=2 * M * 2048 * F)
dY2 = torch.randn(M, 1024, device=dev)
timeit("gemm dG1=dY2.W2 tf32x3", lambda: kn.gemm(dY2, W2, b_mn=True), flops=2 * M * 2048 * 1024)
timeit("gemm dW2=dY2^T.G1 tf32x3 split2", lambda: kn.gemm(dY2, G1, a_mn=True, b_mn=True, split_k=2), flops=2 * M * 2048 * 1024)
timeit("torch fp32 matmul Y1 (cuBLAS, no tf32)", lambda: X @ W1.t(), flops=2 * M * 2048 * F)

h = torch.randn(B, K, C, device=dev).clamp_(min=0)
timeit("adjacency_topk_fwd", lambda: kn.adjacency_topk_fwd(h, nb), nbytes=M * C * 4 + M * K * 4 + 2 * M * nb * 4)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
dalpha = torch.randn_like(alpha)
timeit("adjacency_topk_bwd", lambda: kn.adjacency_topk_bwd(h, idx, alpha, dalpha), nbytes=2 * M * C * 4 + 3 * M * nb * 4)
Y1 = torch.randn(M, 2048, device=dev)
Y2 = torch.randn(M, 1024, device=dev)
q = torch.randn(B, 1024, device=dev)
timeit("graphconv_fwd L1 (relu)", lambda: kn.graphconv_fwd(Y1, idx, alpha, img, gauss, B, K), nbytes=2 * M * 2048 * 4 + 2 * M * nb * 4 + M * 8)
timeit("graphconv_fwd L1 (relu+dropout)", lambda: kn.graphconv_fwd(Y1, idx, alpha, img, gauss, B, K, dropout_p=0.5, seed=1, offset=1), nbytes=2 * M * 2048 * 4 + 2 * M * nb * 4 + M * 8)
timeit("graphconv_pool_fwd L2", lambda: kn.graphconv_pool_fwd(Y2, idx, img, gauss, q, B, K), nbytes=M * 1024 * 4 + M * nb * 4 + M * 8)
dO1 = torch.randn(M, 2048, device=dev)
timeit("graphconv_bwd L1", lambda: kn.graphconv_bwd(Y1, idx, alpha, img, gauss, B, K, dO=dO1), nbytes=3 * M * 2048 * 4 + 3 * M * nb * 4)
pooled, arg, hq = kn.graphconv_pool_fwd(Y2, idx, img, gauss, q, B, K)
dp = torch.randn(B, 1024, device=dev)
timeit("graphconv_bwd L2 (pooled)", lambda: kn.graphconv_bwd(Y2, idx, None, img, gauss, B, K, dpooled=dp, argmax=arg), nbytes=2 * M * 1024 * 4 + M * nb * 4)
timeit("dropout image", lambda: kn.dropout(X, 0.5, 1, 2), nbytes=2 * M * F * 4)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w"), indent=1)
