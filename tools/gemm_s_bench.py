"""Split-bf16 GEMM vs TF32x3 at the VQA2 B=512 shapes (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(name, fn, flops, iters=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:48s} {t:9.1f} us  {flops / t / 1e6:8.1f} TF/s", flush=True)
M, F = 18432, 2052
torch.manual_seed(0)
X = torch.randn(M, F, device=dev).clamp_(min=0); W1 = torch.randn(2048, F, device=dev) * 0.02
W2 = torch.randn(1024, 2048, device=dev) * 0.02; G1 = torch.randn(M, 2048, device=dev); dY1 = torch.randn(M, 2048, device=dev); dY2 = torch.randn(M, 1024, device=dev)
Xs, W1s, W2s, G1s, dY1s, dY2s = (kn.split(t) for t in (X, W1, W2, G1, dY1, dY2))
timeit("split X (151 MB fp32 -> 2 planes)", lambda: kn.split(X), 1)
ref = (X.double()[:512] @ W1.double().t())
for passes in (3, 1):
    for bn in (128, 256):
        timeit(f"Y1 = X.W1^T  passes={passes} bn={bn}", lambda: kn.gemm_s(Xs, W1s, passes=passes, tile_n=bn), 2 * M * 2048 * F)
    timeit(f"Y1 = X.W1^T  passes={passes} bn=256 no pairs", lambda: kn.gemm_s(Xs, W1s, passes=passes, tile_n=256, cluster=False), 2 * M * 2048 * F)
    out = kn.gemm_s(Xs, W1s, passes=passes)
    print("   rel err vs fp64 (first 512 rows):", ((out[:512].double() - ref).abs().max() / ref.abs().max()).item())
    timeit(f"Y2 = G1.W2^T passes={passes}", lambda: kn.gemm_s(G1s, W2s, passes=passes), 2 * M * 1024 * 2048)
    timeit(f"dW1 = dY1^T.X passes={passes}", lambda: kn.gemm_s(dY1s, Xs, a_mn=True, b_mn=True, passes=passes), 2 * M * 2048 * F)
    timeit(f"dG1 = dY2.W2 passes={passes}", lambda: kn.gemm_s(dY2s, W2s, b_mn=True, passes=passes), 2 * M * 2048 * 1024)
    timeit(f"dW2 = dY2^T.G1 passes={passes} split2", lambda: kn.gemm_s(dY2s, G1s, a_mn=True, b_mn=True, passes=passes, split_k=2), 2 * M * 2048 * 1024)
timeit("tf32x3 Y1 (old kernel) bn256", lambda: kn.gemm(X, W1, tile_n=256), 2 * M * 2048 * F)
