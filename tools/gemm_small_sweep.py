"""Tile / split-K sweep of the small-M and short-K products of the step (plain epilogue), L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize(); torch.cuda._sleep(400000); flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
shapes = [("o1   512x3000x1024", 512, 3000, 1024, False, False), ("logit 512x3000x3000", 512, 3000, 3000, False, False),
          ("do1  512x3000x3000 (B mn)", 512, 3000, 3000, False, True), ("dhq  512x1024x3000 (B mn)", 512, 1024, 3000, False, True),
          ("GL2  18432x512x512", 18432, 512, 512, False, False), ("GI   7168x3072x300", 7168, 3072, 300, False, False),
          ("dE   7168x300x3072 (B mn)", 7168, 300, 3072, False, True), ("dWhh 3072x1024x7168 (A,B mn)", 3072, 1024, 7168, True, True)]
for name, M, N, K, amn, bmn in shapes:
    a = torch.randn((K, M) if amn else (M, K), device=dev); b = torch.randn((K, N) if bmn else (N, K), device=dev)
    As, Bs = kn.split(a), kn.split(b)
    res = []
    for tile in (64, 128, 256):
        for sk in (1, 2, 3, 4, 6):
            if sk > 1 and K // 64 < 2 * sk: continue
            out = torch.zeros(M, N, device=dev)
            try:
                t = timeit(lambda: kn.gemm_s(As, Bs, a_mn=amn, b_mn=bmn, out=out, tile_n=tile, split_k=sk))
            except Exception as e:
                continue
            res.append((t, tile, sk))
    auto = timeit(lambda: kn.gemm_s(As, Bs, a_mn=amn, b_mn=bmn))
    res.sort()
    fl = 2.0 * M * N * K
    print(f"{name:32s} auto {auto:6.1f} us ({fl / auto / 1e6:5.0f} TF/s) | best " + "  ".join(f"bn{tl}/k{sk}: {t:.1f}" for t, tl, sk in res[:4]), flush=True)
