"""Tile / split-K sweep for the two per-step GRU products (M = B = 512)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
B, H = 512, 1024
h = kn.split(torch.randn(B, H, device=dev)); W = kn.split(torch.randn(3 * H, H, device=dev) * 0.03); dg = kn.split(torch.randn(B, 3 * H, device=dev))
out_f = torch.zeros(B, 3 * H, device=dev); out_b = torch.zeros(B, H, device=dev)
for bn in (64, 128, 256):
    for sk in (1, 2, 4, 8):
        tf = timeit(lambda: kn.gemm_s(h, W, out=out_f, accumulate=True, split_k=sk, tile_n=bn))
        tb = timeit(lambda: kn.gemm_s(dg, W, b_mn=True, out=out_b, accumulate=True, split_k=sk, tile_n=bn))
        print(f"bn={bn:3d} split_k={sk}:  fwd h.Whh^T {tf:6.1f} us   bwd dg.Whh {tb:6.1f} us", flush=True)
