"""Is the aggregate kernel DRAM-pattern bound?  Time it with inputs L2-resident (small batch, no flush) vs flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
K, F, nb, nk = 36, 2052, 16, 8
for B in (64, 128, 512):
    M = B * K
    torch.manual_seed(0)
    img = torch.rand(B, K, F, device=dev); gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
    h = torch.randn(B, K, 512, device=dev).clamp_(min=0)
    adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
    Ys = kn.split(torch.randn(M, 2048, device=dev))
    fn = lambda: kn.graphconv_fwd_s(Ys, idx, alpha, img, gauss, B, K)
    for mode in ("flushed", "L2-warm"):
        ts = []
        for _ in range(7):
            if mode == "flushed": flush.zero_()
            else: fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[3]
        print(f"B={B:4d} {mode:8s}: {t:7.1f} us  -> {2 * M * 2048 * 4 / t / 1e3:7.1f} GB/s  ({t / (B * 16) * 1e3:6.1f} ns per tile-launch slot)", flush=True)
