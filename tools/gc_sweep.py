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
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
B, K, F, nk = 512, 36, 2052, 8
torch.manual_seed(0)
img = torch.rand(B, K, F, device=dev); gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
for nb in (1, 4, 16):
    for out in (256, 1024, 2048, 4096):
        idx = torch.stack([torch.randperm(K, device=dev)[:nb] for _ in range(B * K)]).view(B, K, nb).int().contiguous()
        alpha = torch.softmax(torch.randn(B, K, nb, device=dev), -1)
        Y = torch.randn(B * K, out, device=dev)
        t = timeit(lambda: kn.graphconv_fwd(Y, idx, alpha, img, gauss, B, K))
        print(f"nb={nb:2d} out={out:5d}: {t:8.1f} us   ({2*B*K*out*4/t/1e3:7.1f} GB/s)", flush=True)
