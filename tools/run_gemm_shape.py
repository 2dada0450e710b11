"""One split-bf16 GEMM shape, a few launches (ncu target):  python tools/run_gemm_shape.py M N K [passes] [--bias] [--split-out]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn
dev = torch.device("cuda:0")
torch.manual_seed(0)
M, N, K = (int(x) for x in sys.argv[1:4])
passes = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 3
Xs = kn.split(torch.randn(M, K, device=dev)); Ws = kn.split(torch.randn(N, K, device=dev) * 0.02)
bias = torch.randn(N, device=dev) if "--bias" in sys.argv else None
out = kn.empty_split(M, N, dev, passes == 3) if "--split-out" in sys.argv else None
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for _ in range(3):
    flush.zero_()
    if out is not None:
        kn.gemm_s(Xs, Ws, bias=bias, out_split=out, want_f32=False, passes=passes)
    else:
        kn.gemm_s(Xs, Ws, bias=bias, passes=passes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(1e6)); e0.record()
for _ in range(10):
    kn.gemm_s(Xs, Ws, bias=bias, passes=passes)
e1.record(); torch.cuda.synchronize()
print(f"M={M} N={N} K={K} passes={passes}: {e0.elapsed_time(e1) * 100:.1f} us per launch (10 back-to-back, warm L2)")
