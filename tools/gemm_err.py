import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
from vqa_b200 import kernels as kn
torch.manual_seed(0)
def err(o, r): return ((o.double() - r).abs().max() / r.abs().max()).item()
M, N = 512, 512
for K in (64, 256, 1024, 2052, 4096):
    a = torch.randn(M, K, device="cuda").clamp_(min=0); b = torch.randn(N, K, device="cuda") * 0.02
    ref = a.double() @ b.double().t()
    line = [f"K={K:5d} torch_fp32 {err(a @ b.t(), ref):.1e}"]
    for prec, nm in ((1, "tf32"), (0, "x3")):
        for s in (1, 2, 4, 8, 16):
            if K // 32 < s: continue
            line.append(f"{nm}/s{s} {err(kn.gemm(a, b, precision=prec, split_k=s), ref):.1e}")
    print("  ".join(line))
# positive x positive (worst case for truncation bias): h h^T-like
a = torch.randn(M, 2048, device="cuda").clamp_(min=0); b = torch.randn(N, 2048, device="cuda").clamp_(min=0)
ref = a.double() @ b.double().t()
print("pos*pos K=2048:", "torch", f"{err(a @ b.t(), ref):.1e}", *[f"x3/s{s} {err(kn.gemm(a, b, split_k=s), ref):.1e}" for s in (1, 2, 4, 8, 16, 32)])
