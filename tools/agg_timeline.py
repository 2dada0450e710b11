"""DEBUG: per-role clock64 timeline of CTA 0 of the persistent aggregate kernel (VQA_AGG_DBG must include 0x4000)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn, _cabi
dev = torch.device("cuda:0")
B, K, F, nb, nk = 512, 36, 2052, 16, 8
M = B * K
torch.manual_seed(0)
img = torch.rand(B, K, F, device=dev); gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
h = torch.randn(B, K, 512, device=dev).clamp_(min=0)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
Y1s = kn.split(torch.randn(M, 2048, device=dev)); Y2s = kn.split(torch.randn(M, 1024, device=dev)); q = torch.randn(B, 1024, device=dev)
ec1 = kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K); ec2 = kn.graphconv_edge_coef(idx, None, img, gauss, B, K)
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for _ in range(3):
    flush.zero_()
    if mode == "fwd": kn.graphconv_fwd_s(Y1s, idx, alpha, img, gauss, B, K, dropout_p=0.5, seed=1, offset=1, ec=ec1)
    elif mode == "pool": kn.graphconv_pool_fwd_s(Y2s, idx, img, gauss, q, B, K, ec=ec2)
    else: kn.graphconv_bwd_data_s(Y1s, idx, alpha, img, gauss, B, K, ec=ec1)
torch.cuda.synchronize()
N = 96
buf = (C.c_longlong * (10 * N))()
lib = _cabi.load()
lib.vqa_debug_agg_timeline.argtypes = [C.c_void_p]
rc = lib.vqa_debug_agg_timeline(buf)
assert rc == 0, rc
t = [[buf[r * N + i] for i in range(N)] for r in range(10)]
t0 = min(x for r in t for x in r if x > 0)
names = ["prod_issue", "mma_tempty", "mma_full", "mma_commit", "epi_tfull", "epi_arrive", "st_sfull", "st_done"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for g in range(56):
    print(f"{g:4d} " + " ".join(f"{(t[r][g] - t0) if t[r][g] else -1:10d}" for r in range(8)))
print("item  bld_start  bld_cempty   bld_cfull    bld_end")
for n in range(16):
    print(f"{n:4d} " + " ".join(f"{(x - t0) if x else -1:10d}" for x in (t[8][2 * n], t[8][2 * n + 1], t[9][2 * n], t[9][2 * n + 1])))
