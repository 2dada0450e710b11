"""Where the backward edge kernel (`edge_p_kernel`, csrc/graphconv_mma.cu) waits: cycles each role of a CTA spends on each barrier.

Needs the tracing build of the library (never the shipped one: the counters are compiled out of it):

    make -C vqa-project_b200/csrc EXTRA=-DVQA_EDGE_TRACE BUILD=build_trace OUT=../vqa_b200/libvqa_trace.so
    python tools/edge_trace.py [--shape vqa2|med|k100] > gpurun_out/edge_trace.txt

The role that (almost) never waits is the bottleneck: the TMA producer waits for free stages, the MMA issuer for filled stages and
for a free accumulator, the TMEM readers for a finished accumulator, the A-tile builders (pooled upstream) for free stages."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
from vqa_b200 import _cabi  # noqa: E402

_cabi.LIB_PATH = os.path.join(ROOT, "vqa-project_b200", "vqa_b200", "libvqa_trace.so")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from vqa_b200 import kernels as kn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="vqa2")
args = ap.parse_args()
dev = torch.device("cuda:0")
B, K, nb, nk, H = {"vqa2": (512, 36, 16, 8, 1024), "med": (512, 51, 19, 8, 1024), "k100": (1024, 100, 32, 8, 1024)}[args.shape]
M = B * K
torch.manual_seed(0)
img = torch.rand(B, K, 12, device=dev)
gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
h = torch.randn(B, K, 512, device=dev).clamp_(min=0)
adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
Y1 = kn.split(torch.randn(M, 2 * H, device=dev))
Y2 = kn.split(torch.randn(M, H, device=dev))
dO1 = kn.split(torch.randn(M, 2 * H, device=dev))
q = torch.randn(B, H, device=dev)
ec2 = kn.graphconv_edge_coef(idx, None, img, gauss, B, K)
pooled, arg, hq = kn.graphconv_pool_fwd_s(Y2, idx, img, gauss, q, B, K, ec=ec2)
dp = torch.randn(B, H, device=dev)
lib = _cabi.load()
lib.vqa_debug_edge_trace.argtypes = [C.c_void_p, C.c_int]
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def trace(name, fn):
    for _ in range(3):
        fn()
    flush.zero_()
    torch.cuda.synchronize()
    fn()
    out = np.zeros((148, 8), dtype=np.int64)
    assert lib.vqa_debug_edge_trace(out.ctypes.data, 148) == 0
    med = np.median(out, axis=0)
    print(f"== {name} ({args.shape}: B={B} K={K} nb={nb} nk={nk}): CTA life {med[0]:.0f} cycles for {med[7]:.0f} units "
          f"= {med[0] / max(med[7], 1):.0f} cycles per unit")
    for slot, nm in ((1, "TMA producer waits for a free stage"), (2, "MMA issuer waits for a filled stage"), (3, "MMA issuer waits for a free accumulator"),
                     (4, "TMEM readers wait for a finished accumulator"), (5, "A-tile builders wait for a free stage")):
        print(f"   {nm:46s} {med[slot]:9.0f} cycles = {100 * med[slot] / max(med[0], 1):5.1f} % of the CTA's life")


trace("edge_p<plain> layer 1", lambda: kn.graphconv_bwd_edges_s(Y1, idx, alpha, img, gauss, B, K, dOs=dO1))
trace("edge_p<pooled> layer 2", lambda: kn.graphconv_bwd_edges_s(Y2, idx, None, img, gauss, B, K, dpooled=dp, argmax=arg))
