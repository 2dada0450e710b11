"""Phase timeline of the backward edge kernel (`edge_p_kernel`, csrc/graphconv_mma.cu), per CTA, from %globaltimer stamps.

Needs the tracing build of the library (never the shipped one: the stamps are compiled out of it):

    make -C vqa-project_b200/csrc EXTRA=-DVQA_EDGE_TRACE BUILD=build_trace OUT=../vqa_b200/libvqa_trace.so
    python tools/edge_trace.py [--shape vqa2|med|k100] > gpurun_out/edge_trace.txt

Prints, for the plain (layer 1) and pooled (layer 2) variants: the kernel's span, CTAs per SM, and the median / p90 length of each
phase of a CTA (set-up, the wait for each Gaussian kernel's accumulator, the edge finish loop, the reduction)."""
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
    out = np.zeros((B, 16), dtype=np.uint64)
    assert lib.vqa_debug_edge_trace(out.ctypes.data, B) == 0
    t = out[:, :13].astype(np.int64)
    t0 = t[:, 0].min()
    span = (t[:, 12].max() - t0) / 1e3
    sm = out[:, 15].astype(np.int64)
    print(f"== {name} ({args.shape}: B={B} K={K} nb={nb} nk={nk}): span {span:.1f} us, {len(set(sm.tolist()))} SMs, CTA life median "
          f"{np.median(t[:, 12] - t[:, 0]) / 1e3:.1f} us")
    names = ["set-up"] + [f"acc k={k} ready" for k in range(8)] + ["main loop drained", "finish edge loop", "reduction + exit"]
    prev = t[:, 0]
    for i, nm in enumerate(names, start=1):
        d = (t[:, i] - prev) / 1e3
        print(f"   {nm:22s} median {np.median(d):7.2f} us   p90 {np.percentile(d, 90):7.2f}   max {d.max():7.2f}")
        prev = t[:, i]
    # concurrency: how many CTAs were alive at the median start time of the second wave
    starts = np.sort(t[:, 0] - t0) / 1e3
    print(f"   CTA start times: first {starts[0]:.1f}, #296 {starts[min(295, B - 1)]:.1f}, #297 {starts[min(296, B - 1)]:.1f}, last {starts[-1]:.1f} us")


trace("edge_p<plain> layer 1", lambda: kn.graphconv_bwd_edges_s(Y1, idx, alpha, img, gauss, B, K, dOs=dO1))
trace("edge_p<pooled> layer 2", lambda: kn.graphconv_bwd_edges_s(Y2, idx, None, img, gauss, B, K, dpooled=dp, argmax=arg))
