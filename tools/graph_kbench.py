"""Stand-alone timings of the graph kernels (adjacency / top-k, the aggregates, the edge kernels) at the shapes of BASELINE.json
configs[1], [3] and [4]: CUDA events on the launching stream, L2 flushed (256 MB write) before every timed launch, median of 7.
Prints one JSON line per kernel with its algorithmic bytes and the fraction of the measured HBM peak.

    python tools/graph_kbench.py [--only adjacency] [--shapes vqa2,med,k100] > gpurun_out/graph_kbench.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch  # noqa: E402
from vqa_b200 import kernels as kn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--shapes", default="vqa2,med,k100")
args = ap.parse_args()
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
SPIN = int(400e-6 * 1.9e9)               # ~0.4 ms of SM clock
peak = 6544.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(shape, name, fn, nbytes, iters=7, warm=2):
    if args.only and args.only not in name:
        return
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(SPIN)          # the host needs ~50 us to allocate outputs and launch: let it run ahead of the GPU, so that the
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # bracket holds device time only
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(json.dumps({"shape": shape, "kernel": name, "us": round(t, 2), "GB/s": round(nbytes / t / 1e3, 1), "frac": round(nbytes / t / 1e3 / peak, 4),
                      "algorithmic_bytes": nbytes}), flush=True)


SHAPES = {"vqa2": (512, 36, 16, 8, 1024), "med": (512, 51, 19, 8, 1024), "k100": (1024, 100, 32, 8, 1024)}
for sname in args.shapes.split(","):
    B, K, nb, nk, H = SHAPES[sname]
    M, C = B * K, 512
    torch.manual_seed(0)
    img = torch.rand(B, K, 12, device=dev)
    gauss = torch.rand(4 * nk, device=dev) * 0.9 + 0.1
    h = torch.randn(B, K, C, device=dev).clamp_(min=0)
    tag = f"{sname} B={B} K={K} nb={nb}"
    timeit(tag, "adjacency_topk_fwd", lambda: kn.adjacency_topk_fwd(h, nb), M * C * 4 + M * K * 4 + 2 * M * nb * 4)
    adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
    dalpha = torch.randn_like(alpha)
    timeit(tag, "adjacency_topk_bwd", lambda: kn.adjacency_topk_bwd(h, idx, alpha, dalpha), 2 * M * C * 4 + 3 * M * nb * 4)
    Y1 = kn.split(torch.randn(M, 2 * H, device=dev))
    Y2 = kn.split(torch.randn(M, H, device=dev))
    dO1 = kn.split(torch.randn(M, 2 * H, device=dev))
    q = torch.randn(B, H, device=dev)
    ec1 = kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K)
    ec2 = kn.graphconv_edge_coef(idx, None, img, gauss, B, K)
    timeit(tag, "edge_coef", lambda: kn.graphconv_edge_coef(idx, alpha, img, gauss, B, K), M * nb * nk * 4 + 3 * M * nb * 4)
    b1 = 2 * M * 2 * H * 4 + 2 * M * nb * 4 + M * 16
    timeit(tag, "graphconv_mma_fwd L1 relu+dropout", lambda: kn.graphconv_fwd_s(Y1, idx, alpha, img, gauss, B, K, relu=True, dropout_p=0.5, seed=1, offset=1, ec=ec1), b1)
    timeit(tag, "graphconv_mma_fwd L1 relu", lambda: kn.graphconv_fwd_s(Y1, idx, alpha, img, gauss, B, K, relu=True, ec=ec1), b1)
    timeit(tag, "graphconv_mma_pool_fwd L2", lambda: kn.graphconv_pool_fwd_s(Y2, idx, img, gauss, q, B, K, ec=ec2), M * H * 4 + M * nb * 4 + M * 16 + 4 * B * H * 4)
    timeit(tag, "graphconv_mma_bwd_data L1", lambda: kn.graphconv_bwd_data_s(dO1, idx, alpha, img, gauss, B, K, ec=ec1), b1)
    timeit(tag, "graphconv_mma_bwd_edges L1", lambda: kn.graphconv_bwd_edges_s(Y1, idx, alpha, img, gauss, B, K, dOs=dO1), 2 * M * 2 * H * 4 + 3 * M * nb * 4)
    pooled, arg, hq = kn.graphconv_pool_fwd_s(Y2, idx, img, gauss, q, B, K, ec=ec2)
    dp = torch.randn(B, H, device=dev)
    timeit(tag, "graphconv_mma_bwd_edges L2 pooled", lambda: kn.graphconv_bwd_edges_s(Y2, idx, None, img, gauss, B, K, dpooled=dp, argmax=arg), M * H * 4 + 2 * M * nb * 4)
    timeit(tag, "graphconv_pool_bwd_data L2", lambda: kn.graphconv_pool_bwd_data_s(dp, arg, idx, ec2, B, K, H), M * H * 4 + M * nb * nk * 4)
    del Y1, Y2, dO1

# column sums of the step tail (bias gradients): GRU gate gradients (T*B, 3H), graph-learner hidden gradients (B*K, 512), logits
for rows, cols in ((7168, 3072), (18432, 512), (512, 3000), (512, 32)):
    x = torch.randn(rows, cols, device=dev)
    timeit("tail", f"colsum {rows}x{cols}", lambda: kn.colsum(x), rows * cols * 4 + cols * 4)
