"""Phase timeline of the one-tile-per-CTA split-bf16 GEMM (`gemm_bf16s_kernel`) for one shape, from clock64 stamps per CTA.

Needs the tracing build of the library (never the shipped one: the stamps are compiled out of it):

    make -C vqa-project_b200/csrc EXTRA=-DVQA_GEMM_TRACE BUILD=build_trace OUT=../vqa_b200/libvqa_trace.so
    python tools/gemm_timeline.py 512 3072 1024 [tile_n] > gpurun_out/gemm_timeline.txt      # the GRU's per-step product
"""
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

M, N, K = (int(x) for x in sys.argv[1:4])
tile = int(sys.argv[4]) if len(sys.argv) > 4 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
Xs = kn.split(torch.randn(M, K, device=dev)); Ws = kn.split(torch.randn(N, K, device=dev) * 0.02)
lib = _cabi.load()
lib.vqa_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int]
for _ in range(5):
    kn.gemm_s(Xs, Ws, tile_n=tile)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(1e6)); e0.record()
for _ in range(20):
    kn.gemm_s(Xs, Ws, tile_n=tile)
e1.record(); torch.cuda.synchronize()
n = ((M + 127) // 128) * ((N + tile - 1) // tile)
out = np.zeros((n, 8), dtype=np.int64)
assert lib.vqa_debug_gemm_trace(out.ctypes.data, n) == 0
t = out[:, :7] - out[:, :1]
mhz = 1965.0
print(f"M={M} N={N} K={K} tile_n={tile}: {n} CTAs, {e0.elapsed_time(e1) * 50:.1f} us per launch (20 back-to-back, warm L2); CTA timeline, median cycles (us at {mhz:.0f} MHz):")
for i, nm in enumerate(("start", "set-up done (barriers, TMEM, cluster sync)", "first stage landed (MMA issuer)", "last MMA + commit issued", "accumulator complete (epilogue woke)",
                        "epilogue stores issued", "all warps at the end")):
    m = np.median(t[:, i])
    print(f"   {nm:46s} {m:9.0f}  ({m / mhz:6.2f} us)   p90 {np.percentile(t[:, i], 90):9.0f}")
