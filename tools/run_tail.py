"""Stand-alone timing of the step-tail and batch-assembly kernels at the VQA2 B=512 sizes (CUDA events, launching stream):
adam_flat_kernel over the model's 31 M parameters, the criterion's forward/backward over (512, 3000) logits, gather_image over
512 x 36 x 2048 rows (fp32 and bf16 tables), scatter_targets.  Prints one JSON line per kernel with the algorithmic bytes and the
achieved GB/s against MEASURED_PEAKS.json.  Also the ncu target for these kernels (`ncu --set full -k regex:...`)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
from vqa_b200 import kernels as kn  # noqa: E402
from vqa_b200.ddp import GradReducer  # noqa: E402
from vqa_b200.loss import MultiLabelSoftMarginLoss  # noqa: E402
from vqa_b200.optim import FlatAdam  # noqa: E402
from vqa_b200.synthetic import WORKLOADS, make_wemb  # noqa: E402
import sparse_graph_model as M  # noqa: E402

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=int(os.environ.get("VQA_TAIL_REPS", "20")), flush_l2=True):
    ts = []
    for _ in range(reps + 3):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[3:] or ts)
    return ts[len(ts) // 2]


def report(name, us, nbytes, note=""):
    gbs = nbytes / us * 1e-3
    print(json.dumps({"kernel": name, "us": round(us, 2), "algorithmic_bytes": int(nbytes), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / peak, 3), "note": note}), flush=True)


w = WORKLOADS["vqa2_b512"]
torch.manual_seed(0)
model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev)
red = GradReducer(model.parameters())
opt = FlatAdam(red, lr=1e-4)
red.flat.normal_()
for p in red.params:
    p.grad = red._view(p)
n = sum(p.numel() for p in red.params)
report("adam_flat_kernel", timed(opt.step), 28 * n, f"{n} parameters, read p/g/m/v + write p/m/v")
topt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
report("torch fused Adam (multi_tensor_apply, for comparison)", timed(topt.step), 28 * n, "2 launches")

x = torch.randn(w.batch, w.out_dim, device=dev)
y = (torch.rand(w.batch, w.out_dim, device=dev) > 0.999).float()
go = torch.ones((), device=dev)
sc = 1.0 / x.numel()
report("mlsm_loss_fwd_kernel", timed(lambda: kn.mlsm_loss_fwd(x, y, sc)), 8 * x.numel())
report("mlsm_loss_bwd_kernel", timed(lambda: kn.mlsm_loss_bwd(x, y, go, sc)), 12 * x.numel())
xr = x.clone().requires_grad_()
crit = torch.nn.MultiLabelSoftMarginLoss()


def torch_loss():
    xr.grad = None
    crit(xr, y).backward()


report("torch MultiLabelSoftMarginLoss fwd+bwd (for comparison)", timed(torch_loss), 20 * x.numel(), "~20 launches, includes host launch time")

n_img, K, D = 1536, w.n_obj, w.feat_dim - 4
err = torch.zeros(1, dtype=torch.int32, device=dev)
rows = torch.randint(0, n_img, (w.batch,), device=dev)
boxes = torch.rand(n_img, K, 4, device=dev)
for dt, s in ((torch.float32, 4), (torch.bfloat16, 2)):
    table = torch.randn(n_img, K, D, device=dev).clamp_(min=0).to(dt)
    report(f"gather_image_kernel<{'bf16' if s == 2 else 'f32'}>", timed(lambda: kn.gather_image(table, boxes, rows, err)),
           w.batch * K * (D * s + 16 + (D + 4) * 4), "read rows + boxes, write (B,K,D+4) fp32")
    del table
ptr = torch.arange(0, 2 * w.batch + 1, 2, device=dev)
ids = torch.randint(0, w.out_dim, (2 * w.batch,), dtype=torch.int32, device=dev)
vals = torch.rand(2 * w.batch, device=dev)
report("scatter_targets_kernel", timed(lambda: kn.scatter_targets(ptr, ids, vals, w.batch, w.out_dim, err)), 4 * w.batch * w.out_dim)
assert err.item() == 0
