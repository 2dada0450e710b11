"""Per-call GEMM trace of one eager training step at VQA2 B=512: shape, operand majors, time, TF/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200 import kernels as kn, ops
from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
from vqa_b200.ddp import GradReducer
import sparse_graph_model as M
if len(sys.argv) > 1: ops.set_precision(sys.argv[1])
dev = torch.device("cuda:0")
w = WORKLOADS["vqa2_b512"]
torch.manual_seed(1000)
model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev).train()
crit = torch.nn.MultiLabelSoftMarginLoss(); red = GradReducer(model.parameters())
hb = make_batch(w); b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()}
def step():
    red.zero_grad(); logits, _, _ = model(b["question"], b["image"], b["K"], b["qlen"])
    crit(logits, b["target"]).backward(); red.finish()
for _ in range(3): step()
trace = []
orig = kn._call
def traced(name, *args):
    if name in ("vqa_gemm_bf16s", "vqa_gemm_f32"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *args); e1.record()
        if name == "vqa_gemm_bf16s": M_, N_, K_, amn, bmn, sk = args[13], args[14], args[15], args[3], args[7], args[27]
        else: M_, N_, K_, amn, bmn, sk = args[8], args[9], args[10], args[2], args[5], args[20]
        trace.append((name, M_, N_, K_, amn, bmn, sk, e0, e1))
    else:
        orig(name, *args)
kn._call = traced
step(); torch.cuda.synchronize()
tot = 0.0
agg = {}
for name, M_, N_, K_, amn, bmn, sk, e0, e1 in trace:
    us = e0.elapsed_time(e1) * 1e3
    key = (name[4:], M_, N_, K_, amn, bmn, sk)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
for key, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    name, M_, N_, K_, amn, bmn, sk = key
    tot += us
    print(f"{name:12s} M={M_:6d} N={N_:5d} K={K_:6d} a_mn={amn} b_mn={bmn} splitk={sk:2d}  x{cnt:3d} {us:9.1f} us total {us/cnt:8.1f} us each  {2.0*M_*N_*K_*cnt/us/1e6:7.1f} TF/s")
print(f"total GEMM time (event-bracketed, includes launch gaps): {tot:.1f} us")
