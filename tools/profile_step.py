"""torch.profiler breakdown of one training step (GPU kernel time by name, CPU-vs-GPU bound check, H2D bandwidth)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
from vqa_b200.ddp import GradReducer
import sparse_graph_model as M
dev = torch.device("cuda:0")
w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "vqa2_b512"]
torch.manual_seed(1000)
model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev).train()
crit = torch.nn.MultiLabelSoftMarginLoss()
red = GradReducer(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
hb = make_batch(w)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()}
def step():
    red.zero_grad()
    logits, _, _ = model(b["question"], b["image"], b["K"], b["qlen"])
    loss = crit(logits, b["target"]); loss.backward(); red.finish(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t_cpu = time.perf_counter() - t0          # CPU time to ENQUEUE 5 steps
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"enqueue {t_cpu/5*1e3:.2f} ms/step, wall {t_all/5*1e3:.2f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3, e.count / 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] if hasattr(prof.key_averages()[0], "device_type") else []
ka = prof.key_averages()
kern = {}
for e in prof.events():
    if e.device_type.name == "CUDA":
        k = e.name[:90]
        d = kern.setdefault(k, [0.0, 0]); d[0] += e.device_time / 3 if hasattr(e, "device_time") else e.cuda_time / 3; d[1] += 1
tot = sum(v[0] for v in kern.values())
print(f"GPU kernel time per step: {tot/1e3:.3f} ms")
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:48]:
    print(f"{v[0]:10.1f} us  x{v[1]/3:5.1f}  {k}")
# H2D bandwidth
img = hb["image"].pin_memory()
d = torch.empty_like(img, device=dev)
for _ in range(2): d.copy_(img, non_blocking=True)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(img, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"H2D pinned {img.numel()*4/1e6:.0f} MB in {e0.elapsed_time(e1):.2f} ms = {img.numel()*4/e0.elapsed_time(e1)/1e6:.1f} GB/s")
