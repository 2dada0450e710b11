"""BASELINE config[4]: inference sweep, model.eval(), B=4096, K=100 boxes, top-k=32 (stresses adjacency / top-k / gather kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch
from vqa_b200.synthetic import WORKLOADS, make_batch, make_wemb
import sparse_graph_model as M
dev = torch.device("cuda:0")
w = WORKLOADS["eval_k100"]
torch.manual_seed(1000)
model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev).eval()
model.max_question_len = w.max_qlen
b = make_batch(w, seed=7)
q, img, K = b["question"].to(dev), b["image"].to(dev), b["K"].to(dev)
qlen = torch.tensor([int(x) for x in b["qlen"]], dtype=torch.int32, device=dev)
with torch.no_grad():
    for _ in range(2):
        logits, adj, arg = model(q, img, K, qlen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        logits, adj, arg = model(q, img, K, qlen)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"eval_k100: B={w.batch} K={w.n_obj} nb={w.neighbourhood}: {ms:.2f} ms per batch = {w.batch / ms * 1e3:.0f} questions/s; logits {tuple(logits.shape)} finite={bool(torch.isfinite(logits).all())} adjacency {tuple(adj.shape)}")
