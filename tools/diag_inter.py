"""Compare backward intermediates of the CUDA path with fp64 oracle autograd on the `small` golden workload."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "vqa-project_b200")]
from conftest import load_golden, golden_params
from oracle import vqa_oracle as O
from vqa_b200.synthetic import WORKLOADS, make_wemb
import sparse_graph_model as M
from vqa_b200 import kernels as kn
name = sys.argv[1] if len(sys.argv) > 1 else "small"
g = load_golden(name); w = WORKLOADS[name]
def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-300)).item()
# ---------------- oracle fp64 with retained intermediates
p = {k: v.double().requires_grad_(True) for k, v in golden_params(g).items()}
q = torch.from_numpy(g["in.question"]); img = torch.from_numpy(g["in.image"]).double(); qlen = [int(x) for x in g["in.qlen"]]
tgt = torch.from_numpy(g["in.target"]).double()
cen = O.box_centres(img); pseudo = O.polar_pseudo_coordinates(cen)
qenc = O.gru_last_hidden(p["wembed.weight"][q], qlen, p); qenc.retain_grad()
nodes = torch.cat((img, qenc.unsqueeze(1).expand(-1, w.n_obj, -1)), -1)
h1 = torch.relu(O.wn_linear(nodes, p, "adjacency_1.edge_layer_1")); h1.retain_grad()
h2 = torch.relu(O.wn_linear(h1, p, "adjacency_1.edge_layer_2")); h2.retain_grad()
adj = h2 @ h2.transpose(1, 2); adj.retain_grad()
alpha, idx = O.select_neighbourhood(adj, w.neighbourhood); alpha.retain_grad()
nbrp = O.gather_pseudo(pseudo, idx)
g1 = torch.relu(O.graph_convolution(alpha.unsqueeze(-1) * O.gather_neighbours(img, idx), nbrp, p, "graph_convolution_1", w.n_kernels)); g1.retain_grad()
g2 = torch.relu(O.graph_convolution(O.gather_neighbours(g1, idx), nbrp, p, "graph_convolution_2", w.n_kernels))
pooled, arg = g2.max(1); pooled.retain_grad()
hq = torch.relu(qenc) * pooled; hq.retain_grad()
hid = torch.relu(O.wn_linear(hq, p, "out_1")); hid.retain_grad()
logits = O.wn_linear(hid, p, "out_2")
O.multilabel_soft_margin_loss(logits, tgt).backward()
# ---------------- CUDA path with recorded intermediates
rec = {}
def wrap(fn_name):
    orig = getattr(kn, fn_name)
    def f(*a, **k):
        out = orig(*a, **k)
        rec.setdefault(fn_name, []).append((a, k, out))
        return out
    setattr(kn, fn_name, f)
for fn in ("graphconv_bwd", "adjacency_topk_bwd", "gate_bwd", "adjacency_topk_fwd", "segment_sum"):
    wrap(fn)
model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()); model.load_state_dict(golden_params(g)); model = model.cuda().train()
qd = q.cuda(); imd = torch.from_numpy(g["in.image"]).cuda(); Kt = torch.full((qd.shape[0], 1), w.n_obj).cuda()
ql = [torch.tensor(x) for x in qlen]
lg, ad, ar = model(qd, imd, Kt, ql)
torch.nn.MultiLabelSoftMarginLoss()(lg, torch.from_numpy(g["in.target"]).cuda()).backward()
B, K = imd.shape[:2]
_, _, (adj_c, idx_c, alpha_c) = rec["adjacency_topk_fwd"][0]
order_c = idx_c.long().argsort(-1).cpu(); order_o = idx.argsort(-1)
print("idx sets equal:", torch.equal(torch.gather(idx_c.long().cpu(), -1, order_c), torch.gather(idx, -1, order_o)))
print("alpha (sorted by idx):", rel(torch.gather(alpha_c.cpu(), -1, order_c), torch.gather(alpha, -1, order_o)))
print("alpha max per row stats: mean top alpha", alpha.max(-1).values.mean().item(), "min", alpha.max(-1).values.min().item())
(a2, k2, (dY2, _, dgs2)), (a1, k1, (dY1, dalpha_c, dgs1)) = rec["graphconv_bwd"]
print("dpooled:", rel(k2["dpooled"], pooled.grad * 0 + pooled.grad))   # upstream of layer 2 (already gated/masked)
print("dalpha (sorted):", rel(torch.gather(dalpha_c.cpu(), -1, order_c), torch.gather(alpha.grad, -1, order_o)))
print("dG1 (dO of layer 1) vs oracle g1.grad (masked):", rel(k1["dO"].view(B, K, -1), g1.grad * (g1 > 0)))
(aa, ka, dh2_c) = rec["adjacency_topk_bwd"][0]
print("dh2 (masked):", rel(dh2_c.view(B, K, -1), h2.grad * (h2 > 0)))
print("dadj passed:", aa[4] is None if len(aa) > 4 else ka)
dA_total = adj.grad
print("oracle |dA| max", dA_total.abs().max().item())
print("dq total:", rel(model.q_gru.weight_hh_l0.grad, p["q_gru.weight_hh_l0"].grad), "(gru whh grad)")
(ag, kg, (dpool_c, dq_c)) = rec["gate_bwd"][0]
print("dq gate only vs oracle total qenc.grad:", rel(dq_c, qenc.grad))
print("h1 grad check via db1:", rel(model.adjacency_1.edge_layer_1.bias.grad, p["adjacency_1.edge_layer_1.bias"].grad))
# ---------------- isolate kernel error from upstream error
print("---- isolation")
print("qenc fwd:", rel(model.encode_question(qd, ql), qenc))
Y1_c = a1[0].view(B, K, -1); dO_c = k1["dO"].view(B, K, -1)
Wc1 = torch.cat([p[f"graph_convolution_1.conv_weights.{i}.weight"] for i in range(w.n_kernels)])
Y1_x = img @ Wc1.detach().t()
print("Y1:", rel(Y1_c, Y1_x))
dO_x = (g1.grad * (g1 > 0)).detach()
mm = ((dO_c.cpu() != 0) != (dO_x != 0))
print("dO mask mismatches:", int(mm.sum()), "of", mm.numel(), " dO err where masks agree:", rel(torch.where(mm, torch.zeros_like(dO_x), dO_c.cpu().double()), torch.where(mm, torch.zeros_like(dO_x), dO_x)))
def dalpha_from(dO, Y1, al):
    nk = w.n_kernels
    wts = O.gaussian_kernel_weights(nbrp, {k: v.detach() for k, v in p.items()}, "graph_convolution_1").view(B, K, -1, nk)
    D = Y1.shape[-1] // nk
    nb_ = O.gather_neighbours(Y1, idx).view(B, K, -1, nk, D)
    P = (dO.view(B, K, 1, nk, D) * nb_).sum(-1)
    return (wts * P).sum(-1)
da_mine64 = dalpha_from(dO_c.cpu().double(), Y1_c.cpu().double(), None)
da_exact = dalpha_from(dO_x, Y1_x, None)
srt = lambda t, o: torch.gather(t, -1, o)
# my kernel's dalpha is in my idx order; oracle idx order differs -> compare in sorted-by-index order
da_c_sorted = srt(dalpha_c.cpu().double(), order_c)
print("oracle alpha.grad vs fp64 recompute from exact inputs:", rel(srt(da_exact, order_o), srt(alpha.grad, order_o)))
print("my dalpha vs fp64 recompute from MY inputs (kernel error):", rel(da_c_sorted, srt(da_mine64, order_o)))
print("fp64 recompute from MY inputs vs exact (upstream error):", rel(srt(da_mine64, order_o), srt(da_exact, order_o)))
dv_exact = alpha.detach() * (alpha.grad - (alpha.detach() * alpha.grad).sum(-1, keepdim=True))
print("|dalpha|max %.3e  |dv|max %.3e" % (alpha.grad.abs().max().item(), dv_exact.abs().max().item()))
