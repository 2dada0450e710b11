import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "vqa-project_b200")]
from conftest import load_golden, golden_params, rel_err
from vqa_b200.synthetic import WORKLOADS, make_wemb
import sparse_graph_model as M
from vqa_b200 import kernels as kn
from vqa_b200 import ops
torch.backends.cudnn.allow_tf32 = os.environ.get('GRU_TF32', '0') == '1'
ops.set_precision(os.environ.get("VQA_PREC", "fp32"))
for name in sys.argv[1:] or ("tiny", "small"):
    g = load_golden(name); w = WORKLOADS[name]
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs())
    model.load_state_dict(golden_params(g)); model = model.cuda().train()
    q = torch.from_numpy(g["in.question"]).cuda(); img = torch.from_numpy(g["in.image"]).cuda()
    qlen = [torch.tensor(int(x)) for x in g["in.qlen"]]; tgt = torch.from_numpy(g["in.target"]).cuda()
    K = torch.full((q.shape[0], 1), img.shape[1], dtype=torch.int64).cuda()
    logits, adj, arg = model(q, img, K, qlen)
    idx, _ = kn.topk_softmax(adj.detach(), w.neighbourhood)
    got = idx.long().sort(-1).values.cpu().numpy()
    print(name, "topk rows differing:", (got != g["nbr.idx_sorted"]).any(-1).sum(), "of", got.shape[0] * got.shape[1],
          "argmax differing:", (arg.cpu().numpy() != g["out.h_max_indices"]).sum(), "of", arg.numel())
    loss = torch.nn.MultiLabelSoftMarginLoss()(logits, tgt); loss.backward()
    for k, v in model.named_parameters():
        print(f"   {k:55s} {rel_err(v.grad.cpu(), g['grad.' + k]):.2e}  |g|max {np.abs(g['grad.' + k]).max():.2e}")
