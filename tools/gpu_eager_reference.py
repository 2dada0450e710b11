"""The reference ALGORITHM in eager PyTorch on the same GPU (SURVEY.md 8d's "same-box competitor"): the oracle restatement
(`oracle/vqa_oracle.py`: reference operation order - concatenate, materialised neighbourhoods, aggregate-first patch operator,
nk separate linears - on torch's stock CUDA kernels) timed at the bench workload, fp32, with Adam, CUDA events.

    python tools/gpu_eager_reference.py [--workload vqa2_b512] [--batch 512] [--steps 20] [--warmup 5]

A study tool (tools/ is outside the product package; nothing here is on the product path).  Prints one JSON line.  The reference
materialises (B, K, nb, F) neighbourhood tensors: 2.4 GB per layer-1 copy at B = 512, so the full batch needs ~20 GB.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
from oracle import vqa_oracle as O  # noqa: E402
from vqa_b200.synthetic import WORKLOADS, make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vqa2_b512")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--tf32", action="store_true", help="allow TF32 matmuls (torch's default is full fp32)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    w = WORKLOADS[args.workload]
    B = args.batch or w.batch
    params = O.init_params(w.vocab, w.emb_dim, w.feat_dim, w.hid_dim, w.out_dim, w.n_kernels)
    leaves = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(leaves.values()), lr=1e-4, fused=True)
    batches = []
    for i in range(3):
        b = make_batch(w, seed=1000 + i, batch=B)
        batches.append((b["question"].to(dev), b["image"].to(dev), [int(x) for x in b["qlen"]], b["target"].to(dev)))

    def step(i):
        q, img, qlen, tgt = batches[i % 3]
        opt.zero_grad(set_to_none=True)
        logits, _, _ = O.forward(leaves, q, img, qlen, w.neighbourhood, w.n_kernels, dropout_p=w.dropout, training=True)
        loss = O.multilabel_soft_margin_loss(logits, tgt)
        loss.backward()
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "reference algorithm, eager PyTorch CUDA (oracle port)", "workload": w.name, "batch": B, "ms_per_step": round(ms, 3),
                      "questions_per_s": round(B / ms * 1e3, 1), "matmul": "tf32" if args.tf32 else "fp32", "steps": args.steps,
                      "peak_memory_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "final_loss": float(loss.detach())}))


if __name__ == "__main__":
    main()
