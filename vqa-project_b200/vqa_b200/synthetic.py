"""Synthetic VQA batches with the layout the reference's data layer produces.

Layout contract (reference ``torch_dataset.py:148-164``, ``utils.py:22-31``, ``collate_fn``):
``question`` int64 (B,100) zero padded and sorted by descending length, ``image`` float32
(B,K,F) whose last 4 columns are xyxy boxes normalised to [0,1], ``K`` int64 (B,1) all equal,
``qlen`` a python list of 0-d int64 tensors, soft answer targets (B,A) summing to 1 per row.
Features are ``clamp(N(0,1), 0)`` (post-ReLU RoI features, ~50 % zeros).  SURVEY.md 8(d).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List

import torch


@dataclass(frozen=True)
class Workload:
    """One row of BASELINE.json ``configs`` expressed as shapes."""
    name: str
    batch: int
    n_obj: int          # K, nodes per image
    feat_dim: int       # F, includes the 4 box columns
    hid_dim: int = 1024
    emb_dim: int = 300
    out_dim: int = 3000
    vocab: int = 20000
    n_kernels: int = 8
    neighbourhood: int = 16
    max_qlen: int = 14
    q_width: int = 100
    dropout: float = 0.5

    def model_kwargs(self) -> Dict:
        return dict(vocab_size=self.vocab, emb_dim=self.emb_dim, feat_dim=self.feat_dim,
                    hid_dim=self.hid_dim, out_dim=self.out_dim, dropout=self.dropout,
                    n_kernels=self.n_kernels, neighbourhood_size=self.neighbourhood,
                    n_obj=self.n_obj)

    def asdict(self) -> Dict:
        return asdict(self)


WORKLOADS = {
    # BASELINE.json configs[0..4]
    "vqa2_b64": Workload("vqa2_b64", 64, 36, 2052),
    "vqa2_b512": Workload("vqa2_b512", 512, 36, 2052),
    "med_b512": Workload("med_b512", 512, 51, 1028, out_dim=512, neighbourhood=19, max_qlen=15, dropout=0.4),
    "med_b8": Workload("med_b8", 8, 51, 1028, out_dim=512, neighbourhood=19, max_qlen=15, dropout=0.4),
    "eval_k100": Workload("eval_k100", 4096, 100, 2052, neighbourhood=32, dropout=0.0),
    # small shapes for tests
    "tiny": Workload("tiny", 3, 12, 20, hid_dim=16, emb_dim=8, out_dim=24, vocab=50, n_kernels=4,
                     neighbourhood=5, max_qlen=6, q_width=10, dropout=0.0),
    "small": Workload("small", 8, 36, 132, hid_dim=128, emb_dim=32, out_dim=200, vocab=300,
                      n_kernels=8, neighbourhood=16, max_qlen=14, q_width=20, dropout=0.0),
    # smallest shape whose graph convolutions take the tensor-core aggregate path ((2*hid/nk) % 128 == 0 and (hid/nk) % 128 == 0)
    "medium": Workload("medium", 6, 36, 68, hid_dim=512, emb_dim=32, out_dim=120, vocab=200,
                       n_kernels=4, neighbourhood=16, max_qlen=9, q_width=12, dropout=0.0),
}


def make_batch(w: Workload, seed: int = 1000, batch: int | None = None) -> Dict[str, object]:
    """CPU tensors for one batch; deterministic in (workload, seed, batch)."""
    g = torch.Generator().manual_seed(seed)
    b = w.batch if batch is None else batch
    feats = torch.randn(b, w.n_obj, w.feat_dim - 4, generator=g).clamp_(min=0)
    xy1 = torch.rand(b, w.n_obj, 2, generator=g) * 0.7
    wh = torch.rand(b, w.n_obj, 2, generator=g) * 0.25 + 0.05
    xy2 = (xy1 + wh).clamp_(max=1.0)
    image = torch.cat((feats, xy1, xy2), dim=-1).contiguous()

    qlen = torch.randint(3, w.max_qlen + 1, (b,), generator=g)
    qlen, _ = torch.sort(qlen, descending=True)
    question = torch.zeros(b, w.q_width, dtype=torch.int64)
    toks = torch.randint(1, w.vocab, (b, w.q_width), generator=g)
    mask = torch.arange(w.q_width).unsqueeze(0) < qlen.unsqueeze(1)
    question[mask] = toks[mask]

    target = torch.zeros(b, w.out_dim)
    n_ans = torch.randint(1, 4, (b,), generator=g)
    for i in range(b):
        ids = torch.randperm(w.out_dim, generator=g)[: int(n_ans[i])]
        sc = torch.rand(len(ids), generator=g) + 0.1
        target[i, ids] = sc / sc.sum()
    return dict(
        question=question,
        image=image,
        K=torch.full((b, 1), w.n_obj, dtype=torch.int64),
        qlen=[q for q in qlen],          # list of 0-d int64 tensors, as utils.batch_to_cuda builds it
        target=target,
    )


def make_wemb(w: Workload, seed: int = 1000):
    """Stand-in for the GloVe matrix (numpy float32 (V, emb)), as ``Model.__init__`` expects."""
    g = torch.Generator().manual_seed(seed + 1)
    return (0.4 * torch.randn(w.vocab, w.emb_dim, generator=g)).numpy()
