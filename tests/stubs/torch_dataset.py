"""Stand-in for the reference's ``torch_dataset`` module, for the driver drop-in test ONLY (tests/test_dropin_gpu.py).

The reference's data layer reads zarr / pandas / pickle files that do not exist in this image; the drivers
(`run.py:29-31`) only need ``VQA_Dataset_Test`` / ``VQA_Dataset`` objects with the attributes they read (``q_words``,
``feat_dim``, ``n_answers``, ``pretrained_wemb``), items in the tuple order of ``torch_dataset.py:164`` and ``collate_fn``
(``:27-31``: sort by the LAST tuple element - the dataset index - then ``default_collate``).  Items are synthetic
(``vqa_b200.synthetic``), deterministic in the index."""
import numpy as np
import torch
from torch.utils.data import Dataset, dataloader

from vqa_b200.synthetic import Workload, make_batch, make_wemb

# smallest widths on which both graph convolutions take the tensor-core aggregate path
SHAPE = Workload("dropin", 24, 36, 68, hid_dim=512, emb_dim=32, out_dim=120, vocab=200, n_kernels=4, neighbourhood=16,
                 max_qlen=9, q_width=100, dropout=0.0)


def collate_fn(batch):
    batch.sort(key=lambda x: x[-1], reverse=True)
    return dataloader.default_collate(batch)


class VQA_Dataset_Test(Dataset):
    def __init__(self, data_dir, emb_dim=300, train=True):
        w = SHAPE
        assert emb_dim == w.emb_dim, "run the driver with --emb %d" % w.emb_dim
        b = make_batch(w, seed=2024)
        g = torch.Generator().manual_seed(7)
        self.q = b["question"].numpy()
        self.a = b["target"].numpy()
        self.votes = (torch.randint(0, 11, b["target"].shape, generator=g).float() * (b["target"] > 0)).numpy()
        self.img = b["image"].numpy()
        self.qlen = [int(x) for x in b["qlen"]]
        self.q_words, self.feat_dim, self.n_answers = w.vocab, w.feat_dim, w.out_dim
        self.pretrained_wemb = make_wemb(w)
        self.n_questions = w.batch

    def __len__(self):
        return self.n_questions

    def __getitem__(self, idx):
        return (self.q[idx], self.a[idx], self.votes[idx], np.asarray(1000 + idx).reshape(-1), self.img[idx],
                np.asarray(SHAPE.n_obj).reshape(1), self.qlen[idx], idx)


VQA_Dataset = VQA_Dataset_Test
