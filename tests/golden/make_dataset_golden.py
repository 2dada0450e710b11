"""Generate ``dataset_small.npz`` from the UNMODIFIED reference dataset code.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_dataset_golden.py

``torch_dataset.py`` imports ``zarr`` (absent from this image) at module level but only ``VQA_Dataset.__init__`` uses it, so an
empty stand-in module satisfies the import.  A ``VQA_Dataset`` is created without running ``__init__`` (which needs the real data
files) and given the attributes ``__getitem__`` reads - ``vqa``, ``q_wtoi``, ``a_wtoi``, ``n_answers``, ``i_feat``, ``bbox`` (dicts
of arrays standing in for the zarr groups) and ``sizes`` (a pandas DataFrame with one column per image id, as ``pd.read_csv`` of
the size table gives) - from the miniature dataset of ``tests/shard_fixture.py`` with K = 36 boxes (the reference hard-codes 36).
Every item and two collated batches (the reference's own ``collate_fn``) are stored, and the same for ``ImageclefDataset`` and
``MimicDataset`` (K = 51 boxes, string image keys); ``oracle/dataset_oracle.py`` is pinned to them.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import shard_fixture as SF  # noqa: E402

REF = "/root/reference"
GOLDEN_ARGS = dict(n_images=5, n_questions=19, K=36, D=16, n_answers=11, seed=3)
MEDICAL_ARGS = dict(n_images=6, n_questions=17, K=51, D=16, n_answers=9, seed=5)
BATCHES = ([3, 17, 9, 0, 12], [18, 1, 2, 7])
MEDICAL_BATCHES = ([3, 16, 9, 0, 12, 5], [1, 2, 7])


def main():
    sys.modules.setdefault("zarr", types.ModuleType("zarr"))
    sys.path.insert(0, REF)
    import torch_dataset as T
    assert T.__file__.startswith(REF)
    sys.path.remove(REF)
    ds = SF.make_dataset(**GOLDEN_ARGS)
    obj = T.VQA_Dataset.__new__(T.VQA_Dataset)
    obj.vqa, obj.q_wtoi, obj.a_wtoi, obj.n_answers = ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["n_answers"]
    obj.i_feat, obj.bbox = ds["i_feat"], {k: v.copy() for k, v in ds["bbox"].items()}
    obj.sizes = pd.DataFrame({k: v for k, v in ds["sizes"].items()})
    obj.n_questions = len(ds["records"])
    out = {}
    names = ("q", "a", "n_votes", "qid", "i", "k", "qlen", "idx")
    items = []
    for n in range(len(obj)):
        obj.bbox = {k: v.copy() for k, v in ds["bbox"].items()}          # __getitem__ scales in place: a zarr read would be fresh
        items.append(obj[n])
    for j, name in enumerate(names):
        out["item." + name] = np.stack([np.asarray(it[j]) for it in items])
    for b, idx in enumerate(BATCHES):
        obj.bbox = None
        batch = []
        for n in idx:
            obj.bbox = {k: v.copy() for k, v in ds["bbox"].items()}
            batch.append(obj[n])
        col = T.collate_fn(batch)
        for j, name in enumerate(names):
            out[f"batch{b}.{name}"] = col[j].numpy()
    # ---- the medical datasets: ImageclefDataset.__getitem__ and MimicDataset.__getitem__ (its own override), same construction
    for variant, cls in (("imageclef", T.ImageclefDataset), ("mimic", T.MimicDataset)):
        ds = SF.make_medical_dataset(variant=variant, **MEDICAL_ARGS)
        obj = cls.__new__(cls)
        obj.vqa, obj.q_wtoi, obj.a_wtoi, obj.n_answers = ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["n_answers"]
        obj.i_feat = ds["i_feat"]
        obj.sizes = pd.DataFrame({k: v for k, v in ds["sizes"].items()})
        obj.n_questions = len(ds["records"])

        def med_item(n):
            obj.bbox = {k: v.copy() for k, v in ds["bbox"].items()}
            return obj[n]

        items = [med_item(n) for n in range(len(obj))]
        for j, name in enumerate(names[:-1]):
            out[f"{variant}.item.{name}"] = np.stack([np.asarray(it[j]) for it in items])
        out[f"{variant}.item.iid"] = np.array([it[7] for it in items])
        for b, idx in enumerate(MEDICAL_BATCHES):
            col = T.collate_fn([med_item(n) for n in idx])
            for j, name in enumerate(names[:-1]):
                out[f"{variant}.batch{b}.{name}"] = col[j].numpy()
            out[f"{variant}.batch{b}.iid"] = np.array(col[7])
            out[f"{variant}.batch{b}.src"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "dataset_small.npz"), **out)
    print("wrote dataset_small.npz:", {k: v.shape for k, v in out.items() if k.startswith("item.")})


if __name__ == "__main__":
    main()
