"""Convert a reference dataset into feature shards (vqa_b200.shards), using the reference's OWN loader classes to read the files.

    PYTHONPATH=/path/to/vqa-project python tools/convert_dataset.py --dataset vqa2 --data-dir data --split train \
        --out /data/shards/vqa2_train [--feature-dtype bf16]

Needs what the reference needs to open its data (``zarr``, ``pandas``, the GloVe file the dataset constructors read).  The dataset
object is built exactly as ``run.py:357`` / ``run_imageclef.py`` / ``run_mimic.py`` build it; the conversion applies its
``__getitem__`` rules to every question once (``shards.from_reference_records``).  The word-embedding matrix the constructor
prepares (``pretrained_wemb``, passed to ``Model`` at ``run.py:376``) is saved next to the shards as ``pretrained_wemb.npy``.
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vqa-project_b200"))
from vqa_b200 import shards  # noqa: E402


def convert(dataset, out_dir: str, variant: str, feature_dtype: str = "f32"):
    """``dataset``: a loaded ``VQA_Dataset`` / ``ImageclefDataset`` / ``MimicDataset`` (or any object with their attributes)."""
    meta = shards.from_reference_records(dataset.vqa, dataset.q_wtoi, dataset.a_wtoi, dataset.i_feat, dataset.bbox, dataset.sizes, out_dir,
                                         n_answers=dataset.n_answers, n_obj=36 if variant == "vqa2" else None, variant=variant,
                                         feature_dtype=feature_dtype)
    wemb = getattr(dataset, "pretrained_wemb", None)
    if wemb is not None:
        np.save(os.path.join(out_dir, "pretrained_wemb.npy"), np.asarray(wemb, dtype=np.float32))
    return meta


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dataset", required=True, choices=list(shards.VARIANTS))
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--split", default="train", choices=["train", "val"])
    ap.add_argument("--emb", type=int, default=300)
    ap.add_argument("--out", required=True)
    ap.add_argument("--feature-dtype", default="f32", choices=["f32", "bf16"])
    args = ap.parse_args()
    import torch_dataset as T                       # the reference's module: put its checkout on PYTHONPATH
    train = args.split == "train"
    if args.dataset == "vqa2":
        ds = T.VQA_Dataset(args.data_dir, args.emb, train=train)
    else:
        cls = T.ImageclefDataset if args.dataset == "imageclef" else T.MimicDataset
        ds = cls(types.SimpleNamespace(data_dir=args.data_dir, emb=args.emb), train=train)
    meta = convert(ds, args.out, args.dataset, args.feature_dtype)
    print(f"wrote {args.out}: {meta}")


if __name__ == "__main__":
    main()
