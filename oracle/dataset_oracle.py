"""CPU oracle for the INPUT side of the path (dataset item + collate)  --  TEST INFRASTRUCTURE ONLY.

A restatement of ``VQA_Dataset.__getitem__`` (``torch_dataset.py:105-164``) and ``collate_fn`` (``:27-31``) over plain Python
containers (dicts in place of the zarr groups / the pandas size table), used as the checker of ``vqa_b200.shards`` and
``csrc/loader.cu``.  Only ``tests/`` import it; the product package never does.

Parity status: **pinned**.  The reference has no tests or fixtures for its dataset code, so the pin is that code itself:
``tests/golden/make_dataset_golden.py`` imports the unmodified ``torch_dataset.py`` from ``/root/reference`` in the build
container (with an empty stand-in for the absent ``zarr`` package, which only ``__init__`` uses), fills a ``VQA_Dataset`` with the
miniature dataset of ``tests/shard_fixture.py`` and commits every ``__getitem__`` result and two collated batches as
``tests/golden/dataset_small.npz``; ``tests/test_shards_cpu.py::test_dataset_oracle_matches_reference_golden`` checks this file
against them bit for bit.
"""
from __future__ import annotations

import numpy as np
from torch.utils.data import dataloader


def reference_item(ds, idx, q_width=100):
    """``torch_dataset.py:105-164`` in its own statement order.  ``ds``: dict with ``records`` (the question json), ``q_wtoi``,
    ``a_wtoi``, ``i_feat`` / ``bbox`` / ``sizes`` keyed by ``str(image_id)``, ``n_answers`` and ``K`` (the reference hard-codes 36)."""
    rec = ds["records"][idx]
    qlen = len(rec["question_toked"])                                   # :108
    q = [0] * q_width                                                   # :109 (width 100 in the reference)
    for i, w in enumerate(rec["question_toked"]):                       # :110-114 unseen word -> 0
        try:
            q[i] = ds["q_wtoi"][w]
        except KeyError:
            q[i] = 0
    a = np.zeros(ds["n_answers"], dtype=np.float32)                     # :117-122 soft labels; unseen answer skipped
    for w, c in rec["answers_w_scores"]:
        try:
            a[ds["a_wtoi"][w]] = c
        except KeyError:
            continue
    n_votes = np.zeros(ds["n_answers"], dtype=np.float32)               # :125-130 vote counts
    for w, c in rec["answers"]:
        try:
            n_votes[ds["a_wtoi"][w]] = c
        except KeyError:
            continue
    qid = rec["question_id"]                                            # :133
    iid = rec["image_id"]                                               # :136-139
    img = ds["i_feat"][str(iid)]
    bboxes = np.array(ds["bbox"][str(iid)])                             # (a zarr read returns a fresh array)
    imsize = ds["sizes"][str(iid)]
    if np.logical_not(np.isfinite(img)).sum() > 0:                      # :141-142
        raise ValueError
    k = ds["K"]                                                         # :145
    for i in range(k):                                                  # :148-154 boxes / (w, h, w, h), in place, in the boxes' dtype
        bb = bboxes[i]
        bb[0] /= imsize[0]
        bb[1] /= imsize[1]
        bb[2] /= imsize[0]
        bb[3] /= imsize[1]
        bboxes[i] = bb
    return (np.asarray(q), np.asarray(a).reshape(-1), np.asarray(n_votes).reshape(-1), np.asarray(qid).reshape(-1),   # :157-164
            np.concatenate([img, bboxes], axis=1), np.asarray(k).reshape(1), qlen, idx)


def reference_item_medical(ds, idx, q_width=100, variant="imageclef"):
    """``ImageclefDataset.__getitem__`` (``torch_dataset.py:236-291``) and ``MimicDataset.__getitem__`` (``:356-417``): as above
    except that every box of the image is used and the last element is the image key; ImageCLEF keys images by
    ``image_id + '.jpg'`` (``:268``) and iterates ``answers`` as a dict (``:258``), MIMIC does neither (``:375,385``)."""
    clef = variant == "imageclef"
    rec = ds["records"][idx]
    qlen = len(rec["question_toked"])
    q = [0] * q_width
    for i, w in enumerate(rec["question_toked"]):
        try:
            q[i] = ds["q_wtoi"][w]
        except KeyError:
            q[i] = 0
    a = np.zeros(ds["n_answers"], dtype=np.float32)
    for w, c in rec["answers_w_scores"]:
        try:
            a[ds["a_wtoi"][w]] = c
        except KeyError:
            continue
    n_votes = np.zeros(ds["n_answers"], dtype=np.float32)
    for w, c in (rec["answers"].items() if clef else rec["answers"]):   # :258 a dict (ImageCLEF) / :375 pairs (MIMIC)
        try:
            n_votes[ds["a_wtoi"][w]] = c
        except KeyError:
            continue
    qid = rec["question_id"]
    iid = rec["image_id"] + ".jpg" if clef else rec["image_id"]         # :268 / :385
    img = ds["i_feat"][str(iid)]
    bboxes = np.array(ds["bbox"][str(iid)])
    imsize = ds["sizes"][str(iid)]
    if np.logical_not(np.isfinite(img)).sum() > 0:
        raise ValueError
    for i in range(bboxes.shape[0]):                                    # :279-285 every box
        bb = bboxes[i]
        bb[0] /= imsize[0]
        bb[1] /= imsize[1]
        bb[2] /= imsize[0]
        bb[3] /= imsize[1]
        bboxes[i] = bb
    return (np.asarray(q), np.asarray(a).reshape(-1), np.asarray(n_votes).reshape(-1), np.asarray(qid).reshape(-1),
            np.concatenate([img, bboxes], axis=1), np.asarray(bboxes.shape[0]).reshape(1), qlen, iid)


def reference_collate(batch):
    """``torch_dataset.py:27-31``: sorts by the LAST tuple element - the dataset index (``:164``), not the question length its comment
    speaks of - in descending order, then ``default_collate``."""
    batch.sort(key=lambda x: x[-1], reverse=True)
    return dataloader.default_collate(batch)
