"""A miniature dataset in the shape ``VQA_Dataset.__init__`` loads it (torch_dataset.py:35-75) plus a restatement of
``__getitem__`` / ``collate_fn`` (:27-31, :105-164) used ONLY as the checker of vqa_b200.shards."""
import numpy as np
import torch
from torch.utils.data import dataloader


def make_dataset(n_images=7, n_questions=23, K=6, D=16, n_answers=11, seed=3):
    rng = np.random.RandomState(seed)
    words = [f"w{i}" for i in range(30)]
    q_wtoi = {w: i + 1 for i, w in enumerate(words[:24])}              # w24.. are unseen -> 0
    ans = [f"a{i}" for i in range(14)]
    a_wtoi = {w: i + 1 for i, w in enumerate(ans[:n_answers - 1])}     # a10.. are unseen -> skipped
    image_ids = [100 + 7 * i for i in range(n_images)]
    i_feat = {str(i): np.maximum(rng.randn(K, D), 0).astype(np.float32) for i in image_ids}
    sizes = {str(i): np.array([int(rng.randint(200, 640)), int(rng.randint(200, 480))]) for i in image_ids}
    bbox = {}
    for i in image_ids:
        w, h = sizes[str(i)]
        xy1 = rng.rand(K, 2) * [w * 0.7, h * 0.7]
        wh = rng.rand(K, 2) * [w * 0.25, h * 0.25] + 5
        bbox[str(i)] = np.concatenate([xy1, np.minimum(xy1 + wh, [w, h])], axis=1).astype(np.float32)
    records = []
    for n in range(n_questions):
        ql = int(rng.randint(1, 9))
        toks = [words[int(rng.randint(0, 30))] for _ in range(ql)]
        picks = [ans[int(rng.randint(0, 14))] for _ in range(int(rng.randint(1, 5)))]
        if n % 5 == 0:
            picks.append(picks[0])                                      # a repeated answer: the last value wins
        records.append(dict(question_toked=toks, image_id=image_ids[int(rng.randint(0, n_images))], question_id=9000 + 3 * n,
                            answers_w_scores=[(p, round(float(rng.rand()), 3)) for p in picks],
                            answers=[(p, float(rng.randint(1, 11))) for p in picks]))
    return dict(records=records, q_wtoi=q_wtoi, a_wtoi=a_wtoi, i_feat=i_feat, bbox=bbox, sizes=sizes, n_answers=n_answers, K=K, D=D)


def reference_item(ds, idx, q_width=100):
    """torch_dataset.py:105-164, statement by statement (k = number of boxes of the fixture)."""
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
    for w, c in rec["answers"]:
        try:
            n_votes[ds["a_wtoi"][w]] = c
        except KeyError:
            continue
    qid = rec["question_id"]
    iid = rec["image_id"]
    img = ds["i_feat"][str(iid)]
    bboxes = np.array(ds["bbox"][str(iid)])                             # (zarr reads return fresh arrays)
    imsize = ds["sizes"][str(iid)]
    if np.logical_not(np.isfinite(img)).sum() > 0:
        raise ValueError
    k = ds["K"]
    for i in range(k):
        bb = bboxes[i]
        bb[0] /= imsize[0]
        bb[1] /= imsize[1]
        bb[2] /= imsize[0]
        bb[3] /= imsize[1]
        bboxes[i] = bb
    return (np.asarray(q), np.asarray(a).reshape(-1), np.asarray(n_votes).reshape(-1), np.asarray(qid).reshape(-1),
            np.concatenate([img, bboxes], axis=1), np.asarray(k).reshape(1), qlen, idx)


def reference_collate(batch):
    batch.sort(key=lambda x: x[-1], reverse=True)                       # torch_dataset.py:27-31
    return dataloader.default_collate(batch)
