"""A miniature dataset in the shape ``VQA_Dataset.__init__`` loads it (torch_dataset.py:35-75): question records, word/answer
dictionaries, per-image features, boxes and sizes keyed by ``str(image_id)``.  Deterministic in its arguments; shared by the shard
tests and by ``tests/golden/make_dataset_golden.py``.  The checker that turns it into items/batches is ``oracle/dataset_oracle.py``."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.dataset_oracle import reference_collate, reference_item, reference_item_medical  # noqa: E402,F401  (re-exported for the tests)


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


def make_medical_dataset(n_images=6, n_questions=17, K=51, D=16, n_answers=9, seed=5, variant="imageclef"):
    """The same in the shape ``ImageclefDataset`` (string image ids keyed as ``id + '.jpg'``, ``answers`` a dict of vote counts) or
    ``MimicDataset`` (string ids used as they are, ``answers`` a list of pairs) load it; K boxes per image, all used."""
    ds = make_dataset(n_images=n_images, n_questions=n_questions, K=K, D=D, n_answers=n_answers, seed=seed)
    ren = {}
    for j, old in enumerate(sorted(ds["i_feat"], key=int)):
        ren[old] = f"synpic{(j * 7919) % 1000:03d}"
    suffix = ".jpg" if variant == "imageclef" else ""
    for name in ("i_feat", "bbox", "sizes"):
        ds[name] = {ren[k] + suffix: v for k, v in ds[name].items()}
    for r in ds["records"]:
        r["image_id"] = ren[str(r["image_id"])]
        if variant == "imageclef":
            r["answers"] = {w: c for w, c in r["answers"]}
    return ds
