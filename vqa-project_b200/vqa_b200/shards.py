"""Feature shards and a batch loader that assembles batches ON the GPU.

The reference's input side (``torch_dataset.py:105-164`` + ``collate_fn :27-31`` + ``utils.batch_to_cuda :22-31``) does, per QUESTION,
a zarr read of the image's (K, D) features, a Python loop scaling the K boxes by the image size, a concatenate, two dense
(n_answers,) vectors filled from Python lists, and then copies ~164 MB per 512 questions to the GPU.  At >10^5 questions/s per
GPU that cannot keep up, and at 8 GPUs the host link is the step's bound.

Here the dataset is converted ONCE into flat arrays (``write_shards`` / ``from_reference_records``):

    meta.json            sizes, dtypes
    features.bin         (n_images, K, D)  fp32, or bf16 bit patterns (uint16) with ``feature_dtype="bf16"``
    boxes.bin            (n_images, K, 4)  fp32 xyxy ALREADY divided by the image size (torch_dataset.py:148-154)
    questions.npy        (n_questions, q_width) int32, zero padded        qlen.npy / image_row.npy / qid.npy  (n_questions,)
    ans_{ptr,id,val}.npy CSR of the soft labels  (torch_dataset.py:117-122)
    vote_{ptr,id,val}.npy CSR of the vote counts (torch_dataset.py:125-130)

and ``ShardLoader`` keeps the feature table RESIDENT in HBM (VQA2 trainval: 123 k images x 36 x 2048 = 36 GB fp32 / 18 GB bf16 of
the B200's 180 GB), so a batch costs a few KB of host->device traffic (row indices, tokens, CSR triplets) and two kernels
(``csrc/loader.cu``: ``gather_image``, ``scatter_targets``).  ``resident=False`` streams the batch's rows through pinned memory
instead (tables larger than HBM).  Batches come out in the reference's tuple order ``(q, a, n_votes, qid, i, k, qlen, idx)`` with
the dtypes/shapes ``default_collate`` gives, ``q, a, n_votes, i, k`` already on the device - ``utils.batch_to_cuda`` passes them
through.  One process per GPU: ``rank`` / ``world`` split every global batch into disjoint per-rank slices (no exchange).

fp32 shards reproduce the reference's batches bit for bit; bf16 shards round the FEATURES (not the boxes) to bf16 - exactly the
rounding ``--precision bf16`` applies to them anyway, and a stated loss of input precision in fp32 mode.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterator, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

FORMAT_VERSION = 1
VARIANTS = ("vqa2", "imageclef", "mimic")          # VQA_Dataset / ImageclefDataset / MimicDataset of torch_dataset.py
_NPY = ("questions", "qlen", "image_row", "qid", "ans_ptr", "ans_id", "ans_val", "vote_ptr", "vote_id", "vote_val")


# ------------------------------------------------------------------------------------------------------ writing
def _csr(rows: Sequence[Sequence[Tuple[int, float]]]):
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    for i, r in enumerate(rows):
        ptr[i + 1] = ptr[i] + len(r)
    ids = np.fromiter((a for r in rows for a, _ in r), dtype=np.int32, count=int(ptr[-1]))
    val = np.fromiter((c for r in rows for _, c in r), dtype=np.float32, count=int(ptr[-1]))
    return ptr, ids, val


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even, as ``Tensor.to(torch.bfloat16)``), returned as uint16 bit patterns."""
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16)


def write_shards(out_dir: str, *, features: Optional[np.ndarray] = None, boxes: Optional[np.ndarray] = None, questions: np.ndarray, qlen: np.ndarray,
                 image_row: np.ndarray, qid: np.ndarray, answers: Sequence[Sequence[Tuple[int, float]]],
                 votes: Sequence[Sequence[Tuple[int, float]]], n_answers: int, feature_dtype: str = "f32",
                 image_keys: Optional[Sequence[str]] = None, variant: str = "vqa2", image_iter=None) -> Dict:
    """Write one shard directory.  ``features`` (n_images, K, D) fp32, ``boxes`` (n_images, K, 4) normalised xyxy,
    ``questions`` (n_questions, q_width) token ids, ``answers`` / ``votes``: per question a list of (answer id, value).
    ``image_keys`` (one string per image row; the medical datasets return it as the batch's last element) and ``variant``
    (``"vqa2"`` / ``"imageclef"`` / ``"mimic"``: which reference dataset class the batches mirror) are recorded for the loader.
    Large tables: pass ``features=None, boxes=None`` and ``image_iter`` = an iterable of per-image ``(features (K, D), boxes (K, 4))``
    pairs in row order; they are streamed to disk one image at a time (VQA2 trainval is 36 GB of features)."""
    if feature_dtype not in ("f32", "bf16"):
        raise ValueError(f"feature_dtype must be 'f32' or 'bf16', got {feature_dtype!r}")
    if variant not in VARIANTS:
        raise ValueError(f"variant must be one of {VARIANTS}, got {variant!r}")
    nq = len(questions)
    if not (len(qlen) == len(image_row) == len(qid) == len(answers) == len(votes) == nq):
        raise ValueError("questions, qlen, image_row, qid, answers and votes must have one entry per question")
    if image_iter is None:
        features = np.asarray(features)
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        if features.ndim != 3 or boxes.shape != (features.shape[0], features.shape[1], 4):
            raise ValueError(f"features must be (n_images, K, D) and boxes (n_images, K, 4), got {features.shape} / {boxes.shape}")
        image_iter = zip(features, boxes)
    elif features is not None or boxes is not None:
        raise ValueError("give either features + boxes or image_iter")
    os.makedirs(out_dir, exist_ok=True)
    n_img, K, D = 0, None, None
    with open(os.path.join(out_dir, "features.bin"), "wb") as ff, open(os.path.join(out_dir, "boxes.bin"), "wb") as fb:
        for f, b in image_iter:
            f = np.ascontiguousarray(f, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            if K is None:
                K, D = f.shape
                if D % 8:
                    raise ValueError(f"feature width must be a multiple of 8 (16-byte rows in either dtype), got {D}")
            if f.shape != (K, D) or b.shape != (K, 4):
                raise ValueError(f"image row {n_img}: features {f.shape} / boxes {b.shape}, expected {(K, D)} / {(K, 4)}")
            if not np.isfinite(f).all():
                raise ValueError("non-finite image features")        # the reference raises per item (torch_dataset.py:141-142)
            (to_bf16_bits(f) if feature_dtype == "bf16" else f).tofile(ff)
            b.tofile(fb)
            n_img += 1
    if n_img == 0:
        raise ValueError("no images")
    if variant != "vqa2" and image_keys is None:
        raise ValueError(f"variant={variant!r} needs image_keys (the batches carry the image key as their last element)")
    if image_keys is not None and len(image_keys) != n_img:
        raise ValueError("image_keys must have one entry per image row")
    image_row = np.asarray(image_row, dtype=np.int64)
    if nq and (image_row.min() < 0 or image_row.max() >= n_img):
        raise ValueError("image_row out of range")
    ap, ai, av = _csr(answers)
    vp, vi, vv = _csr(votes)
    for ids in (ai, vi):
        if ids.size and (ids.min() < 0 or ids.max() >= n_answers):
            raise ValueError("answer id out of range")
    arrays = dict(questions=np.ascontiguousarray(questions, dtype=np.int32), qlen=np.asarray(qlen, dtype=np.int32),
                  image_row=image_row, qid=np.asarray(qid, dtype=np.int64), ans_ptr=ap, ans_id=ai, ans_val=av, vote_ptr=vp,
                  vote_id=vi, vote_val=vv)
    for k, v in arrays.items():
        np.save(os.path.join(out_dir, k + ".npy"), v)
    meta = dict(format=FORMAT_VERSION, n_images=int(n_img), n_obj=int(K), feat_width=int(D), feat_dim=int(D + 4),
                feature_dtype=feature_dtype, n_questions=int(nq), q_width=int(arrays["questions"].shape[1]) if nq else 0,
                n_answers=int(n_answers), variant=variant)
    if image_keys is not None:
        with open(os.path.join(out_dir, "image_keys.json"), "w") as f:
            json.dump([str(k) for k in image_keys], f)
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def from_reference_records(records: Sequence[Mapping], q_wtoi: Mapping[str, int], a_wtoi: Mapping[str, int], i_feat: Mapping,
                           bbox: Mapping, sizes: Mapping, out_dir: str, *, n_answers: int, n_obj: Optional[int] = 36, q_width: int = 100,
                           feature_dtype: str = "f32", variant: str = "vqa2") -> Dict:
    """Convert what ``VQA_Dataset.__init__`` loads (``torch_dataset.py:35-75``: the question json ``records``, the two word->index
    dictionaries, the zarr groups ``i_feat`` / ``bbox`` and the image-size table ``sizes``, all indexed by the image key) into
    shards, applying ``__getitem__``'s rules (``:105-164``): unseen question words -> 0, unseen answers skipped, a repeated answer
    keeps its last value, boxes divided by (w, h, w, h), features must be finite.

    ``variant="imageclef"`` / ``"mimic"`` follow ``ImageclefDataset.__getitem__`` (``:236-291``) / ``MimicDataset.__getitem__``
    (``:356-417``) instead: every box of the image is used (``n_obj=None``: taken from the data; it must be the same for all images -
    the model reads one K per batch) and the batch's last element is the image key, not the dataset index; ImageCLEF additionally
    keys images by ``image_id + '.jpg'`` and stores ``answers`` as a dict."""
    if variant not in VARIANTS:
        raise ValueError(f"variant must be one of {VARIANTS}, got {variant!r}")
    medical = variant != "vqa2"

    def key_of(r):
        return str(r["image_id"]) + ".jpg" if variant == "imageclef" else str(r["image_id"])

    keys: List[str] = []
    row_of: Dict[str, int] = {}
    for r in records:
        k = key_of(r)
        if k not in row_of:
            row_of[k] = len(keys)
            keys.append(k)
    state = {"n_obj": n_obj}

    def images():                                                       # streamed: one image in memory at a time
        for k in keys:
            f = np.asarray(i_feat[k], dtype=np.float32)
            b = np.array(np.asarray(bbox[k]), dtype=np.float32)         # a copy: the reference scales in place
            if medical and state["n_obj"] is None:
                state["n_obj"] = b.shape[0]
            n = state["n_obj"]
            if b.shape[0] < n or f.shape[0] < n or (medical and (b.shape[0] != n or f.shape[0] != n)):
                raise ValueError(f"image {k}: {b.shape[0]} boxes / {f.shape[0]} feature rows, expected {n}")
            f, b = f[:n], b[:n]
            w, h = (float(x) for x in np.asarray(sizes[k]).reshape(-1)[:2])
            b[:, 0] /= w
            b[:, 1] /= h
            b[:, 2] /= w
            b[:, 3] /= h
            yield f, b

    questions = np.zeros((len(records), q_width), dtype=np.int32)
    qlen = np.zeros(len(records), dtype=np.int32)
    answers, votes = [], []
    for n, r in enumerate(records):
        toks = r["question_toked"]
        qlen[n] = len(toks)
        for i, wd in enumerate(toks):
            questions[n, i] = q_wtoi.get(wd, 0)
        vote_pairs = r["answers"].items() if variant == "imageclef" else r["answers"]
        answers.append([(a_wtoi[wd], float(c)) for wd, c in r["answers_w_scores"] if wd in a_wtoi])
        votes.append([(a_wtoi[wd], float(c)) for wd, c in vote_pairs if wd in a_wtoi])
    return write_shards(out_dir, image_iter=images(), questions=questions, qlen=qlen,
                        image_row=np.array([row_of[key_of(r)] for r in records], dtype=np.int64),
                        qid=np.array([r["question_id"] for r in records], dtype=np.int64), answers=answers, votes=votes,
                        n_answers=n_answers, feature_dtype=feature_dtype, image_keys=keys, variant=variant)


# ------------------------------------------------------------------------------------------------------ reading
class ShardSet:
    """Memory-mapped view of one shard directory (host side; no CUDA needed)."""

    def __init__(self, path: str):
        with open(os.path.join(path, "meta.json")) as f:
            self.meta = json.load(f)
        if self.meta.get("format") != FORMAT_VERSION:
            raise ValueError(f"{path}: shard format {self.meta.get('format')} != {FORMAT_VERSION}")
        m = self.meta
        self.path = path
        self.n_images, self.n_obj, self.feat_width, self.n_answers = m["n_images"], m["n_obj"], m["feat_width"], m["n_answers"]
        self.n_questions, self.q_width, self.bf16 = m["n_questions"], m["q_width"], m["feature_dtype"] == "bf16"
        self.variant = m.get("variant", "vqa2")
        kp = os.path.join(path, "image_keys.json")
        self.image_keys: Optional[List[str]] = None
        if os.path.exists(kp):
            with open(kp) as f:
                self.image_keys = json.load(f)
        if self.variant not in VARIANTS or (self.variant != "vqa2" and self.image_keys is None):
            raise ValueError(f"{path}: unknown variant {self.variant!r} or medical shards without image_keys.json")
        shape = (self.n_images, self.n_obj, self.feat_width)
        self.features = np.memmap(os.path.join(path, "features.bin"), dtype=np.uint16 if self.bf16 else np.float32, mode="r", shape=shape)
        self.boxes = np.memmap(os.path.join(path, "boxes.bin"), dtype=np.float32, mode="r", shape=(self.n_images, self.n_obj, 4))
        for k in _NPY:
            setattr(self, k, np.load(os.path.join(path, k + ".npy"), mmap_mode="r"))
        if len(self.questions) != self.n_questions or len(self.ans_ptr) != self.n_questions + 1 or len(self.vote_ptr) != self.n_questions + 1:
            raise ValueError(f"{path}: array lengths disagree with meta.json")

    def __len__(self) -> int:
        return self.n_questions

    def csr_rows(self, which: str, idx: np.ndarray):
        """CSR slice for the questions ``idx`` (in that order): (ptr (B+1,) int64, ids int32, vals fp32)."""
        ptr, ids, val = getattr(self, which + "_ptr"), getattr(self, which + "_id"), getattr(self, which + "_val")
        start, stop = np.asarray(ptr[idx]), np.asarray(ptr[idx + 1])
        cnt = stop - start
        out_ptr = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(cnt, out=out_ptr[1:])
        take = np.repeat(start - out_ptr[:-1], cnt) + np.arange(out_ptr[-1], dtype=np.int64)
        return out_ptr, np.asarray(ids[take], dtype=np.int32), np.asarray(val[take], dtype=np.float32)

    def dense_item(self, n: int):
        """One question the way ``VQA_Dataset.__getitem__`` returns it (host, fp32) - the checker's view of the shards."""
        a = np.zeros(self.n_answers, dtype=np.float32)
        v = np.zeros(self.n_answers, dtype=np.float32)
        for e in range(self.ans_ptr[n], self.ans_ptr[n + 1]):
            a[self.ans_id[e]] = self.ans_val[e]
        for e in range(self.vote_ptr[n], self.vote_ptr[n + 1]):
            v[self.vote_id[e]] = self.vote_val[e]
        r = int(self.image_row[n])
        f = np.asarray(self.features[r])
        if self.bf16:
            f = (f.astype(np.uint32) << 16).view(np.float32)
        img = np.concatenate([f, np.asarray(self.boxes[r])], axis=1)
        return (np.asarray(self.questions[n], dtype=np.int64), a, v, np.asarray(self.qid[n]).reshape(-1), img,
                np.asarray(self.n_obj).reshape(1), int(self.qlen[n]), self.last_element(n))

    def last_element(self, n: int):
        """What the reference item carries last: the dataset index (``torch_dataset.py:164``) or, for the medical datasets, the
        image key (``:291``) - also the key ``collate_fn`` sorts a batch by."""
        return self.image_keys[int(self.image_row[n])] if self.variant != "vqa2" else n


def epoch_batches(n_questions: int, batch_size: int, *, epoch: int = 0, shuffle: bool = True, seed: int = 1000, rank: int = 0,
                  world: int = 1, drop_last: bool = False) -> List[np.ndarray]:
    """Question indices of every batch of one epoch for ``rank``: the (seeded, rank-independent) permutation is cut into global
    batches of ``batch_size * world`` and each is dealt out in ``world`` contiguous slices, so ranks never share a question.
    A short last global batch is split as evenly as possible (or dropped)."""
    if shuffle:
        perm = torch.randperm(n_questions, generator=torch.Generator().manual_seed(seed + epoch)).numpy()
    else:
        perm = np.arange(n_questions, dtype=np.int64)
    gb = batch_size * world
    out = []
    for s in range(0, n_questions, gb):
        chunk = perm[s:s + gb]
        if len(chunk) < gb and (drop_last or len(chunk) < world):
            break
        per = len(chunk) // world                                  # equal counts on every rank (a remainder < world is dropped)
        out.append(np.asarray(chunk[rank * per:(rank + 1) * per], dtype=np.int64))
    return out


def order_batch(idx: np.ndarray, qlen: np.ndarray, order: str, keys: Optional[Sequence[str]] = None) -> np.ndarray:
    """In-batch order.  ``"reference"``: what ``collate_fn`` does - ``batch.sort(key=lambda x: x[-1], reverse=True)`` with the
    dataset index as the last tuple element (``torch_dataset.py:27-31,164``), i.e. descending INDEX.  ``"qlen"``: descending
    question length (what that function's comment intends), stable - lets the GRU's row-tile gate skip finished tiles.
    ``keys``: per entry of ``idx`` the string the medical datasets carry last (their ``collate_fn`` order is by that string)."""
    if order == "reference":
        if keys is not None:                                            # medical datasets: the last element is the image key (a str)
            pos = sorted(range(len(idx)), key=lambda j: keys[j], reverse=True)      # list.sort semantics: stable, also reversed
            return idx[np.asarray(pos, dtype=np.int64)]
        return idx[np.argsort(-idx, kind="stable")]
    if order == "qlen":
        return idx[np.argsort(-np.asarray(qlen[idx], dtype=np.int64), kind="stable")]
    if order == "none":
        return idx
    raise ValueError(f"order must be 'reference', 'qlen' or 'none', got {order!r}")


class ShardLoader:
    """Iterates one epoch of device-assembled batches ``(q, a, n_votes, qid, i, k, qlen, idx)`` (see the module docstring)."""

    def __init__(self, path: str, batch_size: int, device="cuda", *, shuffle: bool = True, seed: int = 1000, rank: int = 0,
                 world: int = 1, drop_last: bool = False, resident: bool = True, order: str = "qlen", upload_rows: int = 4096):
        from . import kernels as kn                                   # loads libvqa_sm100.so: raises if it is missing
        self._kn = kn
        self.set = ShardSet(path)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ShardLoader assembles batches on the GPU: device must be a CUDA device (no CPU fallback)")
        self.batch_size, self.shuffle, self.seed, self.rank, self.world = batch_size, shuffle, seed, rank, world
        self.drop_last, self.order, self.resident, self.epoch = drop_last, order, resident, 0
        order_batch(np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32), order)      # validates ``order``
        s = self.set
        fdt = torch.bfloat16 if s.bf16 else torch.float32
        self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
        if resident:
            self.features = torch.empty((s.n_images, s.n_obj, s.feat_width), dtype=fdt, device=self.device)
            self.boxes = torch.empty((s.n_images, s.n_obj, 4), dtype=torch.float32, device=self.device)
            for r0 in range(0, s.n_images, upload_rows):                # bounded host staging: upload_rows images at a time
                r1 = min(s.n_images, r0 + upload_rows)
                self.features[r0:r1].copy_(self._as_tensor(s.features[r0:r1], s.bf16))
                self.boxes[r0:r1].copy_(self._as_tensor(s.boxes[r0:r1], False))
        else:
            self._pin_f = torch.empty((batch_size, s.n_obj, s.feat_width), dtype=fdt).pin_memory()
            self._pin_b = torch.empty((batch_size, s.n_obj, 4), dtype=torch.float32).pin_memory()
            self._staged = torch.cuda.Event()
            self._staged.record()

    @staticmethod
    def _as_tensor(a: np.ndarray, bf16: bool) -> torch.Tensor:
        a = np.array(a, copy=True, order="C")                          # a writable host copy (the maps are read-only)
        return torch.from_numpy(a.view(np.int16)).view(torch.bfloat16) if bf16 else torch.from_numpy(a)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def batches(self) -> List[np.ndarray]:
        return [self.ordered(b) for b in
                epoch_batches(len(self.set), self.batch_size, epoch=self.epoch, shuffle=self.shuffle, seed=self.seed,
                              rank=self.rank, world=self.world, drop_last=self.drop_last)]

    def ordered(self, idx: np.ndarray) -> np.ndarray:
        """``idx`` in this loader's in-batch order (``order_batch``; the medical variant's reference order is by image key)."""
        keys = [self.set.last_element(int(n)) for n in idx] if self.set.variant != "vqa2" and self.order == "reference" else None
        return order_batch(idx, self.set.qlen, self.order, keys)

    def __len__(self) -> int:
        return len(epoch_batches(len(self.set), self.batch_size, epoch=self.epoch, shuffle=False, rank=self.rank, world=self.world,
                                 drop_last=self.drop_last))

    def assemble(self, idx: np.ndarray, image_out: Optional[torch.Tensor] = None):
        """One batch for the question indices ``idx`` (already ordered), enqueued on the current stream.  ``image_out``: a (B, K, D+4)
        fp32 device tensor the image batch is gathered INTO (``engine.TrainStep.input_slot("image")``: the step then reads it in
        place instead of copying 151 MB once more); ignored when its shape does not fit (the short last batch of an epoch)."""
        s, dev, kn = self.set, self.device, self._kn
        B = len(idx)
        sidx = np.sort(idx)                                             # memmap reads in file order
        back = np.searchsorted(sidx, idx)
        q = torch.from_numpy(np.asarray(s.questions[sidx])[back].astype(np.int64)).to(dev, non_blocking=True)
        rows_np = np.asarray(s.image_row[sidx])[back]
        if self.resident:
            rows = torch.from_numpy(rows_np).to(dev, non_blocking=True)
            image = kn.gather_image(self.features, self.boxes, rows, self._err, out=self._fits(image_out, B))
        else:
            self._staged.synchronize()                                  # the previous batch's H2D copy has left the staging buffers
            urows, inv = np.unique(rows_np, return_inverse=True)         # questions of one image share one staged row
            n = len(urows)
            self._pin_f[:n].copy_(self._as_tensor(s.features[urows], s.bf16))
            self._pin_b[:n].copy_(self._as_tensor(s.boxes[urows], False))
            f_dev = self._pin_f[:n].to(dev, non_blocking=True)
            b_dev = self._pin_b[:n].to(dev, non_blocking=True)
            self._staged.record()
            image = kn.gather_image(f_dev, b_dev, torch.from_numpy(inv.astype(np.int64)).to(dev, non_blocking=True), self._err,
                                    out=self._fits(image_out, B))
        ap, ai, av = s.csr_rows("ans", idx)
        vp, vi, vv = s.csr_rows("vote", idx)
        a = kn.scatter_targets(*(torch.from_numpy(x).to(dev, non_blocking=True) for x in (ap, ai, av)), B, s.n_answers, self._err)
        n_votes = kn.scatter_targets(*(torch.from_numpy(x).to(dev, non_blocking=True) for x in (vp, vi, vv)), B, s.n_answers, self._err)
        qid = torch.from_numpy(np.asarray(s.qid[sidx])[back].reshape(B, 1).astype(np.int64))
        k = torch.full((B, 1), s.n_obj, dtype=torch.int64, device=dev)
        qlen = torch.from_numpy(np.asarray(s.qlen[sidx])[back].astype(np.int64))
        last = [s.last_element(int(n)) for n in idx] if s.variant != "vqa2" else torch.from_numpy(idx.astype(np.int64))
        return q, a, n_votes, qid, image, k, qlen, last

    def _fits(self, image_out, B):
        s = self.set
        want = (B, s.n_obj, int(s.features.shape[-1]) + 4)
        ok = image_out is not None and tuple(image_out.shape) == want and image_out.dtype == torch.float32 and image_out.is_contiguous()
        return image_out if ok else None

    def check_errors(self) -> None:
        """Raise if a kernel met an index outside its table (synchronises; called at the end of every epoch)."""
        e = int(self._err.item())
        if e:
            self._err.zero_()
            raise RuntimeError(f"ShardLoader: {'image row' if e == 1 else 'answer id'} out of range in the shard arrays")

    def __iter__(self) -> Iterator:
        for idx in self.batches():
            yield self.assemble(idx)
        self.check_errors()
        self.epoch += 1
