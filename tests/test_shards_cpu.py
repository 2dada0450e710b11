"""Host side of the shard pipeline (vqa_b200/shards.py) against a restatement of the reference's dataset code (shard_fixture.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import shard_fixture as SF  # noqa: E402
from vqa_b200 import shards  # noqa: E402


@pytest.fixture(scope="module")
def ds():
    return SF.make_dataset()


def _convert(ds, path, dtype="f32", q_width=100):
    return shards.from_reference_records(ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["i_feat"], ds["bbox"], ds["sizes"], str(path),
                                         n_answers=ds["n_answers"], n_obj=ds["K"], q_width=q_width, feature_dtype=dtype)


def test_fp32_shards_reproduce_every_reference_item_bit_for_bit(ds, tmp_path):
    meta = _convert(ds, tmp_path)
    s = shards.ShardSet(str(tmp_path))
    assert meta["n_questions"] == len(ds["records"]) == len(s) and meta["feat_dim"] == ds["D"] + 4
    assert s.n_images == len({r["image_id"] for r in ds["records"]})
    for n in range(len(s)):
        ref, got = SF.reference_item(ds, n), s.dense_item(n)
        for j, (r, g) in enumerate(zip(ref, got)):
            if isinstance(r, np.ndarray):
                assert r.shape == g.shape and np.array_equal(r, g), (n, j)
            else:
                assert r == g, (n, j)
        assert got[4].dtype == np.float32 and got[0].dtype == np.int64


def test_bf16_shards_round_features_only(ds, tmp_path):
    _convert(ds, tmp_path, "bf16")
    s = shards.ShardSet(str(tmp_path))
    assert s.bf16 and s.features.dtype == np.uint16
    for n in range(len(s)):
        ref, got = SF.reference_item(ds, n), s.dense_item(n)
        D = ds["D"]
        want = torch.from_numpy(ref[4][:, :D].copy()).to(torch.bfloat16).float().numpy()
        assert np.array_equal(got[4][:, :D], want)
        assert np.array_equal(got[4][:, D:], ref[4][:, D:])            # boxes stay fp32
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_csr_rows_in_batch_order(ds, tmp_path):
    _convert(ds, tmp_path)
    s = shards.ShardSet(str(tmp_path))
    idx = np.array([5, 0, 22, 7, 7 + 1], dtype=np.int64)
    for which, col in (("ans", 1), ("vote", 2)):
        ptr, ids, val = s.csr_rows(which, idx)
        assert ptr.dtype == np.int64 and ids.dtype == np.int32 and val.dtype == np.float32 and ptr[0] == 0
        for b, n in enumerate(idx):
            dense = np.zeros(s.n_answers, dtype=np.float32)
            for e in range(ptr[b], ptr[b + 1]):
                dense[ids[e]] = val[e]
            assert np.array_equal(dense, SF.reference_item(ds, int(n))[col])


def test_epoch_batches_partition_the_epoch_across_ranks():
    n, bs, world = 103, 8, 3
    seen = []
    for rank in range(world):
        bl = shards.epoch_batches(n, bs, epoch=2, seed=5, rank=rank, world=world)
        assert all(len(b) == len(bl[0]) for b in bl[:-1]) and len(bl[0]) == bs
        seen.append(bl)
    assert len({len(b) for b in seen}) == 1                            # same number of steps on every rank (no hung collective)
    for step in zip(*seen):
        assert len({len(b) for b in step}) == 1                        # and the same batch size per step
    flat = np.concatenate([np.concatenate(b) for b in seen])
    assert len(np.unique(flat)) == len(flat) and len(flat) >= n - (n % (bs * world)) and len(flat) <= n
    again = shards.epoch_batches(n, bs, epoch=2, seed=5, rank=1, world=world)
    assert all(np.array_equal(a, b) for a, b in zip(again, seen[1]))
    other = shards.epoch_batches(n, bs, epoch=3, seed=5, rank=1, world=world)
    assert not all(np.array_equal(a, b) for a, b in zip(other, seen[1]))
    full = shards.epoch_batches(n, bs, shuffle=False, drop_last=True)
    assert len(full) == n // bs and np.array_equal(np.concatenate(full), np.arange(n // bs * bs))
    assert sum(len(b) for b in shards.epoch_batches(n, bs, shuffle=False)) == n


def test_in_batch_order_modes(ds):
    qlen = np.array([len(r["question_toked"]) for r in ds["records"]], dtype=np.int32)
    idx = np.array([3, 17, 9, 0, 21, 12], dtype=np.int64)
    ref = SF.reference_collate([SF.reference_item(ds, int(n)) for n in idx])
    assert np.array_equal(shards.order_batch(idx, qlen, "reference"), ref[7].numpy())      # collate_fn sorts by the LAST element: idx
    byq = shards.order_batch(idx, qlen, "qlen")
    assert sorted(byq.tolist()) == sorted(idx.tolist()) and np.all(np.diff(qlen[byq]) <= 0)
    for a, b in zip(byq[:-1], byq[1:]):                                # stable: ties keep the sampling order
        if qlen[a] == qlen[b]:
            assert list(idx).index(a) < list(idx).index(b)
    assert np.array_equal(shards.order_batch(idx, qlen, "none"), idx)
    with pytest.raises(ValueError):
        shards.order_batch(idx, qlen, "random")


def test_writer_rejects_bad_input(tmp_path):
    ok = dict(features=np.zeros((2, 3, 8), np.float32), boxes=np.zeros((2, 3, 4), np.float32), questions=np.zeros((1, 10), np.int32),
              qlen=[1], image_row=[0], qid=[1], answers=[[(1, 0.5)]], votes=[[(1, 3.0)]], n_answers=4)
    shards.write_shards(str(tmp_path / "ok"), **ok)
    bad = dict(ok, features=np.full((2, 3, 8), np.nan, np.float32))
    with pytest.raises(ValueError, match="non-finite"):
        shards.write_shards(str(tmp_path / "a"), **bad)
    with pytest.raises(ValueError, match="image_row"):
        shards.write_shards(str(tmp_path / "b"), **dict(ok, image_row=[2]))
    with pytest.raises(ValueError, match="answer id"):
        shards.write_shards(str(tmp_path / "c"), **dict(ok, answers=[[(4, 1.0)]]))
    with pytest.raises(ValueError, match="multiple of 8"):
        shards.write_shards(str(tmp_path / "d"), **dict(ok, features=np.zeros((2, 3, 12), np.float32)))
    with pytest.raises(ValueError, match="feature_dtype"):
        shards.write_shards(str(tmp_path / "e"), **dict(ok, feature_dtype="fp16"))
    with pytest.raises(RuntimeError, match="CUDA"):
        shards.ShardLoader(str(tmp_path / "ok"), 1, device="cpu")


# ---------------------------------------------------------------------------------------------- pinned to the reference's own code
GOLDEN_ARGS = dict(n_images=5, n_questions=19, K=36, D=16, n_answers=11, seed=3)      # as tests/golden/make_dataset_golden.py
NAMES = ("q", "a", "n_votes", "qid", "i", "k", "qlen", "idx")


def _golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_small.npz"))
    return {k: z[k] for k in z.files}


def test_dataset_oracle_matches_reference_golden():
    """oracle/dataset_oracle.py against what the unmodified torch_dataset.py returned for the same miniature dataset."""
    g, ds = _golden(), SF.make_dataset(**GOLDEN_ARGS)
    assert g["item.i"].dtype == np.float32 and g["item.q"].dtype == np.int64 and g["item.a"].dtype == np.float32
    for n in range(len(ds["records"])):
        it = SF.reference_item(ds, n)
        for j, name in enumerate(NAMES):
            want = g["item." + name][n]
            got = np.asarray(it[j])
            assert got.shape == want.shape and np.array_equal(got, want), (n, name)
            assert got.dtype == want.dtype or name in ("qlen", "idx"), (n, name, got.dtype, want.dtype)
    for b, idx in enumerate(([3, 17, 9, 0, 12], [18, 1, 2, 7])):
        col = SF.reference_collate([SF.reference_item(ds, n) for n in idx])
        for j, name in enumerate(NAMES):
            want = g[f"batch{b}.{name}"]
            assert tuple(col[j].shape) == want.shape and np.array_equal(col[j].numpy(), want), (b, name)
            assert col[j].numpy().dtype == want.dtype, (b, name)
        assert list(g[f"batch{b}.idx"]) == sorted(idx, reverse=True)     # the reference sorts by index, not by length


def test_shards_reproduce_the_reference_golden_items(tmp_path):
    """The converter + shard reader against the reference's own __getitem__ outputs (K = 36, the reference's hard-coded box count)."""
    g, ds = _golden(), SF.make_dataset(**GOLDEN_ARGS)
    _convert(ds, tmp_path)
    s = shards.ShardSet(str(tmp_path))
    for n in range(len(s)):
        it = s.dense_item(n)
        for j, name in enumerate(NAMES):
            assert np.array_equal(np.asarray(it[j]), g["item." + name][n]), (n, name)
    for b in range(2):
        idx = g[f"batch{b}.idx"]
        assert np.array_equal(shards.order_batch(np.array(sorted(idx.tolist())), s.qlen, "reference"), idx)


MEDICAL_ARGS = dict(n_images=6, n_questions=17, K=51, D=16, n_answers=9, seed=5)


@pytest.mark.parametrize("variant", ["imageclef", "mimic"])
def test_medical_variants_match_reference_golden(tmp_path, variant):
    """ImageclefDataset / MimicDataset items (string image key last, every box used, ImageCLEF's dict answers and '.jpg' keys):
    the oracle restatement, the converter + shard reader and the in-batch order against the unmodified reference's outputs."""
    g = _golden()
    ds = SF.make_medical_dataset(variant=variant, **MEDICAL_ARGS)
    meta = shards.from_reference_records(ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["i_feat"], ds["bbox"], ds["sizes"], str(tmp_path),
                                         n_answers=ds["n_answers"], n_obj=None, variant=variant)
    assert meta["n_obj"] == 51 and meta["variant"] == variant
    s = shards.ShardSet(str(tmp_path))
    for n in range(len(s)):
        ref, got = SF.reference_item_medical(ds, n, variant=variant), s.dense_item(n)
        for j, name in enumerate(NAMES[:-1]):
            want = g[f"{variant}.item.{name}"][n]
            assert np.array_equal(np.asarray(ref[j]), want) and np.array_equal(np.asarray(got[j]), want), (n, name)
        assert ref[7] == got[7] == str(g[f"{variant}.item.iid"][n])
    for b in range(2):
        src = g[f"{variant}.batch{b}.src"]
        col = SF.reference_collate([SF.reference_item_medical(ds, int(n), variant=variant) for n in src])
        assert list(col[7]) == list(g[f"{variant}.batch{b}.iid"])                     # sorted by the key string, descending, stable
        assert np.array_equal(col[4].numpy(), g[f"{variant}.batch{b}.i"]) and np.array_equal(col[0].numpy(), g[f"{variant}.batch{b}.q"])
        keys = [s.last_element(int(n)) for n in src]
        order = shards.order_batch(np.asarray(src, dtype=np.int64), s.qlen, "reference", keys)
        assert [s.last_element(int(n)) for n in order] == list(g[f"{variant}.batch{b}.iid"])
        assert np.array_equal(np.stack([s.dense_item(int(n))[4] for n in order]), g[f"{variant}.batch{b}.i"])
        assert np.array_equal(np.stack([s.dense_item(int(n))[1] for n in order]), g[f"{variant}.batch{b}.a"])
        assert np.array_equal(np.stack([s.dense_item(int(n))[2] for n in order]), g[f"{variant}.batch{b}.n_votes"])


def test_medical_converter_rejects_ragged_box_counts(tmp_path):
    ds = SF.make_medical_dataset(variant="mimic", **MEDICAL_ARGS)
    k = next(iter(ds["bbox"]))
    ds["bbox"][k] = ds["bbox"][k][:40]
    ds["i_feat"][k] = ds["i_feat"][k][:40]
    with pytest.raises(ValueError, match="boxes"):
        shards.from_reference_records(ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["i_feat"], ds["bbox"], ds["sizes"], str(tmp_path),
                                      n_answers=ds["n_answers"], n_obj=None, variant="mimic")
    with pytest.raises(ValueError, match="variant"):
        shards.from_reference_records([], {}, {}, {}, {}, {}, str(tmp_path), n_answers=3, variant="clevr")


def test_convert_dataset_tool_takes_a_loaded_reference_dataset_object(ds, tmp_path):
    """tools/convert_dataset.py: the attributes VQA_Dataset.__init__ sets (torch_dataset.py:35-102) are all it reads."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("convert_dataset", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                "tools", "convert_dataset.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    obj = types.SimpleNamespace(vqa=ds["records"], q_wtoi=ds["q_wtoi"], a_wtoi=ds["a_wtoi"], i_feat=ds["i_feat"], bbox=ds["bbox"],
                                sizes=ds["sizes"], n_answers=ds["n_answers"], pretrained_wemb=np.ones((25, 8), np.float32))
    with pytest.raises(ValueError):                                    # the fixture has 6 boxes per image, VQA2 needs 36
        tool.convert(obj, str(tmp_path / "x"), "vqa2")
    meta = tool.convert(obj, str(tmp_path / "y"), "mimic")
    assert meta["n_obj"] == ds["K"] and meta["variant"] == "mimic"
    assert np.load(tmp_path / "y" / "pretrained_wemb.npy").shape == (25, 8)
    assert len(shards.ShardSet(str(tmp_path / "y"))) == len(ds["records"])


# ---------------------------------------------------------------------------------------------- properties (hypothesis)
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 400), bs=st.integers(1, 33), world=st.integers(1, 8), epoch=st.integers(0, 3), drop_last=st.booleans())
def test_epoch_batches_properties(n, bs, world, epoch, drop_last):
    per_rank = [shards.epoch_batches(n, bs, epoch=epoch, seed=9, rank=r, world=world, drop_last=drop_last) for r in range(world)]
    steps = len(per_rank[0])
    assert all(len(b) == steps for b in per_rank)                      # every rank takes the same number of steps ...
    for t in range(steps):
        sizes = {len(per_rank[r][t]) for r in range(world)}
        assert len(sizes) == 1 and 1 <= next(iter(sizes)) <= bs        # ... with the same batch size (collectives stay aligned)
    flat = np.concatenate([np.concatenate(b) for b in per_rank if b]) if steps else np.zeros(0, dtype=np.int64)
    assert len(np.unique(flat)) == len(flat) and (flat.size == 0 or (flat.min() >= 0 and flat.max() < n))
    full = (n // (bs * world)) * bs * world
    assert len(flat) >= full and (not drop_last or len(flat) == full) and len(flat) > n - max(bs * world, world) - world


@settings(max_examples=40, deadline=None)
@given(data=st.data())
def test_csr_rows_and_batch_order_properties(data, tmp_path_factory):
    nq = data.draw(st.integers(1, 30))
    A = data.draw(st.integers(2, 12))
    rows = [[(data.draw(st.integers(0, A - 1)), float(data.draw(st.integers(0, 10)))) for _ in range(data.draw(st.integers(0, 4)))]
            for _ in range(nq)]
    qlen = np.array([data.draw(st.integers(1, 14)) for _ in range(nq)], dtype=np.int32)
    path = str(tmp_path_factory.mktemp("prop"))
    shards.write_shards(path, features=np.zeros((1, 2, 8), np.float32), boxes=np.zeros((1, 2, 4), np.float32),
                        questions=np.zeros((nq, 4), np.int32), qlen=qlen, image_row=np.zeros(nq, np.int64), qid=np.arange(nq),
                        answers=rows, votes=rows[::-1], n_answers=A)
    s = shards.ShardSet(path)
    idx = np.array(data.draw(st.permutations(list(range(nq))))[:data.draw(st.integers(1, nq))], dtype=np.int64)
    ptr, ids, val = s.csr_rows("ans", idx)
    assert ptr[0] == 0 and ptr[-1] == len(ids) == len(val) == sum(len(rows[n]) for n in idx)
    for b, n in enumerate(idx):
        assert [(int(i), float(v)) for i, v in zip(ids[ptr[b]:ptr[b + 1]], val[ptr[b]:ptr[b + 1]])] == rows[n]
    vp, vi, _ = s.csr_rows("vote", idx)
    assert [int(x) for x in vi[vp[0]:vp[1]]] == [a for a, _ in rows[::-1][idx[0]]]
    byq = shards.order_batch(idx, s.qlen, "qlen")
    assert sorted(byq.tolist()) == sorted(idx.tolist()) and np.all(np.diff(qlen[byq]) <= 0)
    ref = shards.order_batch(idx, s.qlen, "reference")
    assert ref.tolist() == sorted(idx.tolist(), reverse=True)


def test_writer_streams_images_from_an_iterator(tmp_path):
    rng = np.random.RandomState(1)
    feats, boxes = rng.rand(5, 3, 8).astype(np.float32), rng.rand(5, 3, 4).astype(np.float32)
    common = dict(questions=np.zeros((2, 6), np.int32), qlen=[1, 2], image_row=[4, 0], qid=[7, 8], answers=[[], [(1, 1.0)]],
                  votes=[[], []], n_answers=3)
    shards.write_shards(str(tmp_path / "a"), features=feats, boxes=boxes, **common)
    shards.write_shards(str(tmp_path / "b"), image_iter=((feats[i], boxes[i]) for i in range(5)), **common)
    a, b = shards.ShardSet(str(tmp_path / "a")), shards.ShardSet(str(tmp_path / "b"))
    assert a.meta == b.meta and np.array_equal(a.features, b.features) and np.array_equal(a.boxes, b.boxes) and np.array_equal(a.features, feats)
    with pytest.raises(ValueError, match="either"):
        shards.write_shards(str(tmp_path / "c"), features=feats, boxes=boxes, image_iter=iter([]), **common)
    with pytest.raises(ValueError, match="no images"):
        shards.write_shards(str(tmp_path / "d"), image_iter=iter([]), **common)
    with pytest.raises(ValueError, match="image row 1"):
        shards.write_shards(str(tmp_path / "e"), image_iter=iter([(feats[0], boxes[0]), (feats[1][:2], boxes[1][:2])]), **common)
