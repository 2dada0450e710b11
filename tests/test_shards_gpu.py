"""Device-side batch assembly (csrc/loader.cu through vqa_b200.shards.ShardLoader) against the reference's dataset + collate code
restated in shard_fixture.py.  Index/byte work: everything is compared bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import shard_fixture as SF  # noqa: E402
from vqa_b200 import shards  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _convert(ds, path, dtype="f32", q_width=100):
    return shards.from_reference_records(ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["i_feat"], ds["bbox"], ds["sizes"], str(path),
                                         n_answers=ds["n_answers"], n_obj=ds["K"], q_width=q_width, feature_dtype=dtype)


def _check_batch(ds, batch, bf16):
    q, a, n_votes, qid, img, k, qlen, idx = batch
    ref = SF.reference_collate([SF.reference_item(ds, int(n)) for n in idx])          # sorts by idx descending ...
    inv = {int(n): j for j, n in enumerate(ref[7])}
    perm = torch.tensor([inv[int(n)] for n in idx])                                   # ... so pick its rows in the loader's order
    rq, ra, rv, rqid, ri, rk, rqlen, ridx = (t[perm] for t in ref)
    assert q.is_cuda and a.is_cuda and n_votes.is_cuda and img.is_cuda and k.is_cuda
    assert (q.dtype, a.dtype, n_votes.dtype, img.dtype, k.dtype) == (rq.dtype, ra.dtype, rv.dtype, ri.dtype, rk.dtype)
    assert (qid.dtype, qlen.dtype, idx.dtype) == (rqid.dtype, rqlen.dtype, ridx.dtype) and not qlen.is_cuda
    for got, want in ((q, rq), (a, ra), (n_votes, rv), (qid, rqid), (k, rk), (qlen, rqlen), (idx, ridx)):
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
    D = ds["D"]
    if bf16:
        ri = torch.cat((ri[..., :D].to(torch.bfloat16).float(), ri[..., D:]), dim=-1)
    assert img.shape == ri.shape and torch.equal(img.cpu(), ri)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("resident", [True, False])
@pytest.mark.parametrize("order", ["qlen", "reference"])
def test_loader_batches_equal_the_reference_collate(tmp_path, dtype, resident, order):
    ds = SF.make_dataset(n_images=9, n_questions=41, K=6, D=16)
    _convert(ds, tmp_path, dtype)
    ld = shards.ShardLoader(str(tmp_path), 8, DEV, seed=11, resident=resident, order=order)
    seen = []
    for ep in range(2):
        n = 0
        for batch in ld:
            _check_batch(ds, batch, dtype == "bf16")
            if order == "reference":
                assert torch.equal(batch[7], torch.sort(batch[7], descending=True).values)
            else:
                assert torch.all(batch[6][:-1] >= batch[6][1:])
            seen.append(batch[7])
            n += 1
        assert n == len(ld) == 6 and ld.epoch == ep + 1
    first, second = torch.cat(seen[:6]), torch.cat(seen[6:])
    assert sorted(first.tolist()) == list(range(41)) == sorted(second.tolist()) and not torch.equal(first, second)


def test_two_ranks_see_disjoint_questions(tmp_path):
    ds = SF.make_dataset(n_images=5, n_questions=37)
    _convert(ds, tmp_path)
    per_rank = []
    for rank in range(2):
        ld = shards.ShardLoader(str(tmp_path), 4, DEV, seed=2, rank=rank, world=2)
        per_rank.append([b[7] for b in ld])
    assert len(per_rank[0]) == len(per_rank[1])
    assert all(len(a) == len(b) for a, b in zip(*per_rank))
    everything = torch.cat(per_rank[0] + per_rank[1])
    assert len(set(everything.tolist())) == len(everything) >= 36


def test_large_rows_and_vqa2_width(tmp_path):
    """K=36 x D=2048 rows (the VQA2 layout) through both storage types, several questions per image."""
    rng = np.random.RandomState(0)
    n_img, K, D, nq, A = 5, 36, 2048, 40, 3000
    feats = np.maximum(rng.randn(n_img, K, D), 0).astype(np.float32)
    boxes = rng.rand(n_img, K, 4).astype(np.float32)
    rows = rng.randint(0, n_img, nq)
    answers = [[(int(rng.randint(0, A)), float(np.float32(rng.rand()))) for _ in range(rng.randint(0, 4))] for _ in range(nq)]
    for dtype in ("f32", "bf16"):
        p = tmp_path / dtype
        shards.write_shards(str(p), features=feats, boxes=boxes, questions=rng.randint(0, 50, (nq, 100)), qlen=rng.randint(1, 15, nq),
                            image_row=rows, qid=np.arange(nq), answers=answers, votes=answers, n_answers=A, feature_dtype=dtype)
        ld = shards.ShardLoader(str(p), 16, DEV, shuffle=False, order="none")
        for batch in ld:
            idx = batch[7].numpy()
            want = torch.from_numpy(feats[rows[idx]])
            if dtype == "bf16":
                want = want.to(torch.bfloat16).float()
            assert torch.equal(batch[4][..., :D].cpu(), want)
            assert torch.equal(batch[4][..., D:].cpu(), torch.from_numpy(boxes[rows[idx]]))
            dense = np.zeros((len(idx), A), np.float32)
            for b, n in enumerate(idx):
                for a, c in answers[n]:
                    dense[b, a] = c
            assert torch.equal(batch[1].cpu(), torch.from_numpy(dense)) and torch.equal(batch[2].cpu(), torch.from_numpy(dense))


def test_out_of_range_indices_are_flagged_not_read():
    from vqa_b200 import kernels as kn
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    feats = torch.ones(3, 2, 8, device=DEV)
    boxes = torch.ones(3, 2, 4, device=DEV)
    out = kn.gather_image(feats, boxes, torch.tensor([0, 3, -1, 2], device=DEV), err)
    assert err.item() == 1 and out[1].abs().sum().item() == 0 and out[2].abs().sum().item() == 0 and out[0].min().item() == 1
    err.zero_()
    ptr = torch.tensor([0, 2, 3], device=DEV)
    t = kn.scatter_targets(ptr, torch.tensor([1, 7, 0], dtype=torch.int32, device=DEV), torch.tensor([0.5, 0.25, 1.0], device=DEV), 2, 4, err)
    assert err.item() == 2 and t.tolist() == [[0, 0.5, 0, 0], [1.0, 0, 0, 0]]
    with pytest.raises(RuntimeError):
        kn.gather_image(feats.double(), boxes, torch.tensor([0], device=DEV), err)
    with pytest.raises(RuntimeError):
        kn.gather_image(feats.cpu(), boxes.cpu(), torch.tensor([0]), err)


def test_loader_batch_drives_the_model_like_the_reference_loop(tmp_path):
    """run.py:425-431 with the loader in place of DataLoader: batch_to_cuda passes the device tensors through, Model.forward takes
    them, and the logits equal those of the same questions fed from a host-collated reference batch."""
    import sparse_graph_model as M
    import utils as U
    ds = SF.make_dataset(n_images=6, n_questions=16, K=12, D=16, n_answers=24)
    _convert(ds, tmp_path, q_width=10)
    torch.manual_seed(0)
    model = M.Model(vocab_size=40, emb_dim=8, feat_dim=20, hid_dim=16, out_dim=24, pretrained_wemb=np.zeros((40, 8), np.float32) + 0.1,
                    dropout=0.0, n_kernels=4, neighbourhood_size=5, n_obj=12).to(DEV).eval()
    ld = shards.ShardLoader(str(tmp_path), 8, DEV, seed=1)
    for batch in ld:
        q, a, n_votes, i, k, qlen = U.batch_to_cuda(batch)
        assert i.data_ptr() == batch[4].data_ptr()                     # no copy of what is already on the device
        logits, _, _ = model(q, i, k, qlen)
        ref = SF.reference_collate([SF.reference_item(ds, int(n), q_width=10) for n in batch[7]])
        rq, ra, rv, ri, rk, rqlen = U.batch_to_cuda(ref)
        ref_logits, _, _ = model(rq, ri, rk, rqlen)
        inv = {int(n): j for j, n in enumerate(ref[7])}
        perm = torch.tensor([inv[int(n)] for n in batch[7]], device=DEV)
        assert torch.allclose(logits, ref_logits[perm], rtol=1e-5, atol=1e-6)
        assert U.total_vqa_score(logits, n_votes) == pytest.approx(U.total_vqa_score(ref_logits, rv), abs=1e-5)


def test_loader_reproduces_the_reference_golden_batches(tmp_path):
    """Device-assembled batches against tests/golden/dataset_small.npz - the outputs of the unmodified torch_dataset.py
    (__getitem__ + collate_fn) on the same miniature dataset (tests/golden/make_dataset_golden.py)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_small.npz"))
    ds = SF.make_dataset(n_images=5, n_questions=19, K=36, D=16, n_answers=11, seed=3)
    _convert(ds, tmp_path)
    ld = shards.ShardLoader(str(tmp_path), 5, DEV, order="reference")
    names = ("q", "a", "n_votes", "qid", "i", "k", "qlen", "idx")
    for b in range(2):
        idx = shards.order_batch(np.sort(z[f"batch{b}.idx"]), ld.set.qlen, "reference")
        batch = ld.assemble(idx)
        for j, name in enumerate(names):
            want = torch.from_numpy(z[f"batch{b}.{name}"])
            assert batch[j].dtype == want.dtype and batch[j].shape == want.shape and torch.equal(batch[j].cpu(), want), (b, name)
    ld.check_errors()


@pytest.mark.parametrize("variant", ["imageclef", "mimic"])
def test_medical_loader_reproduces_the_reference_golden_batches(tmp_path, variant):
    """ImageclefDataset / MimicDataset batches (K = 51 boxes, image key as the last element, collate order by that key) against the
    unmodified reference's outputs in tests/golden/dataset_small.npz."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_small.npz"))
    ds = SF.make_medical_dataset(variant=variant, n_images=6, n_questions=17, K=51, D=16, n_answers=9, seed=5)
    shards.from_reference_records(ds["records"], ds["q_wtoi"], ds["a_wtoi"], ds["i_feat"], ds["bbox"], ds["sizes"], str(tmp_path),
                                  n_answers=ds["n_answers"], n_obj=None, variant=variant)
    ld = shards.ShardLoader(str(tmp_path), 6, DEV, order="reference")
    names = ("q", "a", "n_votes", "qid", "i", "k", "qlen")
    for b in range(2):
        src = z[f"{variant}.batch{b}.src"].astype(np.int64)
        batch = ld.assemble(ld.ordered(src))
        for j, name in enumerate(names):
            want = torch.from_numpy(z[f"{variant}.batch{b}.{name}"])
            assert batch[j].dtype == want.dtype and batch[j].shape == want.shape and torch.equal(batch[j].cpu(), want), (b, name)
        assert list(batch[7]) == [str(x) for x in z[f"{variant}.batch{b}.iid"]]
    ld.check_errors()


def test_loader_gathers_the_image_batch_into_the_training_steps_input_slot(tmp_path):
    """engine.TrainStep.input_slot + ShardLoader.assemble(image_out=...): from the second step on the image batch is gathered
    straight into the captured step's idle input buffer (no second 151 MB copy at the VQA2 shapes) and training is unchanged."""
    import sparse_graph_model as M
    from vqa_b200.ddp import GradReducer
    from vqa_b200.engine import TrainStep
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    from vqa_b200.optim import FlatAdam
    ds = SF.make_dataset(n_images=7, n_questions=40, K=12, D=16, n_answers=24)
    _convert(ds, tmp_path, q_width=10)
    runs = {}
    for direct in (False, True):
        torch.manual_seed(0)
        model = M.Model(vocab_size=40, emb_dim=8, feat_dim=20, hid_dim=32, out_dim=24, pretrained_wemb=np.zeros((40, 8), np.float32) + 0.1,
                        dropout=0.0, n_kernels=4, neighbourhood_size=5, n_obj=12).to(DEV).train()
        model.max_question_len = 10
        red = GradReducer(model.parameters())
        step = TrainStep(model, FlatAdam(red, lr=1e-3), MultiLabelSoftMarginLoss(), reducer=red, use_graph=True, seed=3)
        ld = shards.ShardLoader(str(tmp_path), 8, DEV, shuffle=False, order="none")
        losses, in_place = [], 0
        for ep in range(2):
            for idx in ld.batches():
                with torch.cuda.stream(step.copy_stream):
                    slot = step.input_slot("image") if direct else None
                    q, a, _nv, _qid, image, k, qlen, _ = ld.assemble(idx, image_out=slot)
                in_place += int(slot is not None and image.data_ptr() == slot.data_ptr())
                losses.append(step(q, image, k, qlen, a).item())
        ld.check_errors()
        step.close()
        red.remove()
        runs[direct] = (losses, in_place)
    assert runs[False][1] == 0 and runs[True][1] == len(runs[True][0]) - 1        # every step but the one that builds the graphs
    assert runs[True][0] == pytest.approx(runs[False][0], rel=1e-6)
