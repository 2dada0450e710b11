"""Drop-in host utilities (vqa-project_b200/utils.py) against restatements of the reference's loops (utils.py:22-31, 47-55)."""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("vqa_dropin_utils", os.path.join(ROOT, "vqa-project_b200", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _reference_loop(logits, n_votes):
    score = 0
    _, oix = logits.data.max(1)
    for i, pred in enumerate(oix):            # the reference's per-sample loop with one .item() each
        score += min(n_votes[i, pred].item() / 3, 1)
    return score


def test_total_vqa_score_equals_the_reference_loop():
    u = _load()
    g = torch.Generator().manual_seed(3)
    for B, A in [(1, 5), (64, 3000), (513, 77)]:
        logits = torch.randn(B, A, generator=g)
        votes = torch.randint(0, 11, (B, A), generator=g).float()
        got = u.total_vqa_score(logits, votes)
        assert isinstance(got, float)
        assert abs(got - _reference_loop(logits, votes)) < 1e-9 * max(1, B)


def test_shadow_module_reexports_the_reference_utils_when_it_is_on_the_path(tmp_path, monkeypatch):
    (tmp_path / "utils.py").write_text("def total_vqa_score(a, b):\n    return -1\n\ndef batch_to_cuda(batch):\n    return None\n\n"
                                       "def save(args, model, path, name):\n    return 'ref-save'\n")
    other = tmp_path / "other"
    other.mkdir()
    (other / "utils.py").write_text("raise SystemExit('an unrelated utils.py must never be executed')\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.syspath_prepend(str(other))                                 # found first, skipped by its text, not run
    u = _load()
    assert u.save(None, None, None, None) == "ref-save"                     # re-exported
    assert u.total_vqa_score(torch.zeros(2, 3), torch.ones(2, 3) * 3) == 2.0   # replaced


def test_shadow_module_fails_loudly_without_a_reference_utils(tmp_path, monkeypatch):
    import pytest
    spec = importlib.util.spec_from_file_location("vqa_dropin_utils_probe", os.path.join(ROOT, "vqa-project_b200", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    monkeypatch.setattr(sys, "path", [str(tmp_path)])
    monkeypatch.setenv("VQA_REFERENCE_DIR", str(tmp_path))
    monkeypatch.setattr(os.path, "isfile", lambda p: False)
    with pytest.raises(ImportError):
        spec.loader.exec_module(mod)
