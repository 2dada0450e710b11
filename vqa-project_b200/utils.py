"""Drop-in for the two host utilities the drivers call once per training step.

Shadows the reference's ``utils`` module the same way ``sparse_graph_model`` / ``layers`` are shadowed: everything the
reference module defines (``save``, ``xyxy2xywh`` ...) is re-exported from the next ``utils.py`` on ``sys.path`` (the
reference checkout), only the two per-step functions are replaced:

* ``total_vqa_score(logits, n_votes_batch)`` - reference ``utils.py:47-55`` loops over the batch in Python and calls
  ``.item()`` once per sample (B host syncs per step: at ~4 ms per B=512 step that loop alone costs more than the step).
  Here: one gather / clamp / sum on the device in fp64 and ONE read-back.  Same value (sum of min(votes/3, 1)), same
  Python-float return type.
* ``batch_to_cuda(batch)`` - reference ``utils.py:22-31`` issues five synchronous pageable ``.cuda()`` copies.  Here the
  tensors are pinned (if they are not already) and copied with ``non_blocking=True``; same return tuple.
"""
import importlib.util
import os
import sys

import torch


def _find_reference_utils():
    """Path of the reference's own ``utils.py``: ``$VQA_REFERENCE_DIR``, else the first LATER ``sys.path`` entry whose ``utils.py``
    defines both per-step functions (decided by reading the file - nothing is executed to find out), else the repo's
    reference install ``baseline/_ref``."""
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ["VQA_REFERENCE_DIR"]] if os.environ.get("VQA_REFERENCE_DIR") else []
    cands += [d or "." for d in sys.path] + [os.path.join(os.path.dirname(here), "baseline", "_ref")]
    for d in cands:
        cand = os.path.join(d, "utils.py")
        if os.path.isfile(cand) and os.path.abspath(d) != here:
            try:
                text = open(cand, encoding="utf-8", errors="replace").read()
            except OSError:
                continue
            if "def total_vqa_score(" in text and "def batch_to_cuda(" in text:
                return cand
    return None


def _load_reference_utils():
    cand = _find_reference_utils()
    if cand is None:
        raise ImportError("vqa-project_b200/utils.py shadows the reference's `utils` module and re-exports everything else it defines "
                          "(save, xyxy2xywh, ...): put the reference checkout on sys.path behind vqa-project_b200, or set VQA_REFERENCE_DIR")
    spec = importlib.util.spec_from_file_location("_reference_utils", cand)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)              # errors propagate: a half-imported reference module must not go unnoticed
    return mod


_ref = _load_reference_utils()
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("_")})


def total_vqa_score(logits, n_votes_batch):
    """Total VQA score of a batch as assessed by the challenge: sum_i min(n_votes[i, argmax_j logits[i, j]] / 3, 1)."""
    oix = logits.detach().argmax(dim=1, keepdim=True)
    votes = n_votes_batch.detach().gather(1, oix).squeeze(1).double()
    score = float((votes / 3.0).clamp_(max=1.0).sum())        # the one read-back of the step: device-side error flags ride along
    if logits.is_cuda:
        from vqa_b200 import kernels
        kernels.check_device_errors(logits.device)
    return score


def batch_to_cuda(batch, volatile=False):
    """Moves a dataset batch (torch_dataset.py:164 tuple) onto the GPU: pinned staging, asynchronous copies."""
    def dev(t):
        if not t.is_cuda and not t.is_pinned():
            t = t.pin_memory()
        return t.cuda(non_blocking=True)
    q, a, n_votes, i, k = (dev(batch[j]) for j in (0, 1, 2, 4, 5))
    qlen = list(batch[6])
    return q, a, n_votes, i, k, qlen
