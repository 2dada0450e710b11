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


def _load_reference_utils():
    here = os.path.dirname(os.path.abspath(__file__))
    for d in sys.path:
        cand = os.path.join(d or ".", "utils.py")
        if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != here:
            spec = importlib.util.spec_from_file_location("_reference_utils", cand)
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except Exception:                 # not the reference's utils (or its imports are missing): keep looking
                continue
            if hasattr(mod, "total_vqa_score"):
                return mod
    return None


_ref = _load_reference_utils()
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("_")})


def total_vqa_score(logits, n_votes_batch):
    """Total VQA score of a batch as assessed by the challenge: sum_i min(n_votes[i, argmax_j logits[i, j]] / 3, 1)."""
    oix = logits.detach().argmax(dim=1, keepdim=True)
    votes = n_votes_batch.detach().gather(1, oix).squeeze(1).double()
    return float((votes / 3.0).clamp_(max=1.0).sum())


def batch_to_cuda(batch, volatile=False):
    """Moves a dataset batch (torch_dataset.py:164 tuple) onto the GPU: pinned staging, asynchronous copies."""
    def dev(t):
        if not t.is_cuda and not t.is_pinned():
            t = t.pin_memory()
        return t.cuda(non_blocking=True)
    q, a, n_votes, i, k = (dev(batch[j]) for j in (0, 1, 2, 4, 5))
    qlen = list(batch[6])
    return q, a, n_votes, i, k, qlen
