"""Harness of the driver drop-in test: runs the reference's UNMODIFIED ``run.trainval`` (`run.py:343-471`) in this process with

    --impl b200       sys.path = [tests/stubs, vqa-project_b200, <reference>]   (our modules shadow sparse_graph_model / layers / utils)
    --impl reference  sys.path = [tests/stubs, <reference>]                     (the reference's own modules, eager PyTorch CUDA)

and prints one JSON line: every loss value the driver's criterion produced, the VQA scores it accumulated, where the modules
came from, how many kernels of libvqa_sm100.so were launched, and the checkpoint file the driver wrote."""
import argparse
import glob
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["b200", "reference"])
    ap.add_argument("--ref", required=True)
    ap.add_argument("--save_dir", required=True)
    ap.add_argument("--dropout", default="0.0")
    a = ap.parse_args()
    warnings.filterwarnings("ignore")
    import torch
    torch.cuda.init()                       # run.py:33 pins CUDA_VISIBLE_DEVICES='1' at import; the context exists by then ...
    torch.cuda.device_count()
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    pkg = os.path.join(ROOT, "vqa-project_b200")
    sys.path[:0] = [os.path.join(HERE, "stubs")] + ([pkg] if a.impl == "b200" else []) + [a.ref]
    if a.impl == "reference":
        sys.path.append(pkg)                # only so that the stub dataset can import vqa_b200.synthetic (no module name clashes: it is last)
    losses, scores = [], []
    crit_forward = torch.nn.MultiLabelSoftMarginLoss.forward

    def recording_forward(self, x, y):      # observes the driver's criterion; does not change what it computes
        out = crit_forward(self, x, y)
        losses.append(out)
        return out
    torch.nn.MultiLabelSoftMarginLoss.forward = recording_forward
    sys.argv = ["run.py", "--bsize", "8", "--ep", "1", "--hid", "512", "--emb", "32", "--n_obj", "36", "--neighbourhood_size", "16",
                "--n_kernels", "4", "--dropout", a.dropout, "--save_dir", a.save_dir, "--log_interval", "1", "--model_path", "/nonexistent"]
    import run                              # the reference driver, unmodified
    if visible is None:                     # ... and the caller's device selection is put back (a one-GPU box has no device 1:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)   # torch.cuda.device_count() re-reads the variable and would report 0 devices)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = visible
    import sparse_graph_model
    import layers
    import utils
    score_fn = run.total_vqa_score

    def recording_score(logits, votes):
        s = score_fn(logits, votes)
        scores.append(float(s))
        return s
    run.total_vqa_score = recording_score
    args, _, unparsed = run.input_args()
    assert not unparsed, unparsed
    run.trainval(args)
    torch.cuda.synchronize()
    launches = 0
    if a.impl == "b200":
        from vqa_b200 import kernels
        launches = kernels.LAUNCHES
    ckpt = glob.glob(os.path.join(a.save_dir, "vqa_36_4_16_*.pt"))
    print("DROPIN " + json.dumps({"impl": a.impl, "losses": [float(l) for l in losses], "scores": scores, "launches": launches, "checkpoint": ckpt,
                                  "modules": {m.__name__: os.path.dirname(os.path.abspath(m.__file__)) for m in (run, sparse_graph_model, layers, utils)}}))


if __name__ == "__main__":
    main()
