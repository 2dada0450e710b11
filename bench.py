#!/usr/bin/env python
"""Train-step throughput of the conditioned-graph VQA hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload vqa2_b512]

One "step" = zero_grad -> Model.forward -> MultiLabelSoftMarginLoss -> backward (-> bucketed NCCL all-reduce when
N > 1) -> Adam step, on one synthetic batch of BASELINE.json configs[1] (VQA2: B=512/GPU, K=36, F=2052, nb=16,
nk=8, 3000 answers, dropout 0.5).  Prints ONE JSON line (see the task contract): `value` = questions/s with
inputs resident in HBM, `e2e` = the same through the public Model API with pinned-host inputs copied every step,
`roofline` = the layer-1 graph-convolution kernel against the measured HBM peak, `cpu_baseline` = the oracle port
on the host cores.  `--impl reference` times the reference algorithm (oracle port, PyTorch CPU) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vqa-project_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="vqa2_b512")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_strict", "bf16"],
                    help="fp32: split-bf16 3-pass tcgen05 GEMMs (fp32-grade, the parity mode); bf16: 1-pass bf16 tensor-core GEMMs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bucket-mb", type=int, default=32, help="gradient all-reduce bucket size (N > 1)")
    ap.add_argument("--tail", default="fused", choices=["fused", "torch"],
                    help="criterion + optimiser: our single-launch kernels (vqa_b200.loss / vqa_b200.optim) or torch's modules")
    ap.add_argument("--no-resident-table", action="store_true", help="skip the ShardLoader (feature table in HBM) end-to-end leg")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--gemm-table", default="", help="write one line per dense product of a step (M, N, K, passes, event-bracketed us, TFLOP/s, fraction of the bf16 peak) to this file (diagnosis)")
    ap.add_argument("--kernel-totals", default="", help="also run 10 steps under torch.profiler on rank 0 and write per-kernel device time totals to this file (diagnosis)")
    ap.add_argument("--quick", action="store_true", help="device-resident and end-to-end legs only (scaling experiments)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="questions per CPU step (bounded sample of the workload)")
    return ap.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed `ncu --set full` captures
# (profiles/), keyed by workload: a capture of one workload says nothing about another, so anything else reports null
NCU_TRAFFIC = {("vqa2_b512", "vqa_graphconv_mma_fwd"): 267.9e6}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


def load_reference_model_module():
    """The UNMODIFIED reference (`sparse_graph_model.py` + `layers.py`) from the git-ignored install baseline/_ref (written by
    __graft_entry__.build() in the build container, shipped with the snapshot), imported without disturbing the drop-in modules
    of the same names.  None when the install is absent."""
    d = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(d, "sparse_graph_model.py")):
        return None
    import importlib
    import warnings
    names = ("layers", "sparse_graph_model")
    saved = {n: sys.modules.pop(n, None) for n in names}
    sys.path.insert(0, d)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            importlib.import_module("layers")
            ref = importlib.import_module("sparse_graph_model")
        assert os.path.dirname(os.path.abspath(ref.__file__)) == d
    finally:
        sys.path.remove(d)
        for n in names:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    return ref


def reference_train_steps(workload, sample, steps, warmup, device, threads=None, tf32=False):
    """The reference's own train step - its Model, nn.MultiLabelSoftMarginLoss, torch.optim.Adam, the loop body of run.py:425-460 -
    through its stock PyTorch code path on `device` ("cpu": the host cores; "cuda": eager PyTorch CUDA, the same-box competitor
    of SURVEY.md 8d).  Returns (seconds per step list, last loss, peak device memory in GB)."""
    import warnings
    from vqa_b200.synthetic import make_batch, make_wemb
    ref = load_reference_model_module()
    if ref is None:
        return None
    warnings.filterwarnings("ignore")
    if threads:
        torch.set_num_threads(threads)
    w = workload
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        torch.manual_seed(1000)
        model = ref.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(device).train()
        crit = torch.nn.MultiLabelSoftMarginLoss().to(device)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        batches = []
        for i in range(2):
            b = make_batch(w, seed=1000 + i, batch=sample)
            batches.append((b["question"].to(device), b["image"].to(device), b["K"].to(device), b["qlen"], b["target"].to(device)))
        cuda = torch.device(device).type == "cuda"
        if cuda:
            torch.cuda.reset_peak_memory_stats()
        times = []
        for it in range(warmup + steps):
            q, img, K, qlen, tgt = batches[it % 2]
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, _, _ = model(q, img, K, qlen)
            loss = crit(out, tgt)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if cuda:
                torch.cuda.synchronize()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        peak = torch.cuda.max_memory_allocated() / 2 ** 30 if cuda else None
        last = float(loss.detach())
        del model, opt, batches, loss, out
        if cuda:
            torch.cuda.empty_cache()
        return times, last, peak
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_train_steps(workload, sample, steps, warmup, threads):
    """-> (times, last loss, kind).  kind "reference": the unmodified reference on the host cores (baseline/_ref);
    kind "port": the oracle restatement of its algorithm (oracle/vqa_oracle.py, PyTorch CPU autograd) when no install is present."""
    r = reference_train_steps(workload, sample, steps, warmup, "cpu", threads=threads)
    if r is not None:
        return r[0], r[1], "reference"
    from oracle import vqa_oracle as O
    from vqa_b200.synthetic import make_batch
    torch.set_num_threads(threads)
    w = workload
    params = O.init_params(w.vocab, w.emb_dim, w.feat_dim, w.hid_dim, w.out_dim, w.n_kernels)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(leaves.values()), lr=1e-4)
    batch = make_batch(w, seed=1000, batch=sample)
    q, img, tgt = batch["question"], batch["image"], batch["target"]
    qlen = [int(x) for x in batch["qlen"]]
    gen = torch.Generator().manual_seed(0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        logits, _, _ = O.forward(leaves, q, img, qlen, w.neighbourhood, w.n_kernels, dropout_p=w.dropout,
                                 training=True, generator=gen)
        loss = O.multilabel_soft_margin_loss(logits, tgt)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times, float(loss.detach()), "port"


def run_reference(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times, _, kind = cpu_train_steps(workload, args.cpu_sample, args.steps, args.warmup, cores)
    total = sum(times)
    value = args.cpu_sample * len(times) / total
    how = ("the unmodified reference (baseline/_ref: sparse_graph_model.Model + nn.MultiLabelSoftMarginLoss + Adam, run.py:425-460 loop body), PyTorch CPU"
           if kind == "reference" else "PyTorch CPU autograd through oracle/vqa_oracle.py (no reference install present)")
    sample = f"{args.cpu_sample} of {workload.batch} questions per step, {len(times)} steps, {how}"
    line = {
        "impl": "reference", "metric": "train questions/sec (fwd+bwd)", "value": round(value, 3), "unit": "questions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(workload, args, 1) | {"cpu_sample_batch": args.cpu_sample},
        "cpu_baseline": {"value": round(value, 3), "unit": "questions/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(w, args, world):
    family = "medical-VQA (ImageCLEF/MIMIC feature shapes)" if w.name.startswith("med") else "VQA2"
    return {"workload": f"{w.name}: {family} conditioned-graph train step, per-GPU batch {w.batch}, K={w.n_obj} boxes x {w.feat_dim}-d, "
                        f"<= {w.max_qlen}-token questions, top-k={w.neighbourhood}, {w.n_kernels} Gaussian kernels, {w.out_dim} answers, dropout {w.dropout}",
            "global_batch": w.batch * world, "step": "zero_grad+forward+MultiLabelSoftMarginLoss+backward+allreduce+Adam (" + ("criterion and Adam as single kernels of libvqa_sm100.so" if args.tail == "fused" else "torch criterion and fused Adam") + "), " + ("eager launches" if args.no_graph else "one CUDA-graph replay per step (vqa_b200.engine.TrainStep)"),
            "parallelism": f"dp{world}", "gru": "padded masked recurrence on split-bf16 x3 tcgen05 GEMMs (fp32-grade)", "gemm_precision": {"fp32": "split-bf16 x3 passes (fp32-grade, rel err ~1e-5)", "fp32_strict": "split-bf16 x3 passes + chunk-promoted 3xTF32 graph-learner forward",
                                   "bf16": "bf16 x1 pass (graph-learner forward x3 passes)"}[args.precision],
            "l2_policy": f"inputs larger than L2 (image batch {w.batch * w.n_obj * w.feat_dim * 4 / 1e6:.0f} MB > 126 MB), 3 rotating batches"
                         if w.batch * w.n_obj * w.feat_dim * 4 > 126e6 else
                         f"3 rotating batches ({3 * w.batch * w.n_obj * w.feat_dim * 4 / 1e6:.0f} MB) + the step's own ~1 GB of activations and 0.8 GB of optimizer traffic between two uses of any input"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region through NVML (the library nvidia-smi itself
    uses; no subprocess, so sampling does not perturb the launching thread)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:   # keep the bench alive; the JSON line records that sampling failed
            self.err = repr(e)

    def summary(self):
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
               "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------------ B200 arm
def _dbg(msg):
    if os.environ.get("VQA_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_b200(args, workload):
    import torch.distributed as dist
    from vqa_b200 import kernels as kn, ops
    from vqa_b200.ddp import GradReducer, broadcast_parameters
    from vqa_b200.engine import TrainStep
    from vqa_b200.synthetic import make_batch, make_wemb
    import sparse_graph_model as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 on their own (NCCL prints its version banner
    # there on the first communicator) are pointed at stderr, the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the vqa_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "unbound"
    try:        # run this rank (and first-touch its pinned batches) on the CPUs next to its GPU: 8 ranks x 158 MB of H2D per step
        import pynvml                                      # otherwise all cross one socket link
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        numa = f"nvml cpu affinity of gpu {phys}: {len(os.sched_getaffinity(0))} cpus"
    except Exception as e:                                 # pragma: no cover - affinity is an optimisation, never a requirement
        numa = f"unbound ({type(e).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(args.precision)
    w = workload
    _dbg("process group up")

    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev)
    model.max_question_len = w.max_qlen
    broadcast_parameters(model)
    _dbg("parameters broadcast")
    model.train()
    reducer = GradReducer(model.parameters(), bucket_bytes=args.bucket_mb << 20)
    if args.tail == "fused":             # criterion and optimiser of run.py:382,392 as our own kernels (csrc/train_step.cu)
        from vqa_b200.loss import MultiLabelSoftMarginLoss
        from vqa_b200.optim import FlatAdam
        criterion = MultiLabelSoftMarginLoss()
        opt = FlatAdam(reducer, lr=1e-4)
    else:                                # torch's modules on the same flat gradient buffer (A/B)
        criterion = torch.nn.MultiLabelSoftMarginLoss()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=not args.no_graph)
    torch.manual_seed(1234 + rank)       # rank-offset dropout streams
    step = TrainStep(model, opt, criterion, reducer=reducer, use_graph=not args.no_graph, seed=1234 + rank)

    # host batches (pinned) and device-resident copies; rank-offset seeds = different data per rank
    NB = 3
    keys = ("question", "image", "K", "qlen", "target")
    host = []
    for i in range(NB):
        b = make_batch(w, seed=1000 + 17 * rank + i)
        b["qlen"] = torch.tensor([int(x) for x in b["qlen"]], dtype=torch.int32)
        host.append({k: b[k].pin_memory() for k in keys})
    resident = [{k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def run(b):
        return step(b["question"], b["image"], b["K"], b["qlen"], b["target"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ---- device-resident throughput ------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        run(resident[i % NB])
        if i == 0:
            torch.cuda.synchronize()
            _dbg("first step (capture) done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(resident[i % NB])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    _dbg(f"resident loop done: {ms_total / args.steps:.3f} ms/step")
    launches = step.launches_per_step
    sampler.stop_flag = True
    ms_step = ms_total / args.steps
    value = w.batch * world / (ms_step * 1e-3)
    final_loss = float(loss.detach())

    if args.kernel_totals:                # diagnosis: where does the device time of a step go (kernels inside the graph replays included)
        from torch.profiler import profile, ProfilerActivity
        barrier()
        if rank == 0:
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                for i in range(10):
                    run(resident[i % NB])
                torch.cuda.synchronize()
            rows = [(e.key, e.count, getattr(e, "device_time_total", getattr(e, "cuda_time_total", 0))) for e in prof.key_averages()]
            rows = sorted([r for r in rows if r[2] > 0], key=lambda r: -r[2])
            with open(args.kernel_totals, "w") as f:
                json.dump({"n_gpus": world, "steps": 10, "ms_per_step": ms_step, "kernels": [{"name": k[:120], "count": c, "us_per_step": round(t / 10, 2)} for k, c, t in rows]}, f, indent=1)
        else:
            for i in range(10):
                run(resident[i % NB])
            torch.cuda.synchronize()
        barrier()

    # ---- end to end through the public API: pinned host -> device every step, loss read back ---------------
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        """Every step: H2D of that step's batch from pinned memory (overlapping the previous step), the step, and a D2H read of
        its loss.  The read is software-pipelined by one step - the copy of loss i is enqueued right after step i, the host
        picks it up while step i+1 runs - so the host never idles the GPU between replays."""
        last = None
        for i in range(n):
            hb = host[i % NB]
            if i == 0:
                step.prefetch(hb["question"], hb["image"], hb["K"], hb["qlen"], hb["target"])
            loss = run(hb)                                  # waits for this batch's H2D, replays the step
            loss_host[i & 1].copy_(loss.detach().reshape(()), non_blocking=True)   # D2H read of the step's result, every step
            loss_ev[i & 1].record()
            if i + 1 < n:
                nb_ = host[(i + 1) % NB]                    # next step's H2D overlaps this step's compute
                step.prefetch(nb_["question"], nb_["image"], nb_["K"], nb_["qlen"], nb_["target"])
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                last = float(loss_host[(i - 1) & 1])
        loss_ev[(n - 1) & 1].synchronize()
        return float(loss_host[(n - 1) & 1])

    e2e_loop(max(3, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last_loss = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    _dbg(f"e2e loop done: {ms_e2e:.3f} ms/step")
    e2e_value = w.batch * world / (ms_e2e * 1e-3)

    # ---- what the host link can deliver: every rank copies its pinned image batch to its GPU at the same time, nothing else running
    probe = torch.empty_like(step.slots[0]["image"])
    for _ in range(2):
        probe.copy_(host[0]["image"], non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        probe.copy_(host[i % NB]["image"], non_blocking=True)
    e1.record()
    barrier()
    ms_copy = max_over_ranks(e0.elapsed_time(e1)) / 10
    img_bytes = probe.numel() * 4
    h2d_link = {"GB/s_per_gpu_all_ranks_copying": round(img_bytes / (ms_copy * 1e-3) / 1e9, 1), "copy_ms_of_one_image_batch": round(ms_copy, 3),
                "bytes": img_bytes, "ranks_copying_at_once": world,
                "note": "pinned host -> device, one cudaMemcpyAsync per batch, slowest rank; a step cannot be shorter end to end than this copy when the batch comes from the host"}
    del probe

    # ---- end to end with the feature table RESIDENT in HBM (vqa_b200.shards.ShardLoader): the same three batches written as one
    # shard directory; every step the host sends row indices, tokens and CSR answer triplets (a few hundred KB) and two kernels
    # assemble the batch on the device; loss read back every step.  Reported next to "e2e", never instead of it. ----------------
    e2e_table = None
    if not args.no_resident_table and not args.no_graph and not args.quick:
        import shutil
        import tempfile
        from vqa_b200 import shards
        tmp = tempfile.mkdtemp(prefix=f"vqa_shards_r{rank}_")
        try:
            D = w.feat_dim - 4
            img = torch.cat([hb["image"] for hb in host]).numpy()
            tgt = torch.cat([hb["target"] for hb in host])
            ans = [[] for _ in range(tgt.shape[0])]
            for r, c in tgt.nonzero().tolist():
                ans[r].append((c, float(tgt[r, c])))
            shards.write_shards(tmp, features=img[..., :D], boxes=img[..., D:], questions=torch.cat([hb["question"] for hb in host]).numpy(),
                                qlen=torch.cat([hb["qlen"] for hb in host]).numpy(), image_row=np.arange(tgt.shape[0]), qid=np.arange(tgt.shape[0]),
                                answers=ans, votes=ans, n_answers=w.out_dim)
            del img
            ld = shards.ShardLoader(tmp, w.batch, dev, shuffle=False, order="none")      # uploads the table once (untimed, like a dataset load)
            order = ld.batches()
            table_bytes = ld.features.numel() * ld.features.element_size() + ld.boxes.numel() * 4

            def stage(i):
                """Assemble batch i on the copy stream (index H2D + gather/scatter kernels) and hand it to the step's idle slot."""
                with torch.cuda.stream(step.copy_stream):                # the image batch is gathered straight into the step's idle input slot
                    q, a, _nv, _qid, image, k, qlen, _idx = ld.assemble(order[i % len(order)], image_out=step.input_slot("image"))
                step.prefetch(q, image, k, qlen, a)
                return q, image, k, qlen, a

            def table_loop(n):
                cur = stage(0)
                for i in range(n):
                    loss = step(*cur)
                    loss_host[i & 1].copy_(loss.detach().reshape(()), non_blocking=True)
                    loss_ev[i & 1].record()
                    if i + 1 < n:
                        cur = stage(i + 1)
                    if i > 0:
                        loss_ev[(i - 1) & 1].synchronize()
                loss_ev[(n - 1) & 1].synchronize()
                return float(loss_host[(n - 1) & 1])

            table_loop(max(3, args.warmup))
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tl = table_loop(args.steps)
            e1.record()
            barrier()
            ms_tab = max_over_ranks(e0.elapsed_time(e1)) / args.steps
            ld.check_errors()
            B = w.batch
            nnz = int(ld.set.ans_ptr[B])
            idx_bytes = B * w.q_width * 8 + B * 8 + 2 * ((B + 1) * 8 + nnz * 8) + B * 4
            e2e_table = {"value": round(B * world / (ms_tab * 1e-3), 1), "unit": "questions/s", "ms_per_step": round(ms_tab, 4),
                         "h2d_bytes_per_step": idx_bytes, "d2h_bytes_per_step": 4, "table_bytes_in_hbm": int(table_bytes), "last_loss": tl,
                         "note": "vqa_b200.shards.ShardLoader, feature table resident in HBM (fp32): per step the host sends tokens, image "
                                 "rows and CSR answer triplets, gather_image / scatter_targets assemble the batch on the device"}
            del ld
            _dbg(f"resident-table loop done: {ms_tab:.3f} ms/step")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    # ---- the same step in bf16 mode (BASELINE config[1] names both precisions): single-pass tensor-core products outside the
    # graph-learner chain; same model, optimizer and batches, its own captured graph --------------------------------
    bf16_mode = None
    if args.precision == "fp32" and not args.no_graph and not args.quick:
        ops.set_precision("bf16")
        step16 = TrainStep(model, opt, criterion, reducer=reducer, use_graph=True, seed=4321 + rank)
        for i in range(max(args.warmup, 3)):
            step16(*(resident[i % NB][k] for k in keys))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            l16 = step16(*(resident[i % NB][k] for k in keys))
        e1.record()
        barrier()
        ms16 = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        bf16_mode = {"value": round(w.batch * world / (ms16 * 1e-3), 1), "unit": "questions/s", "ms_per_step": round(ms16, 4),
                     "final_loss": float(l16.detach()),
                     "note": "--precision bf16: 1-pass bf16 tcgen05 products (graph-learner chain stays 3-pass), device-resident batches"}
        ops.set_precision(args.precision)
        _dbg(f"bf16-mode loop done: {ms16:.3f} ms/step")

    # ---- per-kernel CUDA-event times: an eager pass over the same step (individual launches cannot be bracketed inside a
    # graph replay), same inputs, same stream ------------------------------------------------------------------
    GC, ADJ = "vqa_graphconv_fwd_f32", "vqa_adjacency_topk_fwd_f32"
    # every rank runs it (the gradient all-reduce inside the step is a collective); rank 0 reads the timers
    timers = {}
    eager = TrainStep(model, opt, criterion, reducer=reducer, use_graph=False, seed=99)
    for i in range(0 if args.quick else 3):
        eager(*(resident[i % NB][k] for k in keys))
    kn.enable_timing(GC, ADJ, "vqa_graphconv_mma_fwd", "vqa_graphconv_mma_pool_fwd", "vqa_graphconv_mma_bwd_data", "vqa_graphconv_mma_bwd_edges",
                     "vqa_adjacency_topk_bwd_f32", "vqa_gemm_bf16s")
    spin = int(10e-3 * getattr(torch.cuda.get_device_properties(local), "clock_rate", 1.9e6) * 1e3)   # ~10 ms of SM clock
    kn.GEMM_LOG = []
    for i in range(0 if args.quick else 6):
        # the host needs ~5 ms to enqueue one eager step: let it run ahead of the GPU behind a spin kernel, so that every
        # bracketed launch starts the moment its predecessor ends and the events see device time only, no launch gaps
        torch.cuda.synchronize()
        torch.cuda._sleep(spin)
        eager(*(resident[i % NB][k] for k in keys))
    torch.cuda.synchronize()
    timers = {k: [a.elapsed_time(b) for a, b in v] for k, v in kn.TIMERS.items()}
    gemm_log, kn.GEMM_LOG = kn.GEMM_LOG, None
    kn.enable_timing()
    _dbg("eager timing pass done")
    barrier()

    def leave():
        """Multi-rank runs end with a hard exit after a final barrier: tearing down NCCL communicators that captured CUDA
        graphs still reference can block in destroy_process_group (seen at N=2), and there is nothing left to clean up."""
        if world > 1:
            barrier()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        leave()
        return

    # ---- roofline of the graph kernels (algorithmic bytes, SURVEY.md 8d / DESIGN.md) --------------------------
    hbm_peak, tf_peak, peak_src = peaks()
    B, K, nb, H = w.batch, w.n_obj, w.neighbourhood, w.hid_dim
    Mrows = B * K

    def med(xs):
        xs = sorted(xs)
        return xs[len(xs) // 2] if xs else None

    GCM = "vqa_graphconv_mma_fwd"
    gc_name = GCM if timers.get(GCM) else GC
    gc_ms = med(timers.get(gc_name, []))
    gc_bytes = 2 * Mrows * 2 * H * 4 + 2 * Mrows * nb * 4 + Mrows * 16      # read Y1 + write G1 (4 B/element each: fp32 or hi+lo bf16) + idx/alpha + boxes
    roof = None
    extra = {}
    if gc_ms:
        ach = gc_bytes / (gc_ms * 1e-3) / 1e9
        roof = {"kernel": gc_name + " (layer 1: Gaussian weights on selected edges + neighbourhood aggregate + ReLU + dropout)", "bound": "hbm",
                "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4), "traffic": NCU_TRAFFIC.get((w.name, gc_name)),
                "algorithmic_bytes": gc_bytes, "us_per_launch": round(gc_ms * 1e3, 2), "peak_source": peak_src,
                "timed": "CUDA events on the launching stream in an eager pass of the same step, enqueued behind a spin kernel so no host launch gap is inside the bracket (launches inside a graph replay cannot be bracketed)"}
    alg = {ADJ: Mrows * 512 * 4 + Mrows * K * 4 + 2 * Mrows * nb * 4,
           "vqa_graphconv_mma_pool_fwd": Mrows * H * 4 + Mrows * nb * 4 + Mrows * 16 + 4 * B * H * 4,
           "vqa_graphconv_mma_bwd_data": 2 * Mrows * 2 * H * 4 + 2 * Mrows * nb * 4 + Mrows * 16,
           "vqa_adjacency_topk_bwd_f32": 2 * Mrows * 512 * 4 + 3 * Mrows * nb * 4}
    for k, nbytes in alg.items():
        m = med(timers.get(k, []))
        if m:
            extra[k] = {"us": round(m * 1e3, 2), "GB/s": round(nbytes / (m * 1e-3) / 1e9, 1), "frac": round(nbytes / (m * 1e-3) / 1e9 / hbm_peak, 4)}
    ed = sorted(timers.get("vqa_graphconv_mma_bwd_edges", []))
    if ed:   # two launches per step: layer 2 (pooled upstream, reads Y2) then layer 1 (reads dG1 and Y1)
        extra["vqa_graphconv_mma_bwd_edges"] = {"us_layer2_layer1": [round(ed[0] * 1e3, 2), round(ed[-1] * 1e3, 2)],
                                                "GB/s_layer1": round((2 * Mrows * 2 * H * 4) / (ed[-1] * 1e-3) / 1e9, 1)}
    # dense projections: total tensor-core time and rate of the split-bf16 GEMMs in one step
    gm = timers.get("vqa_gemm_bf16s", [])
    roof_gemm = None
    ceiling = None
    if gm:
        per_step = sum(gm) / 6.0                                             # 6 timed eager steps
        passes = 1 if args.precision == "bf16" else 3
        # FLOPs of the products actually issued (kernels.GEMM_LOG: M, N, K, passes per launch).  Row-gated products (the padded GRU
        # recurrence) skip the 128-row tiles whose questions have all ended: counted by the tiles alive at that step.
        qh = [sorted((int(x) for x in hb["qlen"]), reverse=True) for hb in host]
        def alive_rows(t):                                                   # rows of tiles with a question longer than t (batches are length-sorted)
            n = [sum(128 for r0 in range(0, len(q), 128) if q[r0] > t) for q in qh]
            return sum(n) / len(n)
        def gated_rows(M_, gate):                                            # gate: None | step t of a per-step product | -1 = all steps of a (T*B, .) product
            if gate is None:
                return M_
            if gate == -1:
                return sum(min(B, alive_rows(t)) for t in range(M_ // B))
            return min(M_, alive_rows(gate))
        raw = useful = 0.0
        for (M_, N_, K_, ps, gate) in gemm_log:
            rows = gated_rows(M_, gate)
            useful += 2.0 * rows * N_ * K_
            raw += 2.0 * rows * N_ * K_ * ps
        raw, useful = raw / 6.0, useful / 6.0
        ach_tf = raw / (per_step * 1e-3) / 1e12
        if args.gemm_table and len(gm) == len(gemm_log) and len(gm) % 6 == 0:
            n = len(gm) // 6
            with open(args.gemm_table, "w") as f:
                for j in range(n):                                           # launch j of each of the 6 timed steps
                    M_, N_, K_, ps, gate = gemm_log[j]
                    rows = gated_rows(M_, gate)
                    ms = med([gm[i * n + j] for i in range(6)])
                    tf = 2.0 * rows * N_ * K_ * ps / (ms * 1e-3) / 1e12
                    f.write(json.dumps({"launch": j, "M": M_, "N": N_, "K": K_, "passes": ps, "row_gate_step": gate, "live_rows": rows, "us": round(ms * 1e3, 1),
                                        "mma_tflops": round(tf, 1), "frac_of_bf16_peak": round(tf / tf_peak, 3)}) + "\n")
        roof_gemm = {"kernel": "gemm_bf16s_kernel / gemm_bf16s_persistent_kernel (all dense projections of one step, event-bracketed launches incl. gaps)",
                     "bound": "tensor", "ms_per_step": round(per_step, 3), "launches_per_step": len(gm) // 6, "passes": passes,
                     "useful_tflop_per_step": round(useful / 1e12, 4), "mma_tflop_per_step": round(raw / 1e12, 4),
                     "achieved": round(ach_tf, 1), "achieved_useful": round(useful / (per_step * 1e-3) / 1e12, 1),
                     "peak": tf_peak, "frac": round(ach_tf / tf_peak, 4),
                     "unit": "TFLOP/s (bf16 MMA rate incl. the 3 passes; peak = bf16 sustained, measured)" if peak_src == "measured" else "TFLOP/s (fallback peak)"}
        # questions/s against the tensor-core ceiling: every required product once, in bf16, at the sustained peak
        ceil_qps = tf_peak * 1e12 / (useful / B)
        ceiling = {"ceiling_questions_per_s_per_gpu": round(ceil_qps, 0), "useful_gflop_per_question": round(useful / B / 1e9, 3),
                   "frac": round(value / world / ceil_qps, 4),
                   "note": "value per GPU / (measured sustained bf16 peak / dense FLOPs one question needs, forward + backward); the fp32-grade mode issues every product 3 times"}
    cpu = None
    if not args.no_cpu_baseline and world == 1 and not args.quick:      # reported on rank 0 at N = 1 only
        from vqa_b200.synthetic import WORKLOADS
        cores = os.cpu_count() or 1
        t, _, kind = cpu_train_steps(WORKLOADS["vqa2_b64"], args.cpu_sample, 8, 2, cores)      # ~20 s of host work on the 16-core boxes
        v = args.cpu_sample * len(t) / sum(t)
        cpu = {"value": round(v, 3), "unit": "questions/s", "cores": cores, "kind": kind,
               "sample": f"{len(t)} train steps of {args.cpu_sample} questions (BASELINE config[0] shapes), "
                         + ("the unmodified reference from baseline/_ref, PyTorch CPU" if kind == "reference" else "oracle port, PyTorch CPU")}
    # the reference in eager PyTorch CUDA on this same GPU (SURVEY.md 8d "same-box competitor"): its own Model / criterion / Adam at the
    # bench batch, fp32 matmuls and with TF32 allowed; CUDA-synchronised wall clock per step
    eager = None
    if not args.no_cpu_baseline and world == 1 and not args.quick:
        try:
            del resident
            torch.cuda.empty_cache()
            e32 = reference_train_steps(w, w.batch, 10, 3, dev)
            if e32 is not None:
                etf = reference_train_steps(w, w.batch, 10, 3, dev, tf32=True)
                m32, mtf = sorted(e32[0])[len(e32[0]) // 2], sorted(etf[0])[len(etf[0]) // 2]
                eager = {"fp32": {"ms_per_step": round(m32 * 1e3, 2), "questions_per_s": round(w.batch / m32, 1)},
                         "tf32": {"ms_per_step": round(mtf * 1e3, 2), "questions_per_s": round(w.batch / mtf, 1)},
                         "peak_memory_GB": round(e32[2], 1), "batch": w.batch, "speedup_vs_fp32": round(value / (w.batch / m32), 2),
                         "speedup_vs_tf32": round(value / (w.batch / mtf), 2),
                         "what": "the unmodified reference (baseline/_ref) in eager PyTorch CUDA on the same B200: Model.forward + MultiLabelSoftMarginLoss + backward + Adam, "
                                 "same workload and batch, median of 10 synchronised steps after 3 warm-up"}
        except Exception as e:            # e.g. out of memory on a large workload: report, never fail the bench line
            eager = {"error": repr(e)[:300]}

    line = {
        "metric": "train questions/sec (fwd+bwd)", "value": round(value, 1), "unit": "questions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32 (split-bf16 x3 tensor-core products, fp32 accumulate)",
        "data": "synthetic", "config": config_dict(w, args, world),
        "e2e": {"value": round(e2e_value, 1), "unit": "questions/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e, 4), "last_loss": last_loss, "h2d_link": h2d_link},
        "gpu_launches": int(launches),
        "roofline": roof, "roofline_other_kernels": extra, "roofline_gemm": roof_gemm, "tensor_ceiling": ceiling, "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "clocks": sampler.summary(), "final_loss": final_loss, "bf16_mode": bf16_mode,
        "e2e_resident_table": e2e_table,
    }
    line["config"]["host_affinity"] = numa
    if getattr(reducer, "p2p", False):
        line["config"]["gradient_exchange"] = ("NVLink peer memory (symmetric memory): one fused kernel per rank = reduce-scatter of the flat gradient buffer + Adam on "
                                               "the rank's 1/N slice (sharded moments) + all-gather of the updated parameters, bracketed by two flag barriers; no NCCL call in the step")
    else:
        line["config"]["gradient_exchange"] = {"collective": "bucketed NCCL all-reduce of the flat gradient buffer" if world > 1 else "none (one GPU)", "buckets_mb": args.bucket_mb,
                                               "start": "inside backward, as each bucket's last gradient kernel is enqueued" if reducer.early else "when the fused operator's autograd node returns",
                                               "sm_reserve": step.sm_reserve if world > 1 else 0, "NCCL_MAX_CTAS": os.environ.get("NCCL_MAX_CTAS")}
    print(json.dumps(line), file=json_out, flush=True)
    leave()


def run_eval(args, workload):
    """BASELINE.json configs[4]: the inference sweep of run.py::test (`run.py:274-341`) - model.eval(), no_grad, B=4096 questions over
    K=100 boxes, top-k=32 - forward only.  Replicas only at N > 1 (no collective: SURVEY.md 8e).  value = questions/s with the batch
    resident in HBM; e2e = pinned host batch copied every step + the predicted answer ids read back (what the driver's loop consumes)."""
    import torch.distributed as dist
    from vqa_b200 import kernels as kn, ops
    from vqa_b200.synthetic import make_batch, make_wemb
    import sparse_graph_model as M

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the vqa_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(args.precision)
    w = workload
    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs()).to(dev).eval()
    model.max_question_len = w.max_qlen
    NB = 2
    host = []
    for i in range(NB):
        b = make_batch(w, seed=1000 + 17 * rank + i)
        host.append({"question": b["question"].pin_memory(), "image": b["image"].pin_memory(), "K": b["K"].pin_memory(),
                     "qlen": torch.tensor([int(x) for x in b["qlen"]], dtype=torch.int32).pin_memory()})
    resident = [{k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    def fwd(b):
        with torch.no_grad():
            logits, adj, arg = model(b["question"], b["image"], b["K"], b["qlen"])
        return logits

    n0 = kn.LAUNCHES
    fwd(resident[0])
    launches = kn.LAUNCHES - n0
    for i in range(max(args.warmup, 3)):
        fwd(resident[i % NB])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        logits = fwd(resident[i % NB])
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    sampler.stop_flag = True
    value = w.batch * world / (ms_step * 1e-3)

    # end to end: double-buffered device slots, the copy of batch i+1 overlaps the forward of batch i, predictions read back every step
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{k: torch.empty_like(v) for k, v in resident[0].items()} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    pred_host = [torch.empty(w.batch, dtype=torch.int64).pin_memory() for _ in range(2)]
    pred_ev = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            for k, v in host[i % NB].items():
                slots[s][k].copy_(v, non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            freed[s].record()
        stage(0)
        for i in range(n):
            s = i & 1
            torch.cuda.current_stream().wait_event(ready[s])
            if i + 1 < n:
                stage(i + 1)
            lg = fwd(slots[s])
            freed[s].record()
            pred_host[s].copy_(lg.argmax(dim=1), non_blocking=True)
            pred_ev[s].record()
            if i > 0:
                pred_ev[(i - 1) & 1].synchronize()
        pred_ev[(n - 1) & 1].synchronize()
        return int(pred_host[(n - 1) & 1][0])

    e2e_loop(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    # per-kernel event times (eager launches, behind a spin kernel so the host is never inside a bracket)
    names = ("vqa_adjacency_topk_fwd_f32", "vqa_graphconv_mma_fwd", "vqa_graphconv_mma_pool_fwd", "vqa_gemm_bf16s")
    kn.enable_timing(*names)
    kn.GEMM_LOG = []
    spin = int(20e-3 * getattr(torch.cuda.get_device_properties(local), "clock_rate", 1.9e6) * 1e3)
    for i in range(4):
        torch.cuda.synchronize()
        torch.cuda._sleep(spin)
        fwd(resident[i % NB])
    torch.cuda.synchronize()
    timers = {k: [a.elapsed_time(b) for a, b in v] for k, v in kn.TIMERS.items()}
    gemm_log, kn.GEMM_LOG = kn.GEMM_LOG, None
    kn.enable_timing()
    barrier()
    if rank != 0:
        if world > 1:
            barrier()
            os._exit(0)
        return
    hbm_peak, tf_peak, peak_src = peaks()
    B, K, nb, H = w.batch, w.n_obj, w.neighbourhood, w.hid_dim
    Mrows = B * K
    med = lambda xs: sorted(xs)[len(xs) // 2] if xs else None
    alg = {"vqa_graphconv_mma_fwd": 2 * Mrows * 2 * H * 4 + 2 * Mrows * nb * 4 + Mrows * 16,
           "vqa_graphconv_mma_pool_fwd": Mrows * H * 4 + Mrows * nb * 4 + Mrows * 16 + 4 * B * H * 4,
           "vqa_adjacency_topk_fwd_f32": Mrows * 512 * 4 + Mrows * K * 4 + 2 * Mrows * nb * 4}
    per = {}
    for k, nbytes in alg.items():
        m = med(timers.get(k, []))
        if m:
            per[k] = {"us": round(m * 1e3, 1), "GB/s": round(nbytes / (m * 1e-3) / 1e9, 1), "frac": round(nbytes / (m * 1e-3) / 1e9 / hbm_peak, 4),
                      "algorithmic_bytes": nbytes}
    gc = per.get("vqa_graphconv_mma_fwd")
    roof = None
    if gc:
        roof = {"kernel": "vqa_graphconv_mma_fwd (layer 1 aggregate at K=100, nb=32)", "bound": "hbm", "achieved": gc["GB/s"], "peak": hbm_peak, "unit": "GB/s",
                "frac": gc["frac"], "traffic": NCU_TRAFFIC.get((w.name, "vqa_graphconv_mma_fwd")), "algorithmic_bytes": gc["algorithmic_bytes"],
                "us_per_launch": gc["us"], "peak_source": peak_src}
    gm = timers.get("vqa_gemm_bf16s", [])
    roof_gemm = None
    if gm:
        per_step = sum(gm) / 4.0
        useful = sum(2.0 * m_ * n_ * k_ for (m_, n_, k_, ps, g) in gemm_log) / 4.0
        raw = sum(2.0 * m_ * n_ * k_ * ps for (m_, n_, k_, ps, g) in gemm_log) / 4.0
        roof_gemm = {"bound": "tensor", "ms_per_step": round(per_step, 3), "launches_per_step": len(gm) // 4, "mma_tflop_per_step": round(raw / 1e12, 3),
                     "achieved": round(raw / (per_step * 1e-3) / 1e12, 1), "peak": tf_peak, "frac": round(raw / (per_step * 1e-3) / 1e12 / tf_peak, 4),
                     "unit": "TFLOP/s (bf16 MMA rate incl. passes; GRU row gating not subtracted)",
                     "tensor_ceiling_frac": round(value / world / (tf_peak * 1e12 / (useful / B)), 4)}
    line = {"metric": "eval questions/sec (forward only)", "value": round(value, 1), "unit": "questions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32 (split-bf16 x3 tensor-core products, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{w.name}: inference sweep, model.eval(), per-GPU batch {w.batch}, K={w.n_obj} boxes x {w.feat_dim}-d, top-k={w.neighbourhood}, "
                                   f"{w.n_kernels} Gaussian kernels, {w.out_dim} answers (BASELINE.json configs[4])",
                       "global_batch": w.batch * world, "step": "Model.forward under no_grad, eager launches", "parallelism": f"replicas x{world}",
                       "l2_policy": f"inputs larger than L2 (image batch {B * K * w.feat_dim * 4 / 1e9:.2f} GB), 2 rotating batches"},
            "e2e": {"value": round(w.batch * world / (ms_e2e * 1e-3), 1), "unit": "questions/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": w.batch * 8,
                    "ms_per_step": round(ms_e2e, 3)},
            "gpu_launches": int(launches), "roofline": roof, "roofline_other_kernels": per, "roofline_gemm": roof_gemm, "cpu_baseline": None,
            "clocks": sampler.summary()}
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        barrier()
        os._exit(0)


def main():
    args = parse()
    from vqa_b200.synthetic import WORKLOADS
    workload = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, workload)
    elif workload.name.startswith("eval"):
        run_eval(args, workload)
    else:
        run_b200(args, workload)


if __name__ == "__main__":
    main()
