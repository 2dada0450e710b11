"""Step-tail timing at N GPUs of one node (torchrun): flag barrier alone, the fused peer-memory reduce-scatter + Adam + all-gather,
NCCL all-reduce + single-GPU Adam on the same 30.3 M-element flat buffers (the VQA2 model's size).  CUDA events, max over ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_adam_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vqa-project_b200")]
import torch, torch.distributed as dist
from vqa_b200 import kernels as kn
from vqa_b200.ddp import symmetric_empty

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 30_300_000 // 4 * 4
G = torch.zeros(n, device=dev)
P, paddr, h2 = symmetric_empty(n, torch.float32, dev)
F, faddr, h3 = symmetric_empty(64, torch.int32, dev)
G.normal_(); P.normal_()
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
epoch = torch.zeros(1, dtype=torch.int32, device=dev)
lr = torch.full((1,), 1e-4, device=dev); state = torch.zeros(2, dtype=torch.int32, device=dev)
per = (n // 4 + world - 1) // world * 4
lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
chunks = torch.tensor([[P.data_ptr() + 4 * s, s, min(4096, n - s)] for s in range(0, n, 4096)], dtype=torch.int64, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
torch.cuda.synchronize(); dist.barrier()

def timeit(name, fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(iters):
        flush.zero_()
        kn.p2p_barrier(faddr, rank, world, epoch)          # align the ranks on the device before the bracket
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(json.dumps({"world": world, "what": name, "us": round(t.item(), 1)}), flush=True)

CHL = 20; CH = 1 << CHL; row = CH * world
nrows = (n + row - 1) // row; n = nrows * row; n_own = nrows * CH
G = torch.randn(n, device=dev); m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
P, paddr, h2 = symmetric_empty(n, torch.float32, dev)
R_, raddr, h4 = symmetric_empty(world * n_own, torch.float32, dev)
chunks = torch.tensor([[P.data_ptr() + 4 * s, s, min(4096, n - s)] for s in range(0, n, 4096)], dtype=torch.int64, device=dev)
peers_r = [None if q == rank else h4.get_buffer(q, (world * n_own,), torch.float32, 0) for q in range(world)]
def push():
    for d in range(1, world):
        q = (rank + d) % world
        for k in range(nrows):
            peers_r[q][rank * n_own + k * CH:rank * n_own + (k + 1) * CH].copy_(G[(k * world + q) * CH:(k * world + q + 1) * CH], non_blocking=True)
def kernel(paddrs=None):
    kn.adam_flat_p2p(G, R_, n_own, CHL, paddr if paddrs is None else paddrs, m, v, rank, world, lr, 0.9, 0.999, 1e-8, 0.0, 1.0 / world, state)
def fused():
    push()
    kn.p2p_barrier(faddr, rank, world, epoch)
    kernel()
    kn.p2p_barrier(faddr, rank, world, epoch)
def nccl():
    dist.all_reduce(G)
    kn.adam_flat(chunks, G, m, v, lr, 0.9, 0.999, 1e-8, 0.0, 1.0 / world, state)
tmp = torch.empty(n, device=dev)
timeit("read+write 121 MB: regular -> regular (torch copy_)", lambda: tmp.copy_(G))
timeit("read+write 121 MB: symmetric -> regular", lambda: tmp.copy_(P))
timeit("read+write 121 MB: regular -> symmetric", lambda: P.copy_(G))
Gs = kn.split(G[:30_000_000].view(-1, 3000))
timeit("split planes of a regular fp32 matrix (120 MB)", lambda: kn.split(G[:30_000_000].view(-1, 3000)))
timeit("split planes of a symmetric fp32 matrix (120 MB)", lambda: kn.split(P[:30_000_000].view(-1, 3000)))
timeit("flag barrier", lambda: kn.p2p_barrier(faddr, rank, world, epoch))
timeit("copy-engine push of the gradient slices to their owners", push)
timeit("sum + Adam on my slice + parameter stores to all ranks (one kernel)", kernel)
timeit("the same kernel with all parameter stores local", lambda: kernel([paddr[rank]] * world))
if h2.multicast_ptr:
    timeit("the same kernel with ONE multicast store (multimem.st) instead of world stores",
           lambda: kn.adam_flat_p2p(G, R_, n_own, CHL, paddr, m, v, rank, world, lr, 0.9, 0.999, 1e-8, 0.0, 1.0 / world, state, h2.multicast_ptr))
timeit("whole tail, nothing hidden: push + barrier + kernel + barrier", fused)
Gm, gmaddr, h5 = symmetric_empty(n, torch.float32, dev)
Gm.copy_(G)
per_c = (n // 4 + world - 1) // world * 4
lo_c, hi_c = min(n, rank * per_c), min(n, (rank + 1) * per_c)
if h5.multicast_ptr and h2.multicast_ptr:
    def mc():
        kn.adam_flat_mc(h5.multicast_ptr, h2.multicast_ptr, P, m, v, lo_c, hi_c, lr, 0.9, 0.999, 1e-8, 0.0, 1.0 / world, state)
    def mc_tail():
        kn.p2p_barrier(faddr, rank, world, epoch); mc(); kn.p2p_barrier(faddr, rank, world, epoch)
    timeit("multicast kernel: multimem.ld_reduce + Adam + multimem.st", mc)
    timeit("whole multicast tail: barrier + kernel + barrier", mc_tail)
elif rank == 0:
    print(json.dumps({"world": world, "what": "no multicast support on this box"}), flush=True)
timeit("NCCL all-reduce (121 MB) + Adam on the whole buffer", nccl)
timeit("Adam on the whole buffer (N = 1 work)", lambda: kn.adam_flat(chunks, G, m, v, lr, 0.9, 0.999, 1e-8, 0.0, 1.0, state))
timeit("NCCL all-reduce (121 MB) alone", lambda: dist.all_reduce(G))
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
