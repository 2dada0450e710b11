import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm.rendezvous(t, dist.group.WORLD)
if rank == 0:
    print("multicast_ptr", hex(h.multicast_ptr), "signal_pad_size", h.signal_pad_size, flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
