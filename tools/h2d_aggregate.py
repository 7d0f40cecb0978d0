"""What the host side of the box can feed: every rank copies pinned host buffers to its GPU
and back at the same time (the traffic pattern of bench.py's end-to-end leg, no kernels).
Explains the e2e numbers at N > 1: the aggregate host<->device rate, not the GPUs, is the
limit.    torchrun --nproc-per-node N tools/h2d_aggregate.py [GB per rank]   (H2D_NUMA=1: allocate on the GPU's NUMA node)"""
import os
import sys
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
argv = sys.argv[1:]
gb = float(argv[0]) if argv else 2.0
n = int(gb*1e9/8)
numa = None
if os.environ.get("H2D_NUMA"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from compose_b200.pipeline import near_gpu
    numa = near_gpu()
    numa.__enter__()
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n//2, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n//2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0)/3
if rank == 0:
    per = (n + n//2)*8/dt/1e9
    print("numa-local cpus: %s" % (len(numa.cpus) if numa and numa.cpus else None))
    print("ranks %d: %.1f GB in + %.1f GB out per rank per step, %.1f ms -> %.1f GB/s per rank, "
          "%.1f GB/s aggregate" % (world, n*8/1e9, n*4/1e9, 1e3*dt, per, per*world))
if world > 1:
    dist.destroy_process_group()
