"""Aggregate host <-> device DMA ceilings with every GPU of the box copying at once (one process per GPU under
torchrun): what the host's memory system gives the end-to-end path at N GPUs.  Rank 0 prints one line per leg."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 256 << 20
h1 = torch.empty(N, dtype=torch.uint8).pin_memory()
h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
h1.fill_(1); h2.fill_(2)
d1 = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def leg(name, nbytes, fn, n=6):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t)
        print("N=%d  %-40s %7.3f ms  per GPU %6.1f GB/s  aggregate %7.1f GB/s" % (
            world, name, ms, nbytes / ms / 1e6, world * nbytes / ms / 1e6), flush=True)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


leg("DMA host -> device, 256 MiB", N, lambda: d1.copy_(h1, non_blocking=True))
leg("DMA device -> host, 256 MiB", N, lambda: h2.copy_(d2, non_blocking=True))
leg("DMA both directions, 2 x 256 MiB", 2 * N, both)
dist.barrier()
dist.destroy_process_group()
