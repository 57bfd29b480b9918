#!/usr/bin/env python
"""Host<->device copy bandwidth of the box: each direction alone, then both at once (the ceiling of the pipelined C-ABI calls).

    python profiles/pcie_probe.py                                          # one GPU
    python -m torch.distributed.run --nproc-per-node 8 profiles/pcie_probe.py   # all ranks copy AT THE SAME TIME: what 8 pipelined
                                                                                # e2e calls can get from host memory / the PCIe tree
Every rank binds itself to its GPU's NUMA node first (as bench.py does) unless PROBE_NO_BIND is set; rank 0 prints per-rank
minimum / mean and the aggregate."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
numa = None
if not os.environ.get("PROBE_NO_BIND"):
    import bench
    numa = bench.bind_to_gpu_numa_node(torch, local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
h_a.fill_(1); h_b.fill_(2)                                    # first touch after binding
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    times = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t)
    return min(times[1:])


for name, a, b in (("H2D", 1, 0), ("D2H", 0, 1), ("both", 1, 1)):
    t = run(a, b)
    gbs = (a + b) * n / t / 1e9
    if dist is not None:
        v = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        vals = [float(x.item()) for x in allv]
        if rank == 0:
            print(f"{name}: {world} ranks at once: per rank min {min(vals):.1f} / mean {sum(vals) / world:.1f} / max {max(vals):.1f} GB/s, aggregate {sum(vals):.1f} GB/s", flush=True)
    else:
        print(f"{name}: {1e3 * t:.1f} ms for {a + b} GiB -> {gbs:.1f} GB/s", flush=True)
if rank == 0:
    print("numa binding of rank 0:", numa)
if dist is not None:
    dist.destroy_process_group()
