#!/usr/bin/env python3
"""How much host->device bandwidth do N ranks get when they all copy at once?  (The end-to-end path moves 2.9 GB of file
images per rank and step; at N = 8 it scaled 0.889 in round 1.)  Run under torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/h2d_probe.py [GB per rank]
Every rank copies a pinned buffer to its GPU five times, alone (ranks take turns) and then all together (barrier before every copy)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.9
n = int(gb * 1e9)
src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
src.fill_(rank + 1)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")


def copy_ms():
    torch.cuda.synchronize()
    t = time.perf_counter()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3


def bar():
    if world > 1:
        dist.barrier()


copy_ms()
alone = []
for r in range(world):        # one rank at a time
    bar()
    if r == rank:
        alone = [copy_ms() for _ in range(3)]
bar()
together = []
for _ in range(5):
    bar()
    together.append(copy_ms())
a, t = min(alone), sorted(together)[len(together) // 2]
vals = torch.tensor([n / a / 1e6, n / t / 1e6], dtype=torch.float64, device="cuda")
if world > 1:
    lst = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(lst, vals)
else:
    lst = [vals]
if rank == 0:
    al = [float(v[0]) for v in lst]
    tg = [float(v[1]) for v in lst]
    print("ranks %d, %.1f GB per rank; H2D alone GB/s per rank: %s" % (world, gb, " ".join("%.1f" % x for x in al)))
    print("H2D all ranks at once GB/s per rank: %s ; aggregate %.1f GB/s (alone would add up to %.1f)" % (" ".join("%.1f" % x for x in tg), sum(tg), sum(al)))
    try:
        print("cpus: %d, affinity of rank 0: %d cpus" % (os.cpu_count(), len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
if world > 1:
    dist.destroy_process_group()
