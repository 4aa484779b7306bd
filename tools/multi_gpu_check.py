#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): album mode across ranks vs the CPU oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank builds the same seeded corpus and analyses its byte-balanced shard; the engine all-reduces the 12000-bin album
histogram and {peak, error key} over its own NCCL communicator (rg_analyze_album_sharded), and rank 0 compares the merged histogram / album loudness with the sum of the oracle's per-track
histograms (tests/ oracle: checker only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp3rgain_b200 as m  # noqa: E402
from mp3rgain_b200 import sharded  # noqa: E402
import oracle_lib  # noqa: E402
import synth_corpus  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pool = synth_corpus.GroupPool(seed=7, n_groups=64, group_len=16)
imgs = [pool.track(500 + i, 300 + 37 * (i % 5)) for i in range(n_tracks)]
with m.Engine(device=local) as eng:
    sharded.connect(eng, rank, world)       # the engine creates its own NCCL communicator from rank 0's unique id
    alb, mine = sharded.analyze_album_sharded(eng, imgs, rank, world)
    exch_ms = eng.stats().last_ms_exchange
    # a failing file on the LAST rank fails the album on every rank, with that file's index
    broken = list(imgs)
    broken[-1] = np.zeros(len(imgs[-1]), np.uint8)
    try:
        sharded.analyze_album_sharded(eng, broken, rank, world)
        err = None
    except m.ReplayGainError as ex:
        err = (ex.status, ex.index)
ok = True
if rank == 0:
    orc = oracle_lib.load()
    total = np.zeros(12000, np.uint64)
    peak = 0.0
    for im in imgs:
        rc, res, hist, info = orc.analyze_mp3(im)
        assert rc == 0
        total += hist.astype(np.uint64)
        peak = max(peak, res.peak)
    same = bool((total == alb.album_histogram).all())
    o_loud = orc.hist_loudness(total.astype(np.uint32))
    print(f"world {world}: album histogram identical to the oracle sum: {same}; windows {int(total.sum())}; "
          f"album loudness {alb.album_loudness_db} (oracle {o_loud}); peak {alb.album_peak} vs {peak}; shard sizes "
          f"{[len(s) for s in sharded.shard_by_bytes([len(i) for i in imgs], world)]}", flush=True)
    print(f"exchange (NCCL group: all-reduce sum u64[12000] + max u64[2]) {exch_ms:.3f} ms on the engine's stream; failing-file case -> {err}", flush=True)
    ok = same and alb.album_peak == peak and (o_loud is None or o_loud == alb.album_loudness_db)
ok = ok and err == (-5, n_tracks - 1)
if world > 1:
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    ok = int(t.item()) == 0
    dist.destroy_process_group()
print("multi_gpu_check", "ok" if ok else "FAILED", "rank", rank, flush=True)
sys.exit(0 if ok else 1)
