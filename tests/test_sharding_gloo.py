"""CPU tests of the N>1 path's host logic with a world_size-2 gloo group: byte-balanced sharding and the album-mode
exchange (all-reduce(sum) of the 12000-bin histogram + all-reduce(max) of the peak).  Per-rank analysis is done by the
oracle here (no GPU); the merged result must equal the oracle's single-process album merge."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_by_bytes_properties():
    from mp3rgain_b200.sharded import shard_by_bytes
    rng = np.random.default_rng(0)
    sizes = rng.integers(100_000, 6_000_000, 237).tolist()
    for world in (1, 2, 4, 8):
        sh = shard_by_bytes(sizes, world)
        flat = [i for s in sh for i in s]
        assert flat == list(range(len(sizes)))            # a partition, order preserved (src/replaygain.rs:1061)
        loads = [sum(sizes[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= 2 * max(sizes)  # balanced by bytes, not by count
    assert shard_by_bytes([], 4) == [[], [], [], []]
    assert shard_by_bytes([5], 2) in ([[0], []], [[], [0]])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    import synth_corpus
    from mp3rgain_b200.sharded import album_reduce, shard_by_bytes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    imgs = [synth_corpus.make_track(seed=40 + i, n_frames=60 + 25 * i, channels=1 + (i % 2)) for i in range(7)]
    mine = shard_by_bytes([len(i) for i in imgs], world)[rank]
    hist = np.zeros(12000, np.int64)
    peak = 0.0
    for i in mine:
        rc, res, h, info = orc.analyze_mp3(imgs[i])
        hist += h.astype(np.int64)
        peak = max(peak, res.peak)
    th, tp = torch.from_numpy(hist), torch.tensor([peak], dtype=torch.float64)
    album_reduce(th, tp)
    loud = orc.hist_loudness(th.numpy().view(np.uint64))
    if rank == 0:
        hs, ps = [], []
        for im in imgs:
            rc, res, h, info = orc.analyze_mp3(im)
            hs.append(h)
            ps.append(res.peak)
        merged, loud_ref, gain_ref, peak_ref = orc.album_merge(np.stack(hs), ps)
        q.put((bool((th.numpy() == merged.astype(np.int64)).all()), loud == loud_ref, float(tp.item()) == peak_ref, len(mine)))
    dist.destroy_process_group()


def test_album_merge_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] and res[1] and res[2]
    assert 0 < res[3] < 7
