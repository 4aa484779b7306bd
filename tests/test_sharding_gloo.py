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


def _worker_engine(rank, world, port, q):
    """analyze_album_sharded itself, per-rank analysis by the engine built from the kernel sources on the CPU SIMT interpreter
    (tests/dev/simt_emu: device memory is host memory there, so the 'device' tensors of the collective are CPU tensors and
    gloo carries them)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "dev", "simt_emu"))
    import torch
    import torch.distributed as dist
    import build_emu
    from mp3rgain_b200 import _lib
    _lib.LIB_PATH = build_emu.build()
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda *a, **k: None
    _device = torch.device
    torch.device = lambda *a, **k: _device("cpu") if a and a[0] == "cuda" else _device(*a, **k)
    import mp3rgain_b200 as m
    import oracle_lib
    import synth_corpus
    from mp3rgain_b200.sharded import analyze_album_sharded, gain_analyze_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    imgs = [synth_corpus.make_track(seed=140 + i, n_frames=50 + 20 * i, channels=1 + (i % 2), version=3 if i % 3 else 2) for i in range(7)]
    with m.Engine(device=0, exact=False) as eng:
        alb, mine = analyze_album_sharded(eng, imgs, rank, world)
        gst, mine2 = gain_analyze_sharded(eng, imgs, rank, world)
        ok_tracks = all((orc.analyze_mp3(imgs[i])[1].loudness_db == t.loudness_db) for i, t in zip(mine, alb.tracks))
        ok_gain = mine2 == mine and all(orc.gain_analyze(imgs[i])[1].gain_sum == g.gain_sum for i, g in zip(mine, gst))
        bad_rc = None
        try:      # a failing file on ONE rank aborts the album on every rank (src/replaygain.rs:1055)
            broken = list(imgs)
            broken[0] = np.zeros(500, np.uint8)
            analyze_album_sharded(eng, broken, rank, world)
        except m.ReplayGainError as ex:
            bad_rc = ex.status if hasattr(ex, "status") else ex.args[0]
    hs, ps = [], []
    for im in imgs:
        rc, res, h, info = orc.analyze_mp3(im)
        hs.append(h)
        ps.append(res.peak)
    merged, loud_ref, gain_ref, peak_ref = orc.album_merge(np.stack(hs), ps)
    q.put((rank, bool((alb.album_histogram == merged.astype(np.uint64)).all()), alb.album_loudness_db == loud_ref, alb.album_gain_db == gain_ref,
           alb.album_peak == peak_ref, ok_tracks, ok_gain, len(mine), bad_rc))
    dist.destroy_process_group()


@pytest.mark.skipif(__import__("shutil").which("g++") is None or os.uname().machine != "x86_64", reason="needs g++ on x86-64")
def test_analyze_album_sharded_world2_gloo_on_the_interpreted_engine():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "dev", "simt_emu"))
    import build_emu
    build_emu.build()          # once, before the ranks race for it
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_engine, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in got:
        assert all(r[1:7]), r
        assert 0 < r[7] < 7
        assert r[8] is not None and r[8] != 0, r     # both ranks saw the failure, although only rank 0 held the broken file
    assert got[0][7] + got[1][7] == 7
