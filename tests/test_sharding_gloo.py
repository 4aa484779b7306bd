"""CPU tests of the N>1 path with a world_size-2 gloo group: byte-balanced sharding and the album-mode exchange that the
engine owns (rg_analyze_album_sharded: sum of the 12000-bin histograms, max of {peak, error key}).  There is no GPU and no
NCCL here, so the engine is the build of its kernel sources on the CPU SIMT interpreter and the exchange travels through the
engine's host-transport hook (rg_engine_comm_attach_host) over gloo; the merged result must equal the oracle's
single-process album merge, and a failing file must fail the album on every rank with that file's index and status."""
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


def test_shard_offsets_are_the_global_index_of_each_shard():
    from mp3rgain_b200.sharded import shard_by_bytes, shard_offsets
    sizes = [5, 1, 9, 9, 2, 7, 3]
    for world in (1, 2, 3, 8):
        sh = shard_by_bytes(sizes, world)
        offs = shard_offsets(sh)
        assert all(s[0] == o for s, o in zip(sh, offs) if s)   # rg_analyze_album_sharded's shard_offset
        assert offs[0] == 0 and offs == sorted(offs)


def _worker_engine(rank, world, port, q):
    """rg_analyze_album_sharded, per-rank analysis by the engine built from the kernel sources on the CPU SIMT interpreter
    (tests/dev/simt_emu), the exchange through rg_engine_comm_attach_host over gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "dev", "simt_emu"))
    import torch
    import torch.distributed as dist
    import build_emu
    from mp3rgain_b200 import _lib
    _lib.LIB_PATH = build_emu.build()
    import mp3rgain_b200 as m
    import oracle_lib
    import synth_corpus
    from mp3rgain_b200.sharded import analyze_album_sharded, connect, gain_analyze_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    imgs = [synth_corpus.make_track(seed=140 + i, n_frames=50 + 20 * i, channels=1 + (i % 2), version=3 if i % 3 else 2) for i in range(7)]
    with m.Engine(device=0, exact=False) as eng:
        connect(eng, rank, world, transport="host")
        alb, mine = analyze_album_sharded(eng, imgs, rank, world)
        gst, mine2 = gain_analyze_sharded(eng, imgs, rank, world)
        ok_tracks = all((orc.analyze_mp3(imgs[i])[1].loudness_db == t.loudness_db) for i, t in zip(mine, alb.tracks))
        ok_gain = mine2 == mine and all(orc.gain_analyze(imgs[i])[1].gain_sum == g.gain_sum for i, g in zip(mine, gst))
        bad_rc = []
        for bad_at in ([0], [6], [5, 1]):   # a failing file on ONE rank aborts the album on every rank, with the error of the
            try:                            # failing file that comes first in the list (src/replaygain.rs:1055)
                broken = list(imgs)
                for k in bad_at:
                    broken[k] = np.zeros(len(imgs[k]), np.uint8)    # (same sizes: the same shards)
                analyze_album_sharded(eng, broken, rank, world, names=[f"f{k}.mp3" for k in range(7)])
                bad_rc.append(None)
            except m.ReplayGainError as ex:
                bad_rc.append((ex.status, ex.index, str(ex)))
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
        # both ranks saw every failure, although only one of them held the broken file; two broken files: the first one wins
        assert r[8] == [(-5, 0, "Failed to probe format: f0.mp3"), (-5, 6, "Failed to probe format: f6.mp3"),
                        (-5, 1, "Failed to probe format: f1.mp3")], r
    assert got[0][7] + got[1][7] == 7
