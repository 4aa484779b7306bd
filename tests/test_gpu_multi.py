"""GPU tests of album mode across ranks through the C ABI (rg_analyze_album_sharded): with one rank it is analyze_album; with
two engines on ONE GPU the exchange goes through the host-transport hook (two threads, the reduction done here); with two
or more GPUs visible the real thing runs under torchrun over the engine's own NCCL communicator (tools/multi_gpu_check.py).
Reference: the accumulate loop of src/replaygain.rs:1044-1074."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import synth_corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus(n=9):
    return [synth_corpus.make_track(seed=900 + i, n_frames=70 + 30 * i, channels=1 + (i % 2)) for i in range(n)]


def test_one_rank_is_analyze_album(engine_fast, oracle):
    imgs = _corpus()
    a = engine_fast.analyze_album_images(imgs, want_hist=True)
    b = engine_fast.analyze_album_sharded_images(imgs, 0)
    assert (a.album_histogram == b.album_histogram).all()
    assert (a.album_loudness_db, a.album_gain_db, a.album_peak) == (b.album_loudness_db, b.album_gain_db, b.album_peak)
    hs, ps = [], []
    for im in imgs:
        rc, res, h, info = oracle.analyze_mp3(im)
        hs.append(h)
        ps.append(res.peak)
    merged, loud, gain, peak = oracle.album_merge(np.stack(hs), ps)
    assert (b.album_histogram == merged.astype(np.uint64)).all() and b.album_loudness_db == loud and b.album_peak == peak
    import mp3rgain_b200 as m
    broken = list(imgs)
    broken[4] = np.zeros(300, np.uint8)
    with pytest.raises(m.ReplayGainError) as ei:
        engine_fast.analyze_album_sharded_images(broken, 0, names=[f"t{i}.mp3" for i in range(len(imgs))])
    assert ei.value.status == -5 and ei.value.index == 4 and str(ei.value) == "Failed to probe format: t4.mp3"


def test_two_engines_one_gpu_exchange_through_the_host_hook(oracle):
    """Two ranks as two engines on the same device, one thread each: every entry point holds its engine's lock, the exchange
    function is called once per album with the two host buffers, and both ranks return the album of the whole list."""
    if os.environ.get("RG_SIMT_EMU") == "1":
        pytest.skip("the CPU interpreter of tests/dev runs one kernel at a time (process-wide state): two engines cannot overlap there")
    import mp3rgain_b200 as m
    from mp3rgain_b200 import sharded
    imgs = _corpus()
    shards = sharded.shard_by_bytes([len(i) for i in imgs], 2)
    offs = sharded.shard_offsets(shards)
    meet = threading.Barrier(2)
    box = {}

    def xchg_for(rank):
        def xchg(s, mx):
            box[rank] = (s.copy(), mx.copy())
            meet.wait(timeout=120)
            s[:] = box[0][0] + box[1][0]
            mx[:] = np.maximum(box[0][1], box[1][1])
            meet.wait(timeout=120)
        return xchg

    res = [None, None]

    def run(rank, broken_at=None):
        with m.Engine(device=0) as eng:
            eng.comm_attach_host(rank, 2, xchg_for(rank))
            mine = [imgs[i] if i != broken_at else np.zeros(len(imgs[i]), np.uint8) for i in shards[rank]]
            try:
                res[rank] = eng.analyze_album_sharded_images(mine, offs[rank])
            except m.ReplayGainError as ex:
                res[rank] = (ex.status, ex.index)

    for broken_at in (None, shards[1][-1]):
        ths = [threading.Thread(target=run, args=(r, broken_at)) for r in range(2)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=300)
        if broken_at is None:
            hs, ps = [], []
            for im in imgs:
                rc, r, h, info = oracle.analyze_mp3(im)
                hs.append(h)
                ps.append(r.peak)
            merged, loud, gain, peak = oracle.album_merge(np.stack(hs), ps)
            for r in res:
                assert (r.album_histogram == merged.astype(np.uint64)).all() and r.album_loudness_db == loud and r.album_peak == peak
            assert len(res[0].tracks) + len(res[1].tracks) == len(imgs)
        else:
            assert res[0] == res[1] == (-5, broken_at)     # the rank that held only sound files fails the album too


def test_concurrent_calls_on_one_engine_queue_on_its_lock(engine_fast, oracle):
    imgs = _corpus(6)
    want = [oracle.analyze_mp3(im)[1].loudness_db for im in imgs]
    got = [None] * 4

    def work(k):
        res, sts = engine_fast.analyze_images(imgs)
        got[k] = [r.loudness_db for r in res]

    ths = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=300)
    assert all(g == want for g in got)


def test_album_over_nccl_when_two_gpus_are_visible():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "16"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    assert "album histogram identical to the oracle sum: True" in p.stdout
