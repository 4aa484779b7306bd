"""Multi-GPU driver: one process per GPU (torch.distributed), tracks sharded by bytes, no data-path collective in track
mode; album mode adds exactly one exchange -- all-reduce(sum) of the 12000-bin histogram and all-reduce(max) of the peak
(the accumulate loop of the reference, src/replaygain.rs:1053-1066, with the cross-device step in the middle).

torch is plumbing here (device buffers for the collective, process group); the analysis itself runs in librg_engine.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .replaygain import AlbumGainResult, Engine, ReplayGainError, ReplayGainResult, _message

HIST = _lib.HIST_SIZE


def shard_by_bytes(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Partition track indices across `world` ranks by cumulative bytes (not by count: VBR / mixed-length corpora), keeping
    input order inside each shard.  Greedy contiguous split at the byte quantiles; every rank gets a (possibly empty) list."""
    total = float(sum(sizes))
    shards: List[List[int]] = [[] for _ in range(world)]
    acc = 0.0
    for i, s in enumerate(sizes):
        mid = acc + 0.5 * s
        r = min(world - 1, int(world * mid / total)) if total > 0 else i % world
        shards[r].append(i)
        acc += s
    return shards


def album_reduce(hist, peak, group=None):
    """The one collective of the path.  hist: int64 tensor [12000] (u64 counters reinterpreted), peak: float64 tensor [1],
    both resident where the process group's backend wants them (CUDA for NCCL over NVLink, CPU for gloo).  In place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=group)
    return hist, peak


def analyze_album_sharded(engine: Engine, images: Sequence, rank: int, world: int, group=None,
                          track_index: Optional[int] = None) -> Tuple[AlbumGainResult, List[int]]:
    """analyze_album across `world` GPUs.  Every rank passes the SAME list of file images (or at least the same sizes);
    each analyses its byte-balanced shard on its own GPU, the partial album histogram / peak are all-reduced over NCCL,
    and every rank finishes the percentile on its device.  Returns (album result with this rank's tracks, shard indices).
    A failing file aborts the album on every rank (src/replaygain.rs:1055)."""
    import torch
    import torch.distributed as dist
    lib = engine._lib
    mine = shard_by_bytes([len(im) for im in images], world)[rank]
    dev = torch.device("cuda", torch.cuda.current_device())
    d_hist = torch.zeros(HIST, dtype=torch.int64, device=dev)
    d_peak = torch.zeros(1, dtype=torch.float64, device=dev)
    n = len(mine)
    arr, keep = Engine._inputs([images[i] for i in mine], track_index)
    out = (_lib.RgTrackResult * max(n, 1))()
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    cor = C.c_void_p()
    rc = lib.rg_corpus_upload(engine._h, arr, n, C.byref(cor))
    if rc == 0:
        torch.cuda.synchronize()
        rc = lib.rg_analyze_album_shard(engine._h, cor, out, d_hist.data_ptr(), d_peak.data_ptr())
        lib.rg_corpus_free(engine._h, cor)
    status[0] = -rc  # statuses are <= 0: the max over ranks of -rc finds any failure
    if world > 1:
        dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group)
    if int(status.item()) != 0:
        st = -int(status.item())
        raise ReplayGainError(st, _message(st))
    album_reduce(d_hist, d_peak, group)
    torch.cuda.synchronize()   # the collective ran on torch's streams; rg_album_finish reads the buffers on the engine's
    alb = _lib.RgAlbumResult()
    rc = lib.rg_album_finish(engine._h, d_hist.data_ptr(), d_peak.data_ptr(), C.byref(alb))
    if rc != 0:
        raise ReplayGainError(rc, _message(rc))
    tracks = [Engine._result(out[i]) for i in range(n)]
    hist = d_hist.cpu().numpy().view(np.uint64).copy()
    return AlbumGainResult(tracks, alb.album_loudness_db, alb.album_gain_db, alb.album_peak, hist), mine


def gain_analyze_sharded(engine: Engine, images: Sequence, rank: int, world: int):
    """mp3rgain::analyze() (src/lib.rs:470-514) over a file list across `world` GPUs: files are independent, so each rank
    walks its byte-balanced shard (rg_gain_analyze) and there is no collective.  Returns (this rank's RgGainStats list,
    shard indices)."""
    mine = shard_by_bytes([len(im) for im in images], world)[rank]
    return engine.gain_analyze_images([images[i] for i in mine]), mine
