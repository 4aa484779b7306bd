"""Multi-GPU launcher: one process per GPU, tracks sharded by bytes.  Track mode has no data-path collective; album mode has
exactly one exchange, and the ENGINE owns it (rg_analyze_album_sharded: all-reduce(sum) of the 12000-bin histogram +
all-reduce(max) of {peak, error key} as one NCCL group on the engine's stream -- the accumulate loop of the reference,
src/replaygain.rs:1053-1066, with the cross-device step in the middle).  What is left here is the launch protocol: split
the list, hand rank 0's NCCL unique id to the other ranks, call the engine.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .replaygain import AlbumGainResult, Engine


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


def shard_offsets(shards: Sequence[Sequence[int]]) -> List[int]:
    """Global index of each shard's first file (shards are contiguous and in input order)."""
    offs, acc = [], 0
    for s in shards:
        offs.append(acc)
        acc += len(s)
    return offs


def connect(engine: Engine, rank: int, world: int, group=None, transport: str = "nccl"):
    """Make `engine` rank `rank` of `world`.  transport "nccl": the engine creates its own communicator (rank 0's unique id
    is broadcast over the torch.distributed group, any backend); transport "host": the exchange goes through
    torch.distributed on host tensors (gloo: the CPU tests, or a host without NCCL)."""
    if world <= 1:
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    if transport == "host":
        def xchg(sum_u64, max_u64):
            s = torch.from_numpy(sum_u64.view(np.int64))     # counters stay far below 2^63
            m = torch.from_numpy(max_u64.view(np.int64))     # non-negative doubles and error keys (< 2^63) order like int64
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        engine.comm_attach_host(rank, world, xchg)
        return
    box = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(rank, world, box[0])


def analyze_album_sharded(engine: Engine, images: Sequence, rank: int, world: int, track_index: Optional[int] = None,
                          names: Optional[Sequence] = None) -> Tuple[AlbumGainResult, List[int]]:
    """analyze_album across `world` GPUs (call connect() once before).  Every rank passes the SAME list of file images (or at
    least the same sizes); each analyses its byte-balanced shard on its own GPU.  Returns (album result with this rank's
    tracks, shard indices).  A failing file aborts the album on every rank with that file's error (src/replaygain.rs:1055)."""
    shards = shard_by_bytes([len(im) for im in images], world)
    mine = shards[rank]
    return engine.analyze_album_sharded_images([images[i] for i in mine], shard_offsets(shards)[rank], track_index,
                                               names=names), mine


def gain_analyze_sharded(engine: Engine, images: Sequence, rank: int, world: int):
    """mp3rgain::analyze() (src/lib.rs:470-514) over a file list across `world` GPUs: files are independent, so each rank
    walks its byte-balanced shard (rg_gain_analyze) and there is no collective.  Returns (this rank's RgGainStats list,
    shard indices)."""
    mine = shard_by_bytes([len(im) for im in images], world)[rank]
    return engine.gain_analyze_images([images[i] for i in mine]), mine
