"""Read sharding over the GPUs of one box.

Reads are independent (the reference loops over them one at a time, basecall.py:70-123), so the
data path has no collective: every rank decodes its own reads against its own table replica and
only the decoded strings are gathered on the host (the FASTA writer, basecall.py:129-141).
"""
from __future__ import annotations

import heapq

import numpy as np


def lpt_shards(frame_counts, world: int):
    """Longest-processing-time-first assignment of reads to `world` ranks.

    Decode time is proportional to the frame count, so reads are dealt longest first to the
    currently least loaded rank.  Returns a list of `world` int64 index arrays, each in
    descending length order (the order the device work queue wants).  Deterministic.
    """
    fc = np.asarray(frame_counts, dtype=np.int64)
    order = np.argsort(-fc, kind="stable")
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    out = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(int(i))
        heapq.heappush(heap, (load + int(fc[i]), r))
    return [np.asarray(x, dtype=np.int64) for x in out]


def shard_for_rank(frame_counts, rank: int, world: int) -> np.ndarray:
    return lpt_shards(frame_counts, world)[rank]


def gather_strings(local: dict, n_total: int, group=None, dst: int = 0):
    """Host-side gather of {read index: sequence} dicts onto rank `dst`, returned in read order
    (None on the other ranks).  Works on any torch.distributed backend (object gather)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return [local[i] for i in range(n_total)]
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for d in bucket:
        merged.update(d)
    missing = [i for i in range(n_total) if i not in merged]
    if missing:
        raise RuntimeError(f"reads {missing[:5]}... were decoded by no rank")
    return [merged[i] for i in range(n_total)]


def decode_sharded(mats, beam_width, lm, s_threshold, r_threshold, len_context, group=None,
                   decode_fn=None):
    """Decode a list of posterior matrices known to every rank: each rank takes its LPT shard,
    decodes it on its own GPU (`decode.beam_search_batch` unless `decode_fn` is given) and rank 0
    receives all sequences in input order."""
    import torch.distributed as dist

    if decode_fn is None:
        from .decode import beam_search_batch as decode_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_for_rank([len(m) for m in mats], rank, world)
    seqs = decode_fn([mats[i] for i in mine], beam_width, lm, s_threshold, r_threshold, len_context)
    return gather_strings({int(i): s for i, s in zip(mine, seqs)}, len(mats), group)
