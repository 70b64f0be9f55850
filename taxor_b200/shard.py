"""Read sharding across ranks (one process per GPU; index replicated; no data-path collective).

`taxor search` reads are independent (src/main/taxor_search.cpp:214 loops over records; the only shared state is
the output mutex, :308-311), so rank r takes the contiguous block shard_range(n, r, world) and rank 0 concatenates
the per-rank results in read order.  The only communication is this final gather of (small) hit lists, done with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous block of reads for `rank`: sizes differ by at most one, earlier ranks get the larger blocks."""
    base, rem = divmod(int(n_reads), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def concat_results(parts):
    """Concatenates per-shard result dicts (hash_count, threshold, hit_begin, user_bin, count, keep) in shard order."""
    hash_count = np.concatenate([p["hash_count"] for p in parts])
    threshold = np.concatenate([p["threshold"] for p in parts])
    user_bin = np.concatenate([p["user_bin"] for p in parts])
    count = np.concatenate([p["count"] for p in parts])
    keep = np.concatenate([p["keep"] for p in parts])
    begins, base = [], 0
    for p in parts:
        hb = np.asarray(p["hit_begin"], dtype=np.uint64)
        begins.append(hb[:-1] + np.uint64(base))
        base += int(hb[-1])
    hit_begin = np.concatenate(begins + [np.array([base], dtype=np.uint64)])
    return dict(hash_count=hash_count, threshold=threshold, hit_begin=hit_begin, user_bin=user_bin, count=count, keep=keep)


def result_to_dict(res) -> dict:
    return dict(hash_count=res.hash_count, threshold=res.threshold, hit_begin=res.hit_begin, user_bin=res.user_bin,
                count=res.count, keep=res.keep)


def gather_results(local: dict, dist, dst: int = 0):
    """All ranks call this; rank `dst` returns the concatenation in rank (= read) order, the others None."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    return concat_results(bucket)
