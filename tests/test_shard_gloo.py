"""not gpu: the N>1 path (reads sharded over ranks, index replicated, results concatenated in read order) with two
gloo processes on the CPU.  The per-shard "search" is the oracle here; the plumbing (taxor_b200/shard.py) is what
runs under torchrun on the GPUs."""
import os
import socket

import numpy as np
import pytest

from taxor_b200 import shard


def test_shard_range_partition():
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from oracle.oracle import Oracle
    from tests import helpers as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    ds = H.make_dataset(oracle, n_genomes=12, genome_len=30_000, t_max=4)      # index replicated: same seeds on every rank
    reads = H.make_reads(ds, np.random.default_rng(0).integers(300, 4000, 41), err=0.03)
    codes, off = H.reads_to_codes(reads)
    lo, hi = shard.shard_range(reads.n, rank, world)

    def search(a, b):
        r = oracle.search_batch(oracle.make_hixf(ds.arrays), codes[int(off[a]):int(off[b])], off[a:b + 1] - off[a], k=22, s=12, t=5,
                                use_syncmer=True, window_size=20, error_rate=0.1)
        keep = np.zeros(len(r["raw_ub"]), bool)
        return dict(hash_count=r["hash_count"], threshold=r["threshold"], hit_begin=r["raw_off"], user_bin=r["raw_ub"],
                    count=r["raw_cnt"], keep=keep)

    merged = shard.gather_results(search(lo, hi), dist, dst=0)
    if rank == 0:
        whole = search(0, reads.n)
        ok = all(np.array_equal(merged[k], whole[k]) for k in whole)
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path, built_libs):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
