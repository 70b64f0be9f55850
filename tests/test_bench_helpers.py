"""CPU: host-side helpers of bench.py that both arms rely on (no GPU, no CUDA library calls that need a device)."""
import argparse
import sys

import numpy as np

import bench
from taxor_b200 import capi, tools


def test_unpack_2bit_matches_the_library_layout():
    rng = np.random.default_rng(0)
    for n in (0, 1, 31, 32, 33, 64, 1000, 4097):
        codes = rng.integers(0, 4, n, dtype=np.uint8)
        packed = capi.pack_codes([codes])
        assert np.array_equal(bench.unpack_2bit(packed.words, n), codes)
        if n:
            assert np.array_equal(capi.unpack_codes(packed, 0), codes)


def test_workload_presets_and_overrides(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert (a.k, a.s, a.t, a.use_syncmer, a.genomes, a.genome_len, a.t_max) == (22, 12, 5, True, 1000, 40_000_000, 64)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "kmer", "--genomes", "77"])
    a = bench.parse_args()
    assert (a.k, a.use_syncmer, a.window, a.genomes, a.t_max, a.read_len_range) == (20, False, 20, 77, 128, (1000, 50_000))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "deep"])
    a = bench.parse_args()
    assert (a.genomes, a.t_max, a.use_syncmer) == (20000, 64, True)


def test_cpu_hashing_of_the_reference_arm_equals_the_oracle(oracle, tmp_path):
    args = argparse.Namespace(genomes=6, genome_len=30_000, cache=str(tmp_path), min_genome_len=20_000, k=22, s=12, t=5, use_syncmer=True)
    genomes, lens = bench.make_genomes(args)
    ub, n_seg = bench.hash_genomes_cpu(args, genomes, lens)
    assert n_seg == 0 and len(ub) == 6
    for g in range(6):
        ref = tools.genome(1000 + g, int(lens[g]))
        assert np.array_equal(genomes[g], ref)
        exp = oracle.syncmer_hashes(bench.unpack_2bit(ref, lens[g]), 22, 12, 5)
        assert np.array_equal(np.unique(ub[g]), np.sort(exp))


def test_index_depth_from_cached_arrays(oracle, tmp_path):
    from tests import helpers as H
    ds = H.make_dataset(oracle, n_genomes=70, genome_len=20_000, t_max=4)     # 70 bins under t_max 4: at least 3 levels
    info = dict(hash_s=0, build_s=0, n_ixf=ds.hixf.n_ixf, fp_bytes=ds.hixf.fp_bytes, n_hashes=0, reseeds=0)
    d = str(tmp_path / "ix")
    bench.save_index(ds.hixf, d, info)
    ix = bench.LoadedIndex(d)
    assert ix.n_ixf == ds.hixf.n_ixf and ix.n_user_bins == 70 and ix.depth >= 3
    assert all(np.array_equal(a, b) for a, b in zip(ix.data, ds.hixf.data))


def test_gtdb_preset_and_cache_tag(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "gtdb"])
    a = bench.parse_args()
    assert (a.genomes, a.t_max, a.t_max_lower, a.reads, a.genome_len) == (102_400, 4096, 16, 250_000, 100_000)
    d, _ = bench.index_cache_paths(a)
    assert d.endswith("_tl16")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "gtdb", "--reads", "5000", "--t-max", "1024"])
    a = bench.parse_args()
    assert (a.reads, a.t_max, a.t_max_lower) == (5000, 1024, 16)


def test_at_scale_comparison_flags_every_kind_of_difference(oracle):
    """bench.compare_with_oracle (the parity_at_scale check of every rank): equal answers pass, and a difference in any of hash
    count, threshold, hit offsets, user bins or counts is counted -- exercised with the oracle on both sides"""
    from tests import helpers as H
    ds = H.make_dataset(oracle, n_genomes=20, genome_len=30_000, t_max=4)
    reads = H.make_reads(ds, np.random.default_rng(3).integers(300, 6000, 120), err=0.04)
    codes, off = H.reads_to_codes(reads)
    ora = oracle.search_batch(oracle.make_hixf(ds.arrays), codes, off, k=22, s=12, t=5, use_syncmer=True, window_size=20, error_rate=0.1)

    class G:                                                   # what capi.SearchResult exposes, built from the oracle's raw output
        pass
    g = G()
    g.hash_count, g.threshold = ora["hash_count"].copy(), ora["threshold"].copy()
    g.hit_begin, g.user_bin, g.count = ora["raw_off"].copy(), ora["raw_ub"].copy(), ora["raw_cnt"].copy()
    keep = np.zeros(len(g.count), bool)
    for r in range(reads.n):
        a, b = int(g.hit_begin[r]), int(g.hit_begin[r + 1])
        if b > a:
            keep[a:b] = ~(g.count[a:b].astype(np.float64) < float(g.count[a:b].max()) * 0.8)
    g.keep = keep
    n = reads.n
    assert bench.compare_with_oracle(g, ora, n)["ok"]
    assert len(ora["ub"]) > 20
    for field in ("hash_count", "threshold", "count", "user_bin"):
        saved = getattr(g, field).copy()
        getattr(g, field)[int(np.flatnonzero(keep)[0]) if field in ("count", "user_bin") else 5] += 1
        p = bench.compare_with_oracle(g, ora, n)
        assert not p["ok"] and p["mismatches"] >= 1, field
        setattr(g, field, saved)
    blk = bench.random_access_block({"query_ms": 10.0, "query_bytes": 200 * 1_000_000}, None, type("I", (), {"tbins": np.array([64])})(), None)
    assert abs(blk["achieved_G_rows_per_s"] - 3 * 1_000_000 / 0.010 / 1e9) < 1e-9
