"""-m gpu: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from tests import helpers as H
from taxor_b200 import capi

pytestmark = pytest.mark.gpu


def lowcomplexity_reads(rng):
    """Tie-heavy inputs for the syncmer scan: homopolymers, short tandem repeats, A/T-only, palindromes."""
    seqs = []
    seqs.append(np.zeros(5000, np.uint8))                                   # poly-A
    seqs.append(np.full(3000, 3, np.uint8))                                 # poly-T
    for period in (1, 2, 3, 4, 5, 6, 7, 11, 12, 13):
        unit = rng.integers(0, 4, period, dtype=np.uint8)
        x = np.resize(unit, 4000).copy()
        idx = rng.integers(0, len(x), 12)
        x[idx] = rng.integers(0, 4, len(idx))
        seqs.append(x)
    seqs.append((rng.integers(0, 2, 6000) * 3).astype(np.uint8))            # A/T only
    half = rng.integers(0, 4, 1500, dtype=np.uint8)
    seqs.append(np.concatenate([half, (3 - half)[::-1]]))                   # reverse-complement palindrome
    r = rng.integers(0, 4, 8000, dtype=np.uint8)
    r[2000:2100] = 0
    r[5000:5300] = np.resize(np.array([0, 1], np.uint8), 300)
    seqs.append(r)                                                          # random with embedded repeats
    big = rng.integers(0, 4, 40000, dtype=np.uint8)
    big[1000:9000] = 2                                                      # 8 kb homopolymer across many tiles
    seqs.append(big)
    return seqs


def edge_reads(rng, k):
    seqs = [np.zeros(0, np.uint8), rng.integers(0, 4, 1, dtype=np.uint8), rng.integers(0, 4, k - 1, dtype=np.uint8),
            rng.integers(0, 4, k, dtype=np.uint8), rng.integers(0, 4, k + 1, dtype=np.uint8)]
    for n in (31, 32, 33, 63, 64, 65, 1023 + k - 1, 1024 + k - 1, 1025 + k - 1, 2048 + k, 3000):
        seqs.append(rng.integers(0, 4, n, dtype=np.uint8))
    return seqs


@pytest.mark.parametrize("k,s", [(22, 12), (20, 10), (24, 12), (30, 16), (16, 8), (21, 11), (28, 20), (12, 11)])
def test_syncmer_hash_parity(ctx, oracle, k, s):
    """kernel #1 (fast template or generic fallback) == oracle: raw emission multiset and distinct set per read."""
    t = oracle.t_syncmer(k, s)
    ctx.set_params(k=k, s=s, t=t, use_syncmer=True, window_size=20)
    rng = np.random.default_rng(k * 100 + s)
    seqs = edge_reads(rng, k) + lowcomplexity_reads(rng)
    seqs += [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(500, 12000, 40)]
    reads = capi.pack_codes(seqs)
    off_raw, h_raw = ctx.hash_batch(reads, dedup=False)
    off, h = ctx.hash_batch(reads, dedup=True)
    for i, c in enumerate(seqs):
        exp_raw = oracle.syncmer_hashes_raw(c, k, s, t)
        got_raw = h_raw[int(off_raw[i]):int(off_raw[i + 1])]
        assert np.array_equal(np.sort(got_raw), np.sort(exp_raw)), (i, len(c))
        exp = oracle.syncmer_hashes(c, k, s, t)
        got = h[int(off[i]):int(off[i + 1])]
        assert len(got) == len(exp) and np.array_equal(np.sort(got), np.sort(exp)), (i, len(c))


def test_syncmer_t_variants(ctx, oracle):
    """t is read from the index file, not derived: every legal offset must work (generic kernel)."""
    rng = np.random.default_rng(5)
    seqs = [rng.integers(0, 4, 3000, dtype=np.uint8) for _ in range(6)] + lowcomplexity_reads(rng)[:6]
    reads = capi.pack_codes(seqs)
    for t in (1, 3, 6, 11):
        ctx.set_params(k=22, s=12, t=t, use_syncmer=True, window_size=20)
        off, h = ctx.hash_batch(reads, dedup=True)
        for i, c in enumerate(seqs):
            exp = oracle.syncmer_hashes(c, 22, 12, t)
            assert np.array_equal(np.sort(h[int(off[i]):int(off[i + 1])]), np.sort(exp)), (t, i)


@pytest.mark.parametrize("k", [20, 22, 31, 32, 8])
def test_kmer_hash_parity(ctx, oracle, k):
    """canonical k-mer mode: same values in the same (position) order, duplicates kept."""
    ctx.set_params(k=k, use_syncmer=False, window_size=k)
    rng = np.random.default_rng(k)
    seqs = edge_reads(rng, k) + lowcomplexity_reads(rng)[:4] + [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(100, 9000, 20)]
    reads = capi.pack_codes(seqs)
    off, h = ctx.hash_batch(reads, dedup=True)
    for i, c in enumerate(seqs):
        exp = oracle.kmer_hashes(c, k)
        assert np.array_equal(h[int(off[i]):int(off[i + 1])], exp), (i, len(c))


def test_scaling_filter(ctx, oracle):
    rng = np.random.default_rng(3)
    seqs = [rng.integers(0, 4, 6000, dtype=np.uint8) for _ in range(8)]
    reads = capi.pack_codes(seqs)
    for scaling in (10, 100):
        ctx.set_params(k=22, s=12, t=5, use_syncmer=True, window_size=20, scaling=scaling)
        off, h = ctx.hash_batch(reads, dedup=True)
        for i, c in enumerate(seqs):
            exp = [x for x in oracle.syncmer_hashes(c, 22, 12, 5).tolist() if oracle.scaling_keep(x, scaling)]
            assert np.array_equal(np.sort(h[int(off[i]):int(off[i + 1])]), np.sort(np.array(exp, np.uint64)))
        ctx.set_params(k=20, use_syncmer=False, window_size=20, scaling=scaling)
        off, h = ctx.hash_batch(reads, dedup=True)
        for i, c in enumerate(seqs):
            exp = [x for x in oracle.kmer_hashes(c, 20).tolist() if oracle.scaling_keep(x, scaling)]
            assert np.array_equal(h[int(off[i]):int(off[i + 1])], np.array(exp, np.uint64))


def test_long_sequence_hash(ctx, oracle):
    """a genome-sized input (many tiles, global-memory dedup class)"""
    from taxor_b200 import tools
    n = 700_000
    g = tools.genome(77, n)
    reads = capi.PackedReads(g, np.zeros(1, np.uint64), np.array([n], np.uint32))
    ctx.set_params(k=22, s=12, t=5, use_syncmer=True, window_size=20)
    off, h = ctx.hash_batch(reads, dedup=True)
    exp = oracle.syncmer_hashes(H.codes_of(g, n), 22, 12, 5)
    assert np.array_equal(np.sort(h), np.sort(exp))


@pytest.mark.parametrize("t_max", [8, 64])
def test_bulk_count_parity(ctx, oracle, t_max):
    """kernel #2's probe/count against bulk_count of the oracle, every IXF of a small hierarchy."""
    ds = H.make_dataset(oracle, n_genomes=20, genome_len=40_000, t_max=t_max)
    H.upload(ctx, ds)
    oh = oracle.make_hixf(ds.arrays)
    rng = np.random.default_rng(0)
    member = np.concatenate([ds.hixf._ub[3][:500], ds.hixf._ub[7][:300]])
    values = np.concatenate([member, rng.integers(0, 2**63, 2000, dtype=np.uint64)])
    for x in range(ds.hixf.n_ixf):
        got = ctx.ixf_bulk_count(x, values, int(ds.hixf.bins[x]))
        exp = oracle.ixf_bulk_count(oh, x, values)
        assert np.array_equal(got, exp), x
    assert np.array_equal(ctx.ixf_bulk_count(0, np.zeros(0, np.uint64), int(ds.hixf.bins[0])), np.zeros(int(ds.hixf.bins[0]), np.uint32))


def _search_case(ctx, oracle, ds, reads, **par):
    H.upload(ctx, ds)
    ctx.set_params(k=ds.k, s=ds.s, t=ds.t, use_syncmer=ds.use_syncmer, window_size=par.get("window_size", 20),
                   scaling=par.get("scaling", 1), percentage=par.get("percentage", -1.0), error_rate=par.get("error_rate", 0.1))
    res = ctx.search(reads)
    codes, off = H.reads_to_codes(reads)
    ora = oracle.search_batch(oracle.make_hixf(ds.arrays), codes, off, k=ds.k, s=ds.s, t=ds.t, use_syncmer=ds.use_syncmer,
                              window_size=par.get("window_size", 20), scaling=par.get("scaling", 1),
                              percentage=par.get("percentage", -1.0), error_rate=par.get("error_rate", 0.1))
    H.assert_same_search(res, ora, reads.n)
    return res, ora


@pytest.mark.parametrize("t_max,n_genomes", [(8, 40), (64, 100), (4, 70)])
def test_search_parity_syncmer(ctx, oracle, t_max, n_genomes):
    """whole hot path == oracle: hash counts, thresholds, per-(read,user bin) counts, DFS order, 0.8*max filter."""
    ds = H.make_dataset(oracle, n_genomes=n_genomes, genome_len=50_000, t_max=t_max)
    rng = np.random.default_rng(1)
    lengths = rng.integers(300, 12000, 300)
    reads = H.make_reads(ds, lengths, err=0.05)
    res, ora = _search_case(ctx, oracle, ds, reads, error_rate=0.1)
    assert int(res.hit_begin[-1]) > 50          # the workload really descends and reports
    assert ds.hixf.n_ixf > 1
    # resident path gives the same answer
    h = ctx.upload_reads(reads)
    res2 = ctx.search_resident(h, fetch=True)
    ctx.free_reads(h)
    H.assert_same_search(res2, ora, reads.n)


def test_early_exit_is_exact_and_taken(ctx, oracle):
    """reads that match nothing stop probing once no bin can reach the threshold: same output, fewer probes"""
    ds = H.make_dataset(oracle, n_genomes=100, genome_len=30_000, t_max=64)   # > t_max genomes: no split bins in the root
    rng = np.random.default_rng(31)
    foreign = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(2000, 12000, 150)]   # not in the index
    own = H.make_reads(ds, rng.integers(2000, 12000, 150), err=0.03)
    seqs = foreign + [capi.unpack_codes(own, i) for i in range(own.n)]
    order = rng.permutation(len(seqs))
    reads = capi.pack_codes([seqs[i] for i in order])
    for er in (0.05, 0.1):
        res, ora = _search_case(ctx, oracle, ds, reads, error_rate=er)
        tm = ctx.timing()
        assert int(res.hit_begin[-1]) > 50
        # a foreign read stops after about (1 - ratio) of its ~L/11 hashes: ratio 0.44 / 0.23 for these error rates
        assert tm["skipped_hashes"] > 0.1 * sum(len(x) for x in foreign) / 11, tm


def test_search_threshold_zero_flood(ctx, oracle):
    """H5: reads shorter than k have hash_count 0 -> threshold 0 -> every user bin reported, every merged bin descended."""
    ds = H.make_dataset(oracle, n_genomes=30, genome_len=30_000, t_max=4)
    rng = np.random.default_rng(2)
    seqs = [rng.integers(0, 4, n, dtype=np.uint8) for n in (0, 5, 21, 22, 40, 500)]
    reads = capi.pack_codes(seqs)
    res, ora = _search_case(ctx, oracle, ds, reads, error_rate=0.05)
    assert len(res.hits(0)[0]) == 30 and len(res.hits(1)[0]) == 30


def test_search_percentage_and_batching(ctx, oracle):
    ds = H.make_dataset(oracle, n_genomes=50, genome_len=40_000, t_max=8)
    rng = np.random.default_rng(4)
    reads = H.make_reads(ds, rng.integers(1000, 9000, 500), err=0.03)
    ctx.configure(max_batch_reads=64, max_batch_bases=200_000, n_slots=3)   # many small batches through 3 slots
    try:
        _search_case(ctx, oracle, ds, reads, percentage=0.2)
        ctx.configure(max_batch_reads=64, max_batch_bases=200_000, n_slots=1)
        _search_case(ctx, oracle, ds, reads, percentage=0.05)
    finally:
        ctx.configure()


def test_search_parity_kmer_mode(ctx, oracle):
    """config-4 shape at test size: canonical 20-mers, no dedup, k-mer-model threshold."""
    ds = H.make_dataset(oracle, n_genomes=24, genome_len=60_000, k=20, s=0, t=0, use_syncmer=False, t_max=8)
    rng = np.random.default_rng(6)
    lengths = np.exp(rng.uniform(np.log(1000), np.log(20000), 60)).astype(np.int64)
    reads = H.make_reads(ds, lengths, err=0.02)
    res, ora = _search_case(ctx, oracle, ds, reads, window_size=20, error_rate=0.02)
    assert int(res.hit_begin[-1]) > 10


def test_search_large_rows(ctx, oracle):
    """IXFs wider than 512 technical bins take the CTA-per-item kernel (root T up to 4096 in real layouts)."""
    ds = H.make_dataset(oracle, n_genomes=700, genome_len=3_000, t_max=640, size_jitter=True)
    assert int(ds.hixf.tbins.max()) > 512
    rng = np.random.default_rng(8)
    reads = H.make_reads(ds, rng.integers(400, 1200, 120), err=0.02)
    res, ora = _search_case(ctx, oracle, ds, reads, error_rate=0.1)
    assert int(res.hit_begin[-1]) > 20
    ds2 = H.make_dataset(oracle, n_genomes=1500, genome_len=2_500, t_max=576, size_jitter=True, seed=5000)
    assert ds2.hixf.n_ixf > 1 and int(ds2.hixf.tbins.max()) > 512
    reads2 = H.make_reads(ds2, rng.integers(400, 1000, 100), err=0.02)
    _search_case(ctx, oracle, ds2, reads2, error_rate=0.1)


@pytest.mark.parametrize("t_max,n_genomes,early", [(1024, 1300, "1"), (2048, 2300, "1"), (4096, 4300, "1"), (4096, 4300, "0"), (1536, 1536, "1")])
def test_search_wide_rows(monkeypatch, oracle, t_max, n_genomes, early):
    """GTDB-shaped upper levels (root t_max up to 4096, taxor_build.cpp:173-187): every warp of the CTA probes a share of
    the hash list over whole rows; wide rows in 2 KB passes; exact early exit between hash blocks"""
    monkeypatch.setenv("TXR_EARLY_EXIT", early)
    c = capi.Context(0)
    try:
        ds = H.make_dataset(oracle, n_genomes=n_genomes, genome_len=2_000, t_max=t_max, size_jitter=True, seed=9000 + t_max)
        assert int(ds.hixf.tbins[0]) == t_max
        rng = np.random.default_rng(t_max)
        own = H.make_reads(ds, rng.integers(300, 800, 80), err=0.02)
        foreign = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(300, 6000, 60)]
        seqs = foreign + [capi.unpack_codes(own, i) for i in range(own.n)] + [np.zeros(0, np.uint8), np.zeros(21, np.uint8)]
        reads = capi.pack_codes([seqs[i] for i in rng.permutation(len(seqs))])
        res, ora = _search_case(c, oracle, ds, reads, error_rate=0.1)
        assert int(res.hit_begin[-1]) > 20
        tm = c.timing()
        if early == "1":
            assert tm["skipped_hashes"] > 0, tm
        else:
            assert tm["skipped_hashes"] == 0, tm
        _search_case(c, oracle, ds, reads, percentage=0.5)
    finally:
        c.close()


@pytest.mark.parametrize("env", [{"TXR_FUSE_DEDUP": "1", "TXR_FUSE_MAX_KEYS": "64"}, {"TXR_FUSE_DEDUP": "1", "TXR_FUSE_MAX_KEYS": "300"},
                                 {"TXR_FUSE_DEDUP": "1"}, {}])
@pytest.mark.parametrize("scaling", [1, 5])
def test_fused_distinct_set_variants(monkeypatch, oracle, env, scaling):
    """the distinct set built inside the syncmer kernel: the hand-over of a read to the CTA-per-read kernel (forced early
    by TXR_FUSE_MAX_KEYS), the separate dedup kernel (TXR_FUSE_DEDUP=0) and the default all give the oracle's sets; reads
    made of repeated blocks have most of their hashes more than once"""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    c = capi.Context(0)
    try:
        k, s, t = 22, 12, 5
        c.set_params(k=k, s=s, t=t, use_syncmer=True, window_size=20, scaling=scaling)
        rng = np.random.default_rng(77 + scaling)
        seqs = edge_reads(rng, k) + lowcomplexity_reads(rng)
        seqs += [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(500, 10200, 60)]
        for _ in range(10):                                               # repeated blocks: duplicates inside and across tiles
            block = rng.integers(0, 4, int(rng.integers(50, 1500)), dtype=np.uint8)
            seqs.append(np.resize(block, int(rng.integers(3000, 10000))).copy())
        reads = capi.pack_codes(seqs)
        off, h = c.hash_batch(reads, dedup=True)
        for i, x in enumerate(seqs):
            exp = [v for v in oracle.syncmer_hashes(x, k, s, t).tolist() if scaling == 1 or oracle.scaling_keep(v, scaling)]
            got = h[int(off[i]):int(off[i + 1])]
            assert len(got) == len(exp) and np.array_equal(np.sort(got), np.sort(np.array(exp, np.uint64))), (i, len(x), env)
    finally:
        c.close()


@pytest.mark.parametrize("k,w", [(20, 24), (20, 40), (8, 12), (31, 126), (16, 17), (4, 8)])
def test_minimiser_hash_parity(ctx, oracle, k, w):
    """minimiser windows (window_size > k): same values, same order, same repeats as views::minimiser_hash; the
    low-complexity inputs are all ties (sequential hand-over between lanes), small k makes ties common everywhere."""
    ctx.set_params(k=k, use_syncmer=False, window_size=w)
    rng = np.random.default_rng(k * 131 + w)
    seqs = edge_reads(rng, k) + lowcomplexity_reads(rng)
    seqs += [rng.integers(0, 4, n, dtype=np.uint8) for n in (w - 1, w, w + 1, w + 30, 1023 + w - 1, 1024 + w - 1, 1025 + w - 1, 2048 + w)]
    seqs += [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(100, 12000, 30)]
    reads = capi.pack_codes(seqs)
    off, h = ctx.hash_batch(reads, dedup=True)
    for i, c in enumerate(seqs):
        exp = oracle.minimiser_hashes(c, k, w)
        got = h[int(off[i]):int(off[i + 1])]
        assert len(got) == len(exp) and np.array_equal(got, exp), (i, len(c), len(got), len(exp))


@pytest.mark.parametrize("k,w,scaling", [(20, 32, 1), (20, 24, 10)])
def test_search_parity_minimiser_mode(ctx, oracle, k, w, scaling):
    """minimiser index end to end: FracMinHash threshold per read (hash_count AND read length), duplicates kept."""
    ds = H.make_dataset(oracle, n_genomes=24, genome_len=60_000, k=k, s=0, t=0, use_syncmer=False, t_max=8, window_size=w,
                        scaling=scaling)
    rng = np.random.default_rng(16)
    lengths = np.concatenate([np.exp(rng.uniform(np.log(300), np.log(20000), 80)).astype(np.int64), [0, 5, k - 1, k, k + 1, w - 1, w, w + 1]])
    reads = H.make_reads(ds, np.maximum(lengths, 1), err=0.02)
    res, ora = _search_case(ctx, oracle, ds, reads, window_size=w, error_rate=0.02, scaling=scaling)
    assert int(res.hit_begin[-1]) > 10
    ctx.configure(max_batch_reads=16, max_batch_bases=100_000, n_slots=3)     # several batches through the host round trip
    try:
        _search_case(ctx, oracle, ds, reads, window_size=w, error_rate=0.02, scaling=scaling)
        _search_case(ctx, oracle, ds, reads, window_size=w, percentage=0.3, scaling=scaling)
    finally:
        ctx.configure()


@pytest.fixture()
def ctx_partitioned(monkeypatch):
    """a context that takes the slot-partitioned root level whatever the index size (auto mode needs a >=48 MB segment)"""
    monkeypatch.setenv("TXR_ROOT_PARTITION", "2")
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("t_max,n_genomes", [(8, 40), (64, 100), (128, 300)])
def test_search_parity_partitioned_root(ctx_partitioned, oracle, t_max, n_genomes):
    """root level through hist -> scatter -> probe -> scan (16-bit global counters): same hits, counts, order."""
    ds = H.make_dataset(oracle, n_genomes=n_genomes, genome_len=40_000, t_max=t_max)
    rng = np.random.default_rng(21)
    lengths = np.concatenate([rng.integers(300, 12000, 400), [0, 1, 21, 22, 23]])
    reads = H.make_reads(ds, np.maximum(lengths, 1), err=0.05)
    res, ora = _search_case(ctx_partitioned, oracle, ds, reads, error_rate=0.1)
    assert int(res.hit_begin[-1]) > 50
    tm = ctx_partitioned.timing()
    assert tm["query_launches"] >= 5                 # the five launches of the partitioned root were taken
    ctx_partitioned.configure(max_batch_reads=37, max_batch_bases=300_000, n_slots=3)
    _search_case(ctx_partitioned, oracle, ds, reads, error_rate=0.05)
    _search_case(ctx_partitioned, oracle, ds, reads, percentage=0.1)


@pytest.mark.parametrize("env", [{"TXR_OVERLAP": "1"}, {"TXR_OVERLAP": "1", "TXR_SM_SPLIT": "4:1"}, {"TXR_OVERLAP": "1", "TXR_SM_SPLIT": "2:1"}])
def test_search_parity_overlap_modes(monkeypatch, oracle, env):
    """hash/dedup of batch i+1 beside the probes of batch i (shared SMs, or disjoint SMs through the SM filter):
    many small batches through 3 slots, same answers; k-mer mode too (kmer_kernel is filtered as well)"""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    c = capi.Context(0)
    try:
        ds = H.make_dataset(oracle, n_genomes=80, genome_len=40_000, t_max=16)
        rng = np.random.default_rng(41)
        reads = H.make_reads(ds, rng.integers(300, 9000, 600), err=0.04)
        c.configure(max_batch_reads=50, max_batch_bases=300_000, n_slots=3)
        res, ora = _search_case(c, oracle, ds, reads, error_rate=0.1)
        assert int(res.hit_begin[-1]) > 100
        ds2 = H.make_dataset(oracle, n_genomes=24, genome_len=60_000, k=20, s=0, t=0, use_syncmer=False, t_max=8)
        reads2 = H.make_reads(ds2, rng.integers(500, 6000, 120), err=0.02)
        _search_case(c, oracle, ds2, reads2, window_size=20, error_rate=0.02)
    finally:
        c.close()


@pytest.mark.parametrize("mode", ["syncmer", "syncmer_generic", "kmer", "minimiser", "syncmer_scaled"])
def test_hash_user_bins_build_side(monkeypatch, oracle, mode):
    """compute_hashes of `taxor build` on the GPU: per-user-bin distinct hashes == union of the oracle's per-sequence
    sets, with the sequences cut into many independently hashed segments (every ~1-2 k windows here).  The low-complexity
    inserts put tie runs (homopolymers, tandem repeats, palindromes) right where cuts are attempted: a cut may only be
    taken at a window whose state does not depend on history."""
    monkeypatch.setenv("TXR_SEGMENT_WINDOWS", "1024")
    c = capi.Context(0)
    try:
        par = dict(syncmer=dict(k=22, s=12, t=5, use_syncmer=True, window_size=20),
                   syncmer_generic=dict(k=21, s=11, t=oracle.t_syncmer(21, 11), use_syncmer=True, window_size=20),
                   kmer=dict(k=20, use_syncmer=False, window_size=20),
                   minimiser=dict(k=20, use_syncmer=False, window_size=31),
                   syncmer_scaled=dict(k=22, s=12, t=5, use_syncmer=True, window_size=20, scaling=10))[mode]
        c.set_params(**par)
        rng = np.random.default_rng(sum(mode.encode()))
        seqs, seq_bin = [], []
        n_bins = 9
        for b in range(n_bins):
            if b == 4:
                continue                                               # a user bin without sequences
            for _ in range(int(rng.integers(1, 4))):
                x = rng.integers(0, 4, int(rng.integers(3000, 60000)), dtype=np.uint8)
                for _ in range(int(rng.integers(0, 12))):              # low-complexity inserts at random places
                    at, ln = int(rng.integers(0, len(x) - 2500)), int(rng.integers(30, 2500))
                    kind = rng.integers(0, 3)
                    if kind == 0:
                        x[at:at + ln] = rng.integers(0, 4)
                    elif kind == 1:
                        x[at:at + ln] = np.resize(rng.integers(0, 4, int(rng.integers(2, 7)), dtype=np.uint8), ln)
                    else:
                        half = x[at:at + ln // 2]
                        x[at + ln // 2:at + 2 * (ln // 2)] = (3 - half)[::-1]
                seqs.append(x)
                seq_bin.append(b)
        seqs += [rng.integers(0, 4, n, dtype=np.uint8) for n in (0, 5, 21, 22, 40)]   # short tails in the last bin
        seq_bin += [n_bins - 1] * 5
        packed = capi.pack_codes(seqs)
        off, h, n_seg = c.hash_user_bins(packed, seq_bin, n_bins)
        assert n_seg > 3 * len(seqs)                                    # the sequences really were cut
        k = par["k"]
        for b in range(n_bins):
            exp = set()
            for x, sb in zip(seqs, seq_bin):
                if sb != b:
                    continue
                if par["use_syncmer"]:
                    v = oracle.syncmer_hashes(x, k, par["s"], par["t"])
                elif par["window_size"] > k:
                    v = oracle.minimiser_hashes(x, k, par["window_size"])
                else:
                    v = oracle.kmer_hashes(x, k)
                exp.update(int(t) for t in v.tolist() if oracle.scaling_keep(int(t), par.get("scaling", 1)))
            got = h[int(off[b]):int(off[b + 1])]
            assert len(got) == len(set(got.tolist())), b               # distinct
            assert set(got.tolist()) == exp, (mode, b, len(got), len(exp))
        assert off[4] == off[5]
    finally:
        c.close()


@pytest.mark.parametrize("scheme", [(1, 0, 0, 0, 0), (0, 1, 1, 13, 37), (1, 1, 2, 0, 0)])
@pytest.mark.parametrize("t_max", [16, 1024])
def test_ixf_scheme_variants(oracle, scheme, t_max):
    """the probe arithmetic is a descriptor chosen at txr_index_upload (the SeqAn3 fork that defines it is absent): binary-fuse
    slots and the seed-mix / fingerprint / rotation variants through the small and the wide kernels, bulk_count and the
    whole search, against the oracle's independently written twin switched to the same scheme; bin-major host arrays too"""
    c = capi.Context(0)
    try:
        oracle.set_ixf_scheme(scheme)
        n_genomes = 60 if t_max == 16 else 1100
        ds = H.make_dataset(oracle, n_genomes=n_genomes, genome_len=30_000 if t_max == 16 else 2_000, t_max=t_max, scheme=scheme,
                            seed=4000 + t_max)
        H.upload(c, ds)
        oh = oracle.make_hixf(ds.arrays)
        rng = np.random.default_rng(9)
        values = np.concatenate([ds.hixf._ub[3][:400], ds.hixf._ub[11][:300], rng.integers(0, 2**63, 1500, dtype=np.uint64)])
        for x in range(min(ds.hixf.n_ixf, 4)):
            assert np.array_equal(c.ixf_bulk_count(x, values, int(ds.hixf.bins[x])), oracle.ixf_bulk_count(oh, x, values)), x
        reads = H.make_reads(ds, rng.integers(300, 800 if t_max > 16 else 9000, 150), err=0.03)
        res, ora = _search_case(c, oracle, ds, reads, error_rate=0.1)
        assert int(res.hit_begin[-1]) > 30
        # the same index handed over bin-major (one plain filter after the other): re-laid-out on upload
        hx = ds.hixf
        data = [np.ascontiguousarray(hx.data[i].reshape(int(hx.rows[i]), int(hx.tbins[i]))[:, : int(hx.bins[i])].T).reshape(-1)
                for i in range(hx.n_ixf)]
        c.upload_index(hx.seed, hx.bins, hx.bins, hx.seg_len, data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.n_user_bins,
                       rows=hx.rows, scheme=scheme, layout=1)
        c.set_params(k=ds.k, s=ds.s, t=ds.t, use_syncmer=True, window_size=20, error_rate=0.1)
        H.assert_same_search(c.search(reads), ora, reads.n)
        # the wrong scheme on the same arrays: rejected by the geometry check or simply different counts, never a crash
        wrong = (1 - scheme[0], scheme[1], scheme[2], 21, 42)
        try:
            c.upload_index(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.n_user_bins,
                           rows=hx.rows, scheme=wrong)
        except capi.TaxorError as e:
            assert "geometry" in str(e)
    finally:
        oracle.set_ixf_scheme(None)
        c.close()


def test_index_clone_and_staged_upload(monkeypatch, oracle):
    """txr_index_clone (device-to-device replica; second GPU when there is one, else a second context on GPU 0) and the
    serial upload path (TXR_UPLOAD_THREADS=1) answer exactly like the staged multi-threaded upload; rows that need padding
    (tbins not a multiple of 64 in the source) take the re-striding path of both"""
    import torch
    ds = H.make_dataset(oracle, n_genomes=300, genome_len=30_000, t_max=128)
    rng = np.random.default_rng(5)
    reads = H.make_reads(ds, rng.integers(300, 9000, 300), err=0.04)
    a = capi.Context(0)
    b = capi.Context(1 if torch.cuda.device_count() > 1 else 0)
    monkeypatch.setenv("TXR_UPLOAD_THREADS", "1")
    c = capi.Context(0)
    try:
        res, ora = _search_case(a, oracle, ds, reads, error_rate=0.1)
        assert int(res.hit_begin[-1]) > 50
        b.clone_index_from(a)
        b.set_params(k=ds.k, s=ds.s, t=ds.t, use_syncmer=True, window_size=20, error_rate=0.1)
        H.assert_same_search(b.search(reads), ora, reads.n)
        _search_case(c, oracle, ds, reads, error_rate=0.1)
        # unpadded source rows (tbins == bins, not a multiple of 64): both upload paths pad to 64-byte rows
        hx = ds.hixf
        data, tb = [], []
        for i in range(hx.n_ixf):
            rows = 3 * int(hx.seg_len[i])
            m = hx.data[i].reshape(rows, int(hx.tbins[i]))[:, : int(hx.bins[i])]
            data.append(np.ascontiguousarray(m).reshape(-1))
            tb.append(int(hx.bins[i]))
        for cx in (a, c):
            cx.upload_index(hx.seed, hx.bins, np.array(tb, np.uint64), hx.seg_len, data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.n_user_bins)
            cx.set_params(k=ds.k, s=ds.s, t=ds.t, use_syncmer=True, window_size=20, error_rate=0.1)
            H.assert_same_search(cx.search(reads), ora, reads.n)
    finally:
        a.close()
        b.close()
        c.close()


def test_errors_are_loud(ctx):
    with pytest.raises(capi.TaxorError):
        ctx.set_params(k=20, use_syncmer=False, window_size=19)       # window smaller than k
    with pytest.raises(capi.TaxorError):
        ctx.set_params(k=20, use_syncmer=False, window_size=116)      # 97 k-mers per window: beyond taxor build's limit
    with pytest.raises(capi.TaxorError):
        capi.pack_ascii(["ACGTX"])
    with pytest.raises(capi.TaxorError):
        capi.Context(4096)


def test_baseline_config0_exactly_as_defined(ctx, oracle):
    """BASELINE.json configs[0] as SURVEY 8(d) defines it: 100 genomes x 5,000,000 bp (i.i.d. uniform, genome g from seed
    1000+g), k=22 s=12 t=5 syncmers; 10,000 reads x 10,000 bp (genome, start, strand uniform; 5 % errors, 1/3 each
    substitution / insertion / deletion; read seed 42); --error-rate 0.05 (ratio 0.437803).  The genomes are hashed by the
    GPU build-side entry point (spot-checked against the oracle), the HIXF is peeled on the CPU, every read's hash count,
    threshold, hits and counts are compared with the oracle."""
    from taxor_b200 import tools
    n_g, glen = 100, 5_000_000
    genomes = [tools.genome(1000 + g, glen) for g in range(n_g)]
    lens = [glen] * n_g
    ctx.set_params(k=22, s=12, t=5, use_syncmer=True, window_size=20, error_rate=0.05)
    nw = np.array([len(w) for w in genomes], dtype=np.uint64)
    off = np.zeros(n_g, dtype=np.uint64)
    off[1:] = np.cumsum(nw)[:-1]
    seqs = capi.PackedReads(np.concatenate(genomes), off, np.full(n_g, glen, np.uint32))
    boff, hashes, n_seg = ctx.hash_user_bins(seqs, np.arange(n_g, dtype=np.uint32), n_g)
    ub = [hashes[int(boff[i]):int(boff[i + 1])] for i in range(n_g)]
    assert n_seg > n_g                                           # long genomes were cut into independently hashed segments
    for g in (0, 57):                                            # the GPU's sets against the oracle's scan of the whole genome
        exp = oracle.syncmer_hashes(H.codes_of(genomes[g], glen), 22, 12, 5)
        assert np.array_equal(np.sort(ub[g]), np.sort(exp))
    assert 4.2e7 < sum(len(x) for x in ub) < 4.8e7               # 500 Mbp / 11: about 45 M distinct hashes in the index
    hx = tools.BuiltHixf(ub, t_max=64, seed=1)
    from oracle.oracle import HixfArrays
    arrays = HixfArrays(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.rows)
    ds = H.Dataset(22, 12, 5, True, genomes, lens, hx, arrays)
    words, roff, rlen, _ = tools.simulate_reads(genomes, lens, np.full(10_000, 10_000), 0.05, 42)
    reads = capi.PackedReads(words, roff, rlen)
    codes, coff = H.reads_to_codes(reads)
    oh = oracle.make_hixf(arrays)
    H.upload(ctx, ds)
    for er in (0.05, 0.10):                                      # as defined, and the rate at which reads actually classify
        ctx.set_params(k=22, s=12, t=5, use_syncmer=True, window_size=20, error_rate=er)
        res = ctx.search(reads)
        ora = oracle.search_batch(oh, codes, coff, k=22, s=12, t=5, use_syncmer=True, window_size=20, error_rate=er)
        H.assert_same_search(res, ora, reads.n)
        if er == 0.05:
            assert abs(int(res.threshold[0]) / max(int(res.hash_count[0]), 1) - 0.437803) < 2e-3
        else:
            assert int(res.hit_begin[-1]) > 3000
    hx.close()


def test_sharded_search_over_contexts_equals_one_context(oracle):
    """the Python form of the multi-GPU driver (taxor_b200/shard.py: contiguous blocks, concatenation in read order) on real
    contexts: shards searched on two devices (two contexts on device 0 where there is only one), merged, equal to one context
    searching everything and to the oracle"""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from taxor_b200 import shard
    ds = H.make_dataset(oracle, n_genomes=70, genome_len=40_000, t_max=16)
    rng = np.random.default_rng(77)
    reads = H.make_reads(ds, rng.integers(200, 9000, 1001), err=0.05)
    world = 3
    devs = [r % max(torch.cuda.device_count(), 1) for r in range(world)]
    ctxs = [capi.Context(d) for d in devs]
    try:
        H.upload(ctxs[0], ds)
        for c in ctxs[1:]:
            c.clone_index_from(ctxs[0])
        for c in ctxs:
            c.set_params(k=ds.k, s=ds.s, t=ds.t, use_syncmer=True, window_size=20, error_rate=0.1)
        ctxs[1].reserve(400, 2_000_000)                                      # buffers allocated ahead of the first call: same answers

        def part(r):
            lo, hi = shard.shard_range(reads.n, r, world)
            w0 = int(reads.word_off[lo]) if lo < reads.n else 0
            sub = capi.PackedReads(reads.words[w0:], reads.word_off[lo:hi] - np.uint64(w0), reads.length[lo:hi])
            return shard.result_to_dict(ctxs[r].search(sub))
        with ThreadPoolExecutor(max_workers=world) as ex:                      # one host thread per context, as in the CLI
            parts = list(ex.map(part, range(world)))
        merged = shard.concat_results(parts)
        whole = shard.result_to_dict(ctxs[0].search(reads))
        for k_ in whole:
            assert np.array_equal(merged[k_], whole[k_]), k_
        codes, off = H.reads_to_codes(reads)
        ora = oracle.search_batch(oracle.make_hixf(ds.arrays), codes, off, k=ds.k, s=ds.s, t=ds.t, use_syncmer=True, window_size=20,
                                  error_rate=0.1)
        assert np.array_equal(merged["hit_begin"], ora["raw_off"]) and np.array_equal(merged["user_bin"], ora["raw_ub"])
        assert np.array_equal(merged["count"], ora["raw_cnt"]) and int(merged["hit_begin"][-1]) > 200
    finally:
        for c in ctxs:
            c.close()
