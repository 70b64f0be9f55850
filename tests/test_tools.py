"""not gpu: the CPU tooling (synthetic index generator, read simulator) obeys the HIXF invariants of SURVEY 3.4 and
produces filters without false negatives under the oracle's probe."""
import numpy as np
import pytest

from oracle.oracle import HixfArrays
from taxor_b200 import capi, tools
from tests import helpers as H


def test_builder_invariants(built_libs, oracle):
    rng = np.random.default_rng(5)
    ub = [np.unique(rng.integers(0, 2**63, size=int(n), dtype=np.uint64)) for n in rng.integers(100, 3000, 90)]
    hx = tools.BuiltHixf(ub, t_max=8, seed=3)
    assert hx.n_ixf > 8 and hx.n_user_bins == 90
    seen_ub, children = set(), set()
    for i in range(hx.n_ixf):
        a, b = int(hx.bin_off[i]), int(hx.bin_off[i + 1])
        assert b - a == int(hx.bins[i]) <= 8 and int(hx.tbins[i]) == 64 and hx.data[i].size == 3 * int(hx.seg_len[i]) * 64
        ubs, nxt = hx.bin_to_ub[a:b], hx.next_ixf_id[a:b]
        for j in range(b - a):
            if ubs[j] < 0:                                   # merged bin: -1 and a child IXF (invariants 2, 3)
                assert ubs[j] == -1 and nxt[j] != i and 0 < nxt[j] < hx.n_ixf
                assert int(nxt[j]) not in children
                children.add(int(nxt[j]))
            else:
                assert nxt[j] == i
        # a split user bin occupies consecutive bins; every user bin lives in exactly one IXF (invariant 3)
        runs = [int(u) for j, u in enumerate(ubs) if u >= 0 and (j == 0 or ubs[j - 1] != u)]
        assert len(runs) == len(set(runs)) and not (set(runs) & seen_ub)
        seen_ub |= set(runs)
    assert seen_ub == set(range(90)) and children == set(range(1, hx.n_ixf))   # a tree rooted at IXF 0 (invariant 1)
    # no false negatives: every hash of a user bin is found (count == n) along its DFS path
    arrays = HixfArrays(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub)
    h = oracle.make_hixf(arrays)
    for u in (0, 17, 55, 89):
        res_ub, res_cnt, _ = oracle.bulk_contains(h, ub[u], len(ub[u]))
        assert u in res_ub.tolist() and int(res_cnt[res_ub.tolist().index(u)]) >= len(ub[u])
    # false-positive rate of a foreign key set is ~2^-8 per bin (8-bit fingerprints, threshold.hpp:53)
    foreign = rng.integers(0, 2**63, size=20000, dtype=np.uint64)
    c0 = oracle.ixf_bulk_count(h, 0, foreign)
    assert 0.001 < c0.mean() / 20000 < 0.01
    hx.close()


def test_read_simulator(built_libs):
    g = [tools.genome(1000 + i, 50_000) for i in range(3)]
    assert not np.array_equal(g[0][:100], g[1][:100])
    w1, off, ln, src = tools.simulate_reads(g, [50_000] * 3, [2000] * 10, 0.05, 42)
    w2, _, _, _ = tools.simulate_reads(g, [50_000] * 3, [2000] * 10, 0.05, 42, threads=1)
    assert np.array_equal(w1, w2)                                  # deterministic in (seed, read index)
    w3, _, _, _ = tools.simulate_reads(g, [50_000] * 3, [2000] * 10, 0.05, 43)
    assert not np.array_equal(w1, w3)
    # error-free forward reads are substrings of their genome
    w0, off0, ln0, src0 = tools.simulate_reads(g, [50_000] * 3, [300] * 20, 0.0, 7)
    reads = capi.PackedReads(w0, off0, ln0)
    comp = str.maketrans("ACGT", "TGCA")
    for i in range(20):
        s = "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))
        gs = "".join("ACGT"[c] for c in H.codes_of(g[int(src0[i])], 50_000))
        assert s in gs or s.translate(comp)[::-1] in gs


SCHEMES = [None, (0, 0, 0, 21, 42), (1, 0, 0, 0, 0), (0, 1, 1, 13, 37), (1, 1, 2, 0, 0), (0, 0, 2, 21, 42)]


@pytest.mark.parametrize("scheme", SCHEMES)
def test_ixf_scheme_variants_builder_vs_oracle(oracle, scheme):
    """The probe arithmetic is a descriptor (the fork that defines it is absent): for every candidate scheme the product-side
    builder (ixf_arith.cuh) and the oracle's independently written twin (oracle/ixf_ref.h) must agree -- every key the builder
    stored is found by the oracle in exactly its bin, foreign keys at the 1/256 rate -- and a mismatching scheme must NOT."""
    rng = np.random.default_rng(11)
    ub = [rng.integers(0, 2**63, int(n), dtype=np.uint64) for n in rng.integers(300, 3000, 40)]
    hx = tools.BuiltHixf(ub, t_max=64, seed=5, scheme=scheme)
    from oracle.oracle import HixfArrays
    arrays = HixfArrays(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.rows)
    try:
        oracle.set_ixf_scheme(scheme)
        h = oracle.make_hixf(arrays)
        for u in (0, 7, 39):
            cnt = oracle.ixf_bulk_count(h, 0, hx._ub[u])
            bins = np.flatnonzero(hx.bin_to_ub[: int(hx.bins[0])] == u)
            assert int(cnt[bins].sum()) >= len(hx._ub[u])                      # every stored key is found in its (split) bins
            others = np.delete(cnt, bins)
            assert others.max() < 0.03 * len(hx._ub[u]) + 10                   # false positives ~ 1/256 per bin
        foreign = rng.integers(0, 2**63, 4000, dtype=np.uint64)
        assert oracle.ixf_bulk_count(h, 0, foreign).max() < 120                # ~16 expected; fold-less fingerprints correlate with the slots
        # a different scheme reads garbage from the same arrays (when the geometry allows it to be read at all)
        other = (0, 1, 0, 21, 42) if scheme in (None, (0, 0, 0, 21, 42)) else None
        if scheme is not None and scheme[0] == 1:
            other = (1, 1 - scheme[1], scheme[2], 0, 0)
        if other is not None:
            oracle.set_ixf_scheme(other)
            cnt = oracle.ixf_bulk_count(h, 0, hx._ub[7])
            assert cnt.max() < 0.05 * len(hx._ub[7]) + 10
    finally:
        oracle.set_ixf_scheme(None)
        hx.close()


def test_fuse_geometry_follows_the_published_allocation():
    """binary fuse geometry (segment length a power of two, size factor >= 1.125) as in FastFilter's binary_fuse8_allocate"""
    rng = np.random.default_rng(3)
    for n in (50, 1000, 20_000):
        ub = [rng.integers(0, 2**63, n, dtype=np.uint64) for _ in range(3)]
        hx = tools.BuiltHixf(ub, t_max=4, seed=1, scheme=(1, 0, 0, 0, 0))
        L, rows, cap = int(hx.seg_len[0]), int(hx.rows[0]), int(hx.capacity[0])
        assert L & (L - 1) == 0 and rows % L == 0 and rows >= 3 * L
        import math
        assert L == min(262144, 1 << int(math.floor(math.log(cap) / math.log(3.33) + 2.25)))
        assert rows >= 1.125 * cap - 2 * L and rows <= (max(1.125, 0.875 + 0.25 * math.log(1e6) / math.log(cap)) * cap) + 3 * L
        hx.close()
