"""not gpu: the CPU tooling (synthetic index generator, read simulator) obeys the HIXF invariants of SURVEY 3.4 and
produces filters without false negatives under the oracle's probe."""
import numpy as np

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
