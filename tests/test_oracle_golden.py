"""not gpu: pins the oracle (CPU restatement) against
  (1) the golden vectors generated from the reference's own sources compiled in place (tests/golden/make_golden.py),
  (2) that compiled reference itself, live, where oracle/_ref exists,
  (3) hand-derived known-answer values."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import HixfArrays

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M64 = (1 << 64) - 1
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}


def to_codes(seq):
    return np.array([CODE[c] for c in seq], dtype=np.uint8)


# ---------------------------------------------------------------- known answers
def test_wyhash_known_answers(oracle):
    """ankerl v3.0.1 wyhash::hash(u64) = lo64(x*C) ^ hi64(x*C), C = 0x9E3779B97F4A7C15 -- recomputed with Python big ints."""
    for x in [0, 1, 2, 3, 0xDEADBEEF, 0x0123456789ABCDEF, M64, 1 << 63, 0x3FFFFFFFFFF]:
        p = x * 0x9E3779B97F4A7C15
        assert oracle.wyhash(x) == ((p & M64) ^ (p >> 64))
    assert oracle.wyhash(1) == 0x9E3779B97F4A7C15
    assert oracle.wyhash(0) == 0


def test_scalar_known_answers(oracle):
    assert oracle.t_syncmer(22, 12) == 5 and oracle.t_syncmer(20, 10) == 5      # taxor_build.cpp:510 (integer division)
    assert oracle.t_syncmer(24, 12) == 6 and oracle.t_syncmer(16, 8) == 4
    assert oracle.adjust_seed(20) == 0x8F3F73B5CF                               # adjust_seed.hpp:40-44
    assert oracle.adjust_seed(32) == 0x8F3F73B5CF1C9ADE
    assert oracle.adjust_seed(22) == 0x8F3F73B5CF1C9ADE >> 20
    # seqan3::dna4 collapse of IUPAC
    for ch, r in zip("ACGTUacgtu", [0, 1, 2, 3, 3] * 2):
        assert oracle.dna4_rank(ch) == r
    for ch in "NRWMDHVnrwmdhv":
        assert oracle.dna4_rank(ch) == 0
    for ch in "YSBysb":
        assert oracle.dna4_rank(ch) == 1
    assert oracle.dna4_rank("K") == 2 and oracle.dna4_rank("k") == 2
    for ch in "XZ*-. 0":
        assert oracle.dna4_rank(ch) == -1


def test_threshold_known_answers(oracle):
    # syncmer_model.hpp table lookups quoted in SURVEY 8(c)
    assert oracle.syncmer_match_ratio(22, 0.05) == 0.437803
    assert oracle.syncmer_match_ratio(22, 0.15) == 0.130048
    assert oracle.syncmer_match_ratio(22, 0.04) == 0.50832
    assert oracle.syncmer_match_ratio(22, 0.045) == 0.50832     # ceil: rounds towards the more accurate row
    assert oracle.syncmer_match_ratio(20, 0.0) == 1.0
    t = oracle.thresholder(20, 22, -1.0, 0.05, True)
    assert oracle.threshold_get(t, 907) == int(907 * 0.437803) == 397
    assert oracle.threshold_get(t, 0) == 0                      # reads shorter than k: threshold 0 (flood, H5)
    t = oracle.thresholder(20, 22, 0.3, 0.05, True)
    assert t.kind == 1 and oracle.threshold_get(t, 1000) == 300
    # k-mer model computes in size_t and wraps (threshold.hpp:65)
    t = oracle.thresholder(20, 20, -1.0, 0.2, False)
    assert t.kind == 2 and oracle.threshold_get(t, 10) > (1 << 63)
    # model selection (threshold.hpp:27-48)
    assert oracle.thresholder(24, 20, -1.0, 0.05, False).kind == 0
    assert oracle.thresholder(20, 22, -1.0, 0.05, True).kind == 3
    assert abs(oracle.normal_cdf_inverse(0.975) - 1.96) < 1e-2


def test_scaling_filter_known_answers(oracle):
    assert all(oracle.scaling_keep(h, 1) for h in (0, 5, M64))
    # wyhash(h) <= 2^64/scaling (double compare)
    for h in range(1, 2000, 7):
        v = oracle.wyhash(h)
        assert oracle.scaling_keep(h, 10) == (float(v) <= float(M64) / 10.0)


def test_syncmer_hand_derived(oracle):
    """A 24-mer small enough to scan by hand-coded brute force (independent of the oracle's state machine for the
    untied case): the selected k-mers are those whose minimal canonical s-mer sits at offset t-1."""
    rng = np.random.default_rng(9)
    k, s, t = 22, 12, 5

    def canon(code_slice):
        f = 0
        for c in code_slice:
            f = f * 4 + int(c)
        r = 0
        for c in code_slice[::-1]:
            r = r * 4 + (3 - int(c))
        return min(f, r)

    done = 0
    while done < 20:
        codes = rng.integers(0, 4, 400, dtype=np.uint8)
        exp, tied = [], False
        for j in range(len(codes) - k + 1):
            sm = [canon(codes[j + q:j + q + s]) for q in range(k - s + 1)]
            if len(set(sm)) != len(sm):                          # a tie (e.g. an s-mer next to its reverse complement):
                tied = True                                      # history-dependent, covered by the golden vectors instead
                break
            if int(np.argmin(sm)) == t - 1:
                x = canon(codes[j:j + k])
                p = x * 0x9E3779B97F4A7C15
                exp.append((p & M64) ^ (p >> 64))
        if tied:
            continue
        done += 1
        got = oracle.syncmer_hashes_raw(codes, k, s, t)
        assert got.tolist() == exp


# ---------------------------------------------------------------- golden vectors from the reference's sources
def test_golden_syncmers(oracle):
    cases = json.load(open(os.path.join(G, "syncmer_golden.json")))
    assert len(cases) >= 80
    for c in cases:
        got = oracle.syncmer_hashes(to_codes(c["seq"]), c["k"], c["s"], c["t"])
        assert [str(int(x)) for x in got] == c["hashes"], (c["name"], c["k"], c["s"], c["t"])


def test_golden_thresholds(oracle):
    cases = json.load(open(os.path.join(G, "threshold_golden.json")))
    assert len(cases) >= 400
    for c in cases:
        t = oracle.thresholder(c["window"], c["k"], c["percentage"], c["error_rate"], bool(c["use_syncmer"]))
        assert str(oracle.threshold_get(t, c["count"], c["scaling_factor"])) == c["threshold"], c


def load_dfs_golden():
    z = np.load(os.path.join(G, "dfs_golden.npz"))
    seg, tb = z["seg_len"], z["tbins"]
    data, at = [], 0
    for i in range(len(seg)):
        n = 3 * int(seg[i]) * int(tb[i])
        data.append(np.ascontiguousarray(z["data"][at:at + n]))
        at += n
    arrays = HixfArrays(z["seed"], z["bins"], z["tbins"], z["seg_len"], data, z["bin_off"], z["next_ixf_id"], z["bin_to_ub"])
    return z, arrays


def test_golden_dfs(oracle):
    z, arrays = load_dfs_golden()
    h = oracle.make_hixf(arrays)
    n = int(z["n_queries"])
    assert n >= 100 and arrays.n_ixf >= 3
    for i in range(n):
        ub, cnt, _ = oracle.bulk_contains(h, z[f"q{i}_values"], int(z[f"q{i}_threshold"][0]))
        assert np.array_equal(ub, z[f"q{i}_ub"]) and np.array_equal(cnt, z[f"q{i}_cnt"]), i


# ---------------------------------------------------------------- live against oracle/_ref
def test_live_reference_syncmers(oracle, reference):
    rng = np.random.default_rng(1)
    for it in range(120):
        n = int(rng.integers(0, 2500))
        mode = it % 4
        if mode == 0:
            codes = rng.integers(0, 4, n, dtype=np.uint8)
        elif mode == 1:
            codes = (rng.integers(0, 2, n) * 3).astype(np.uint8)
        elif mode == 2:
            codes = np.resize(rng.integers(0, 4, int(rng.integers(1, 8)), dtype=np.uint8), n).copy()
            if n:
                codes[rng.integers(0, n, max(1, n // 150))] = rng.integers(0, 4, max(1, n // 150))
        else:
            codes = rng.integers(0, 5, n, dtype=np.uint8)          # with N restarts
        for (k, s) in [(22, 12), (20, 10), (30, 8)]:
            t = oracle.t_syncmer(k, s)
            assert np.array_equal(oracle.syncmer_hashes(codes, k, s, t), reference.syncmer_hashes(codes, k, s, t))


def test_live_reference_thresholds(oracle, reference):
    for (w, k, p, e, syn) in [(20, 22, -1, 0.05, 1), (20, 22, -1, 0.1, 1), (20, 20, -1, 0.05, 0), (24, 20, -1, 0.03, 0), (20, 22, 0.42, 0.05, 1)]:
        ot, rt = oracle.thresholder(w, k, p, e, syn), reference.thresholder(w, k, p, e, syn)
        for c in list(range(1, 1500)) + [5000, 49981, 10**6]:
            for sf in (1.0, 0.3):
                assert oracle.threshold_get(ot, c, sf) == reference.threshold_get(rt, c, sf)


def test_live_reference_dfs(oracle, reference):
    z, arrays = load_dfs_golden()
    oh, rh = oracle.make_hixf(arrays), reference.make_hixf(arrays)
    rng = np.random.default_rng(3)
    for _ in range(40):
        vals = rng.integers(0, 2**63, int(rng.integers(0, 400)), dtype=np.uint64)
        for thr in (0, 1, 3, len(vals) // 100 + 1):
            a, b = oracle.bulk_contains(oh, vals, thr), reference.bulk_contains(rh, vals, thr)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    reference.free_hixf(rh)


# ---------------------------------------------------------------------------------------------------------
# minimiser windows (window_size > k): seqan3::views::minimiser_hash, upstream SeqAn3 3.3.0 semantics
# ---------------------------------------------------------------------------------------------------------
def _minimiser_model(values, w_vals):
    """Independent pure-Python statement of views::minimiser (deque of the window, rightmost minimum,
    re-report when the tracked minimiser leaves, strict replacement on arrival)."""
    from collections import deque
    n = len(values)
    if n == 0:
        return []
    W = min(w_vals, n)
    win = deque(values[:W])

    def rightmost(d):
        best, pos = None, 0
        for i, x in enumerate(d):
            if best is None or x <= best:
                best, pos = x, i
        return best, pos
    val, off = rightmost(win)
    out = [val]
    for x in values[W:]:
        win.popleft()
        win.append(x)
        if off == 0:
            val, off = rightmost(win)
            out.append(val)
        elif x < val:
            val, off = x, W - 1
            out.append(val)
        else:
            off -= 1
    return out


def test_minimiser_known_answers(oracle):
    m = {c: i for i, c in enumerate("ACGT")}
    enc = lambda t: np.array([m[c] for c in t], dtype=np.uint8)
    # known answers of upstream SeqAn3's own unit test for views::minimiser_hash (ungapped 4-mers, window 8, seed 0;
    # test/unit/search/views/minimiser_hash_test.cpp, recalled -- the fork is not in /root/reference):
    #   ACGGCGACGTTTAG -> ACGG, CGAC, ACGT, aacg, aaac (lower case = reverse strand);  poly-A -> one 0 per W shifts
    assert oracle.minimiser_hashes(enc("ACGGCGACGTTTAG"), 4, 8, seed=0).tolist() == [26, 97, 27, 6, 1]
    assert oracle.minimiser_hashes(enc("A" * 20), 4, 8, seed=0).tolist() == [0, 0, 0]
    # hand-derived: CCACGTCGACGGTT has the canonical values 81 70 27 109 97 216 97 109 26 22 5; windows of 5:
    # 27 (first window), rescan when it leaves -> RIGHTMOST 97, then the strictly smaller arrivals 26, 22, 5
    assert oracle.minimiser_hashes(enc("CCACGTCGACGGTT"), 4, 8, seed=0).tolist() == [27, 97, 26, 22, 5]
    # shorter than k: nothing; fewer k-mers than a window: ONE window over all of them (ACGT=27, CGTT/aacg=6)
    assert len(oracle.minimiser_hashes(enc("ACG"), 4, 8, seed=0)) == 0
    assert oracle.minimiser_hashes(enc("ACGTT"), 4, 8, seed=0).tolist() == [6]
    # homopolymer: first window, then one report each time the tracked (rightmost) value leaves: every W=5 shifts
    assert oracle.minimiser_hashes(np.zeros(37, np.uint8), 4, 8, seed=0).tolist() == [0] * 6


@pytest.mark.parametrize("k,w", [(4, 8), (6, 7), (8, 20), (20, 24), (20, 40), (31, 126), (12, 12)])
def test_minimiser_against_model(oracle, k, w):
    rng = np.random.default_rng(k * 1000 + w)
    seqs = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in list(rng.integers(0, 400, 30)) + [k - 1, k, k + 1, w - 1, w, w + 1]]
    seqs.append(np.zeros(300, np.uint8))
    seqs.append(np.resize(np.array([0, 1], np.uint8), 500))
    seqs.append(np.resize(np.array([0, 1, 2, 3, 3, 2, 1, 0], np.uint8), 700))
    seqs.append((rng.integers(0, 2, 600) * 3).astype(np.uint8))
    for c in seqs:
        vals = oracle.kmer_hashes(c, k).tolist()
        assert oracle.minimiser_hashes(c, k, w).tolist() == _minimiser_model(vals, w - k + 1), (k, w, len(c))
