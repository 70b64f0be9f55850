"""not gpu: the C-ABI library loads, exports every symbol include/*.h declares, and its host-only entry points
(packing, threshold mirror) agree with the oracle.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from taxor_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(txr_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol(built_libs):
    lib = C.CDLL(built_libs.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(capi.EXPORTED), decl ^ set(capi.EXPORTED)


def test_no_cpu_fallback(built_libs):
    """Without a CUDA device context creation must fail loudly (the product never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.TaxorError) as e:
        capi.Context(0)
    assert "-1" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_touches_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "taxor_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|taxor_oracle|ixf_ref\.h", txt):
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_pack_layout_and_roundtrip(built_libs, oracle):
    # base i sits in word i/32 at bits [62-2*(i%32), 63-2*(i%32)]; one zero pad word
    r = capi.pack_ascii(["ACGT", "T" * 32, "G" * 33, ""])
    w = r.words
    assert int(w[0]) == (0b00011011 << 56) and int(w[1]) == 0
    assert int(w[2]) == (1 << 64) - 1 and int(w[3]) == 0
    assert int(w[4]) == int("10" * 32, 2) and int(w[5]) == 0b10 << 62 and int(w[6]) == 0
    assert list(r.word_off) == [0, 2, 4, 7] and list(r.length) == [4, 32, 33, 0]
    rng = np.random.default_rng(0)
    seqs = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in (0, 1, 31, 32, 33, 64, 1000, 12345)]
    p = capi.pack_codes(seqs)
    for i, s in enumerate(seqs):
        assert np.array_equal(capi.unpack_codes(p, i), s)
    # IUPAC collapse identical to the oracle's dna4 table; illegal characters are rejected
    iupac = "ACGTUNRYSWKMBDHVacgtunryswkmbdhv"
    q = capi.pack_ascii([iupac])
    assert capi.unpack_codes(q, 0).tolist() == [oracle.dna4_rank(c) for c in iupac]
    for bad in ("ACGT-", "ACXGT", "AC GT", "1234"):
        with pytest.raises(capi.TaxorError):
            capi.pack_ascii([bad])
    with pytest.raises(capi.TaxorError):
        capi.pack_codes([np.array([0, 1, 4], np.uint8)])


def test_threshold_mirror_matches_oracle(built_libs, oracle):
    """taxor_b200/csrc/threshold.cpp (product host code) == oracle for every model, incl. the size_t wrap."""
    for (w, k, p, e, syn) in [(20, 22, -1.0, 0.05, True), (20, 22, -1.0, 0.1, True), (20, 22, -1.0, 0.045, True),
                              (20, 20, -1.0, 0.05, False), (20, 20, -1.0, 0.2, False), (22, 22, -1.0, 0.01, False),
                              (24, 20, -1.0, 0.05, False), (20, 22, 0.3, 0.05, True), (20, 20, 1.0, 0.05, False)]:
        ot = oracle.thresholder(w, k, p, e, syn)
        for c in list(range(1, 600)) + [907, 4096, 49981, 10**6]:
            for sf in (1.0, 0.2):
                got = capi.threshold_eval(c, sf, k=k, use_syncmer=syn, window_size=w, percentage=p, error_rate=e)
                assert got == oracle.threshold_get(ot, c, sf), (w, k, p, e, syn, c, sf)
    assert capi.threshold_eval(0, 1.0, k=22, use_syncmer=True, window_size=20, error_rate=0.05) == 0


def _tie_heavy_sequences(rng, n):
    out = []
    for _ in range(n):
        x = rng.integers(0, 4, int(rng.integers(2000, 30000)), dtype=np.uint8)
        for _ in range(int(rng.integers(0, 10))):
            at, ln = int(rng.integers(0, len(x) - 1500)), int(rng.integers(20, 1500))
            kind = rng.integers(0, 3)
            if kind == 0:
                x[at:at + ln] = rng.integers(0, 4)                                        # homopolymer
            elif kind == 1:
                x[at:at + ln] = np.resize(rng.integers(0, 4, int(rng.integers(2, 7)), dtype=np.uint8), ln)   # tandem repeat
            else:
                half = x[at:at + ln // 2].copy()
                x[at + ln // 2:at + 2 * (ln // 2)] = (3 - half)[::-1]                     # reverse-complement palindrome
        out.append(x)
    return out


@pytest.mark.parametrize("mode", ["syncmer", "syncmer_t1", "kmer", "minimiser"])
def test_segment_cuts_are_history_free(oracle, built_libs, mode):
    """txr_plan_segments (host only): hashing the pieces independently gives exactly what one scan of the whole sequence
    gives -- the property txr_hash_user_bins rests on -- checked with the ORACLE's scan on tie-heavy sequences."""
    from taxor_b200 import capi
    rng = np.random.default_rng({"syncmer": 1, "syncmer_t1": 2, "kmer": 3, "minimiser": 4}[mode])
    k, s, t, w = {"syncmer": (22, 12, 5, 20), "syncmer_t1": (16, 8, 1, 20), "kmer": (20, 0, 0, 20), "minimiser": (20, 0, 0, 33)}[mode]
    syn = mode.startswith("syncmer")
    span = k if syn else w
    n_cut = 0
    for x in _tie_heavy_sequences(rng, 40):
        words = capi.pack_codes([x]).words
        cuts = capi.plan_segments(words, len(x), 96, k=k, s=s, use_syncmer=syn, window_size=w)
        assert cuts[0] == 0 and np.all(cuts % 32 == 0) and np.all(np.diff(cuts.astype(np.int64)) > 0)
        n_cut += len(cuts) - 1
        pieces = [x[int(cuts[i]):(int(cuts[i + 1]) + span - 1 if i + 1 < len(cuts) else len(x))] for i in range(len(cuts))]
        if syn:
            whole = oracle.syncmer_hashes_raw(x, k, s, t)
            parts = np.concatenate([oracle.syncmer_hashes_raw(p, k, s, t) for p in pieces])
            assert np.array_equal(whole, parts)                     # same emissions in the same order
        elif w == k:
            assert np.array_equal(oracle.kmer_hashes(x, k), np.concatenate([oracle.kmer_hashes(p, k) for p in pieces]))
        else:
            whole = set(oracle.minimiser_hashes(x, k, w).tolist())
            parts = set(np.concatenate([oracle.minimiser_hashes(p, k, w) for p in pieces]).tolist())
            assert whole == parts                                   # a piece re-reports its first minimiser: equal as sets
    assert n_cut > 200                                              # the sequences really were cut, ties and all


def test_pack_simd_lane_equals_table_lane(built_libs, oracle):
    """txr_pack_2bit packs whole words of plain ACGT/acgt with AVX2+BMI2 (pack_simd.cpp) and everything else through the dna4
    table: both lanes, mixed inside one read (an IUPAC code every few hundred bases, lower case, all lengths around the word
    size), must give the words txr_pack_codes gives for the dna4 ranks; an illegal character is reported at its position"""
    rng = np.random.default_rng(44)
    iupac = "NRWMDHVYSBKUnrykm"
    for n in list(range(0, 70)) + [95, 96, 97, 1000, 4099, 20_001]:
        codes = rng.integers(0, 4, n, dtype=np.uint8)
        chars = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
        lower = rng.random(n) < 0.3
        chars[lower] |= 0x20
        s = bytearray(chars.tobytes())
        for pos in rng.integers(0, max(n, 1), n // 300 + (1 if n > 40 else 0)):
            s[int(pos)] = ord(iupac[int(rng.integers(0, len(iupac)))])
        text = s.decode()
        want = np.array([oracle.dna4_rank(ch) for ch in text], dtype=np.uint8)
        assert (want <= 3).all()
        assert np.array_equal(capi.pack_ascii([text]).words, capi.pack_codes([want]).words), n
    bad = "ACGT" * 50 + "J" + "ACGT" * 50
    with pytest.raises(capi.TaxorError, match="0x4a at position 200"):
        capi.pack_ascii([bad])
