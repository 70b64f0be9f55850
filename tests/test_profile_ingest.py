"""not gpu: SURVEY 8(f) rank 3 -- the head of `taxor profile` (parse_search_results + the three reference-filter rounds) in the
product (txr_profile_*, host code) against the REFERENCE'S OWN src/main/taxor_profile.cpp compiled in place (oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

from taxor_b200 import capi
from tests import helpers as H


def synth_result_file(path, rng, n_reads=4000, n_refs=40, dup_ids=True):
    """A result file shaped like a metagenome: a few dominant references, satellites that share almost all their reads with a
    dominant one (round 3's "explained" case), rare references (round 2), reads with up to 6 hits, unclassified reads, ids with
    spaces, and -- on purpose -- repeated read ids (parse_search_results merges them, a leading "-" survives)."""
    refs = [dict(acc=f"GCF_{i:06d}.1", name=f"Org {i}", taxid=str(1000 + i), length=1_000_000 + 1000 * i,
                 names=f"k__B;g__G{i % 7};s__S{i}", ids=f"2;{100 + i % 7};{1000 + i}") for i in range(n_refs)]
    lines = [H.HEADER]
    weights = np.array([50.0 / (1 + i) for i in range(n_refs)])
    weights /= weights.sum()
    for r in range(n_reads):
        rid = f"read_{r if not dup_ids or r % 97 else r - 1} runid=x ch={r % 5}"
        L = int(rng.integers(500, 20000))
        kind = rng.random()
        if kind < 0.15:
            lines.append(f"{rid}\t-\t-\t-\t-\t{L}\n")
            continue
        main = int(rng.choice(n_refs, p=weights))
        hits = [main]
        if main % 5 == 0 and main + 1 < n_refs and rng.random() < 0.97:      # satellite main+1 shares nearly every read with main
            hits.append(main + 1)
        if kind > 0.7:
            hits += [int(x) for x in rng.choice(n_refs, size=int(rng.integers(1, 5)))]
        seen, order = set(), []
        for h in hits:
            if h not in seen:
                seen.add(h)
                order.append(h)
        hc = int(rng.integers(50, 1800))
        for h in order:
            f = refs[h]
            lines.append("\t".join([rid, f["acc"], f["name"], f["taxid"], str(f["length"]), str(L), str(hc), str(int(rng.integers(10, hc + 1))),
                                    f["names"], f["ids"]]) + "\n")
    open(path, "w").write("".join(lines))


@pytest.mark.parametrize("seed,dup", [(1, True), (2, False), (3, True)])
def test_profile_rounds_match_the_reference_compiled_in_place(tmp_path, reference, seed, dup):
    if not hasattr(reference.lib, "ref_profile_prefilter"):
        pytest.skip("oracle/_ref predates the profile shim")
    rng = np.random.default_rng(seed)
    path = tmp_path / "search.tsv"
    synth_result_file(path, rng, n_reads=3000 + 500 * seed, dup_ids=dup)
    for stage in (0, 1, 2, 3):
        p = capi.Profile()
        p.add_file(path)
        p.filter(stage)
        got = p.text()
        p.close()
        exp = reference.profile_prefilter(path, stage)
        assert got == exp, (stage, next((a, b) for a, b in zip(got.split("\n"), exp.split("\n")) if a != b))
    assert "\nT\t" in exp and exp.count("\nR\t") > 1000
    # round 3 really explained references away and round 2 really dropped some
    after2 = reference.profile_prefilter(path, 2)
    assert after2 != reference.profile_prefilter(path, 1) and exp != after2


def test_profile_in_memory_feed_equals_the_file_round_trip(tmp_path, reference):
    """txr_profile_add_batch on a txr_result (what `taxor search` holds in memory) == the reference parsing the result file
    written from the same result; the result here is hand-made (no GPU needed): two batches, hits with keep == 0, repeated ids"""
    if not hasattr(reference.lib, "ref_profile_prefilter"):
        pytest.skip("oracle/_ref predates the profile shim")
    from taxor_b200 import tools
    rng = np.random.default_rng(7)
    n_ub = 25
    species = tools.default_species(n_ub)
    by_ub = {}
    for i, sp in enumerate(species):
        by_ub.setdefault(sp["user_bin"], i)
    lines = [H.HEADER]
    prof = capi.Profile()
    rid0 = 0
    for batch in range(2):
        n = 1500
        ids = [f"r{rid0 + i if (rid0 + i) % 50 else rid0 + i - 1} extra words" for i in range(n)]
        rid0 += n
        rl = rng.integers(300, 9000, n).astype(np.uint32)
        hc = rng.integers(20, 900, n).astype(np.uint32)
        nh = np.where(rng.random(n) < 0.2, 0, rng.integers(1, 5, n))
        begin = np.zeros(n + 1, np.uint64)
        begin[1:] = np.cumsum(nh)
        tot = int(begin[-1])
        ub = np.zeros(tot, np.int64)
        cnt = np.zeros(tot, np.uint32)
        keep = np.zeros(tot, np.uint8)
        for r in range(n):
            a, b = int(begin[r]), int(begin[r + 1])
            if a == b:
                lines.append(f"{ids[r]}\t-\t-\t-\t-\t{rl[r]}\n")
                continue
            main = int(min(n_ub - 1, rng.geometric(0.25) - 1))
            cand = [main] + [int(x) for x in rng.choice(n_ub, b - a - 1)] if b - a > 1 else [main]
            ub[a:b] = cand
            cnt[a:b] = rng.integers(5, int(hc[r]) + 1, b - a)
            mx = cnt[a:b].max()
            keep[a:b] = ~(cnt[a:b].astype(np.float64) < float(mx) * 0.8)
            for i in range(a, b):
                if keep[i]:
                    sp = species[by_ub[int(ub[i])]]
                    lines.append("\t".join([ids[r], sp["accession_id"], sp["organism_name"], sp["taxid"], str(sp["seq_len"]), str(int(rl[r])),
                                            str(int(hc[r])), str(int(cnt[i])), sp["taxnames_string"], sp["taxid_string"]]) + "\n")
        thr = np.zeros(n, np.uint64)
        res = capi.Result(n, hc.ctypes.data_as(C.POINTER(C.c_uint32)), thr.ctypes.data_as(C.POINTER(C.c_uint64)),
                          begin.ctypes.data_as(C.POINTER(C.c_uint64)), ub.ctypes.data_as(C.POINTER(C.c_int64)),
                          cnt.ctypes.data_as(C.POINTER(C.c_uint32)), keep.ctypes.data_as(C.POINTER(C.c_uint8)))
        prof.add_result(res, ids, rl, species)
    path = tmp_path / "written.tsv"
    open(path, "w").write("".join(lines))
    assert prof.text() == reference.profile_prefilter(path, 0)
    prof.filter(3)
    assert prof.text() == reference.profile_prefilter(path, 3)
    with pytest.raises(capi.TaxorError):
        prof.add_file(path)                                      # a filtered table takes no more input
    prof.close()
    bad = capi.Profile()
    with pytest.raises(capi.TaxorError, match="open"):
        bad.add_file(tmp_path / "missing.tsv")
    open(tmp_path / "short.tsv", "w").write(H.HEADER + "r1\tGCF_1\n")
    with pytest.raises(capi.TaxorError, match="columns"):
        bad.add_file(tmp_path / "short.tsv")
    bad.close()


def test_profile_fuzz_under_sanitizers(tmp_path):
    """tests/fuzz/profile_fuzz.cpp built with -fsanitize=address,undefined: damaged result files through add_file, the three
    filter rounds and both read-out calls; rejected with a message or accepted and consistent, never a crash."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "profile_fuzz")
    build = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                            os.path.join(root, "tests/fuzz/profile_fuzz.cpp"), os.path.join(root, "taxor_b200/csrc/profile_ingest.cpp")],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower():
        pytest.skip("compiler without sanitizer runtimes")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, str(tmp_path / "r.tsv"), "1500", "5"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "fuzz ok" in run.stdout, run.stdout[-500:] + run.stderr[-3000:]
