"""Shared builders for the parity tests: synthetic genomes -> oracle hashes -> HIXF (CPU tooling) -> reads."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from oracle.oracle import HixfArrays
from taxor_b200 import capi, tools


def codes_of(words: np.ndarray, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    capi.lib().txr_unpack_codes(words.ctypes.data, n, out.ctypes.data)
    return out


@dataclass
class Dataset:
    k: int
    s: int
    t: int
    use_syncmer: bool
    genomes: list
    genome_len: list
    hixf: tools.BuiltHixf
    arrays: HixfArrays


def make_dataset(oracle, *, n_genomes=24, genome_len=60_000, k=22, s=12, t=None, use_syncmer=True, t_max=8,
                 seed=1000, size_jitter=True, scaling=1, window_size=None, scheme=None) -> Dataset:
    rng = np.random.default_rng(seed)
    lens = [int(genome_len * (0.5 + rng.random())) if size_jitter else genome_len for _ in range(n_genomes)]
    genomes = [tools.genome(seed + g, lens[g]) for g in range(n_genomes)]
    if t is None:
        t = oracle.t_syncmer(k, s)
    ub = []
    for g in range(n_genomes):
        c = codes_of(genomes[g], lens[g])
        if use_syncmer:
            h = oracle.syncmer_hashes(c, k, s, t)
        elif window_size is not None and window_size > k:
            h = np.unique(oracle.minimiser_hashes(c, k, window_size))     # compute_hashes.cpp:120-124
        else:
            h = np.unique(oracle.kmer_hashes(c, k))
        if scaling > 1:
            h = np.array([x for x in h.tolist() if oracle.scaling_keep(x, scaling)], dtype=np.uint64)
        ub.append(h)
    hx = tools.BuiltHixf(ub, t_max=t_max, seed=seed, scheme=scheme)
    arrays = HixfArrays(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub, hx.rows)
    return Dataset(k, s, t, use_syncmer, genomes, lens, hx, arrays)


def upload(ctx, ds: Dataset) -> None:
    h = ds.hixf
    ctx.upload_index(h.seed, h.bins, h.tbins, h.seg_len, h.data, h.bin_off, h.next_ixf_id, h.bin_to_ub, h.n_user_bins,
                     rows=getattr(h, "rows", None), scheme=getattr(h, "scheme", None))


def make_reads(ds: Dataset, lengths, err=0.05, seed=42) -> capi.PackedReads:
    words, off, ln, _ = tools.simulate_reads(ds.genomes, ds.genome_len, lengths, err, seed)
    return capi.PackedReads(words, off, ln)


def reads_to_codes(reads: capi.PackedReads):
    """(concatenated codes, offsets) for the oracle."""
    parts = [capi.unpack_codes(reads, i) for i in range(reads.n)]
    off = np.zeros(reads.n + 1, dtype=np.uint64)
    if reads.n:
        off[1:] = np.cumsum([len(p) for p in parts])
    codes = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes, dtype=np.uint8), off


def assert_same_search(res, ora, n):
    """GPU result (capi.SearchResult) vs oracle.search_batch dict: bit-exact counts, thresholds, hits, order."""
    assert res.n_reads == n
    assert np.array_equal(res.hash_count, ora["hash_count"])
    assert np.array_equal(res.threshold, ora["threshold"])
    assert np.array_equal(res.hit_begin, ora["raw_off"])
    assert np.array_equal(res.user_bin, ora["raw_ub"])
    assert np.array_equal(res.count, ora["raw_cnt"])
    # 0.8*max filter
    kept_ub, kept_cnt = res.user_bin[res.keep], res.count[res.keep]
    assert np.array_equal(kept_ub, ora["ub"])
    assert np.array_equal(kept_cnt, ora["cnt"])


HEADER = "#QUERY_NAME\tACCESSION\tREFERENCE_NAME\tTAXID\tREF_LEN\tQUERY_LEN\tQHASH_COUNT\tQHASH_MATCH\tTAX_STR\tTAX_ID_STR\n"


def oracle_tsv(oracle, arrays, species, records, *, k, s, t, use_syncmer, window_size=20, scaling=1, percentage=-1.0,
               error_rate=0.04, header=True):
    """The result file `taxor search --threads 1` writes (src/main/taxor_search.cpp:268-306, 343), restated on top of
    the oracle: records = [(id, ascii sequence)]."""
    codes = [np.array([oracle.dna4_rank(c) for c in seq], dtype=np.uint8) for _, seq in records]
    off = np.zeros(len(records) + 1, dtype=np.uint64)
    if records:
        off[1:] = np.cumsum([len(c) for c in codes])
    flat = np.concatenate(codes) if codes else np.zeros(0, np.uint8)
    res = oracle.search_batch(oracle.make_hixf(arrays), np.ascontiguousarray(flat, dtype=np.uint8), off, k=k, s=s, t=t,
                              use_syncmer=use_syncmer, window_size=window_size, scaling=scaling, percentage=percentage,
                              error_rate=error_rate)
    ub_index = {}
    for i, sp in enumerate(species):                       # std::map::emplace: the first entry for a user bin wins (:172-178)
        ub_index.setdefault(sp["user_bin"], i)
    out = [HEADER] if header else []
    for r, (rid, seq) in enumerate(records):
        a, b = int(res["hit_off"][r]), int(res["hit_off"][r + 1])
        if a == b:
            out.append(f"{rid}\t-\t-\t-\t-\t{len(seq)}\n")
            continue
        for i in range(a, b):
            sp = species[ub_index.get(int(res["ub"][i]), 0)]
            out.append("\t".join([rid, sp["accession_id"], sp["organism_name"], sp["taxid"], str(sp["seq_len"]), str(len(seq)),
                                  str(int(res["hash_count"][r])), str(int(res["cnt"][i])), sp["taxnames_string"],
                                  sp["taxid_string"]]) + "\n")
    return "".join(out)
