"""The drop-in `taxor search` driver (taxor_b200/bin/taxor): argument handling on the CPU, whole runs on the GPU."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from taxor_b200 import build as tb, capi, tools
from tests import helpers as H


def run_cli(*args):
    return subprocess.run([tb.CLI_PATH, *args], capture_output=True, text=True, timeout=600)


def test_cli_argument_errors(tmp_path, built_libs):
    """taxor_search.cpp:32-80, 97-151, 380-384: parser errors go to stderr as [TAXOR SEARCH ERROR] and exit -1."""
    assert os.path.exists(tb.CLI_PATH)
    r = run_cli("search", "--query-file", "x.fq")
    assert r.returncode == 255 and "[TAXOR SEARCH ERROR]" in r.stderr and "--index-file" in r.stderr
    r = run_cli("search", "--index-file", str(tmp_path / "nope.hixf"), "--output-file", str(tmp_path / "o.tsv"))
    assert r.returncode == 255 and "The following index file does not exist" in r.stderr and "checking input ... " in r.stdout
    open(tmp_path / "a.hixf", "wb").write(b"x")
    r = run_cli("search", "--index-file", str(tmp_path / "a.hixf"), "--query-file", str(tmp_path / "nope.fq"))
    assert r.returncode == 255 and "The following query file does not exist" in r.stderr
    for bad in (["--threads", "33"], ["--threads", "0"], ["--percentage", "1.5"], ["--error-rate", "-0.1"], ["--error-rate", "abc"],
                ["--no-such-option", "1"]):
        r = run_cli("search", "--index-file", str(tmp_path / "a.hixf"), *bad)
        assert r.returncode == 255 and "[TAXOR SEARCH ERROR]" in r.stderr, bad
    assert run_cli("frobnicate").returncode == 255
    assert run_cli("search", "--help").returncode == 0


def fastq_text(records):
    return "".join(f"@{rid}\n{seq}\n+\n{'I' * len(seq)}\n" for rid, seq in records)


def fasta_text(records, width=70):
    out = []
    for rid, seq in records:
        out.append(f">{rid}\n")
        out.extend(seq[i:i + width] + "\n" for i in range(0, len(seq), width))
        if not seq:
            out.append("\n")
    return "".join(out)


@pytest.mark.gpu
def test_cli_end_to_end_result_file(tmp_path, oracle, built_libs):
    """index file -> reads file -> result file, byte for byte what the reference flow writes (oracle restatement)."""
    ds = H.make_dataset(oracle, n_genomes=36, genome_len=40_000, t_max=8)
    species = tools.default_species(ds.hixf.n_user_bins)
    idx_path = tmp_path / "synthetic.hixf"
    tools.write_hixf(idx_path, ds.hixf, k=22, s=12, t=5, use_syncmer=True, window_size=20, species=species)
    rng = np.random.default_rng(3)
    reads = H.make_reads(ds, rng.integers(200, 9000, 400), err=0.04)
    records = []
    for i in range(reads.n):
        seq = "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))
        if i % 7 == 0:                                         # IUPAC codes and lower case collapse like seqan3::dna4
            seq = seq[:50] + "NNRYKMnnacgtu"[: max(0, len(seq) - 50)] + seq[63:]
        if i % 11 == 0:
            seq = seq.lower()
        records.append((f"read_{i} runid=abc ch={i % 512}", seq))
    records += [("too_short", "ACGTACGT"), ("exactly_k", "ACGTTGCAAGGCTTAACCGGTT"), ("empty", "")]
    fq, fa, fqgz = tmp_path / "r.fastq", tmp_path / "r.fasta", tmp_path / "r.fastq.gz"
    open(fq, "w").write(fastq_text(records))
    open(fa, "w").write(fasta_text(records))
    with gzip.open(fqgz, "wt") as f:
        f.write(fastq_text(records))
    from tests.test_ingest import _write_bgzf
    fqbgzf = tmp_path / "r.bgzf.fastq.gz"
    _write_bgzf(fqbgzf, fastq_text(records).encode(), block=20000)      # blocked gzip: inflated block-parallel
    expect = H.oracle_tsv(oracle, ds.arrays, species, records, k=22, s=12, t=5, use_syncmer=True, error_rate=0.1)
    assert expect.count("\n") > len(records)                   # several reads report more than one reference
    for q in (fq, fa, fqgz, fqbgzf):
        out = tmp_path / (q.name + ".tsv")
        r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(q), "--output-file", str(out), "--error-rate", "0.1",
                    "--threads", "4")
        assert r.returncode == 0, r.stderr
        assert "checking input ... done!" in r.stdout and "use syncmer model" in r.stdout
        assert "CPU time  : " in r.stdout and "Peak RSS  : " in r.stdout
        assert open(out).read() == expect, q.name
    # percentage model; two query files and two index files append under ONE header (taxor_search.cpp:343-358)
    out = tmp_path / "multi.tsv"
    r = run_cli("search", "--index-file", f"{idx_path},{idx_path}", "--query-file", f"{fq},{fa}", "--output-file", str(out),
                "--percentage", "0.15")
    assert r.returncode == 0, r.stderr
    assert "use percentage-model\t0.15" in r.stdout
    body = H.oracle_tsv(oracle, ds.arrays, species, records, k=22, s=12, t=5, use_syncmer=True, percentage=0.15, header=False)
    assert open(out).read() == H.HEADER + body * 4
    # SeqAn3 reader semantics the reference inherits: blanks after '>' are not part of the id; blanks and digits inside FASTA
    # sequence lines are dropped before the alphabet check
    rid, seq = records[5]
    odd = tmp_path / "odd.fasta"
    chunks = [seq[i:i + 60] for i in range(0, len(seq), 60)]
    open(odd, "w").write(">  \t" + rid + "\n" + "\n".join(f"{10 * i + 1:>6} " + " ".join(c[j:j + 10] for j in range(0, len(c), 10))
                                                          for i, c in enumerate(chunks)) + "\n")
    r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(odd), "--output-file", str(tmp_path / "odd.tsv"),
                "--error-rate", "0.1")
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "odd.tsv").read() == H.oracle_tsv(oracle, ds.arrays, species, [records[5]], k=22, s=12, t=5,
                                                             use_syncmer=True, error_rate=0.1)
    # an illegal character aborts loudly
    bad = "ACGT!ACGTACGTACGTACGTACGTACGT"
    open(tmp_path / "bad.fq", "w").write(f"@r1\n{bad}\n+\n{'I' * len(bad)}\n")
    r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(tmp_path / "bad.fq"), "--output-file", str(tmp_path / "b.tsv"))
    assert r.returncode == 255 and "illegal nucleotide" in r.stderr


@pytest.mark.gpu
def test_cli_kmer_mode_index(tmp_path, oracle, built_libs):
    ds = H.make_dataset(oracle, n_genomes=12, genome_len=30_000, k=20, s=0, t=0, use_syncmer=False, t_max=4)
    species = tools.default_species(ds.hixf.n_user_bins)
    idx_path = tmp_path / "kmer.hixf"
    tools.write_hixf(idx_path, ds.hixf, k=20, s=0, t=0, use_syncmer=False, window_size=20, species=species)
    reads = H.make_reads(ds, np.random.default_rng(5).integers(500, 6000, 60), err=0.01)
    records = [(f"r{i}", "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))) for i in range(reads.n)]
    open(tmp_path / "r.fq", "w").write(fastq_text(records))
    out = tmp_path / "o.tsv"
    r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(tmp_path / "r.fq"), "--output-file", str(out),
                "--error-rate", "0.02")
    assert r.returncode == 0 and "use kmer-model" in r.stdout, r.stderr
    assert open(out).read() == H.oracle_tsv(oracle, ds.arrays, species, records, k=20, s=0, t=0, use_syncmer=False, window_size=20,
                                            error_rate=0.02)


@pytest.mark.gpu
def test_cli_minimiser_index(tmp_path, oracle, built_libs):
    """an index built with --window-size > --kmer-size: minimiser hashing, FracMinHash threshold banner and values"""
    ds = H.make_dataset(oracle, n_genomes=12, genome_len=30_000, k=20, s=0, t=0, use_syncmer=False, t_max=4, window_size=28)
    species = tools.default_species(ds.hixf.n_user_bins)
    idx_path = tmp_path / "minimiser.hixf"
    tools.write_hixf(idx_path, ds.hixf, k=20, s=0, t=0, use_syncmer=False, window_size=28, species=species)
    reads = H.make_reads(ds, np.random.default_rng(6).integers(300, 6000, 80), err=0.01)
    records = [(f"r{i}", "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))) for i in range(reads.n)]
    records += [("short", "ACGTACGT"), ("empty", "")]
    open(tmp_path / "r.fa", "w").write(fasta_text(records, width=60))
    out = tmp_path / "o.tsv"
    r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(tmp_path / "r.fa"), "--output-file", str(out),
                "--error-rate", "0.02")
    assert r.returncode == 0 and "use frac minhash" in r.stdout, r.stderr
    got = open(out).read()
    assert got == H.oracle_tsv(oracle, ds.arrays, species, records, k=20, s=0, t=0, use_syncmer=False, window_size=28, error_rate=0.02)
    assert got.count("\t") > 6 * len(records)                   # reads really hit


@pytest.mark.gpu
def test_cli_two_gpus_same_bytes_as_one(tmp_path, oracle, built_libs):
    """reads sharded over two GPUs (one context per device, index uploaded once and cloned device to device, ordered writer):
    the result file is byte-identical to the single-GPU file and to the oracle's; `--gpus 0,0` (two contexts on one device)
    exercises the same multi-context path on a one-GPU box, `--gpus 2` the real thing where two devices exist
    (taxor_search.cpp:214,308-311: reads are independent, only the output is shared)"""
    import torch
    ds = H.make_dataset(oracle, n_genomes=60, genome_len=40_000, t_max=16)
    species = tools.default_species(ds.hixf.n_user_bins)
    idx_path = tmp_path / "two.hixf"
    tools.write_hixf(idx_path, ds.hixf, k=22, s=12, t=5, use_syncmer=True, window_size=20, species=species)
    rng = np.random.default_rng(12)
    reads = H.make_reads(ds, rng.integers(200, 9000, 3000), err=0.04)
    records = [(f"read_{i}", "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))) for i in range(reads.n)]
    fq = tmp_path / "r.fastq"
    open(fq, "w").write(fastq_text(records))
    expect = H.oracle_tsv(oracle, ds.arrays, species, records, k=22, s=12, t=5, use_syncmer=True, error_rate=0.1)
    outs = {}
    specs = ["1", "0,0"] + (["2"] if torch.cuda.device_count() >= 2 else [])
    for spec in specs:
        out = tmp_path / f"o_{spec.replace(',', '_')}.tsv"
        r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(fq), "--output-file", str(out), "--error-rate", "0.1",
                    "--threads", "4", "--gpus", spec)
        assert r.returncode == 0, r.stderr
        outs[spec] = open(out).read()
    assert outs["1"] == expect
    for spec in specs[1:]:
        assert outs[spec] == outs["1"], spec


@pytest.mark.gpu
def test_profile_head_from_gpu_results_equals_reference_on_the_result_file(tmp_path, oracle, reference, built_libs):
    """SURVEY 8(f) rank 3 end to end: the hits of a GPU search, handed to txr_profile_add_batch in memory and filtered, equal
    what the REFERENCE's own taxor_profile.cpp (compiled in place) makes of the result file the CLI wrote for the same reads"""
    import ctypes as C
    if not hasattr(reference.lib, "ref_profile_prefilter"):
        pytest.skip("oracle/_ref predates the profile shim")
    ds = H.make_dataset(oracle, n_genomes=48, genome_len=40_000, t_max=8)
    species = tools.default_species(ds.hixf.n_user_bins)
    idx_path = tmp_path / "p.hixf"
    tools.write_hixf(idx_path, ds.hixf, k=22, s=12, t=5, use_syncmer=True, window_size=20, species=species)
    rng = np.random.default_rng(5)
    reads = H.make_reads(ds, rng.integers(200, 9000, 2500), err=0.06)
    ids = [f"read_{i if i % 40 else max(i - 1, 0)} ch={i % 9}" for i in range(reads.n)]          # some ids repeat, all carry a space
    records = [(ids[i], "".join("ACGT"[c] for c in capi.unpack_codes(reads, i))) for i in range(reads.n)]
    fq = tmp_path / "r.fastq"
    open(fq, "w").write(fastq_text(records))
    out = tmp_path / "o.tsv"
    r = run_cli("search", "--index-file", str(idx_path), "--query-file", str(fq), "--output-file", str(out), "--error-rate", "0.1")
    assert r.returncode == 0, r.stderr
    with capi.Context(0) as ctx:
        H.upload(ctx, ds)
        ctx.set_params(k=22, s=12, t=5, use_syncmer=True, window_size=20, error_rate=0.1)
        res = ctx.search_raw(reads.words.ctypes.data, reads.word_off.ctypes.data, reads.length.ctypes.data, reads.n)
        prof = capi.Profile()
        half = reads.n // 2                                       # two batches: a view of the first half, then of the rest
        first = capi.Result(half, res.hash_count, res.threshold, res.hit_begin, res.user_bin, res.count, res.keep)
        prof.add_result(first, ids[:half], reads.length[:half], species)
        u32, u64 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        rest = capi.Result(reads.n - half, C.cast(C.addressof(res.hash_count.contents) + 4 * half, u32),
                           C.cast(C.addressof(res.threshold.contents) + 8 * half, u64),
                           C.cast(C.addressof(res.hit_begin.contents) + 8 * half, u64), res.user_bin, res.count, res.keep)
        prof.add_result(rest, ids[half:], reads.length[half:], species)
        assert prof.text() == reference.profile_prefilter(out, 0)
        prof.filter(3)
        got = prof.text()
        prof.close()
    assert got == reference.profile_prefilter(out, 3)
    assert got.count("\nH\tGCF_") > 500
