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
