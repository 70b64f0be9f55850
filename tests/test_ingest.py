"""CPU: the CLI's parallel-ingest scanner (taxor_b200/csrc/ingest.cpp) against a plain Python statement of the record
semantics of seqan3::sequence_file_input<fields<id, seq>> (taxor_search.cpp:181-182): FASTA and FASTQ, multi-line
records, CRLF, blank lines, gzip, buffers far smaller than a record (carry-over between raw buffers)."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from taxor_b200 import tools


def _dump(path, target, tmp_path):
    T = tools.tlib()
    T.txs_ingest_dump.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
    T.txs_ingest_dump.restype = C.c_int64
    out = str(tmp_path / "dump.tsv")
    n = T.txs_ingest_dump(str(path).encode(), target, out.encode())
    if n < 0:
        raise RuntimeError(T.txs_last_error().decode())
    recs = []
    with open(out, "rb") as f:
        for line in f.read().split(b"\n")[:-1]:
            i, s = line.split(b"\t")
            recs.append((i, s))
    assert len(recs) == n
    return recs


def _dump_mapped(path, seg, tmp_path):
    T = tools.tlib()
    T.txs_ingest_dump_mapped.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.POINTER(C.c_uint64)]
    T.txs_ingest_dump_mapped.restype = C.c_int64
    out = str(tmp_path / "dump_m.tsv")
    resc = C.c_uint64()
    n = T.txs_ingest_dump_mapped(str(path).encode(), seg, out.encode(), C.byref(resc))
    if n < 0:
        raise RuntimeError(T.txs_last_error().decode())
    recs = []
    with open(out, "rb") as f:
        for line in f.read().split(b"\n")[:-1]:
            i, s = line.split(b"\t")
            recs.append((i, s))
    assert len(recs) == n
    return recs, resc.value


def _rand_seq(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGTNacgtRYKM", dtype=np.uint8), n).tobytes())


def _wrap(seq, width, eol):
    if width == 0 or len(seq) == 0:
        return seq + eol
    return b"".join(seq[i:i + width] + eol for i in range(0, len(seq), width))


def _make(rng, fmt, n, eol=b"\n", width=0, final_eol=True, blank=False):
    recs, blob = [], b""
    for i in range(n):
        L = int(rng.choice([0, 1, 5, 60, 61, 500, 4000, 20000]))
        name = f"read_{i} len={L} some description".encode()
        seq = _rand_seq(rng, L)
        recs.append((name, seq))
        if fmt == "fasta":
            blob += b">" + name + eol + (_wrap(seq, width, eol) if L else b"")
        else:
            qual = bytes(rng.integers(33, 74, L, dtype=np.uint8).tobytes())      # may contain '@' and '+' and '>'
            blob += b"@" + name + eol + _wrap(seq, width, eol) + b"+" + eol + _wrap(qual, width, eol)
        if blank and i % 3 == 0:
            blob += eol
    if not final_eol and blob.endswith(eol):
        blob = blob[:-len(eol)]
    return recs, blob


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
@pytest.mark.parametrize("eol,width,final_eol,blank", [(b"\n", 0, True, False), (b"\r\n", 0, True, True), (b"\n", 60, False, False),
                                                      (b"\r\n", 70, True, True)])
def test_scanner_matches_record_semantics(tmp_path, fmt, eol, width, final_eol, blank):
    rng = np.random.default_rng(hash((fmt, width, final_eol)) % 2**32)
    if fmt == "fastq" and width and not final_eol:
        final_eol = True
    recs, blob = _make(rng, fmt, 120, eol, width, final_eol, blank)
    if fmt == "fastq" and width:
        # wrapped FASTQ: a quality line may start with '@' or '+'; the sequence lines never do
        pass
    p = tmp_path / f"x.{fmt}"
    p.write_bytes(blob)
    for target in (1 << 20, 4096, 64):              # 64 bytes: every record spans many refills
        assert _dump(p, target, tmp_path) == recs
    # mapped path: byte ranges far smaller than a record up to larger than the file; wrapped FASTQ defeats the 4-line
    # guess on purpose -- the in-order check has to rescan, the records must still be exact
    for seg in (37, 1000, 50_000, 1 << 30):
        got, rescans = _dump_mapped(p, seg, tmp_path)
        assert got == recs, seg
        if fmt == "fasta" or width == 0:
            assert rescans == 0, (seg, rescans)       # the guess is exact for FASTA and for 4-line FASTQ
    gz = tmp_path / f"x.{fmt}.gz"
    with gzip.open(gz, "wb") as f:
        f.write(blob)
    assert _dump(gz, 5000, tmp_path) == recs


def test_scanner_errors(tmp_path):
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nIII\n")             # quality shorter than the sequence
    with pytest.raises(RuntimeError, match="quality"):
        _dump(bad, 4096, tmp_path)
    bad.write_bytes(b"@r1\nACGT\n")                     # no '+' line
    with pytest.raises(RuntimeError, match=r"\+"):
        _dump(bad, 4096, tmp_path)
    bad.write_bytes(b"ACGT\n")
    with pytest.raises(RuntimeError, match="does not start"):
        _dump(bad, 4096, tmp_path)
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    assert _dump(empty, 4096, tmp_path) == []
    with pytest.raises(RuntimeError, match="cannot open"):
        _dump(tmp_path / "missing.fa", 4096, tmp_path)


def test_mapped_guess_is_not_fooled_by_quality_lines(tmp_path):
    """quality strings that start with '@' (and '+' lines that could be headers) right at range boundaries"""
    rng = np.random.default_rng(9)
    recs, blob = [], b""
    for i in range(400):
        L = int(rng.integers(30, 90))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L).tobytes())
        qual = b"@" + bytes(rng.choice(np.frombuffer(b"@+>I", dtype=np.uint8), L - 1).tobytes())
        name = b"@r%d" % i                       # header text itself starts with '@' after the marker
        recs.append((name, seq))
        blob += b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n"
    p = tmp_path / "tricky.fq"
    p.write_bytes(blob)
    for seg in (64, 97, 128, 1009):
        got, rescans = _dump_mapped(p, seg, tmp_path)
        assert got == recs and rescans == 0, (seg, rescans)
    assert _dump(p, 4096, tmp_path) == recs


def _write_bgzf(path, blob, block=60000, eof_marker=True):
    """BGZF as bgzip writes it: gzip members of <= 64 KiB with the 'BC' extra field holding the block size - 1"""
    import struct
    import zlib
    with open(path, "wb") as f:
        chunks = [blob[i:i + block] for i in range(0, len(blob), block)] + ([b""] if eof_marker else [])
        for c in chunks:
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            data = co.compress(c) + co.flush()
            bsize = 18 + len(data) + 8
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1))
            f.write(data)
            f.write(struct.pack("<II", zlib.crc32(c) & 0xffffffff, len(c)))


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_bgzf_blocks_are_inflated_in_parallel_and_in_order(tmp_path, fmt):
    rng = np.random.default_rng(77 if fmt == "fasta" else 78)
    recs, blob = _make(rng, fmt, 400, b"\n", 0 if fmt == "fastq" else 70)
    p = tmp_path / f"x.{fmt}.gz"
    _write_bgzf(p, blob)
    assert gzip.open(p, "rb").read() == blob             # it is a valid multi-member gzip file
    for target in (1 << 22, 1 << 16, 70_000, 1000, 64):  # buffers larger / smaller than one 60 kB block
        assert _dump(p, target, tmp_path) == recs, target
    _write_bgzf(p, blob, block=4096, eof_marker=False)    # many small blocks, no EOF marker block
    assert _dump(p, 1 << 20, tmp_path) == recs
    # a flipped payload byte is caught by the block CRC (or by inflate)
    raw = bytearray(p.read_bytes())
    raw[len(raw) // 2] ^= 0x5A
    bad = tmp_path / "bad.gz"
    bad.write_bytes(bytes(raw))
    with pytest.raises(RuntimeError, match="corrupt|start"):
        _dump(bad, 1 << 20, tmp_path)


@pytest.mark.parametrize("fmt", ["fasta", "fastq"])
def test_bzip2_input_through_the_runtime_bound_library(tmp_path, fmt):
    """bzip2 reads (the reference takes them through SeqAn3, taxor_search.cpp:181-182): libbz2.so.1.0 is bound with dlopen --
    single stream, concatenated streams (pbzip2 / cat), small buffers, and a damaged file"""
    import bz2
    rng = np.random.default_rng(91 if fmt == "fasta" else 92)
    recs, blob = _make(rng, fmt, 300, b"\n", 0 if fmt == "fastq" else 60)
    p = tmp_path / f"x.{fmt}.bz2"
    p.write_bytes(bz2.compress(blob, 1))
    for target in (1 << 22, 70_000, 1000):
        assert _dump(p, target, tmp_path) == recs, target
    cut = blob.index(b"\n", len(blob) // 2) + 1
    while blob[cut:cut + 1] not in (b">", b"@") or (fmt == "fastq" and blob[cut - 1:cut] != b"\n"):
        cut = blob.index(b"\n", cut) + 1                   # any line start works: the streams are simply concatenated
        break
    p.write_bytes(bz2.compress(blob[:cut], 9) + bz2.compress(blob[cut:], 5))
    assert _dump(p, 1 << 20, tmp_path) == recs
    raw = bytearray(p.read_bytes())
    raw[len(raw) // 3] ^= 0x5A
    bad = tmp_path / "bad.bz2"
    bad.write_bytes(bytes(raw))
    with pytest.raises(RuntimeError, match="corrupt|start|truncated"):
        _dump(bad, 1 << 20, tmp_path)
    (tmp_path / "short.bz2").write_bytes(bz2.compress(blob, 9)[:-20])
    with pytest.raises(RuntimeError, match="corrupt|truncated"):
        _dump(tmp_path / "short.bz2", 1 << 20, tmp_path)


def test_ingest_fuzz_under_sanitizers(tmp_path):
    """tests/fuzz/ingest_fuzz.cpp built with -fsanitize=address,undefined: mutated FASTA/FASTQ through the streaming and the
    mapped byte-range path; no crash, no sanitizer report, and both paths accept/reject and parse alike."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "ingest_fuzz")
    build = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fopenmp", "-o", exe,
                            os.path.join(root, "tests/fuzz/ingest_fuzz.cpp"), os.path.join(root, "taxor_b200/csrc/ingest.cpp"),
                            os.path.join(root, "taxor_b200/csrc/inflate_fast.cpp"), os.path.join(root, "taxor_b200/csrc/gzip_parallel.cpp"), "-lz", "-ldl", "-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower():
        pytest.skip("compiler without sanitizer runtimes")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, str(tmp_path / "f.txt"), "3000", "11"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "fuzz ok" in run.stdout, run.stdout[-500:] + run.stderr[-3000:]


def test_fast_inflater_against_zlib_under_sanitizers(tmp_path):
    """tests/fuzz/inflate_fuzz.cpp built with -fsanitize=address,undefined: the ingest's own DEFLATE / gzip decoder against zlib --
    every block type, level and strategy, flush points, several members, reads of random size; the whole-buffer (BGZF) form;
    damaged and truncated streams; the carry-less-multiply CRC.  Both builds of the symbol loop (BMI2 and baseline)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "inflate_fuzz")
    build = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                            os.path.join(root, "tests/fuzz/inflate_fuzz.cpp"), os.path.join(root, "taxor_b200/csrc/inflate_fast.cpp"),
                            os.path.join(root, "taxor_b200/csrc/gzip_parallel.cpp"), "-lz", "-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower():
        pytest.skip("compiler without sanitizer runtimes")
    assert build.returncode == 0, build.stderr[-2000:]
    for isa, seed in (("", "21"), ("generic", "22")):
        run = subprocess.run([exe, "120", seed], capture_output=True, text=True, timeout=600, env=dict(os.environ, TAXOR_INFLATE_ISA=isa))
        assert run.returncode == 0 and "fuzz ok" in run.stdout and "runtime error" not in run.stderr, run.stdout[-500:] + run.stderr[-3000:]


def test_gzip_paths_agree(tmp_path):
    """single-member, multi-member and BGZF-less gzip through the mapped fast decoder and through zlib (TAXOR_GZIP=zlib, in a
    fresh process because the choice is read when the scanner opens the file): the same records"""
    import subprocess
    import sys
    rng = np.random.default_rng(11)
    recs = [(b"r%d some text" % i, _rand_seq(rng, int(rng.integers(0, 5000)))) for i in range(300)]
    blob = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in recs)
    one = tmp_path / "one.fq.gz"
    one.write_bytes(gzip.compress(blob, 6))
    cut = len(blob) // 2
    two = tmp_path / "two.fq.gz"
    two.write_bytes(gzip.compress(blob[:cut], 9) + gzip.compress(blob[cut:], 1) + b"\0" * 7)
    for path in (one, two):
        assert _dump(path, 1 << 16, tmp_path) == recs
        code = ("import sys; sys.path.insert(0, %r); from tests import test_ingest as T; import pathlib; "
                "r = T._dump(pathlib.Path(%r), 1 << 16, pathlib.Path(%r)); print(len(r), hash(tuple(r)))" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(path), str(tmp_path)))
        a = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, TAXOR_GZIP="zlib", PYTHONHASHSEED="0"))
        b = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONHASHSEED="0"))
        assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
        assert a.stdout == b.stdout and a.stdout.split()[0] == str(len(recs))


def test_gzip_single_stream_read_by_several_threads(tmp_path):
    """a .gz of more than 4 MB goes through the multi-threaded reader (pieces found by header search, decoded with a symbolic
    window, chained bit-exactly, CRC-checked); with 20 kB pieces this file is cut ~300 times.  Same records as the one-thread
    decoder and as zlib; a flipped byte in the middle is an error, not different records."""
    import subprocess
    import sys
    rng = np.random.default_rng(12)
    recs = []
    for i in range(1500):
        n = int(rng.integers(100, 9000))
        recs.append((b"read%d runid=%d" % (i, int(rng.integers(1 << 30))), _rand_seq(rng, n),
                     bytes((33 + rng.integers(0, 40, n)).astype(np.uint8).tobytes())))
    blob = b"".join(b"@" + i + b"\n" + s + b"\n+\n" + q + b"\n" for i, s, q in recs)
    path = tmp_path / "big.fq.gz"
    path.write_bytes(gzip.compress(blob, 1))
    assert path.stat().st_size > 4 << 20
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from tests import test_ingest as T; import pathlib, hashlib; "
            "r = T._dump(pathlib.Path(sys.argv[1]), 1 << 20, pathlib.Path(%r)); "
            "print(len(r), hashlib.md5(b'|'.join(i + b'/' + s for i, s in r)).hexdigest())" % (root, str(tmp_path)))
    outs = []
    for env in ({"TAXOR_GZIP": "zlib"}, {"TAXOR_GZIP": "serial"}, {"TAXOR_GZIP_PIECE": "20000", "OMP_NUM_THREADS": "4"},
                {"OMP_NUM_THREADS": "3"}):
        r = subprocess.run([sys.executable, "-c", code, str(path)], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert len(set(outs)) == 1 and outs[0].split()[0] == str(len(recs))
    import hashlib
    assert outs[0].split()[1] == hashlib.md5(b"|".join(i + b"/" + s for i, s, _ in recs)).hexdigest()
    bad = bytearray(path.read_bytes())
    bad[len(bad) // 2] ^= 0x10
    (tmp_path / "bad.fq.gz").write_bytes(bytes(bad))
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path / "bad.fq.gz")], capture_output=True, text=True,
                       env=dict(os.environ, TAXOR_GZIP_PIECE="20000", OMP_NUM_THREADS="4"))
    # (a streaming reader hands bytes out before the member's CRC is due, so the damage may surface as a malformed record first)
    assert r.returncode != 0 and "RuntimeError" in r.stderr, r.stderr[-1500:]
