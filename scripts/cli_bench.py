#!/usr/bin/env python
"""End-to-end timing of the drop-in `taxor search` binary (read ingest and index load):
.hixf file -> FASTQ file -> result file, with the phase times the binary prints under TAXOR_TIMING=1.

Uses the index bench.py cached in /dev/shm (run bench.py first, or pass --genome-len for a smaller one), writes it as
a real .hixf, simulates reads into a FASTQ file and runs the CLI with 1 and with N pack threads.  One JSON line."""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def write_bgzf(src, dst, block=65280):
    """bgzip-style blocked gzip (BGZF): one gzip member per 64 KiB block, block size in the 'BC' extra field"""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def member(c):
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        data = co.compress(c) + co.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 18 + len(data) + 8 - 1) + data
                + struct.pack("<II", zlib.crc32(c) & 0xffffffff, len(c)))
    with open(src, "rb") as f, open(dst, "wb") as out, ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        while True:
            big = f.read(block * 4096)
            if not big:
                break
            for m in ex.map(member, [big[i:i + block] for i in range(0, len(big), block)]):
                out.write(m)
        out.write(member(b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--genome-len", type=int, default=40_000_000)
    ap.add_argument("--t-max", type=int, default=64)
    ap.add_argument("--cache", default="/dev/shm/taxor_b200_bench")
    ap.add_argument("--work", default="/dev/shm/taxor_b200_cli")
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--gz", action="store_true", help="also time a gzip-compressed copy of the reads")
    ap.add_argument("--bgzf", action="store_true", help="also time a BGZF (bgzip-style blocked gzip) copy of the reads")
    ap.add_argument("--bz2", action="store_true", help="also time a bzip2 copy of the first --bz2-reads reads (libbz2 is one stream, ~100 MB/s)")
    ap.add_argument("--bz2-reads", type=int, default=20_000)
    ap.add_argument("--gpus", default="1", help="--gpus of the CLI; a comma list of counts runs every one (e.g. 1,2,8)")
    args = ap.parse_args()
    from taxor_b200 import capi, tools
    import taxor_b200

    d, done = bench.index_cache_paths(args)
    if not os.path.exists(done):
        raise SystemExit(f"no cached index in {d}: run bench.py with the same --genomes/--genome-len first")
    ix = bench.LoadedIndex(d)
    os.makedirs(args.work, exist_ok=True)
    hixf = os.path.join(args.work, "bench.hixf")
    t0 = time.time()
    if not os.path.exists(hixf):
        tools.write_hixf(hixf, ix, k=bench.K, s=bench.S, t=bench.T, use_syncmer=True, window_size=20)
    t_write = time.time() - t0

    genomes, lens = bench.make_genomes(args)
    fq = os.path.join(args.work, f"reads_{args.reads}.fastq")
    if not os.path.exists(fq):
        words, off, ln, _ = tools.simulate_reads(genomes, lens, np.full(args.reads, args.read_len, np.uint32), 0.05, seed=4242)
        reads = capi.PackedReads(words, off, ln)
        lut = np.frombuffer(b"ACGT", dtype=np.uint8)
        qual = b"I" * args.read_len
        with open(fq, "wb") as f:
            for i in range(args.reads):
                f.write(b"@read_%d runid=synthetic ch=%d\n" % (i, i % 512))
                f.write(lut[capi.unpack_codes(reads, i)].tobytes())
                f.write(b"\n+\n")
                f.write(qual[: int(ln[i])])
                f.write(b"\n")
    runs = {}
    files = [("fastq", fq)]
    if args.gz:
        gz = fq + ".gz"
        if not os.path.exists(gz):
            subprocess.run(f"gzip -1 -c {fq} > {gz}", shell=True, check=True)
        files.append(("fastq.gz", gz))
    if args.bgzf:
        bg = fq + ".bgzf.gz"
        if not os.path.exists(bg):
            write_bgzf(fq, bg)
        files.append(("fastq.bgzf", bg))
    bz = None
    if args.bz2:
        import bz2
        from concurrent.futures import ThreadPoolExecutor
        bz = fq + f".first{args.bz2_reads}.bz2"
        if not os.path.exists(bz):
            # the first reads of the FASTQ, compressed as concatenated streams of ~64 k lines (what pbzip2 writes)
            with open(fq, "rb") as f:
                lines = [f.readline() for _ in range(4 * args.bz2_reads)]
            chunks = [b"".join(lines[i:i + 65536]) for i in range(0, len(lines), 65536)]
            with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex, open(bz, "wb") as out:
                for c in ex.map(lambda x: bz2.compress(x, 1), chunks):
                    out.write(c)

    def one(tag, path, th, gpus, extra_env=None, n_reads=None):
        out = os.path.join(args.work, "out.tsv")
        env = dict(os.environ, TAXOR_TIMING="1", **(extra_env or {}))
        t0 = time.time()
        r = subprocess.run([taxor_b200.CLI_PATH, "search", "--index-file", hixf, "--query-file", path, "--output-file", out,
                            "--error-rate", "0.1", "--threads", str(th), "--gpus", str(gpus)], capture_output=True, text=True, env=env)
        wall = time.time() - t0
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        m = re.search(r"index load ([\d.e+-]+) s, upload to \d+ GPU\(s\) ([\d.e+-]+) s \(([\d.e+-]+) GB\), ingest\+search\+write ([\d.e+-]+) s", r.stderr)
        load, upload, gb, search = (float(x) for x in m.groups())
        return {"wall_s": round(wall, 2), "index_load_s": load, "index_upload_s": upload, "index_GB": gb, "gpus": gpus,
                "upload_GBps_per_copy": round(gb / max(upload, 1e-9), 1),
                "ingest_search_write_s": search, "file_GB": round(os.path.getsize(path) / 1e9, 2),
                "Mbases_per_s_search_phase": round((n_reads or args.reads) * args.read_len / search / 1e6),
                "hit_lines": sum(1 for line in open(out) if "\t-\t-\t" not in line) - 1}
    gpu_counts = [int(x) for x in str(args.gpus).split(",")]
    for tag, path in files:
        for th in sorted({1, args.threads}):
            runs[f"{tag}_threads{th}"] = one(tag, path, th, gpu_counts[0])
    # index start-up (load_index + replication): the staged multi-threaded upload against the plain
    # pageable copy, and one upload + device-to-device clones against the number of GPUs
    runs["fastq_serial_upload"] = one("fastq", fq, args.threads, gpu_counts[0], {"TXR_UPLOAD_THREADS": "1"})
    for g in gpu_counts[1:]:
        runs[f"fastq_gpus{g}"] = one("fastq", fq, args.threads, g)
    if bz:
        runs[f"fastq.bz2_first{args.bz2_reads}_threads{args.threads}"] = one("fastq.bz2", bz, args.threads, gpu_counts[0], n_reads=args.bz2_reads)
    print(json.dumps({"what": "taxor search CLI end to end (file -> file), 1 GPU", "reads": args.reads, "read_len": args.read_len,
                      "hixf_write_s": round(t_write, 1), "host_cores": os.cpu_count(), "runs": runs}))


if __name__ == "__main__":
    main()
