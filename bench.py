#!/usr/bin/env python
"""bench.py -- `taxor search` hot-path throughput on B200 (BASELINE.json metric: Mbases/s, syncmer hash + HIXF query).

A "step" is one pass of the hot path over this rank's read set (configs[1]: 1 M synthetic 10 kb reads per GPU
against a synthetic 1,000-genome HIXF, k=22 s=12).  Reads are sharded across ranks, the index is replicated,
there is no data-path collective (scaling = weak: per-GPU work is fixed).

  value  : reads resident in HBM when the timed region starts (kernels only), CUDA events on the launch stream
  e2e    : the same through txr_search() from pinned HOST buffers, H2D + kernels + D2H + host ordering per step
  roofline    : kernel #2 (IXF probe/count), algorithmic bytes / CUDA-event duration vs the measured HBM peak
  cpu_baseline: the CPU oracle port (OpenMP, all host cores) on a bounded sample of the same reads (rank 0, N=1)

`--impl reference` times the reference's CPU algorithm (oracle port; see DESIGN.md for why not oracle/_ref)
on the host cores for the same workload, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "taxor search Mbases/s (syncmer hash+HIXF query)"
UNIT = "Mbases/s"
K, S, T = 22, 12, 5

# The headline workload is configs[1].  The other BASELINE configs can be run at full size with --workload (extra
# evidence lines under profiles/, not the driver's bench line): their parameters override the defaults below.
WORKLOADS = {
    "configs1": dict(label="configs[1]"),
    # configs[3]: plain canonical 20-mers (no syncmers, duplicates kept, k-mer-model threshold), viral-scale index,
    # read lengths log-uniform in [1 kb, 50 kb]
    "kmer": dict(label="configs[3]", k=20, s=0, t=0, use_syncmer=False, genomes=5000, genome_len=150_000, min_genome_len=60_000,
                 t_max=128, read_len_range=(1000, 50_000), read_error=0.03, error_rate=0.05),
    # configs[4] shape at 1/10 of its size: 20,000 genomes under t_max=64 give a three-level hierarchy
    "deep": dict(label="configs[4] shape (three-level hierarchy, 10 GB)", genomes=20000, genome_len=2_000_000, t_max=64),
    # configs[4] shape with the WIDE root real GTDB layouts have (taxor build picks t_max up to 4096, taxor_build.cpp:173-187):
    # 102,400 user bins, root of 4096 technical bins, 16-bin IXFs below it (three levels), ~1/20 of the 100 GB the config names
    # (what one box hashes and peels in a few minutes); 250k reads per step because every read streams 11 MB of root rows
    "gtdb": dict(label="configs[4] shape at ~1/20 scale (102,400 user bins, 4096-bin root, three levels)", genomes=102_400,
                 genome_len=100_000, min_genome_len=30_000, t_max=4096, t_max_lower=16, reads=250_000),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs1", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("TAXOR_BENCH_READS", 1_000_000)), help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--genomes", type=int, default=int(os.environ.get("TAXOR_BENCH_GENOMES", 1000)))
    ap.add_argument("--genome-len", type=int, default=int(os.environ.get("TAXOR_BENCH_GENOME_LEN", 40_000_000)),
                    help="mean genome length; 40 Mbp x 1,000 genomes gives the ~10 GB HIXF configs[1] names")
    ap.add_argument("--t-max", type=int, default=64)
    ap.add_argument("--t-max-lower", type=int, default=0, help="bin budget of the IXFs below the root (0: same as --t-max)")
    ap.add_argument("--read-error", type=float, default=0.05)
    ap.add_argument("--error-rate", type=float, default=0.10, help="taxor search --error-rate (see DESIGN.md workload)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-reads", type=int, default=int(os.environ.get("TAXOR_BENCH_BATCH_READS", 0)),
                    help="reads per internal batch (0: library default)")
    ap.add_argument("--cache", default=os.environ.get("TAXOR_BENCH_CACHE", "/dev/shm/taxor_b200_bench"))
    args = ap.parse_args()
    try:                                     # the cache lives in /dev/shm when that is writable, else in the temp dir
        os.makedirs(args.cache, exist_ok=True)
        if not os.access(args.cache, os.W_OK):
            raise OSError
    except OSError:
        import tempfile
        args.cache = os.path.join(tempfile.gettempdir(), "taxor_b200_bench")
        os.makedirs(args.cache, exist_ok=True)
    w = WORKLOADS[args.workload]
    args.k, args.s, args.t, args.use_syncmer = w.get("k", K), w.get("s", S), w.get("t", T), w.get("use_syncmer", True)
    args.window = 20 if args.use_syncmer else args.k
    args.min_genome_len = w.get("min_genome_len", 20_000)
    args.read_len_range = w.get("read_len_range")
    args.label = w["label"]
    explicit = {a.split("=")[0] for a in sys.argv[1:] if a.startswith("--")}
    for key in ("genomes", "genome_len", "t_max", "t_max_lower", "read_error", "error_rate", "reads"):
        if key in w and "--" + key.replace("_", "-") not in explicit:
            setattr(args, key, w[key])
    return args


# ----------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------
def genome_lengths(n, mean, seed=7, min_len=20_000):
    """RefSeq-ABFV-like size mix: log-uniform over a 16x range around the mean (small 'viral' to large genomes)."""
    rng = np.random.default_rng(seed)
    x = np.exp(rng.uniform(np.log(0.25), np.log(4.0), n))
    x = x / x.mean() * mean
    return np.maximum(x.astype(np.int64), min_len)


def make_genomes(args, rank=0, barrier=None):
    """Packed synthetic genomes, generated once (rank 0) into a /dev/shm file that every rank maps: at the default
    size they are 10 GB, too much to hold once per rank."""
    from taxor_b200 import tools
    lens = genome_lengths(args.genomes, args.genome_len, min_len=getattr(args, "min_genome_len", 20_000))
    nw = np.array([tools.packed_words(int(x)) for x in lens], dtype=np.uint64)
    off = np.zeros(args.genomes + 1, dtype=np.uint64)
    off[1:] = np.cumsum(nw)
    d = os.path.join(args.cache, f"genomes_g{args.genomes}_l{args.genome_len}_m{getattr(args, 'min_genome_len', 20_000)}")
    path, done = os.path.join(d, "words.bin"), os.path.join(d, "DONE")
    if rank == 0 and not os.path.exists(done):
        os.makedirs(d, exist_ok=True)
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint64, shape=(int(off[-1]),))
        T = tools.tlib()
        from concurrent.futures import ThreadPoolExecutor

        def gen(g):
            T.txs_genome(1000 + g, int(lens[g]), mm.ctypes.data + 8 * int(off[g]))
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            list(ex.map(gen, range(args.genomes)))
        mm.flush()
        del mm
        open(done, "w").close()
    while not os.path.exists(done):
        time.sleep(0.5)
    if barrier is not None:
        barrier()
    mm = np.load(path, mmap_mode="r")
    return [mm[int(off[g]):int(off[g + 1])] for g in range(args.genomes)], lens


def index_cache_paths(args):
    tag = f"g{args.genomes}_l{args.genome_len}_m{getattr(args, 'min_genome_len', 20_000)}_t{args.t_max}_k{getattr(args, 'k', K)}s{getattr(args, 's', S)}"
    if getattr(args, "t_max_lower", 0):
        tag += f"_tl{args.t_max_lower}"
    d = os.path.join(args.cache, tag)
    return d, os.path.join(d, "DONE")


def unpack_2bit(words, n):
    """2-bit packed bases (first base most significant) -> codes 0..3, pure numpy (the reference arm loads no CUDA library)."""
    w = np.asarray(words[: (int(n) + 31) // 32], dtype=np.uint64)
    shifts = np.arange(62, -2, -2, dtype=np.uint64)
    return ((w[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[: int(n)]


def hash_genomes_cpu(args, genomes, lens):
    """compute_hashes on the host through the oracle (reference arm: no GPU code anywhere on its path)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle
    o = Oracle()

    def one(g):
        c = unpack_2bit(genomes[g], lens[g])      # raw emissions: the builder sorts and de-duplicates every user bin anyway
        return o.syncmer_hashes_raw(c, args.k, args.s, args.t) if args.use_syncmer else o.kmer_hashes(c, args.k)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        return list(ex.map(one, range(len(genomes)))), 0


def hash_genomes_gpu(args, genomes, lens, ctx):
    """build-side hashing on the GPU (txr_hash_user_bins: genomes cut into independent segments, per-genome distinct sets)"""
    from taxor_b200 import capi
    ub = []
    ctx.set_params(k=args.k, s=args.s, t=args.t, use_syncmer=args.use_syncmer, window_size=args.window, error_rate=args.error_rate)
    a, n_seg = 0, 0
    while a < len(genomes):
        b, bases = a, 0
        while b < len(genomes) and (b == a or bases + int(lens[b]) <= 2_500_000_000):
            bases += int(lens[b])
            b += 1
        part = genomes[a:b]
        nw = np.array([len(w) for w in part], dtype=np.uint64)
        off = np.zeros(len(part), dtype=np.uint64)
        off[1:] = np.cumsum(nw)[:-1]
        seqs = capi.PackedReads(np.concatenate(part), off, np.asarray(lens[a:b], dtype=np.uint32))
        o, h, ns = ctx.hash_user_bins(seqs, np.arange(len(part), dtype=np.uint32), len(part))
        n_seg += ns
        for i in range(len(part)):
            ub.append(h[int(o[i]):int(o[i + 1])])
        a = b
    return ub, n_seg


def build_index_arrays(args, genomes, lens, ctx):
    """Hashes every genome (GPU: txr_hash_user_bins; reference arm: the CPU oracle -- same sets), then lays out + peels the
    HIXF on the CPU (tooling)."""
    from taxor_b200 import tools
    t0 = time.time()
    ub, n_seg = hash_genomes_gpu(args, genomes, lens, ctx) if ctx is not None else hash_genomes_cpu(args, genomes, lens)
    t1 = time.time()
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1, which would make the CPU peeling take half an hour
    hx = tools.BuiltHixf(ub, t_max=args.t_max, seed=1, inplace=True, threads=os.cpu_count() or 1, t_max_lower=getattr(args, "t_max_lower", 0))
    del ub
    t2 = time.time()
    info = dict(hash_s=round(t1 - t0, 2), hash_segments=n_seg, build_s=round(t2 - t1, 2), n_ixf=hx.n_ixf, fp_bytes=hx.fp_bytes,
                n_hashes=int(hx.n_keys), reseeds=hx.reseeds)
    return hx, info


def save_index(hx, d, info):
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, "meta.npy"), np.stack([hx.seed, hx.bins, hx.tbins, hx.seg_len]))
    np.save(os.path.join(d, "bin_off.npy"), hx.bin_off)
    np.save(os.path.join(d, "next.npy"), hx.next_ixf_id)
    np.save(os.path.join(d, "ub.npy"), hx.bin_to_ub)
    sizes = np.array([x.size for x in hx.data], dtype=np.uint64)
    np.save(os.path.join(d, "sizes.npy"), sizes)
    with open(os.path.join(d, "fp.bin"), "wb") as f:
        for x in hx.data:
            f.write(memoryview(x))
    with open(os.path.join(d, "info.json"), "w") as f:
        json.dump(dict(info, n_user_bins=hx.n_user_bins), f)
    open(os.path.join(d, "DONE"), "w").close()


class LoadedIndex:
    def __init__(self, d):
        m = np.load(os.path.join(d, "meta.npy"))
        self.seed, self.bins, self.tbins, self.seg_len = m[0], m[1], m[2], m[3]
        self.bin_off = np.load(os.path.join(d, "bin_off.npy"))
        self.next_ixf_id = np.load(os.path.join(d, "next.npy"))
        self.bin_to_ub = np.load(os.path.join(d, "ub.npy"))
        sizes = np.load(os.path.join(d, "sizes.npy"))
        fp = np.memmap(os.path.join(d, "fp.bin"), dtype=np.uint8, mode="r")
        self.data, at = [], 0
        for s in sizes:
            self.data.append(fp[at:at + int(s)])
            at += int(s)
        with open(os.path.join(d, "info.json")) as f:
            self.info = json.load(f)
        self.n_user_bins = self.info["n_user_bins"]
        self.n_ixf = len(self.seed)
        # tree depth: children are created after their parent, so one forward pass over the merged bins is enough
        level = np.ones(self.n_ixf, dtype=np.int64)
        for i in range(self.n_ixf):
            a, b = int(self.bin_off[i]), int(self.bin_off[i + 1])
            ch = self.next_ixf_id[a:b][self.bin_to_ub[a:b] < 0]
            level[ch] = level[i] + 1
        self.depth = int(level.max())
        self.fp_bytes = int(sizes.sum())


def upload_index(ctx, ix):
    data = [np.ascontiguousarray(x) for x in ix.data]
    ctx.upload_index(ix.seed, ix.bins, ix.tbins, ix.seg_len, data, ix.bin_off, ix.next_ixf_id, ix.bin_to_ub, ix.n_user_bins)


class HostArray:
    """plain numpy stand-in for capi.PinnedArray (reference arm)"""

    def __init__(self, n, dtype):
        self.array = np.zeros(int(n), dtype=dtype)
        self.nbytes = self.array.nbytes
        self.ptr = self.array.ctypes.data

    def free(self):
        self.array = None


def make_reads(args, genomes, lens, rank, pinned=True):
    """This rank's shard of the read set, simulated straight into pinned host memory."""
    from taxor_b200 import tools
    if pinned:
        from taxor_b200 import capi
        Arr = capi.PinnedArray
    else:
        Arr = HostArray
    n = args.reads
    if args.read_len_range:
        lo, hi = args.read_len_range
        rl = np.exp(np.random.default_rng(1234 + rank).uniform(np.log(lo), np.log(hi), n)).astype(np.uint32)
    else:
        rl = np.full(n, args.read_len, np.uint32)
    pin = Arr(int(((rl.astype(np.uint64) + 31) // 32 + 1).sum()), np.uint64)
    pin.array[:] = 0
    world = int(os.environ.get("WORLD_SIZE", 1))
    words, off, ln, src = tools.simulate_reads(genomes, lens, rl, args.read_error,
                                               seed=42 + 7919 * rank, out_words=pin.array,
                                               threads=max(1, (os.cpu_count() or 1) // world))
    off_pin = Arr(n, np.uint64)
    off_pin.array[:] = off
    len_pin = Arr(n, np.uint32)
    len_pin.array[:] = ln
    return pin, off_pin, len_pin


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ----------------------------------------------------------------------------------------------------------
_ORACLES = {}


def cpu_oracle(rebuild=True):
    """The CPU arm's library: the oracle restatement compiled ON THIS BOX with -O3 -march=native (oracle/Makefile `native`;
    BASELINE.md section 2).  Falls back to the portable test build if the box has no compiler.  Under torchrun rank 0 builds it
    (rebuild=True) before a barrier and the other ranks load the file (rebuild=False)."""
    if "o" not in _ORACLES:
        from oracle.oracle import Oracle
        try:
            _ORACLES["o"], _ORACLES["build"] = Oracle(native=True, rebuild=rebuild), "-O3 -march=native -ffp-contract=off, built on this box"
        except Exception as e:                                     # no compiler on the box: say so in the line
            _ORACLES["o"], _ORACLES["build"] = Oracle(), f"portable -march=x86-64-v2 build (native build failed: {type(e).__name__})"
    return _ORACLES["o"]


def cpu_run(args, ix, pin_words, off, ln, n_sample, threads, tuned=False):
    """Times oracle.search_batch (restated CPU path, OpenMP over reads) on the first n_sample reads.  tuned: bulk_count with
    software prefetch and 64-bin SIMD compares (same results) instead of the plain restatement of the reference's loop."""
    from oracle.oracle import HixfArrays
    o = cpu_oracle()
    o.set_tuned(tuned)
    arrays = HixfArrays(np.ascontiguousarray(ix.seed), np.ascontiguousarray(ix.bins), np.ascontiguousarray(ix.tbins),
                        np.ascontiguousarray(ix.seg_len), [np.ascontiguousarray(x) for x in ix.data],
                        np.ascontiguousarray(ix.bin_off), np.ascontiguousarray(ix.next_ixf_id), np.ascontiguousarray(ix.bin_to_ub))
    h = o.make_hixf(arrays)
    codes = np.empty(int(ln[:n_sample].astype(np.uint64).sum()), dtype=np.uint8)
    coff = np.zeros(n_sample + 1, dtype=np.uint64)
    at = 0
    for i in range(n_sample):
        c = unpack_2bit(pin_words[int(off[i]):], ln[i])
        codes[at:at + len(c)] = c
        at += len(c)
        coff[i + 1] = at
    t0 = time.perf_counter()
    res = o.search_batch(h, codes, coff, k=args.k, s=args.s, t=args.t, use_syncmer=args.use_syncmer, window_size=args.window,
                         error_rate=args.error_rate,
                         threads=threads, want_raw=False)
    dt = time.perf_counter() - t0
    o.set_tuned(False)
    return at / dt / 1e6, dt, res


def cpu_baseline_block(args, ix, pin, off_pin, len_pin, n_reads, cores, seconds):
    """Both CPU variants on a bounded sample: `port` (the reference's loop restated) and `port_tuned` (prefetch + SIMD).  The
    block's own value is the TUNED one -- the fair yardstick; returns (block, oracle answers for the sample, sample size)."""
    n_s = min(n_reads, 1000)
    v, dt, _ = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_s, cores, tuned=True)
    n_s = int(min(n_reads, max(200, n_s * seconds / max(dt, 1e-3))))
    vt, dtt, ora = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_s, cores, tuned=True)
    n_p = int(min(n_s, max(200, n_s * 0.5 * seconds / max(dtt * 3, 1e-3))))   # the plain port is slower: a shorter sample
    vp, dtp, _ = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_p, cores, tuned=False)
    flags = cpu_oracle().lib.orc_build_flags()
    block = {"value": vt, "unit": UNIT, "cores": cores, "kind": "port", "variant": "port_tuned",
             "sample": f"first {n_s} of the {n_reads} reads, {dtt:.1f} s (restated CPU path, OpenMP over reads, bulk_count with software "
                       f"prefetch + 64-bin SIMD compares; not the reference binary)",
             "build": _ORACLES.get("build"), "simd": "avx512bw" if flags & 1 else "avx2" if flags & 2 else "scalar",
             "variants": {"port_tuned": {"value": vt, "sample_reads": n_s, "seconds": dtt},
                          "port": {"value": vp, "sample_reads": n_p, "seconds": dtp,
                                   "note": "the reference's per-value loop restated as is (no prefetch, scalar compare)"}}}
    return block, ora, n_s


def random_access_block(stage, args, ix, gather):
    """kernel #2 as random row reads per second: 3 per probed hash (per visited IXF), from the device counters of the timed region"""
    if stage["query_ms"] <= 0:
        return None
    # query_bytes = sum Hp * (3 * tbins + 8): the probes of every IXF weigh in with their own row width; the access count
    # is exact for the root-dominated workloads (one row width) and a lower bound otherwise
    row = int(ix.tbins[0])
    probes = stage["query_bytes"] / (3.0 * row + 8.0)
    rate = 3.0 * probes / (stage["query_ms"] / 1e3) / 1e9
    out = {"achieved_G_rows_per_s": rate, "row_bytes_root": row}
    if gather:
        ref = gather["useful_GBps"] / row                                         # G rows/s of the pure one-row-per-probe gather at this width
        out.update({"microbench_G_rows_per_s": ref, "vs_microbench": rate / ref, "microbench_source": gather["source"],
                    "note": "the microbenchmark gathers ONE random row per element; kernel #2 reads three rows of three segments per "
                            "probe and beats it, so it is a yardstick, not a ceiling"})
    return out


def compare_with_oracle(g, ora, n_s):
    """GPU answers (capi.SearchResult) for the first n_s reads against oracle.search_batch for the same reads."""
    nh = int(g.hit_begin[n_s])
    keep = g.keep[:nh]
    kept_per_read = np.concatenate([[0], np.cumsum(keep)])[g.hit_begin[: n_s + 1].astype(np.int64)]
    parity = {"reads": n_s,
              "hash_count_mismatches": int(np.count_nonzero(g.hash_count[:n_s] != ora["hash_count"])),
              "threshold_mismatches": int(np.count_nonzero(g.threshold[:n_s] != ora["threshold"])),
              "hit_offsets_equal": bool(np.array_equal(kept_per_read.astype(np.uint64), ora["hit_off"])),
              "hit_user_bins_equal": bool(np.array_equal(g.user_bin[:nh][keep], ora["ub"])),
              "hit_counts_equal": bool(np.array_equal(g.count[:nh][keep], ora["cnt"])),
              "reported_hits": int(len(ora["ub"]))}
    parity["ok"] = (parity["hash_count_mismatches"] == 0 and parity["threshold_mismatches"] == 0 and parity["hit_offsets_equal"]
                    and parity["hit_user_bins_equal"] and parity["hit_counts_equal"])
    parity["mismatches"] = (parity["hash_count_mismatches"] + parity["threshold_mismatches"] + (not parity["hit_offsets_equal"])
                            + (not parity["hit_user_bins_equal"]) + (not parity["hit_counts_equal"]))
    return parity


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = os.cpu_count() or 1

    if args.impl == "reference" and rank != 0:
        return 0

    import __graft_entry__ as ge
    import taxor_b200
    if not (os.path.exists(taxor_b200.LIB_PATH) and os.path.exists(taxor_b200.TOOLS_PATH)):
        ge.build()
    reference = args.impl == "reference"   # CPU only: no CUDA context, no CUDA library, no torch.cuda on this arm

    dist, ctx, capi = None, None, None
    if not reference:
        from taxor_b200 import capi
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist
            import datetime
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=60))
        ctx = capi.Context(local_rank)
    genomes, lens = make_genomes(args, rank, dist.barrier if dist is not None else None)

    # ---- index: rank 0 builds (hashing on the GPU -- on the CPU for the reference arm -- then CPU peeling), cached in /dev/shm ----
    d, done = index_cache_paths(args)
    if rank == 0 and not os.path.exists(done):
        hx, info = build_index_arrays(args, genomes, lens, ctx)
        save_index(hx, d, info)
        hx.close()
    if rank == 0 and not reference and not args.no_cpu_baseline:
        cpu_oracle(rebuild=True)             # the native CPU-arm library: built once per box, by rank 0, before the barrier
    while not os.path.exists(done):          # the other ranks sleep (no spinning collective) while rank 0 builds
        time.sleep(1.0)
    if dist is not None:
        dist.barrier()
    if rank != 0 and not args.no_cpu_baseline:
        cpu_oracle(rebuild=False)
    ix = LoadedIndex(d)
    if not reference:
        upload_index(ctx, ix)
        ctx.set_params(k=args.k, s=args.s, t=args.t, use_syncmer=args.use_syncmer, window_size=args.window, error_rate=args.error_rate)

    pin, off_pin, len_pin = make_reads(args, genomes, lens, rank, pinned=not reference)
    n_reads = args.reads
    bases_per_step = int(len_pin.array.astype(np.uint64).sum())
    reads = capi.PackedReads(pin.array, off_pin.array, len_pin.array) if not reference else None

    mode = f"k={args.k} s={args.s} t={args.t} syncmers" if args.use_syncmer else f"canonical {args.k}-mers (no syncmers, duplicates kept)"
    rl_txt = (f"{args.read_len_range[0]}-{args.read_len_range[1]} bp (log-uniform, mean {bases_per_step // n_reads})" if args.read_len_range
              else f"{args.read_len} bp")
    workload = {"workload": f"{args.label}: {args.genomes} synthetic genomes (mean {args.genome_len} bp, 16x log-uniform size mix), "
                            f"{mode}, HIXF t_max={args.t_max} ({ix.n_ixf} IXFs, {ix.depth} levels, {ix.fp_bytes / 1e9:.2f} GB fingerprints); "
                            f"{n_reads} reads x {rl_txt} per GPU, {args.read_error:.0%} error, --error-rate {args.error_rate}",
                "reads_per_gpu": n_reads, "read_len": args.read_len, "index_bytes": ix.fp_bytes,
                "parallelism": f"reads sharded over {world} GPU(s), index replicated, no collective",
                "l2": "inputs larger than L2 (packed reads and index each exceed 126 MB); no flush needed"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if reference:
        n_s = min(n_reads, 2000)
        v, dt, _ = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_s, cores, tuned=True)
        # bounded sample per step: all warm-up + timed steps together stay within about two minutes
        per_step = min(args.cpu_seconds, 100.0 / max(args.steps + args.warmup, 1))
        n_s = int(min(n_reads, max(200, n_s * per_step / max(dt, 1e-3))))
        vals = []
        for i in range(args.warmup + args.steps):
            v, dt, _ = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_s, cores, tuned=True)
            if i >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([x[0] for x in vals]))
        n_p = max(200, n_s // 4)
        v_plain, dt_plain, _ = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_p, cores, tuned=False)
        flags = cpu_oracle().lib.orc_build_flags()
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(np.mean([x[1] for x in vals]) * 1e3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "variant": "port_tuned",
                                 "sample": f"{n_s} of the {n_reads} reads per step (restated CPU path, OpenMP over reads, bulk_count with "
                                           f"software prefetch + 64-bin SIMD compares; not the reference binary)",
                                 "build": _ORACLES.get("build"), "simd": "avx512bw" if flags & 1 else "avx2" if flags & 2 else "scalar",
                                 "variants": {"port_tuned": {"value": value, "sample_reads": n_s},
                                              "port": {"value": v_plain, "sample_reads": n_p, "seconds": dt_plain,
                                                       "note": "the reference's per-value loop restated as is (no prefetch, scalar compare)"}}},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    # torchrun exports OMP_NUM_THREADS=1; the library's host pass (ordering the hits of a batch, OpenMP over reads) would then run
    # on one thread per rank.  Give every rank its share of the host cores, as a user launching N processes would.
    omp_threads = int(os.environ.get("TAXOR_BENCH_OMP", 0)) or max(1, cores // max(world, 1))
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(omp_threads)
    except OSError:
        omp_threads = None
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # (1) device-resident: kernels only (all kernels run on one stream in batch order, so the per-stage CUDA events do not overlap)
    mean_len = max(1, bases_per_step // max(n_reads, 1))
    batch_kw = dict(max_batch_reads=args.batch_reads, max_batch_bases=int(args.batch_reads * mean_len * 1.1)) if args.batch_reads else {}
    ctx.configure(n_slots=int(os.environ.get("TAXOR_BENCH_RESIDENT_SLOTS", 2)), **batch_kw)
    h = ctx.upload_reads(reads)
    for _ in range(args.warmup):
        ctx.search_resident(h, fetch=False)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"hash_ms": 0.0, "dedup_ms": 0.0, "query_ms": 0.0, "query_bytes": 0, "hash_bytes": 0, "launches": 0, "query_launches": 0,
             "probe_launches": 0, "skipped_hashes": 0,
             "query_items": 0, "n_hashes": 0}
    with torch.cuda.stream(stream):
        ev0.record()
        for _ in range(args.steps):
            ctx.search_resident(h, fetch=False)
            tm = ctx.timing()
            for kname in ("hash_ms", "dedup_ms", "query_ms", "query_bytes", "query_items"):
                stage[kname] += tm[kname]
            stage["launches"] += tm["hash_launches"] + tm["dedup_launches"] + tm["query_launches"]
            stage["query_launches"] += tm["query_launches"]
            stage["probe_launches"] += tm["probe_launches"]
            stage["skipped_hashes"] += tm["skipped_hashes"]
        ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    resident_ms = float(ms.item())

    # (1b) kernel #2 on its own: the same reads with ONE pipeline slot, which turns the overlap off (the hash stage of the next
    # batch no longer runs beside the probes) -- the stand-alone duration of the kernel the roofline is about.  Not part of `value`.
    ctx.configure(n_slots=1, **batch_kw)
    ctx.search_resident(h, fetch=False)
    alone = {"hash_ms": 0.0, "dedup_ms": 0.0, "query_ms": 0.0, "query_bytes": 0, "probe_launches": 0, "total_ms": 0.0}
    alone_steps = max(1, min(args.steps, 2))
    with torch.cuda.stream(stream):
        ev2a, ev2b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2a.record()
        for _ in range(alone_steps):
            ctx.search_resident(h, fetch=False)
            tm = ctx.timing()
            for kname in ("hash_ms", "dedup_ms", "query_ms", "query_bytes", "probe_launches"):
                alone[kname] += tm[kname]
        ev2b.record()
    torch.cuda.synchronize()
    alone["total_ms"] = ev2a.elapsed_time(ev2b)
    ctx.free_reads(h)

    # (2) end to end from pinned host buffers through txr_search (3 pipeline slots)
    ctx.configure(n_slots=int(os.environ.get("TAXOR_BENCH_E2E_SLOTS", 4)), **batch_kw)
    for _ in range(max(1, min(args.warmup, 2))):
        ctx.search_raw(pin.ptr, off_pin.ptr, len_pin.ptr, n_reads)
    barrier()
    n_hits = 0
    with torch.cuda.stream(stream):
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = ctx.search_raw(pin.ptr, off_pin.ptr, len_pin.ptr, n_reads)
            n_hits = int(r.hit_begin[n_reads])
        ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = torch.tensor([max(ev0.elapsed_time(ev1), 0.0)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms.item())
    tm_e2e = ctx.timing()
    gpu_result = capi.SearchResult(r) if not args.no_cpu_baseline else None  # copy: the view dies with the next call

    # per-rank view of the end-to-end step (the 8-GPU curve is bounded by the host side, not by the kernels): this rank's own
    # step time, the H2D time its batches saw, and -- all ranks at once -- the pure pinned-host -> HBM copy rate of the
    # same buffer, i.e. what the host's memory system gives N GPUs pulling 2.5 GB each at the same moment
    e2e_own_ms = max(ev0.elapsed_time(ev1), 0.0) / args.steps
    probe = torch.empty(pin.nbytes, dtype=torch.uint8, device="cuda")
    hsrc = torch.frombuffer(memoryview(pin.array), dtype=torch.uint8)
    direct = bool(hsrc.is_pinned())           # cudaHostAlloc'ed by the library: torch sees it as pinned and copies straight from it
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev2.record()
        for _ in range(3):
            probe.copy_(hsrc, non_blocking=True)
        ev3.record()
    barrier()
    h2d_gbs = 3 * pin.nbytes / max(ev2.elapsed_time(ev3), 1e-3) / 1e6 if direct else float("nan")
    del probe, hsrc
    rank_stats = torch.tensor([e2e_own_ms, tm_e2e["h2d_ms"], h2d_gbs], device="cuda", dtype=torch.float64)
    if dist is not None:
        allst = [torch.zeros_like(rank_stats) for _ in range(world)]
        dist.all_gather(allst, rank_stats)
        rank_stats_all = [[round(float(v), 2) for v in t.tolist()] for t in allst]
    else:
        rank_stats_all = [[round(float(v), 2) for v in rank_stats.tolist()]]

    # answers checked on the hardware of EVERY rank: a bounded sample of this rank's own shard against the oracle
    # (N > 1: a short sample per rank, mismatches summed over ranks; N == 1: the cpu_baseline sample below serves)
    multi_parity = None
    if world > 1 and not args.no_cpu_baseline:
        n_chk = min(n_reads, 3000)
        _, _, ora_chk = cpu_run(args, ix, pin.array, off_pin.array, len_pin.array, n_chk, max(1, cores // world), tuned=True)
        bad = compare_with_oracle(gpu_result, ora_chk, n_chk)
        t = torch.tensor([bad["mismatches"], n_chk, bad["reported_hits"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        multi_parity = {"reads": int(t[1].item()), "ranks": world, "mismatches": int(t[0].item()), "reported_hits": int(t[2].item()),
                        "ok": int(t[0].item()) == 0, "note": "every rank compares the first reads of ITS shard with the CPU oracle"}

    total_bases = bases_per_step * world
    value = total_bases * args.steps / (resident_ms / 1e3) / 1e6
    e2e_value = total_bases * args.steps / (e2e_ms / 1e3) / 1e6

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        q_gbs = stage["query_bytes"] / (stage["query_ms"] / 1e3) / 1e9 if stage["query_ms"] > 0 else 0.0
        # DRAM traffic of kernel #2 per launch: measured once per round with `ncu --set full` on this workload and
        # stored as a ratio to the algorithmic bytes (profiles/query_traffic.json, written from the capture)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "query_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_algorithmic_byte"] * stage["query_bytes"] / max(stage["probe_launches"], 1)
            traffic_src = tj.get("source")
        except Exception:
            pass
        # the north star's yardstick: random-gather throughput of this GPU for rows of this width, measured by
        # taxor_b200/csrc/microbench/gather_bench2.cu (useful GB/s; every random row costs a whole 128-byte DRAM line)
        gather = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_gather_bench2.json")) as f:
                gb = json.load(f)
            row = int(ix.tbins[0])
            for e in gb["qualifiers_GBps"]:
                if e["row_bytes"] == row:
                    gather = {"row_bytes": row, "useful_GBps": e["nc_noalloc"], "frac": q_gbs / e["nc_noalloc"],
                              "source": "profiles/r1_gather_bench2.json (random rows of the root IXF's width, 8 GiB table)"}
        except Exception:
            pass
        wide = int(ix.tbins[0]) > 512
        roof = {"bound": "hbm", "kernel": ("ixf_query_large_kernel (CTA per item, root rows of %d bytes) + ixf_query_small_kernel below" % int(ix.tbins[0])) if wide
                else "ixf_query_small_kernel (kernel #2, all HIXF levels of a batch)",
                "achieved": q_gbs, "peak": peak, "unit": "GB/s", "frac": q_gbs / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "traffic": traffic, "traffic_source": traffic_src, "random_gather_ceiling": gather,
                "avg_launch_ms": stage["query_ms"] / max(stage["probe_launches"], 1),
                "launches_per_step": stage["probe_launches"] / args.steps,
                "algorithmic_bytes_per_launch": stage["query_bytes"] / max(stage["probe_launches"], 1),
                "algorithmic_bytes_per_step": stage["query_bytes"] / args.steps,
                "bytes_note": "bytes of the probes actually issued (Hp*3*tbins + 8*Hp per visited IXF); probes saved by the exact "
                              "early exit are NOT counted",
                "early_exit_skipped_hashes_per_step": stage["skipped_hashes"] / args.steps,
                # rows narrower than two DRAM lines are bound by random ACCESSES per second, not by bytes (the .L2::64B variant of
                # the gather microbenchmark moves half the bytes in the same time, profiles/r1_gather_bench2_ncu.txt): three row
                # reads per probed hash, against the microbenchmark's rate for rows of this width
                "random_access": random_access_block(stage, args, ix, gather),
                "stage_ms_per_step": {"hash": stage["hash_ms"] / args.steps, "dedup": stage["dedup_ms"] / args.steps,
                                      "query": stage["query_ms"] / args.steps},
                "stage_note": "CUDA-event durations per stage, summed over the batches of a step.  With the overlap on (default where it "
                              "applies) the hash + dedup kernels of batch i+1 run BESIDE the probe kernels of batch i, so the stages "
                              "overlap in time, do not add up to ms_per_step, and each is slower than alone; `standalone` is the same "
                              "workload with one pipeline slot (no overlap): the duration kernel #2 has when it owns the GPU",
                "standalone": {"achieved": alone["query_bytes"] / (alone["query_ms"] / 1e3) / 1e9 if alone["query_ms"] > 0 else None,
                               "frac": alone["query_bytes"] / (alone["query_ms"] / 1e3) / 1e9 / peak if alone["query_ms"] > 0 else None,
                               "unit": "GB/s", "steps": alone_steps, "ms_per_step_serial_schedule": alone["total_ms"] / alone_steps,
                               "stage_ms_per_step": {"hash": alone["hash_ms"] / alone_steps, "dedup": alone["dedup_ms"] / alone_steps,
                                                     "query": alone["query_ms"] / alone_steps},
                               "avg_launch_ms": alone["query_ms"] / max(alone["probe_launches"], 1)}}
        cpu, parity = None, multi_parity
        if world == 1 and not args.no_cpu_baseline:
            cpu, ora, n_s = cpu_baseline_block(args, ix, pin, off_pin, len_pin, n_reads, cores, args.cpu_seconds)
            # parity at full size, for free: the oracle's answers for the CPU sample against the GPU's answers for the same reads
            parity = compare_with_oracle(gpu_result, ora, n_s)
        line = {"metric": METRIC if args.use_syncmer else METRIC.replace("syncmer hash", "k-mer hash"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": resident_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": workload,
                "reads_per_s": n_reads * world * args.steps / (resident_ms / 1e3),
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": int(pin.nbytes + off_pin.nbytes + len_pin.nbytes + 8 * (n_reads + 1)),
                        "d2h_bytes_per_step": int(4 * n_reads + 12 * n_hits + 576 * (-(-n_reads // (args.batch_reads or 262144)) + 2)),
                        "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps, "hits_per_step": n_hits,
                        "stage_ms_per_step_overlapped": {kk: tm_e2e[kk] for kk in ("h2d_ms", "hash_ms", "dedup_ms", "query_ms", "d2h_ms")},
                        "per_rank": {"columns": ["e2e_ms_per_step", "h2d_ms_last_step", "concurrent_h2d_GBps"], "rows": rank_stats_all,
                                     "note": "concurrent_h2d_GBps: all ranks copy their pinned read buffer to HBM at the same moment, nothing "
                                             "else running -- the host-side ceiling of the e2e step at this N (bytes per step / this rate "
                                             "is the time the copies need even when perfectly overlapped)"}},
                "gpu_launches": int(stage["launches"]),
                "roofline": roof, "cpu_baseline": cpu, "parity_at_scale": parity, "clocks": clocks,
                "index_build": ix.info, "host_cores": cores, "host_threads_per_rank": omp_threads}
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            print("PARITY FAILURE at full size: " + json.dumps(parity), file=sys.stderr)
            sys.exit(3)
    pin.free()
    off_pin.free()
    len_pin.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
