"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built, the reference's own sources
compiled in place (oracle/_ref/libtaxor_ref.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg.  Nothing under taxor_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtaxor_ref.so")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/libtaxor_ref.so (only where /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference/src") and not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)


class _IXF(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("bins", C.c_uint64), ("tbins", C.c_uint64), ("seg_len", C.c_uint64),
                ("data", C.c_void_p), ("rows", C.c_uint64)]


class _HIXF(C.Structure):
    _fields_ = [("n_ixf", C.c_uint64), ("ixf", C.POINTER(_IXF)), ("bin_off", C.c_void_p),
                ("next_ixf_id", C.c_void_p), ("bin_to_ub", C.c_void_p)]


class _Thr(C.Structure):
    _fields_ = [("kind", C.c_int), ("kmer_size", C.c_uint8), ("percentage", C.c_double), ("error_rate", C.c_double)]


class _Params(C.Structure):
    _fields_ = [("k", C.c_int), ("s", C.c_int), ("t", C.c_int), ("use_syncmer", C.c_int),
                ("window_size", C.c_uint32), ("scaling", C.c_uint16), ("percentage", C.c_double),
                ("error_rate", C.c_double)]


@dataclass
class HixfArrays:
    """Plain-array view of an HIXF shared by oracle, reference shim and (separately declared) product C-ABI."""
    seed: np.ndarray      # u64[n_ixf]
    bins: np.ndarray      # u64[n_ixf]   counting-vector size
    tbins: np.ndarray     # u64[n_ixf]   stored row width
    seg_len: np.ndarray   # u64[n_ixf]
    data: list            # list of u8 arrays, data[i].size == 3*seg_len[i]*tbins[i]
    bin_off: np.ndarray   # u64[n_ixf+1]
    next_ixf_id: np.ndarray  # i64[sum bins]
    bin_to_ub: np.ndarray    # i64[sum bins]
    rows: np.ndarray = None  # u64[n_ixf] slots per bin; None = 3*seg_len (binary-fuse scheme: (segments+2)*seg_len)

    @property
    def n_ixf(self) -> int:
        return len(self.seed)


NATIVE_SO = os.path.join(HERE, "liboracle_native.so")


def build_native() -> str:
    """`make native` on THIS machine (-O3 -march=native): the CPU-baseline build of bench.py.  Always rebuilt, so that a
    library compiled on another CPU never travels here."""
    subprocess.run(["make", "-C", HERE, "-B", "native"], check=True, capture_output=True)
    return NATIVE_SO


class Oracle:
    def __init__(self, native: bool = False, rebuild: bool = True) -> None:
        """native: the -march=native build made on THIS machine (bench.py's CPU arm); rebuild=False loads the file another
        process of the same machine has just built (the ranks of a torchrun job must not all run make on one file)."""
        build()
        L = self.lib = C.CDLL((build_native() if rebuild or not os.path.exists(NATIVE_SO) else NATIVE_SO) if native else ORACLE_SO)
        L.orc_set_tuned.restype = None
        L.orc_set_tuned.argtypes = [C.c_int]
        L.orc_build_flags.restype = C.c_int
        L.orc_wyhash_u64.restype = C.c_uint64
        L.orc_wyhash_u64.argtypes = [C.c_uint64]
        L.orc_adjust_seed.restype = C.c_uint64
        L.orc_adjust_seed.argtypes = [C.c_uint8]
        L.orc_dna4_rank.restype = C.c_int
        L.orc_dna4_rank.argtypes = [C.c_ubyte]
        L.orc_t_syncmer.restype = C.c_int
        L.orc_t_syncmer.argtypes = [C.c_int, C.c_int]
        L.orc_scaling_keep.restype = C.c_int
        L.orc_scaling_keep.argtypes = [C.c_uint64, C.c_uint16]
        for f in (L.orc_syncmer_hashes, L.orc_syncmer_hashes_raw):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, u64p, C.c_int64]
        L.orc_kmer_hashes.restype = C.c_int64
        L.orc_kmer_hashes.argtypes = [u8p, C.c_int64, C.c_int, C.c_uint64, u64p, C.c_int64]
        L.orc_minimiser_hashes.restype = C.c_int64
        L.orc_minimiser_hashes.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_uint64, u64p, C.c_int64]
        L.orc_threshold_init.restype = None
        L.orc_threshold_init.argtypes = [C.POINTER(_Thr), C.c_uint32, C.c_uint8, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_threshold_get.restype = C.c_uint64
        L.orc_threshold_get.argtypes = [C.POINTER(_Thr), C.c_uint64, C.c_double]
        L.orc_syncmer_match_ratio.restype = C.c_double
        L.orc_syncmer_match_ratio.argtypes = [C.c_uint64, C.c_double]
        L.orc_normal_cdf_inverse.restype = C.c_double
        L.orc_normal_cdf_inverse.argtypes = [C.c_double]
        L.orc_kmer_ci.restype = None
        L.orc_kmer_ci.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_ixf_slots.restype = None
        L.orc_ixf_slots.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint8)] + [C.POINTER(C.c_uint64)] * 3
        L.orc_set_ixf_scheme.restype = None
        L.orc_set_ixf_scheme.argtypes = [C.c_void_p]
        L.orc_ixf_bulk_count.restype = None
        L.orc_ixf_bulk_count.argtypes = [C.POINTER(_IXF), u64p, C.c_uint64, u32p]
        L.orc_hixf_bulk_contains.restype = C.c_int64
        L.orc_hixf_bulk_contains.argtypes = [C.POINTER(_HIXF), u64p, C.c_uint64, C.c_uint64, i64p, u32p, C.c_int64,
                                             C.POINTER(C.c_uint64)]
        L.orc_search_batch.restype = C.c_int
        L.orc_search_batch.argtypes = [C.POINTER(_HIXF), C.POINTER(_Params), u8p, u64p, C.c_uint64, u32p, u64p,
                                       u64p, i64p, u32p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.c_int]

    # ---- scalars
    def wyhash(self, x: int) -> int:
        return self.lib.orc_wyhash_u64(x)

    def adjust_seed(self, k: int) -> int:
        return self.lib.orc_adjust_seed(k)

    def dna4_rank(self, ch: str) -> int:
        return self.lib.orc_dna4_rank(ord(ch))

    def t_syncmer(self, k: int, s: int) -> int:
        return self.lib.orc_t_syncmer(k, s)

    def scaling_keep(self, h: int, scaling: int) -> bool:
        return bool(self.lib.orc_scaling_keep(h, scaling))

    # ---- hashing
    def _hash_call(self, fn, codes, *args):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        cap = max(16, len(codes))
        out = np.empty(cap, dtype=np.uint64)
        n = fn(codes, len(codes), *args, out, cap)
        assert n >= 0
        return out[:n].copy()

    def syncmer_hashes(self, codes, k, s, t):
        return self._hash_call(self.lib.orc_syncmer_hashes, codes, k, s, t)

    def syncmer_hashes_raw(self, codes, k, s, t):
        return self._hash_call(self.lib.orc_syncmer_hashes_raw, codes, k, s, t)

    def kmer_hashes(self, codes, k, seed=None):
        seed = self.adjust_seed(k) if seed is None else seed
        return self._hash_call(self.lib.orc_kmer_hashes, codes, k, seed)

    def minimiser_hashes(self, codes, k, w, seed=None):
        seed = self.adjust_seed(k) if seed is None else seed
        return self._hash_call(self.lib.orc_minimiser_hashes, codes, k, w, seed)

    # ---- thresholds
    def thresholder(self, window_size, kmer_size, percentage, error_rate, use_syncmer, fracminhash=False):
        t = _Thr()
        self.lib.orc_threshold_init(C.byref(t), window_size, kmer_size, percentage, error_rate, int(use_syncmer),
                                    int(fracminhash))
        return t

    def threshold_get(self, thr, count, scaling_factor=1.0):
        return self.lib.orc_threshold_get(C.byref(thr), count, scaling_factor)

    def syncmer_match_ratio(self, k, e):
        return self.lib.orc_syncmer_match_ratio(k, e)

    def normal_cdf_inverse(self, p):
        return self.lib.orc_normal_cdf_inverse(p)

    def kmer_ci(self, r, k, n, conf=0.95):
        lo, hi = C.c_uint64(), C.c_uint64()
        self.lib.orc_kmer_ci(r, k, n, conf, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    # ---- IXF / HIXF
    def ixf_slots(self, key, seed, seg_len):
        f = C.c_uint8()
        p = [C.c_uint64() for _ in range(3)]
        self.lib.orc_ixf_slots(key, seed, seg_len, C.byref(f), *[C.byref(x) for x in p])
        return f.value, p[0].value, p[1].value, p[2].value

    def set_tuned(self, on: bool) -> None:
        """prefetching / SIMD bulk_count (same results) -- the port_tuned CPU baseline"""
        self.lib.orc_set_tuned(int(on))

    def set_ixf_scheme(self, scheme=None):
        """Switch the (unpinned) probe arithmetic: None = the prototype's, else (slots, mix, fingerprint, rot1, rot2).
        Process-global; tests that change it restore it."""
        if scheme is None:
            self.lib.orc_set_ixf_scheme(None)
        else:
            self.lib.orc_set_ixf_scheme(np.array([int(x) for x in scheme], dtype=np.uint32).ctypes.data)

    def make_hixf(self, a: HixfArrays):
        ixfs = (_IXF * a.n_ixf)()
        for i in range(a.n_ixf):
            rows = 3 * int(a.seg_len[i]) if a.rows is None else int(a.rows[i])
            assert a.data[i].dtype == np.uint8 and a.data[i].size == rows * int(a.tbins[i])
            ixfs[i] = _IXF(int(a.seed[i]), int(a.bins[i]), int(a.tbins[i]), int(a.seg_len[i]), a.data[i].ctypes.data, rows)
        h = _HIXF(a.n_ixf, ixfs, a.bin_off.ctypes.data, a.next_ixf_id.ctypes.data, a.bin_to_ub.ctypes.data)
        h._keep = (ixfs, a)
        return h

    def ixf_bulk_count(self, h, ixf_idx, values):
        x = h.ixf[ixf_idx]
        counts = np.zeros(x.bins, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        self.lib.orc_ixf_bulk_count(C.byref(x), values, len(values), counts)
        return counts

    def bulk_contains(self, h, values, threshold):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        cap = 64
        while True:
            ub = np.empty(cap, dtype=np.int64)
            cnt = np.empty(cap, dtype=np.uint32)
            vb = C.c_uint64(0)
            n = self.lib.orc_hixf_bulk_contains(C.byref(h), values, len(values), threshold, ub, cnt, cap, C.byref(vb))
            if n >= 0:
                return ub[:n].copy(), cnt[:n].copy(), vb.value
            cap = -n

    def search_batch(self, h, codes, off, *, k, s, t, use_syncmer, window_size, scaling=1, percentage=-1.0,
                     error_rate=0.04, threads=0, want_raw=True):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        p = _Params(k, s, t, int(use_syncmer), window_size, scaling, percentage, error_rate)
        hc = np.zeros(n, dtype=np.uint32)
        thr = np.zeros(n, dtype=np.uint64)
        cap = max(1024, 8 * n)
        while True:
            hit_off = np.zeros(n + 1, dtype=np.uint64)
            ub = np.empty(cap, dtype=np.int64)
            cnt = np.empty(cap, dtype=np.uint32)
            if want_raw:
                raw_off = np.zeros(n + 1, dtype=np.uint64)
                rub = np.empty(cap, dtype=np.int64)
                rcnt = np.empty(cap, dtype=np.uint32)
                rargs = (raw_off.ctypes.data, rub.ctypes.data, rcnt.ctypes.data, cap)
            else:
                rargs = (None, None, None, 0)
            vb = C.c_uint64(0)
            rc = self.lib.orc_search_batch(C.byref(h), C.byref(p), codes, off, n, hc, thr, hit_off, ub, cnt, cap,
                                           *rargs, C.byref(vb), threads)
            if rc == 0:
                break
            cap = int(max(hit_off[n], raw_off[n] if want_raw else 0)) + 16
        res = dict(hash_count=hc, threshold=thr, hit_off=hit_off, ub=ub[: int(hit_off[n])].copy(),
                   cnt=cnt[: int(hit_off[n])].copy(), visited_bytes=vb.value)
        if want_raw:
            res.update(raw_off=raw_off, raw_ub=rub[: int(raw_off[n])].copy(), raw_cnt=rcnt[: int(raw_off[n])].copy())
        return res


class Reference:
    """The reference's own syncmer.cpp / hixf.hpp DFS / threshold models, compiled in place (oracle/_ref)."""

    def __init__(self) -> None:
        build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_seq_to_syncmers.restype = C.c_int64
        L.ref_seq_to_syncmers.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, u64p, C.c_int64]
        L.ref_wyhash_stub.restype = C.c_uint64
        L.ref_wyhash_stub.argtypes = [C.c_uint64]
        L.ref_hixf_new.restype = C.c_void_p
        L.ref_hixf_new.argtypes = [C.c_uint64, u64p, u64p, u64p, u64p, C.POINTER(C.c_void_p), u64p, i64p, i64p]
        L.ref_hixf_free.argtypes = [C.c_void_p]
        L.ref_bulk_contains.restype = C.c_int64
        L.ref_bulk_contains.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_uint64, i64p, u32p, C.c_int64]
        L.ref_threshold_new.restype = C.c_void_p
        L.ref_threshold_new.argtypes = [C.c_uint32, C.c_uint8, C.c_double, C.c_double, C.c_int, C.c_int]
        L.ref_threshold_free.argtypes = [C.c_void_p]
        L.ref_threshold_get.restype = C.c_uint64
        L.ref_threshold_get.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
        L.ref_syncmer_match_ratio.restype = C.c_double
        L.ref_syncmer_match_ratio.argtypes = [C.c_uint64, C.c_double]
        L.ref_normal_cdf_inverse.restype = C.c_double
        L.ref_normal_cdf_inverse.argtypes = [C.c_double]
        L.ref_kmer_ci.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_adjust_seed.restype = C.c_uint64
        L.ref_adjust_seed.argtypes = [C.c_uint8]
        if hasattr(L, "ref_profile_prefilter"):
            L.ref_profile_prefilter.restype = C.c_void_p
            L.ref_profile_prefilter.argtypes = [C.c_char_p, C.c_int]
            L.ref_profile_free.argtypes = [C.c_void_p]

    def profile_prefilter(self, search_file, stage: int = 3) -> str:
        """the reference's own parse_search_results + filter rounds (src/main/taxor_profile.cpp:93-462, compiled in place) on a
        result file; stage 0..3 = how many rounds; the table in the line format of txr_profile_text"""
        p = self.lib.ref_profile_prefilter(str(search_file).encode(), stage)
        if not p:
            raise RuntimeError("ref_profile_prefilter failed")
        try:
            return C.string_at(p).decode()
        finally:
            self.lib.ref_profile_free(p)

    def syncmer_hashes(self, codes, k, s, t):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        cap = max(16, len(codes))
        out = np.empty(cap, dtype=np.uint64)
        n = self.lib.ref_seq_to_syncmers(codes, len(codes), k, s, t, out, cap)
        assert n >= 0
        return out[:n].copy()

    def make_hixf(self, a: HixfArrays):
        ptrs = (C.c_void_p * a.n_ixf)(*[d.ctypes.data for d in a.data])
        h = self.lib.ref_hixf_new(a.n_ixf, a.seed, a.bins, a.tbins, a.seg_len, ptrs, a.bin_off, a.next_ixf_id, a.bin_to_ub)
        return h

    def free_hixf(self, h):
        self.lib.ref_hixf_free(h)

    def bulk_contains(self, h, values, threshold):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        cap = 64
        while True:
            ub = np.empty(cap, dtype=np.int64)
            cnt = np.empty(cap, dtype=np.uint32)
            n = self.lib.ref_bulk_contains(h, values, len(values), threshold, ub, cnt, cap)
            if n >= 0:
                return ub[:n].copy(), cnt[:n].copy()
            cap = -n

    def thresholder(self, window_size, kmer_size, percentage, error_rate, use_syncmer, fracminhash=False):
        return self.lib.ref_threshold_new(window_size, kmer_size, percentage, error_rate, int(use_syncmer), int(fracminhash))

    def threshold_get(self, thr, count, scaling_factor=1.0):
        return self.lib.ref_threshold_get(thr, count, scaling_factor)

    def kmer_ci(self, r, k, n, conf=0.95):
        lo, hi = C.c_uint64(), C.c_uint64()
        self.lib.ref_kmer_ci(r, k, n, conf, C.byref(lo), C.byref(hi))
        return lo.value, hi.value
