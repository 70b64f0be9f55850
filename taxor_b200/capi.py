"""ctypes binding of include/taxor_b200.h (libtaxor_b200.so).  No CPU fallback: every call needs the CUDA library."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .build import LIB_PATH

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


class TaxorError(RuntimeError):
    pass


class IxfView(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("bins", C.c_uint64), ("tbins", C.c_uint64), ("seg_len", C.c_uint64),
                ("fp", C.c_void_p), ("rows", C.c_uint64)]


class IxfScheme(C.Structure):
    """txr_ixf_scheme: the probe arithmetic an index was built with (all zero = the prototype's)."""
    _fields_ = [("slots", C.c_uint32), ("mix", C.c_uint32), ("fingerprint", C.c_uint32), ("rot1", C.c_uint32), ("rot2", C.c_uint32),
                ("layout", C.c_uint32)]


class HixfView(C.Structure):
    _fields_ = [("n_ixf", C.c_uint64), ("ixf", C.POINTER(IxfView)), ("bin_off", C.c_void_p),
                ("next_ixf_id", C.c_void_p), ("bin_to_user_bin", C.c_void_p), ("n_user_bins", C.c_uint64),
                ("scheme", C.POINTER(IxfScheme))]


class Params(C.Structure):
    _fields_ = [("kmer_size", C.c_uint8), ("syncmer_size", C.c_uint8), ("t_syncmer", C.c_uint8),
                ("use_syncmer", C.c_uint8), ("window_size", C.c_uint32), ("scaling", C.c_uint16),
                ("percentage", C.c_double), ("error_rate", C.c_double)]


class Result(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("hash_count", C.POINTER(C.c_uint32)), ("threshold", C.POINTER(C.c_uint64)),
                ("hit_begin", C.POINTER(C.c_uint64)), ("user_bin", C.POINTER(C.c_int64)),
                ("count", C.POINTER(C.c_uint32)), ("keep", C.POINTER(C.c_uint8))]


class BinHashes(C.Structure):
    _fields_ = [("n_bins", C.c_uint64), ("bin_off", C.POINTER(C.c_uint64)), ("hashes", C.POINTER(C.c_uint64)),
                ("n_segments", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("hash_ms", C.c_float), ("dedup_ms", C.c_float), ("query_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("query_launches", C.c_uint64),
                ("hash_launches", C.c_uint64), ("dedup_launches", C.c_uint64), ("query_items", C.c_uint64),
                ("query_bytes", C.c_uint64), ("hash_bytes", C.c_uint64), ("n_hashes", C.c_uint64),
                ("skipped_hashes", C.c_uint64), ("probe_launches", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_LIB = None


def lib():
    """Loads libtaxor_b200.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise TaxorError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.txr_last_error.restype = C.c_char_p
    L.txr_version.restype = C.c_char_p
    L.txr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.txr_ctx_destroy.argtypes = [vp]
    L.txr_ctx_destroy.restype = None
    L.txr_ctx_set_stream.argtypes = [vp, vp]
    L.txr_ctx_configure.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
    L.txr_ctx_reserve.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.txr_index_upload.argtypes = [vp, C.POINTER(HixfView)]
    L.txr_index_clone.argtypes = [vp, vp]
    L.txr_params_set.argtypes = [vp, C.POINTER(Params)]
    L.txr_threshold_get.argtypes = [vp, C.c_uint64, C.c_double, C.POINTER(C.c_uint64)]
    L.txr_threshold_eval.argtypes = [C.POINTER(Params), C.c_uint64, C.c_double, C.POINTER(C.c_uint64)]
    L.txr_packed_words.argtypes = [C.c_uint64]
    L.txr_packed_words.restype = C.c_uint64
    L.txr_pack_2bit.argtypes = [C.c_char_p, C.c_uint64, vp]
    L.txr_pack_codes.argtypes = [vp, C.c_uint64, vp]
    L.txr_unpack_codes.argtypes = [vp, C.c_uint64, vp]
    L.txr_host_alloc.argtypes = [C.c_size_t]
    L.txr_host_alloc.restype = vp
    L.txr_ctx_host_alloc.argtypes = [vp, C.c_size_t]
    L.txr_ctx_host_alloc.restype = vp
    L.txr_host_free.argtypes = [vp]
    L.txr_host_free.restype = None
    L.txr_search.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(Result)]
    L.txr_reads_upload.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(vp)]
    L.txr_reads_free.argtypes = [vp, vp]
    L.txr_reads_free.restype = None
    L.txr_search_resident.argtypes = [vp, vp, C.c_int, C.POINTER(Result)]
    L.txr_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.txr_hash_batch.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.txr_hash_user_bins.argtypes = [vp, vp, vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(BinHashes)]
    L.txr_plan_segments.argtypes = [C.POINTER(Params), vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.txr_ixf_bulk_count.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp]
    L.txr_profile_last_error.restype = C.c_char_p
    L.txr_profile_create.argtypes = [C.POINTER(vp)]
    L.txr_profile_destroy.argtypes = [vp]
    L.txr_profile_destroy.restype = None
    L.txr_profile_add_batch.argtypes = [vp, C.POINTER(Result), vp, vp, vp, C.c_uint64]
    L.txr_profile_add_file.argtypes = [vp, C.c_char_p]
    L.txr_profile_filter.argtypes = [vp, C.c_int]
    L.txr_profile_get.argtypes = [vp, vp]
    L.txr_profile_text.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64)]
    _LIB = L
    return L


EXPORTED = ["txr_last_error", "txr_version", "txr_ctx_create", "txr_ctx_destroy", "txr_ctx_set_stream", "txr_ctx_configure",
            "txr_ctx_reserve", "txr_index_upload", "txr_index_clone", "txr_params_set", "txr_threshold_get", "txr_threshold_eval", "txr_packed_words", "txr_pack_2bit",
            "txr_pack_codes", "txr_unpack_codes", "txr_host_alloc", "txr_ctx_host_alloc", "txr_host_free", "txr_search",
            "txr_reads_upload", "txr_reads_free", "txr_search_resident", "txr_get_timing", "txr_hash_batch",
            "txr_hash_user_bins", "txr_plan_segments", "txr_ixf_bulk_count",
            "txr_profile_last_error", "txr_profile_create", "txr_profile_destroy", "txr_profile_add_batch", "txr_profile_add_file",
            "txr_profile_filter", "txr_profile_get", "txr_profile_text"]


def _check(rc: int) -> None:
    if rc != 0:
        raise TaxorError(f"taxor_b200 error {rc}: {lib().txr_last_error().decode()}")


@dataclass
class PackedReads:
    """2-bit packed reads in the library layout (see include/taxor_b200.h)."""
    words: np.ndarray     # u64
    word_off: np.ndarray  # u64[n]
    length: np.ndarray    # u32[n]

    @property
    def n(self) -> int:
        return len(self.length)

    @property
    def n_bases(self) -> int:
        return int(self.length.astype(np.uint64).sum())


def packed_words(n_bases: int) -> int:
    return (int(n_bases) + 31) // 32 + 1


def pack_codes(code_arrays) -> PackedReads:
    """Packs a list of uint8 code arrays (0..3) with the library's own packer."""
    L = lib()
    lens = np.array([len(a) for a in code_arrays], dtype=np.uint32)
    nw = np.array([packed_words(x) for x in lens], dtype=np.uint64)
    off = np.zeros(len(lens), dtype=np.uint64)
    if len(lens) > 1:
        off[1:] = np.cumsum(nw)[:-1]
    words = np.zeros(int(nw.sum()) if len(lens) else 1, dtype=np.uint64)
    for i, a in enumerate(code_arrays):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        _check(L.txr_pack_codes(a.ctypes.data, len(a), words[int(off[i]):].ctypes.data))
    return PackedReads(words, off, lens)


def pack_ascii(seqs) -> PackedReads:
    L = lib()
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    nw = np.array([packed_words(x) for x in lens], dtype=np.uint64)
    off = np.zeros(len(lens), dtype=np.uint64)
    if len(lens) > 1:
        off[1:] = np.cumsum(nw)[:-1]
    words = np.zeros(int(nw.sum()) if len(lens) else 1, dtype=np.uint64)
    for i, s in enumerate(seqs):
        b = s.encode() if isinstance(s, str) else bytes(s)
        _check(L.txr_pack_2bit(b, len(b), words[int(off[i]):].ctypes.data))
    return PackedReads(words, off, lens)


def unpack_codes(reads: PackedReads, i: int) -> np.ndarray:
    out = np.empty(int(reads.length[i]), dtype=np.uint8)
    _check(lib().txr_unpack_codes(reads.words[int(reads.word_off[i]):].ctypes.data, len(out), out.ctypes.data))
    return out


class PinnedArray:
    """numpy view over txr_host_alloc'ed (pinned) memory."""

    def __init__(self, n: int, dtype) -> None:
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        self.ptr = lib().txr_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise TaxorError(lib().txr_last_error().decode())
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self) -> None:
        if self.ptr:
            self.array = None
            lib().txr_host_free(self.ptr)
            self.ptr = None


class SearchResult:
    """Owning numpy copy of a txr_result."""

    def __init__(self, r: Result) -> None:
        n = int(r.n_reads)
        self.n_reads = n
        self.hash_count = np.ctypeslib.as_array(r.hash_count, (n,)).copy() if n else np.zeros(0, np.uint32)
        self.threshold = np.ctypeslib.as_array(r.threshold, (n,)).copy() if n else np.zeros(0, np.uint64)
        self.hit_begin = np.ctypeslib.as_array(r.hit_begin, (n + 1,)).copy()
        nh = int(self.hit_begin[n])
        self.user_bin = np.ctypeslib.as_array(r.user_bin, (nh,)).copy() if nh else np.zeros(0, np.int64)
        self.count = np.ctypeslib.as_array(r.count, (nh,)).copy() if nh else np.zeros(0, np.uint32)
        self.keep = np.ctypeslib.as_array(r.keep, (nh,)).copy().astype(bool) if nh else np.zeros(0, bool)

    def hits(self, i: int, filtered: bool = False):
        a, b = int(self.hit_begin[i]), int(self.hit_begin[i + 1])
        ub, cnt = self.user_bin[a:b], self.count[a:b]
        if filtered:
            k = self.keep[a:b]
            return ub[k], cnt[k]
        return ub, cnt


def plan_segments(words: np.ndarray, length: int, target_windows: int, *, k, s=0, use_syncmer=True, window_size=None) -> np.ndarray:
    """txr_plan_segments (host only): the window indices at which a sequence may be cut into independently hashed pieces."""
    p = Params(k, s, 0, int(use_syncmer), k if window_size is None else window_size, 1, -1.0, 0.04)
    cap = int(length) // max(int(target_windows), 1) + 2
    cuts = np.zeros(cap, dtype=np.uint64)
    n = C.c_uint64()
    _check(lib().txr_plan_segments(C.byref(p), words.ctypes.data, int(length), int(target_windows), cuts.ctypes.data, cap, C.byref(n)))
    return cuts[: n.value].copy()


def threshold_eval(count: int, scaling_factor: float = 1.0, *, k, use_syncmer=True, window_size=None, percentage=-1.0,
                   error_rate=0.04) -> int:
    """hixf::threshold::threshold::get without a GPU context (host mirror of the reference's threshold models)."""
    p = Params(k, 0, 0, int(use_syncmer), k if window_size is None else window_size, 1, percentage, error_rate)
    out = C.c_uint64()
    _check(lib().txr_threshold_eval(C.byref(p), count, scaling_factor, C.byref(out)))
    return out.value


class Context:
    """One GPU context (txr_ctx)."""

    def __init__(self, device: int = 0) -> None:
        self._L = lib()
        self._h = C.c_void_p()
        _check(self._L.txr_ctx_create(device, C.byref(self._h)))
        self._keep = None

    def close(self) -> None:
        if self._h:
            self._L.txr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int) -> None:
        _check(self._L.txr_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def configure(self, max_batch_reads=262144, max_batch_bases=3_000_000_000, n_slots=3):
        _check(self._L.txr_ctx_configure(self._h, max_batch_reads, max_batch_bases, n_slots))

    def upload_index(self, seed, bins, tbins, seg_len, data, bin_off, next_ixf_id, bin_to_ub, n_user_bins, rows=None, scheme=None,
                     layout=0):
        """scheme: None or (slots, mix, fingerprint, rot1, rot2); layout 1 = the host arrays are bin-major"""
        n = len(seed)
        ixfs = (IxfView * n)()
        for i in range(n):
            d = data[i]
            ptr = d.ctypes.data if isinstance(d, np.ndarray) else int(d)
            ixfs[i] = IxfView(int(seed[i]), int(bins[i]), int(tbins[i]), int(seg_len[i]), ptr, 0 if rows is None else int(rows[i]))
        bin_off = np.ascontiguousarray(bin_off, dtype=np.uint64)
        nx = np.ascontiguousarray(next_ixf_id, dtype=np.int64)
        ub = np.ascontiguousarray(bin_to_ub, dtype=np.int64)
        sch = None
        if scheme is not None or layout:
            sc = tuple(scheme) if scheme is not None else (0, 0, 0, 21, 42)
            sch = C.pointer(IxfScheme(*[int(x) for x in sc], int(layout)))
        v = HixfView(n, ixfs, bin_off.ctypes.data, nx.ctypes.data, ub.ctypes.data, int(n_user_bins), sch)
        _check(self._L.txr_index_upload(self._h, C.byref(v)))

    def reserve(self, n_reads: int, n_bases: int) -> None:
        """txr_ctx_reserve: allocate the per-slot buffers of a search of this size now instead of inside the first call"""
        _check(self._L.txr_ctx_reserve(self._h, int(n_reads), int(n_bases)))

    def clone_index_from(self, other: "Context") -> None:
        """txr_index_clone: device-to-device replica of the index resident in `other` (NVLink between peer GPUs)."""
        _check(self._L.txr_index_clone(self._h, other._h))

    def set_params(self, *, k, s=0, t=0, use_syncmer=True, window_size=None, scaling=1, percentage=-1.0, error_rate=0.04):
        p = Params(k, s, t, int(use_syncmer), k if window_size is None else window_size, scaling, percentage, error_rate)
        _check(self._L.txr_params_set(self._h, C.byref(p)))

    def threshold(self, count: int, scaling_factor: float = 1.0) -> int:
        out = C.c_uint64()
        _check(self._L.txr_threshold_get(self._h, count, scaling_factor, C.byref(out)))
        return out.value

    def search(self, reads: PackedReads) -> SearchResult:
        r = Result()
        _check(self._L.txr_search(self._h, reads.words.ctypes.data, reads.word_off.ctypes.data,
                                  reads.length.ctypes.data, reads.n, C.byref(r)))
        return SearchResult(r)

    def search_raw(self, words_ptr, off_ptr, len_ptr, n) -> Result:
        """txr_search on raw host pointers; the returned view is valid until the next call."""
        r = Result()
        _check(self._L.txr_search(self._h, words_ptr, off_ptr, len_ptr, n, C.byref(r)))
        return r

    def upload_reads(self, reads: PackedReads):
        h = C.c_void_p()
        _check(self._L.txr_reads_upload(self._h, reads.words.ctypes.data, reads.word_off.ctypes.data,
                                        reads.length.ctypes.data, reads.n, C.byref(h)))
        return h

    def free_reads(self, h) -> None:
        self._L.txr_reads_free(self._h, h)

    def search_resident(self, h, fetch: bool = True):
        r = Result()
        _check(self._L.txr_search_resident(self._h, h, int(fetch), C.byref(r)))
        return SearchResult(r) if fetch else None

    def timing(self) -> dict:
        t = Timing()
        _check(self._L.txr_get_timing(self._h, C.byref(t)))
        return t.as_dict()

    def hash_user_bins(self, seqs: PackedReads, seq_bin, n_bins: int):
        """txr_hash_user_bins: distinct hashes per user bin (compute_hashes of `taxor build`).  Returns (bin_off, hashes,
        n_segments); the arrays are copies."""
        seq_bin = np.ascontiguousarray(seq_bin, dtype=np.uint32)
        r = BinHashes()
        _check(self._L.txr_hash_user_bins(self._h, seqs.words.ctypes.data, seqs.word_off.ctypes.data, seqs.length.ctypes.data,
                                          seqs.n, seq_bin.ctypes.data, n_bins, C.byref(r)))
        off = np.ctypeslib.as_array(r.bin_off, (n_bins + 1,)).copy()
        total = int(off[-1])
        hashes = np.ctypeslib.as_array(r.hashes, (total,)).copy() if total else np.zeros(0, np.uint64)
        return off, hashes, int(r.n_segments)

    def hash_batch(self, reads: PackedReads, dedup: bool = True):
        off, hs = C.c_void_p(), C.c_void_p()
        _check(self._L.txr_hash_batch(self._h, reads.words.ctypes.data, reads.word_off.ctypes.data,
                                      reads.length.ctypes.data, reads.n, int(dedup), C.byref(off), C.byref(hs)))
        o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), (reads.n + 1,)).copy()
        total = int(o[-1])
        h = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint64)), (total,)).copy() if total else np.zeros(0, np.uint64)
        return o, h

    def ixf_bulk_count(self, ixf_idx: int, values, bins: int) -> np.ndarray:
        values = np.ascontiguousarray(values, dtype=np.uint64)
        counts = np.zeros(bins, dtype=np.uint32)
        _check(self._L.txr_ixf_bulk_count(self._h, ixf_idx, values.ctypes.data, len(values), counts.ctypes.data))
        return counts


class ProfileSpecies(C.Structure):
    _fields_ = [("user_bin", C.c_uint64), ("seq_len", C.c_uint64), ("accession_id", C.c_char_p), ("taxid", C.c_char_p),
                ("taxnames_string", C.c_char_p), ("taxid_string", C.c_char_p)]


class Profile:
    """txr_profile: the head of `taxor profile` (parse + the three reference-filter rounds), fed from txr_result batches or from
    a result file.  Host code only -- usable without a GPU."""

    def __init__(self) -> None:
        self._L = lib()
        self._h = C.c_void_p()
        _check(self._L.txr_profile_create(C.byref(self._h)))

    def _ok(self, rc):
        if rc != 0:
            raise TaxorError(f"taxor_b200 error {rc}: {self._L.txr_profile_last_error().decode()}")

    def add_result(self, result: "Result", read_ids, read_len, species) -> None:
        """result: the ctypes Result of a search call (still valid); species: list of dicts as in tools.default_species"""
        n = int(result.n_reads)
        ids = (C.c_char_p * n)(*[s.encode() if isinstance(s, str) else bytes(s) for s in read_ids])
        ln = np.ascontiguousarray(read_len, dtype=np.uint32)
        sp = (ProfileSpecies * len(species))(*[ProfileSpecies(int(x["user_bin"]), int(x["seq_len"]), x["accession_id"].encode(),
                                                              x["taxid"].encode(), x["taxnames_string"].encode(),
                                                              x["taxid_string"].encode()) for x in species])
        self._ok(self._L.txr_profile_add_batch(self._h, C.byref(result), ids, ln.ctypes.data, sp, len(species)))

    def add_file(self, path) -> None:
        self._ok(self._L.txr_profile_add_file(self._h, str(path).encode()))

    def filter(self, rounds: int = 3) -> None:
        self._ok(self._L.txr_profile_filter(self._h, rounds))

    def text(self) -> str:
        t, n = C.c_char_p(), C.c_uint64()
        self._ok(self._L.txr_profile_text(self._h, C.byref(t), C.byref(n)))
        return C.string_at(t, n.value).decode()

    def close(self) -> None:
        if self._h:
            self._L.txr_profile_destroy(self._h)
            self._h = C.c_void_p()
