"""ctypes binding of libtaxor_tools.so: CPU tooling for tests and benchmarks (synthetic genomes / reads, layout +
XOR-filter construction of a valid HIXF).  Not on the search path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import TOOLS_PATH

_T = None


def tlib():
    global _T
    if _T is not None:
        return _T
    if not os.path.exists(TOOLS_PATH):
        raise RuntimeError(f"{TOOLS_PATH} is missing: build with taxor_b200.build_all()")
    T = C.CDLL(TOOLS_PATH)
    vp = C.c_void_p
    T.txs_packed_words.argtypes = [C.c_uint64]
    T.txs_packed_words.restype = C.c_uint64
    T.txs_genome.argtypes = [C.c_uint64, C.c_uint64, vp]
    T.txs_genome.restype = None
    T.txs_reads.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_double, C.c_uint64, vp, vp, vp, C.c_int]
    T.txs_sort_unique_many.argtypes = [vp, vp, C.c_uint64, C.c_int]
    T.txs_sort_unique_many.restype = None
    T.txs_hixf_build.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
    T.txs_hixf_build.restype = vp
    T.txs_hixf_build2.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
    T.txs_hixf_build2.restype = vp
    T.txs_hixf_free.argtypes = [vp]
    T.txs_hixf_free.restype = None
    T.txs_hixf_n_ixf.argtypes = [vp]
    T.txs_hixf_n_ixf.restype = C.c_uint64
    T.txs_hixf_reseeds.argtypes = [vp]
    T.txs_hixf_reseeds.restype = C.c_uint64
    T.txs_hixf_arrays.argtypes = [vp] + [C.POINTER(vp)] * 8
    T.txs_hixf_arrays.restype = None
    T.txs_last_error.restype = C.c_char_p
    T.txs_hixf_write.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16,
                                 C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    T.txs_hixf_open.argtypes = [C.c_char_p, C.c_char_p]
    T.txs_hixf_open.restype = vp
    T.txs_hixf_open2.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    T.txs_hixf_open2.restype = vp
    T.txs_hixf_open_note.restype = C.c_char_p
    T.txs_hixf_ixf_rows.argtypes = [vp, C.c_uint64]
    T.txs_hixf_ixf_rows.restype = C.c_uint64
    T.txs_hixf_build3.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, vp]
    T.txs_hixf_build3.restype = vp
    T.txs_hixf_rows.argtypes = [vp]
    T.txs_hixf_rows.restype = vp
    T.txs_hixf_capacity.argtypes = [vp]
    T.txs_hixf_capacity.restype = vp
    T.txs_hixf_close.argtypes = [vp]
    T.txs_hixf_close.restype = None
    T.txs_hixf_info.argtypes = [vp, vp]
    T.txs_hixf_info.restype = None
    T.txs_hixf_ixf.argtypes = [vp, C.c_uint64] + [C.POINTER(C.c_uint64)] * 4 + [C.POINTER(vp)] * 3
    T.txs_hixf_ixf.restype = None
    T.txs_hixf_species_field.argtypes = [vp, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    T.txs_hixf_species_field.restype = C.c_char_p
    _T = T
    return T


def packed_words(n: int) -> int:
    return (int(n) + 31) // 32 + 1


def genome(seed: int, length: int) -> np.ndarray:
    """Uniform i.i.d. genome (2-bit packed, library layout)."""
    w = np.zeros(packed_words(length), dtype=np.uint64)
    tlib().txs_genome(seed, length, w.ctypes.data)
    return w


def simulate_reads(genomes, genome_len, read_len, err: float, seed: int, out_words=None, threads: int = 0):
    """ONT-like reads (sub/ins/del at err/3 each).  Returns (words, word_off, length, source genome)."""
    n = len(read_len)
    read_len = np.ascontiguousarray(read_len, dtype=np.uint32)
    nw = (read_len.astype(np.uint64) + 31) // 32 + 1
    off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        off[1:] = np.cumsum(nw)[:-1]
    total = int(nw.sum())
    words = out_words if out_words is not None else np.zeros(total, dtype=np.uint64)
    assert len(words) >= total
    gptr = (C.c_void_p * len(genomes))(*[g.ctypes.data for g in genomes])
    glen = np.ascontiguousarray(genome_len, dtype=np.uint64)
    src = np.zeros(n, dtype=np.uint32)
    rc = tlib().txs_reads(gptr, glen.ctypes.data, len(genomes), n, read_len.ctypes.data, err, seed, words.ctypes.data,
                          off.ctypes.data, src.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError("txs_reads failed (genome shorter than a read?)")
    return words, off, read_len, src


class BuiltHixf:
    """An HIXF built by the CPU tooling; exposes the plain arrays every consumer takes."""

    def __init__(self, ub_hashes, t_max: int = 64, seed: int = 1, threads: int = 0, inplace: bool = False,
                 t_max_lower: int = 0, scheme=None) -> None:
        """scheme: None (the prototype's arithmetic) or (slots, mix, fingerprint, rot1, rot2) as in txr_ixf_scheme"""
        # sorted distinct key sets (in place on private copies -- or on the caller's arrays with inplace=True, which
        # halves the footprint of a multi-GB build -- parallel over user bins)
        if inplace:
            self._ub = [h if (h.dtype == np.uint64 and h.flags.c_contiguous and h.flags.writeable) else np.array(h, dtype=np.uint64)
                        for h in ub_hashes]
        else:
            self._ub = [np.array(h, dtype=np.uint64, copy=True) for h in ub_hashes]
        n = len(self._ub)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in self._ub])
        cnt = np.array([len(a) for a in self._ub], dtype=np.uint64)
        tlib().txs_sort_unique_many(ptrs, cnt.ctypes.data, n, threads)
        self._ub = [a[: int(c)] for a, c in zip(self._ub, cnt)]
        self.n_keys = int(cnt.sum())
        self.scheme = None if scheme is None else tuple(int(v) for v in scheme)
        sch = None if scheme is None else np.array(self.scheme, dtype=np.uint32)
        self._h = tlib().txs_hixf_build3(ptrs, cnt.ctypes.data, n, t_max, t_max_lower, seed, threads, None if sch is None else sch.ctypes.data)
        if not self._h:
            raise RuntimeError("txs_hixf_build failed")
        self.n_user_bins = n
        T = tlib()
        k = int(T.txs_hixf_n_ixf(self._h))
        self.reseeds = int(T.txs_hixf_reseeds(self._h))
        outs = [C.c_void_p() for _ in range(8)]
        T.txs_hixf_arrays(self._h, *[C.byref(o) for o in outs])

        def arr(p, n_, ct, dt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (n_,)).astype(dt, copy=True)

        self.seed = arr(outs[0], k, C.c_uint64, np.uint64)
        self.bins = arr(outs[1], k, C.c_uint64, np.uint64)
        self.tbins = arr(outs[2], k, C.c_uint64, np.uint64)
        self.seg_len = arr(outs[3], k, C.c_uint64, np.uint64)
        self.rows = arr(T.txs_hixf_rows(self._h), k, C.c_uint64, np.uint64)
        self.capacity = arr(T.txs_hixf_capacity(self._h), k, C.c_uint64, np.uint64)
        dptr = np.ctypeslib.as_array(C.cast(outs[4], C.POINTER(C.c_uint64)), (k,)).copy()
        self.bin_off = arr(outs[5], k + 1, C.c_uint64, np.uint64)
        nb = int(self.bin_off[-1])
        self.next_ixf_id = arr(outs[6], nb, C.c_int64, np.int64)
        self.bin_to_ub = arr(outs[7], nb, C.c_int64, np.int64)
        # zero-copy views of the fingerprint arrays (owned by the native object)
        self.data = []
        for i in range(k):
            size = int(self.rows[i]) * int(self.tbins[i])
            buf = (C.c_uint8 * size).from_address(int(dptr[i]))
            self.data.append(np.frombuffer(buf, dtype=np.uint8))

    @property
    def n_ixf(self) -> int:
        return len(self.seed)

    @property
    def fp_bytes(self) -> int:
        return int(sum(d.size for d in self.data))

    def close(self) -> None:
        if self._h:
            self.data = []
            tlib().txs_hixf_free(self._h)
            self._h = None


def _cstr_array(strings):
    bufs = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    return (C.c_char_p * len(bufs))(*bufs), bufs


def write_hixf(path, hx, *, k, s, t, use_syncmer=True, window_size=20, scaling=1, species=None, record_spec=""):
    """Writes the arrays of `hx` (BuiltHixf or anything with the same attributes) as a `.hixf` file in the cereal
    binary layout of SURVEY Appendix A.  species: list of dicts (organism_name, accession_id, taxid, taxnames_string,
    taxid_string, user_bin, seq_len); default: one synthetic species per user bin."""
    n_ub = int(hx.n_user_bins)
    if species is None:
        species = default_species(n_ub)
    names, _k1 = _cstr_array([f"/synthetic/genome_{i}.fna" for i in range(n_ub)])
    cols = {f: _cstr_array([sp[f] for sp in species]) for f in ("organism_name", "accession_id", "taxid", "taxnames_string", "taxid_string")}
    sp_ub = np.array([sp["user_bin"] for sp in species], dtype=np.uint64)
    sp_len = np.array([sp["seq_len"] for sp in species], dtype=np.uint64)
    data = [np.ascontiguousarray(d) for d in hx.data]
    dptr = (C.c_void_p * len(data))(*[d.ctypes.data for d in data])
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.bin_off)]
    nxt = np.ascontiguousarray(hx.next_ixf_id, dtype=np.int64)
    ub = np.ascontiguousarray(hx.bin_to_ub, dtype=np.int64)
    rows = np.ascontiguousarray(hx.rows, dtype=np.uint64) if getattr(hx, "rows", None) is not None else None
    cap = np.ascontiguousarray(hx.capacity, dtype=np.uint64) if getattr(hx, "capacity", None) is not None else None
    rc = tlib().txs_hixf_write(str(path).encode(), record_spec.encode(), window_size, k, s, t, int(use_syncmer), scaling, len(data),
                               arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, dptr,
                               arrs[4].ctypes.data, nxt.ctypes.data, ub.ctypes.data, n_ub, names, len(species),
                               cols["organism_name"][0], cols["accession_id"][0], cols["taxid"][0], cols["taxnames_string"][0],
                               cols["taxid_string"][0], sp_ub.ctypes.data, sp_len.ctypes.data,
                               rows.ctypes.data if rows is not None else None, cap.ctypes.data if cap is not None else None)
    if rc != 0:
        raise RuntimeError(tlib().txs_last_error().decode())


def default_species(n_ub):
    """One species per user bin, listed in REVERSE user-bin order so that the user_bin -> species lookup is exercised."""
    return [dict(organism_name=f"Synthetica species{u}", accession_id=f"GCF_{u:09d}.1", taxid=str(100000 + u),
                 taxnames_string=f"k__Bacteria;p__Synth;g__Synthetica;s__Synthetica species{u}",
                 taxid_string=f"2;1239;{100000 + u}", user_bin=u, seq_len=1000 + 7 * u) for u in reversed(range(n_ub))]


class HixfFile:
    """A `.hixf` file opened with the library's reader (mmap); exposes the same arrays as BuiltHixf."""

    def __init__(self, path, record_spec="", scheme=""):
        T = tlib()
        self._h = T.txs_hixf_open2(str(path).encode(), record_spec.encode(), scheme.encode())
        if not self._h:
            raise RuntimeError(T.txs_last_error().decode())
        self.record_spec = T.txs_last_error().decode()
        self.note = T.txs_hixf_open_note().decode()
        info = np.zeros(14, dtype=np.uint64)
        T.txs_hixf_info(self._h, info.ctypes.data)
        (self.version, self.window_size, self.shape_size, self.shape_bits, self.k, self.s, self.t, self.parts, self.use_syncmer,
         self.scaling, self.compressed, n_ixf, self.n_user_bins, self.n_species) = [int(x) for x in info]
        self.seed, self.bins, self.tbins, self.seg_len, self.rows = (np.zeros(n_ixf, np.uint64) for _ in range(5))
        self.data, nxt, ub, off = [], [], [], [0]
        for i in range(n_ixf):
            sc = [C.c_uint64() for _ in range(4)]
            ptr = [C.c_void_p() for _ in range(3)]
            T.txs_hixf_ixf(self._h, i, *[C.byref(x) for x in sc], *[C.byref(x) for x in ptr])
            self.seed[i], self.bins[i], self.tbins[i], self.seg_len[i] = [x.value for x in sc]
            rows_i = int(T.txs_hixf_ixf_rows(self._h, i))
            self.rows[i] = rows_i
            size = rows_i * sc[2].value
            self.data.append(np.frombuffer((C.c_uint8 * size).from_address(ptr[0].value), dtype=np.uint8))
            nb = sc[1].value
            nxt.append(np.ctypeslib.as_array(C.cast(ptr[1], C.POINTER(C.c_int64)), (nb,)).copy())
            ub.append(np.ctypeslib.as_array(C.cast(ptr[2], C.POINTER(C.c_int64)), (nb,)).copy())
            off.append(off[-1] + nb)
        self.bin_off = np.array(off, dtype=np.uint64)
        self.next_ixf_id = np.concatenate(nxt)
        self.bin_to_ub = np.concatenate(ub)
        self.species = []
        for i in range(self.n_species):
            u, ln = C.c_uint64(), C.c_uint64()
            f = [T.txs_hixf_species_field(self._h, i, j, C.byref(u), C.byref(ln)).decode() for j in range(5)]
            self.species.append(dict(organism_name=f[0], accession_id=f[1], taxid=f[2], taxnames_string=f[3], taxid_string=f[4],
                                     user_bin=u.value, seq_len=ln.value))

    @property
    def n_ixf(self):
        return len(self.seed)

    def close(self):
        if self._h:
            self.data = []
            tlib().txs_hixf_close(self._h)
            self._h = None
