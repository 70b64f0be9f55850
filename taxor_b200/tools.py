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
    T.txs_hixf_build.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
    T.txs_hixf_build.restype = vp
    T.txs_hixf_free.argtypes = [vp]
    T.txs_hixf_free.restype = None
    T.txs_hixf_n_ixf.argtypes = [vp]
    T.txs_hixf_n_ixf.restype = C.c_uint64
    T.txs_hixf_reseeds.argtypes = [vp]
    T.txs_hixf_reseeds.restype = C.c_uint64
    T.txs_hixf_arrays.argtypes = [vp] + [C.POINTER(vp)] * 8
    T.txs_hixf_arrays.restype = None
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

    def __init__(self, ub_hashes, t_max: int = 64, seed: int = 1, threads: int = 0) -> None:
        self._ub = [np.ascontiguousarray(np.unique(h), dtype=np.uint64) for h in ub_hashes]
        n = len(self._ub)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in self._ub])
        cnt = np.array([len(a) for a in self._ub], dtype=np.uint64)
        self._h = tlib().txs_hixf_build(ptrs, cnt.ctypes.data, n, t_max, seed, threads)
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
        dptr = np.ctypeslib.as_array(C.cast(outs[4], C.POINTER(C.c_uint64)), (k,)).copy()
        self.bin_off = arr(outs[5], k + 1, C.c_uint64, np.uint64)
        nb = int(self.bin_off[-1])
        self.next_ixf_id = arr(outs[6], nb, C.c_int64, np.int64)
        self.bin_to_ub = arr(outs[7], nb, C.c_int64, np.int64)
        # zero-copy views of the fingerprint arrays (owned by the native object)
        self.data = []
        for i in range(k):
            size = 3 * int(self.seg_len[i]) * int(self.tbins[i])
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
