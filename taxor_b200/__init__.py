"""taxor_b200 -- B200-native (sm_100a) implementation of the `taxor search` hot path.

The product is the C-ABI shared library ``libtaxor_b200.so`` (``include/taxor_b200.h``) built from
``taxor_b200/csrc``.  This package is the thin Python mirror of that ABI used by the tests and ``bench.py``:

* :mod:`taxor_b200.capi`  -- ctypes binding of every ``txr_*`` entry point (fails loudly when the library or a
  CUDA device is missing; there is no CPU fallback).
* :mod:`taxor_b200.tools` -- CPU tooling (synthetic genomes/reads, XOR-filter construction, ``.hixf`` I/O).
"""
from .build import build_all, CLI_PATH, LIB_PATH, TOOLS_PATH  # noqa: F401

__all__ = ["build_all", "CLI_PATH", "LIB_PATH", "TOOLS_PATH"]
