"""In-tree build of the native libraries (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libtaxor_b200.so")
TOOLS_PATH = os.path.join(HERE, "libtaxor_tools.so")
CLI_PATH = os.path.join(HERE, "bin", "taxor")


def build_all(force: bool = False, verbose: bool = False) -> None:
    """``make -C taxor_b200/csrc``: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", "8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building taxor_b200 failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
