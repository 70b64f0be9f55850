import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries_are_current():
    """`make` for the product libraries and the oracle before the first test: a fresh checkout has no .so files (they are
    git-ignored) and an edited source must not be tested through a stale binary.  Test infrastructure only -- the
    product itself never builds or falls back at run time."""
    import taxor_b200
    taxor_b200.build_all()
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own sources compiled in place (oracle/_ref); skipped where it was never built."""
    from oracle.oracle import Reference, REF_SO
    if not os.path.exists(REF_SO) and not os.path.isdir("/root/reference/src"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Reference()


@pytest.fixture(scope="session")
def built_libs():
    import taxor_b200
    if not (os.path.exists(taxor_b200.LIB_PATH) and os.path.exists(taxor_b200.TOOLS_PATH)):
        taxor_b200.build_all()
    return taxor_b200


@pytest.fixture(scope="session")
def ctx(built_libs):
    from taxor_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
