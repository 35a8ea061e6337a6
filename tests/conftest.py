import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The reference troy:: CPU library (oracle/_ref) — the checker."""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libtroyref.so not built (run __graft_entry__.build())")
    return r


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a real GPU."""
    import troy_b200 as tb
    tb.KernelProvider.initialize(0)
    return tb


@pytest.fixture(scope="session")
def emul_lib():
    """tests/emul: the same kernel sources executed sequentially on the CPU (debug aid; never the product)."""
    import __graft_entry__ as g
    g.build_emul()
    import troy_b200 as tb
    from troy_b200 import lib as tlib
    # the package itself only ever loads troy_b200/libtroy_b200.so; the swap is done from here, by the test harness
    prev, tlib._library = tlib._library, tlib.Library(os.path.join(ROOT, "tests", "emul", "libtroy_b200_emul.so"))
    tb.KernelProvider.initialize(0)
    yield tb
    tlib._library = prev
