import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import wn_oracle_py as o
    o.build_oracle()
    o.lib()
    return o


@pytest.fixture(scope="session")
def capi():
    import wn_pkg
    mod = wn_pkg.load_capi()
    if not os.path.exists(mod.LIB_PATH):          # fresh checkout: compile the CUDA library first
        import __graft_entry__ as g
        g.build()
    return mod


@pytest.fixture(scope="session")
def ctx(capi):
    """One GPU context for the whole session.  Fails loudly without the CUDA library."""
    c = capi.Context(0)
    yield c
    c.close()
