import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "blackhole-simulation_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def built():
    """The in-tree CUDA library (built by __graft_entry__.build()); tests never fall back to anything else."""
    import __graft_entry__ as ge
    import gravitas_b200 as g
    if not os.path.exists(g.lib_path()):
        ge.build()
    g.lib()
    return g


@pytest.fixture(scope="session")
def renderer(built):
    r = built.KerrRenderer(device=0)
    r.init()
    yield r
    r.cleanup()
