import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    import __graft_entry__ as entry
    entry.build_oracle()
    from oracle import binding as ob
    return ob.load_port()


@pytest.fixture(scope="session")
def ref_oracle():
    """The reference's own Serial build, when it exists on this machine (else None)."""
    import __graft_entry__ as entry
    entry.build_oracle()
    from oracle import binding as ob
    return ob.load_reference()
