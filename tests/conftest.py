import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Port, build
    build(ref=False)
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build; only where oracle/_ref/libsrslte_ref.so exists."""
    from oracle.bindings import Ref
    if not Ref.available():
        pytest.skip("reference library not built (oracle/build_ref.sh needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    import srsran_b200 as b
    c = b.Context(0)
    yield c
    c.close()
