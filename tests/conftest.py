import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader
    return loader.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref); tests that need it skip when it was never built."""
    from oracle import loader
    lib = loader.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libmpeghref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def so_lib():
    """libmpeghb200.so loaded (host-side entry points only; no context)."""
    from libmpegh_b200.lib import Lib, SO_PATH, build
    if not os.path.exists(SO_PATH):
        build()
    return Lib()


@pytest.fixture(scope="session")
def gpu_lib():
    """libmpeghb200.so with a context on cuda:0.  No fallback: raises when there is no GPU."""
    from libmpegh_b200.lib import Lib
    lib = Lib().create(0)
    yield lib
    lib.close()
