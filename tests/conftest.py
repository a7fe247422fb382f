import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cu-bens_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import cubens_b200
        return cubens_b200.load_library().cb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must run the CUDA path - they fail (not skip) when it is missing on a GPU run."""
    import cubens_b200
    lib = cubens_b200.load_library()
    assert lib.cb_device_count() > 0, "no CUDA device: -m gpu tests need the B200"
    return lib


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Test infrastructure only."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libcubens_ref.so not built (run make -C oracle ref)")
    refbind.lib()
    return refbind
