import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import zxc_ctypes as z
    if not (os.path.exists(z.PRODUCT_SO) and os.path.exists(z.ORACLE_SO)
            and os.path.exists(os.path.join(ROOT, "oracle", "libzxc_corpus.so"))):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def libs():
    """(product, oracle, reference-or-None)"""
    _ensure_built()
    import zxc_ctypes as z
    prod = z.ZxcLib(z.PRODUCT_SO)
    orc = z.Oracle()
    ref = z.ZxcLib(z.REF_SO) if z.have_ref() else None
    return prod, orc, ref


@pytest.fixture(scope="session")
def prod(libs):
    return libs[0]


@pytest.fixture(scope="session")
def orc(libs):
    return libs[1]


@pytest.fixture(scope="session")
def ref(libs):
    if libs[2] is None:
        pytest.skip("oracle/_ref/libzxc_ref.so not built (reference sources absent)")
    return libs[2]


def has_cuda():
    try:
        import ctypes
        import zxc_ctypes as z
        lib = ctypes.CDLL(z.PRODUCT_SO)
        return lib.zxc_b200_device_count() > 0
    except Exception:
        return False
