"""pytest configuration: registers the `gpu` marker and puts the oracle (checker) and the ctypes
binding of the product library on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import pyoracle
    if not pyoracle.available("port"):
        pyoracle.build(("port",))
    return pyoracle


@pytest.fixture(scope="session")
def ref():
    import pyoracle
    if not pyoracle.available("ref"):
        if os.path.isdir("/root/reference/lib"):
            pyoracle.build(("ref",))
        else:
            pytest.skip("oracle/_ref/libedm_ref.so not built and /root/reference absent")
    return pyoracle
