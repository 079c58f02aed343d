import importlib
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
    from oracle.oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reftwin():
    from oracle.oracle import RefTwin

    try:
        return RefTwin()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libmicref.so not built (reference tree absent)")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("medical-image-codec_b200.synth")


@pytest.fixture(scope="session")
def mic():
    """The product package (loads libmicgpu.so; fails loudly if it is missing)."""
    import __graft_entry__ as g

    g.build()
    return importlib.import_module("medical-image-codec_b200")


GOLDEN = os.path.join(ROOT, "tests", "golden")
