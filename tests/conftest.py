import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib.load_oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import oracle_lib
    r = oracle_lib.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libredref.so not built (reference tree absent)")
    return r


@pytest.fixture(scope="session")
def rnlib():
    from rednoise_b200 import build, capi
    build.build()
    return capi


@pytest.fixture(scope="session")
def ctx(rnlib):
    c = rnlib.Context(0)
    yield c
    c.close()
