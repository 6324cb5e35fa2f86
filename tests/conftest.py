import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    from opticalflowcontainer_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def engine_factory(built_lib):
    import opticalflowcontainer_b200 as ofb
    made = []

    def make(w, h, batch=1):
        e = ofb.FlowEngine(w, h, batch, 0)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()
