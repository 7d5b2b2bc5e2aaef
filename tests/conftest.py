import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


from pygsti_b200.fixtures import Case  # noqa: E402


@pytest.fixture(scope="session")
def load_case():
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = Case(name)
        return cache[name]
    return _load


@pytest.fixture(scope="session")
def gpu_ctx():
    from pygsti_b200 import engine
    ctx = engine.Context(0)
    yield ctx
    ctx.close()
