import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with -m gpu on the GPU box")


def golden_names():
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt") and f.startswith("c"))


@pytest.fixture(scope="session")
def edge_cases():
    import torch

    return torch.load(os.path.join(GOLDEN_DIR, "edge_cases.pt"), weights_only=False)


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library; built on demand here (nvcc cross-compiles), prebuilt on the GPU box."""
    from cogstream_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib.load()
