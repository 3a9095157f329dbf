import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# opt-in kernel variants that the suite must keep exercising (read once by the library):
# the 4-CTA-cluster form of blm_gemm_sampled is used for M >= 2048 rows when this is set
os.environ.setdefault("BLM_SAMPLED_CLUSTER", "1")
# ... and the cluster-multicast / staggered form of the persistent LSTM kernel for batches >= 256 rows
os.environ.setdefault("BLM_LSTM_CLUSTER", "1")
os.environ.setdefault("BLM_LSTM_STAGGER", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
    return load
