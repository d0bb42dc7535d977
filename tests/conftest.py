import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference mounted (build container only)")
    config.addinivalue_line("markers", "gpu2: needs torchrun with 2 GPUs (not part of the default -m gpu run)")


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    has_ref = os.path.isdir("/root/reference/models")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_err(a, b):
    """max-norm relative error: ||a-b||_inf / ||b||_inf (the tolerance BASELINE.json north_star states is 1e-3)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = float(b.abs().max())
    return float((a - b).abs().max()) / (denom if denom > 0 else 1.0)
