"""pytest configuration: path setup + the `gpu` marker.

`-m "not gpu"`  runs on the CPU-only build container: oracle vs golden
                vectors, host logic, C-ABI loads and exports every symbol.
`-m gpu`        the parity tests proper, on a B200, through the C-ABI.
Nothing here reads /root/reference at run time (it does not exist on the GPU box).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "face-gan-tts_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def kats():
    """Golden vectors produced by the reference's compiled core.pyx."""
    return np.load(os.path.join(ROOT, "tests", "golden", "mas_kats.npz"))


@pytest.fixture(scope="session")
def logprior_fixture():
    return np.load(os.path.join(ROOT, "tests", "golden", "logprior_small.npz"))
