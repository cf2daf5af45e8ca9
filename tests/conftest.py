"""Shared fixtures.  `-m gpu` tests call the CUDA library through the C-ABI; everything else is CPU-only."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "deepseek.cpp_b200"))
sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo():
    return REPO


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(REPO, "tests", "golden")


@pytest.fixture(scope="session")
def ckpt(tmp_path_factory):
    """Mint (once per session) a synthetic .dseek checkpoint: ckpt(preset, quant, **overrides) -> dir."""
    import mint
    cache = {}
    root = tmp_path_factory.mktemp("ckpt")

    def get(preset, quant, **kw):
        key = (preset, quant, tuple(sorted(kw.items())))
        if key not in cache:
            d = str(root / ("_".join([preset, quant] + [f"{k}{v}" for k, v in sorted(kw.items())])))
            mint.mint(d, preset, quant, **kw)
            cache[key] = d
        return cache[key]

    return get


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
