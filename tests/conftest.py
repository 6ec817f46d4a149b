import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree CUDA library (built on demand; nvcc cross-compiles without a GPU)."""
    from syntex_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def raw_weights():
    from syntex_b200.synthetic import synthetic_vgg_weights
    return synthetic_vgg_weights(0)


@pytest.fixture(scope="session")
def exemplar():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "exemplar_flowerbeds0008_256.npz"))["image"]
