import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_stage1_cases():
    return sorted(f[len("stage1_"):-4] for f in os.listdir(GOLDEN) if f.startswith("stage1_") and f.endswith(".npz"))


def slide_for(g):
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    return SyntheticSlide(int(g["w0"]), int(g["h0"]), seed=int(g["seed"]), with_lesion=bool(int(g["with_mask"])))


@pytest.fixture(scope="session")
def built_lib():
    from ss25_hierarchical_multiscale_image_classification_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()
