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


class DarkSlide:
    """A slide of dark uniform noise (every level generated independently, exact 2**level downsamples) with a small
    lesion rectangle: used for the 'slide smaller than one patch' golden cases, where the synthetic tissue layout
    leaves no survivor.  Same duck type as ``SyntheticSlide``."""

    def __new__(cls, w0, h0, seed, with_lesion=True, n_levels=4):
        from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide, hash32
        levels = []
        for l in range(n_levels):
            w, h = w0 >> l, h0 >> l
            x = np.arange(w, dtype=np.uint32)[None, :, None]
            y = np.arange(h, dtype=np.uint32)[:, None, None]
            c = np.arange(3, dtype=np.uint32)[None, None, :]
            levels.append((40 + hash32(x, y, c, seed * 8 + l) % np.uint32(130)).astype(np.uint8))
        s = SyntheticSlide(levels=levels, seed=seed, with_lesion=with_lesion)

        def lesion_mask(level, _s=s):
            if not with_lesion:
                return None
            w, h = _s.level_dimensions[level]
            m = np.zeros((h, w), np.uint8)
            m[h // 3: h // 3 + max(2, h // 10), w // 4: w // 4 + max(2, w // 8)] = 255
            return m

        s.lesion_mask = lesion_mask
        return s


def make_slide(w0, h0, seed, with_mask):
    """seed >= 0: the synthetic tissue slide; seed < 0: ``DarkSlide(-seed)``."""
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    if seed < 0:
        return DarkSlide(w0, h0, -seed, with_lesion=with_mask)
    return SyntheticSlide(w0, h0, seed=seed, with_lesion=with_mask)


def slide_for(g):
    return make_slide(int(g["w0"]), int(g["h0"]), int(g["seed"]), bool(int(g["with_mask"])))


@pytest.fixture(scope="session")
def built_lib():
    from ss25_hierarchical_multiscale_image_classification_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()
