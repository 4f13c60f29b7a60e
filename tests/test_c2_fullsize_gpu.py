"""BASELINE configs[1] at its FULL size with non-overlapping patches (level-0 16384 x 16384, P = 1792,
S = 1792 -> 10 x 10 = 100 candidates): every candidate is small enough in number for the oracle to check ALL of them --
keep / reject, label and every byte of the Pillow-exact 224 x 224 resize -- on the very image bench.py times.  (The
S = 224 variant of the same image, 5476 candidates, is checked on a sample by bench.py's parity block and through
size-independent properties in test_stage1_gpu.py.)"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def test_configs1_full_size_nonoverlapping_patches_equal_oracle():
    import bench
    from oracle import hipac_oracle as orc
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor

    img_h, msk_h, _ = bench.build_slab(1, 0, pinned=False)
    want = orc.extract_patches_oracle(img_h.numpy(), msk_h.numpy(), bench.LEVEL, stride=bench.PATCH)
    assert want["candidates"] == 100 and 0 < len(want["coords"]) < 100 and 0 < int(want["labels"].sum()) < len(want["coords"])
    for mode in ("auto", "direct"):     # the read-once streaming pass and the general per-patch path
        out = extract_patches_tensor(img_h.cuda(), msk_h.cuda(), bench.LEVEL, stride=bench.PATCH, layout="s2d16", want_u8=True, mode=mode)
        torch.cuda.synchronize()
        assert out.candidates == 100
        assert np.array_equal(out.coords.cpu().numpy(), want["coords"]), mode
        assert np.array_equal(out.labels.cpu().numpy(), want["labels"]), mode
        assert np.array_equal(out.images_u8.cpu().numpy(), want["images"]), mode
