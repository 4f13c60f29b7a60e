"""The stage-1 kernels under -DHIPAC_DEBUG_BOUNDS (device-side bounds assertions on every staged copy, plane read and
plane write): compute-sanitizer is not available on the B200 pool, so the library's own assertions stand in.  Runs in
a subprocess (a failed device assertion traps and poisons the CUDA context) with HIPAC_DEBUG_BOUNDS=1, which makes
``_lib.lib()`` load ``libhipac_b200_dbg.so``; the results must equal the oracle's."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
from oracle import hipac_oracle as orc
from ss25_hierarchical_multiscale_image_classification_b200 import _lib
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
assert _lib.lib()._name.endswith("libhipac_b200_dbg.so"), _lib.lib()._name
rng = np.random.default_rng(1)
cases = [(0, None, 5953, 1800), (0, None, 3100, 2600), (1, None, 2977, 1000), (2, None, 449, 225), (1, None, 100, 90),
         (0, None, 8, 3000), (2, 32, 700, 600), (3, None, 1117, 903), (0, 1792, 4000, 3700), (1, 333, 2100, 1900)]
for level, stride, w, h in cases:
    img = rng.integers(150, 256, size=(h, w, 3), dtype=np.uint8)
    img[h // 8: 3 * h // 4, w // 6: 5 * w // 6] = rng.integers(0, 256, size=(3 * h // 4 - h // 8, 5 * w // 6 - w // 6, 3), dtype=np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[h - 1, w - 1] = 1
    want = orc.extract_patches_oracle(img, mask, level, stride=stride)
    for mode in ("auto", "fused_legacy") if (stride is None or stride %% 32 == 0) else ("auto",):
        out = extract_patches_tensor(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda(), level, stride=stride,
                                     layout="s2d16", want_u8=True, mode=mode)
        torch.cuda.synchronize()
        assert np.array_equal(out.coords.cpu().numpy(), want["coords"]), (level, stride, w, h, mode)
        assert np.array_equal(out.labels.cpu().numpy(), want["labels"])
        assert np.array_equal(out.images_u8.cpu().numpy(), want["images"])
print("BOUNDS_OK")
""" % ROOT


def test_stage1_kernels_pass_their_bounds_assertions():
    from ss25_hierarchical_multiscale_image_classification_b200 import _lib
    _lib.build(debug_bounds=True)          # no-op when the library beside the sources is fresh
    env = dict(os.environ, HIPAC_DEBUG_BOUNDS="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "BOUNDS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
