"""The CPU oracle against vectors frozen from the REAL reference (tests/golden/make_golden.py)
and against the installed Pillow / torchvision directly."""
import zlib

import numpy as np
import pytest

from conftest import golden_stage1_cases, load_golden, slide_for
from oracle import hipac_oracle as orc


@pytest.mark.parametrize("case", golden_stage1_cases())
def test_stage1_matches_reference_fixture(case):
    g = load_golden(f"stage1_{case}.npz")
    slide = slide_for(g)
    level = int(g["level"])
    stride = None if int(g["stride"]) < 0 else int(g["stride"])
    out = orc.extract_patches_oracle(slide.level_array(level), slide.lesion_mask(level), level, stride=stride)
    assert np.array_equal(out["coords"], g["coords"])          # emission order, bit exact
    assert np.array_equal(out["labels"], g["labels"])
    crc = np.array([zlib.crc32(im.tobytes()) for im in out["images"]], dtype=np.uint32)
    assert np.array_equal(crc, g["crc32"])                      # every resized patch, bit exact
    for k, idx in enumerate(g["full_idx"]):
        assert np.array_equal(out["images"][idx], g["full"][k])
    names = [orc.patch_name(slide.name, x, y, l) for (x, y), l in zip(out["coords"], out["labels"])]
    assert names == [str(n) for n in g["names"]]


def test_candidate_counts_known_answers():
    # 4000x3000 level-0 -> 252 candidates at the CLI stride; 4096^2 level-3 -> 361
    assert len(orc.candidate_grid(4000, 3000, 0)[2]) == 252
    assert len(orc.candidate_grid(4096, 4096, 3)[2]) == 361
    # notebook 02 arithmetic: 97792 x 221184 at non-overlapping 1792 -> 54 x 123 full tiles (6642)
    assert (97792 // 1792) * (221184 // 1792) == 6642
    p, s, grid = orc.candidate_grid(97792 // 16, 221184 // 16, 0, stride=1792)
    assert (p, s) == (1792, 1792)


def test_stride_default_is_224_at_every_level():
    for level, p in orc.PATCH_SIZES.items():
        assert orc.patch_and_stride(level) == (p, 224)
        assert orc.patch_and_stride(level, stride=p) == (p, p)


def test_pillow_coefficients_known_answers():
    xmin, cnt, kk = orc.pil_bilinear_coeffs(448, 224)
    assert list(kk[5, :4]) == [int(0.125 * 2**22), int(0.375 * 2**22), int(0.375 * 2**22), int(0.125 * 2**22)]
    for scale in (2, 4, 8):
        xmin, cnt, kk = orc.pil_bilinear_coeffs(224 * scale, 224)
        assert cnt[0] == cnt[-1] == 3 * scale // 2 and set(cnt[1:-1]) == {2 * scale}
        assert xmin[0] == 0 and all(xmin[1:] == scale * np.arange(1, 224) - scale // 2)
        assert all(kk[1:-1].sum(axis=1) == 2**22)               # interior windows sum to exactly 1.0
        m = np.arange(scale)
        tri = np.concatenate([2 * m + 1, (2 * m + 1)[::-1]]) * (2**22 // (2 * scale * scale))
        assert np.array_equal(kk[7, :2 * scale], tri)


@pytest.mark.parametrize("p", [448, 896, 1792])
def test_resize_matches_installed_pillow(p):
    from PIL import Image
    rng = np.random.default_rng(p)
    img = rng.integers(0, 256, size=(p, p, 3), dtype=np.uint8)
    img[: p // 3, p // 2:] = 255                                 # white padding region
    want = np.array(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(orc.pil_resize_bilinear(img), want)


def test_normalize_matches_torchvision():
    import torch
    from torchvision import transforms
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8)
    img[0, :256 // 3 + 1].flat[:256] = np.arange(256, dtype=np.uint8)[: img[0, :256 // 3 + 1].size]
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    want = tf(Image.fromarray(img)).permute(1, 2, 0).numpy()
    got = orc.normalize_u8(img)
    assert np.array_equal(got, want)
    bf = orc.to_bf16_bits(got)
    assert np.array_equal(bf, torch.from_numpy(got).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))


def test_tissue_threshold_integer_form_equals_float_mean():
    rng = np.random.default_rng(1)
    for _ in range(200):
        p = 224
        base = rng.integers(236, 245)
        patch = np.clip(rng.integers(base - 3, base + 4, size=(p, p, 3)), 0, 255).astype(np.uint8)
        assert orc.tissue_keep(patch) == (not (np.mean(patch) > 240))
    edge = np.full((224, 224, 3), 240, np.uint8)
    assert orc.tissue_keep(edge)
    edge[0, 0, 0] = 241
    assert not orc.tissue_keep(edge)


def test_stage2_oracle_matches_reference_fixture():
    g = load_golden("stage2_features.npz")
    net = orc.make_resnet18(seed=int(g["seed"]), classifier=True)
    wsum = float(sum(v.double().abs().sum() for v in net.state_dict().values()))
    assert abs(wsum - float(g["weight_abs_sum"])) < 1e-6 * wsum
    feats, logits = orc.resnet18_features_fp32(net, g["images"])
    # same fp32 math, possibly different batch blocking -> tiny float noise only
    np.testing.assert_allclose(feats, g["features"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-4, atol=1e-5)
