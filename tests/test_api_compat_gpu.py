"""The reference's entry points end to end on the GPU: extract_patches (PNG mode), extract_features,
the PNG-free CLI path and the model wrappers, against the CPU oracle."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import hipac_oracle as orc
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
from test_api_compat_cpu import XML

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _setup_tree(tmp_path, with_xml=True):
    img = tmp_path / "data" / "camelyon16" / "train" / "img"
    ann = tmp_path / "data" / "camelyon16" / "train" / "mask" / "annotations"
    os.makedirs(img), os.makedirs(ann)
    (img / "tumor_900.tif").write_bytes(b"")
    (img / "notes.txt").write_text("ignored")
    if with_xml:
        (ann / "tumor_900.xml").write_text(XML)
    slide = SyntheticSlide(12000, 9000, seed=1234, name="tumor_900")
    return slide, (lambda path: slide)


def _metrics(got, ref):
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    return cos.min(), (np.abs(got - ref).max(1) / np.abs(ref).max(1)).max()


@pytest.mark.parametrize("level", [3, 2])
def test_extract_patches_png_mode_matches_oracle(tmp_path, monkeypatch, capsys, level):
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches, parse_xml_mask
    slide, opener = _setup_tree(tmp_path)
    monkeypatch.chdir(tmp_path)
    extract_patches(level=level, slide_opener=opener, max_slab_bytes=3 << 20)     # several row slabs
    out = capsys.readouterr().out
    mask = np.asarray(parse_xml_mask(str(tmp_path / "data/camelyon16/train/mask/annotations/tumor_900.xml"),
                                     slide.level_dimensions[level], slide))
    want = orc.extract_patches_oracle(slide.level_array(level), mask, level, want_images=False)
    names = sorted(orc.patch_name("tumor_900", x, y, l) for (x, y), l in zip(want["coords"], want["labels"]))
    d = tmp_path / "data" / "camelyon16" / "patches" / f"level_{level}" / "tumor_900"
    assert sorted(os.listdir(d)) == names and want["labels"].sum() > 0
    assert f"Patch extraction complete for tumor_900.tif at level {level}. Total patches: {len(names)}" in out
    p = want["patch"]
    for (x, y), l in list(zip(want["coords"], want["labels"]))[:: max(1, len(names) // 6)]:
        png = np.asarray(Image.open(d / orc.patch_name("tumor_900", x, y, l)))
        assert np.array_equal(png, orc.padded_patch(slide.level_array(level), int(x), int(y), p))
    extract_patches(level=level, slide_opener=opener)                              # idempotent: skip-if-exists
    assert "already extracted, skipping" in capsys.readouterr().out


def test_extract_features_artefacts_and_fused_cli(tmp_path, monkeypatch, capsys):
    from ss25_hierarchical_multiscale_image_classification_b200 import main as cli
    slide, opener = _setup_tree(tmp_path)
    monkeypatch.chdir(tmp_path)
    level = 2
    cli.main(["--patch", "--patch_level", str(level)], slide_opener=opener)
    torch.manual_seed(7)
    cli.main(["--extract_features", "--patch_level", str(level)])
    feats = np.load(f"patch_features_{level}.npy")
    labels = np.load(f"patch_labels_{level}.npy")
    paths = open(f"patch_paths_{level}.txt").read().split()
    assert feats.dtype == np.float32 and feats.shape == (len(paths), 512) and labels.dtype == np.int64
    assert [int("_tumor" in os.path.basename(p)) for p in paths] == labels.tolist()
    # same seeded weights on the oracle side
    from ss25_hierarchical_multiscale_image_classification_b200.models import ResNet18FeatureExtractor
    torch.manual_seed(7)
    model = ResNet18FeatureExtractor()
    import torchvision
    net = torchvision.models.resnet18(weights=None)
    net.load_state_dict(model._tv_state(), strict=False)
    net.fc = torch.nn.Linear(512, 2)
    net.eval()
    imgs = np.stack([orc.pil_resize_bilinear(np.asarray(Image.open(p).convert("RGB"))) for p in paths])
    ref, _ = orc.resnet18_features_fp32(net, imgs)
    cos, maxrel = _metrics(feats, ref)
    assert cos >= 0.9995 and maxrel <= 1e-2, (cos, maxrel)
    # PNG-free path produces the same artefacts (same paths; features from the same weights)
    os.rename(f"patch_features_{level}.npy", "png_features.npy")
    torch.manual_seed(7)
    cli.main(["--fused", "--patch_level", str(level)], slide_opener=opener)
    feats2 = np.load(f"patch_features_{level}.npy")
    paths2 = open(f"patch_paths_{level}.txt").read().split()
    assert sorted(paths2) == sorted(paths)
    order = [paths2.index(p) for p in paths]
    assert np.array_equal(feats2[order], np.load("png_features.npy"))              # identical kernels, identical inputs
    assert np.array_equal(np.load(f"patch_labels_{level}.npy")[order], labels)


def test_model_wrappers_keep_reference_state_dict_names_and_forward():
    import torchvision
    from ss25_hierarchical_multiscale_image_classification_b200.models import (ResNet18Classifier, ResNet18ClassifierSIMCLR,
                                                                                ResNet18FeatureExtractor, UnifiedResNet)
    tv = torchvision.models.resnet18(weights=None)
    fe = ResNet18FeatureExtractor()
    keys = set(fe.state_dict())
    assert "features.0.weight" in keys and "features.7.1.bn2.running_var" in keys and not any("fc" in k for k in keys)
    clf = ResNet18Classifier()
    assert set(clf.state_dict()) == {f"model.{k}" for k in tv.state_dict() if not k.startswith("fc.")} | {"model.fc.weight", "model.fc.bias"}
    assert clf.state_dict()["model.fc.weight"].shape == (2, 512)
    assert "encoder.conv1.weight" in ResNet18ClassifierSIMCLR().state_dict()
    assert isinstance(UnifiedResNet().model.fc, torch.nn.Identity)
    # forward with the reference's signature: float NCHW, normalised
    net = orc.make_resnet18(seed=3, classifier=True)
    clf.load_state_dict({f"model.{k}": v for k, v in net.state_dict().items()})
    clf = clf.cuda()
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(6, 224, 224, 3), dtype=np.uint8)
    x = torch.from_numpy(orc.normalize_u8(imgs)).permute(0, 3, 1, 2).cuda()
    logits = clf(x).cpu().numpy()
    ref_f, ref_l = orc.resnet18_features_fp32(net, imgs)
    assert np.abs(logits - ref_l).max() <= 2e-2 * np.abs(ref_l).max() + 1e-3
    uni = UnifiedResNet(classifier=False)
    uni.load_state_dict({f"model.{k}": v for k, v in net.state_dict().items() if not k.startswith("fc.")})
    f = uni.cuda()(x).cpu().numpy()
    cos, maxrel = _metrics(f, ref_f)
    assert cos >= 0.9995 and maxrel <= 1e-2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        clf(x.cpu())
    clf.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        clf(x)
