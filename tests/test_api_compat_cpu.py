"""Host-side pieces of the drop-in API that need no GPU: XML mask parsing, PatchDataset, CLI flag handling."""
import os

import numpy as np
import pytest
from PIL import Image, ImageDraw

from ss25_hierarchical_multiscale_image_classification_b200.datasets import PatchDataset
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing.extract import parse_xml_mask
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide

XML = """<?xml version="1.0"?>
<ASAP_Annotations>
  <Annotations>
    <Annotation Name="_0" Type="Polygon" PartOfGroup="Tumor" Color="#F4FA58">
      <Coordinates>
        <Coordinate Order="0" X="1003.7" Y="2001.2" />
        <Coordinate Order="1" X="3999.9" Y="2100.0" />
        <Coordinate Order="2" X="3500.5" Y="5200.4" />
        <Coordinate Order="3" X="900.0" Y="4800.9" />
      </Coordinates>
    </Annotation>
    <Annotation Name="_1" Type="Polygon" PartOfGroup="Tumor" Color="#F4FA58">
      <Coordinates>
        <Coordinate Order="0" X="7000" Y="7000" />
        <Coordinate Order="1" X="7900" Y="7050" />
        <Coordinate Order="2" X="bad" Y="7050" />
        <Coordinate Order="3" X="7500" Y="7900" />
      </Coordinates>
    </Annotation>
  </Annotations>
  <AnnotationGroups />
</ASAP_Annotations>
"""


def test_parse_xml_mask_scales_truncates_and_fills(tmp_path, capsys):
    # reference src/main.py:386-409: scale = level_dims / level0_dims, int() truncation, fill = outline = 255
    slide = SyntheticSlide(12000, 9000, seed=1)
    xml = tmp_path / "tumor_900.xml"
    xml.write_text(XML)
    level = 3
    dims = slide.level_dimensions[level]
    mask = parse_xml_mask(str(xml), dims, slide)
    assert mask.mode == "L" and mask.size == dims
    want = Image.new("L", dims, 0)
    d = ImageDraw.Draw(want)
    s = dims[0] / 12000, dims[1] / 9000
    d.polygon([(int(x * s[0]), int(y * s[1])) for x, y in [(1003.7, 2001.2), (3999.9, 2100.0), (3500.5, 5200.4), (900.0, 4800.9)]],
              outline=255, fill=255)
    d.polygon([(int(x * s[0]), int(y * s[1])) for x, y in [(7000, 7000), (7900, 7050), (7500, 7900)]], outline=255, fill=255)
    assert np.array_equal(np.asarray(mask), np.asarray(want))
    assert set(np.unique(np.asarray(mask))) == {0, 255}
    assert "Could not parse coordinate" in capsys.readouterr().out          # the bad coordinate is skipped, not fatal


def test_parse_xml_mask_syntax_error_returns_none(tmp_path, capsys):
    bad = tmp_path / "x.xml"
    bad.write_text("<ASAP_Annotations><Annotations>")
    assert parse_xml_mask(str(bad), (100, 100), SyntheticSlide(800, 800)) is None
    assert "[ERROR]" in capsys.readouterr().out


def _make_tree(root):
    rng = np.random.default_rng(0)
    names = {"tumor_001": ["tumor_001_x0_y0_tumor.png", "tumor_001_x224_y0_normal.png", "tumor_001_x448_y0_normal.png"],
             "normal_002": ["normal_002_x0_y224_normal.png", "normal_002_x0_y448_normal.png", "normal_002_x9_y9_weird.png"]}
    for slide, files in names.items():
        os.makedirs(root / slide)
        for f in files:
            Image.fromarray(rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8)).save(root / slide / f)


def test_patch_dataset_contract(tmp_path, capsys):
    _make_tree(tmp_path)
    ds = PatchDataset(str(tmp_path), transform=lambda im: np.asarray(im).mean(), seed=3)
    out = capsys.readouterr().out
    assert "Could not determine label from filename: normal_002_x9_y9_weird.png" in out
    assert len(ds) == 5 and ds.get_class_counts() == {0: 4, 1: 1}
    img, label, path = ds[0]
    assert isinstance(img, float) and label == (1 if "_tumor" in os.path.basename(path) else 0)
    assert sorted(ds.labels) == [0, 0, 0, 0, 1]
    # same seed -> same order; balanced sampling takes min class count from each class
    assert PatchDataset(str(tmp_path), seed=3).image_paths == ds.image_paths
    bal = PatchDataset(str(tmp_path), balanced=True, seed=1)
    assert bal.get_class_counts() == {0: 1, 1: 1}
    only = PatchDataset(str(tmp_path), slide_names=["normal_002"], seed=1)
    assert len(only) == 2 and all("normal_002" in p for p in only.image_paths)
    t = PatchDataset(str(tmp_path), tumor_transform=lambda im: "T", normal_transform=lambda im: "N", seed=1)
    assert all(t[i][0] == ("T" if t[i][1] == 1 else "N") for i in range(len(t)))


def test_cli_rejects_unknown_flags_and_gates_on_data(tmp_path, monkeypatch, capsys):
    from ss25_hierarchical_multiscale_image_classification_b200 import main as cli
    monkeypatch.chdir(tmp_path)
    with pytest.raises(SystemExit) as e:
        cli.main(["--patch", "--base_dir", "x"])            # README-only flag of the reference: rejected there too
    assert e.value.code == 1 and "Unknown command line arguments: base_dir" in capsys.readouterr().out
    cli.main(["--patch", "--patch_level", "2"])
    assert "Images must be downloaded before extracting patches." in capsys.readouterr().out
    cli.main(["--extract_features"])
    assert "Patches must be extracted at level 3 before extracting features." in capsys.readouterr().out
    assert not cli.features_extracted(3)


def test_heatmap_and_froc_csv(tmp_path):
    import torch
    from ss25_hierarchical_multiscale_image_classification_b200 import heatmap as hm
    coords = torch.tensor([[0, 0], [448, 224], [224, 672]], dtype=torch.int32)
    logits = torch.tensor([[2.0, 0.0], [0.0, 2.0], [0.0, 0.0]])
    grid = hm.heatmap(coords, logits, width=1000, height=900, stride=224, fill=-1.0)
    assert grid.shape == (5, 5)
    p = torch.softmax(logits, 1)[:, 1]
    assert torch.allclose(grid[0, 0], p[0]) and torch.allclose(grid[1, 2], p[1]) and torch.allclose(grid[3, 1], p[2])
    assert int((grid == -1.0).sum()) == 22
    n = hm.write_froc_csv(str(tmp_path / "t.csv"), coords, logits, level=2, patch=448, threshold=0.4)
    rows = [l.split(",") for l in open(tmp_path / "t.csv").read().split()]
    assert n == 2 and rows[0][1:] == [str((448 + 224) * 4), str((224 + 224) * 4)]   # level-0 coordinates of the centre
    assert abs(float(rows[0][0]) - float(p[1])) < 1e-5


def test_mil_classifier_cpu_contract():
    """MILClassifier keeps the reference's constructor, state-dict keys and (logits, attention) return on the CPU path."""
    import torch
    from ss25_hierarchical_multiscale_image_classification_b200.models import MILAttentionPooling, MILClassifier
    torch.manual_seed(0)
    bag = torch.randn(17, 512)
    m = MILClassifier(512, num_classes=2, pooling="attention")
    logits, attn = m(bag)
    assert logits.shape == (2,) and attn.shape == (17, 1) and abs(float(attn.sum()) - 1) < 1e-5
    pooled, a = MILAttentionPooling(512)(bag)
    assert pooled.shape == (512,) and a.shape == (17, 1)
    for pooling in ("mean", "max"):
        lg, at = MILClassifier(512, pooling=pooling)(bag)
        assert lg.shape == (2,) and at is None
    import pytest
    with pytest.raises(ValueError):
        MILClassifier(512, pooling="median")
