"""hipac_polygon_fill (csrc/polygon.cu, through the C ABI) vs the polygon oracle and the installed Pillow: bit-exact masks."""
import numpy as np
import pytest

from oracle import pil_polygon as pp
from test_polygon_oracle import pil_mask, random_polygon

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _gpu(polys, w, h, **kw):
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing.lesion_mask import rasterize_polygons
    return rasterize_polygons(polys, w, h, "cuda", **kw).cpu().numpy()


def test_random_polygons_equal_pillow_and_oracle():
    rng = np.random.default_rng(11)
    for t in range(600):
        w, h = int(rng.integers(8, 140)), int(rng.integers(8, 140))
        xy = random_polygon(rng, t, w, h)
        got = _gpu([xy], w, h, check=(t % 50 == 0))
        want = pil_mask([xy], w, h)
        assert np.array_equal(got, want), (w, h, xy, np.argwhere(got != want)[:5].tolist())
        if t % 10 == 0:
            assert np.array_equal(got, pp.polygon_mask([xy], w, h))


def test_many_annotations_large_mask_and_row_windows():
    """A slide-like case: dozens of overlapping annotations with hundreds of vertices on a 3000 x 2500 mask; the row-window
    form (a slab of the level, what scan_slide uses) equals the corresponding rows of the full rasterisation."""
    rng = np.random.default_rng(3)
    w, h = 3000, 2500
    polys = []
    for t in range(40):
        cx, cy = rng.uniform(-100, w + 100), rng.uniform(-100, h + 100)
        nv = int(rng.integers(30, 900))
        ang = np.linspace(0, 2 * np.pi, nv, endpoint=False)
        r = (0.03 + 0.1 * rng.random()) * w * (1 + 0.3 * np.sin(5 * ang + rng.uniform(0, 6)) + 0.05 * rng.standard_normal(nv))
        polys.append([(int(cx + r[i] * np.cos(ang[i])), int(cy + r[i] * np.sin(ang[i]))) for i in range(nv)])
    want = pil_mask(polys, w, h)
    got = _gpu(polys, w, h)
    assert got.shape == (h, w) and np.array_equal(got, want)
    assert 0.02 < (want > 0).mean() < 0.9
    for y0, n in ((0, 700), (693, 1001), (2499, 1), (1200, 1300)):
        assert np.array_equal(_gpu(polys, w, h, y_begin=y0, n_rows=n), want[y0:y0 + n]), (y0, n)


def test_degenerate_polygons_and_errors():
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing.lesion_mask import rasterize_polygons
    for polys in ([[(3, 3), (3, 3)]], [[(1, 1), (5, 1)]], [[(1, 1), (1, 6)]], [[(2, 2), (6, 2), (6, 2), (2, 2)]], [],
                  [[(-5, -5), (20, -5), (20, 20), (-5, 20)]], [[(0, 0), (8, 8), (8, 0), (0, 8)]]):
        assert np.array_equal(_gpu(polys, 9, 9), pil_mask(polys, 9, 9)), polys
    with pytest.raises(TypeError):
        rasterize_polygons([[(3, 3)]], 9, 9, "cuda")                    # Pillow raises the same for a single vertex
    # a scan line crossing more edges than the kernel's row buffer holds is reported, not silently mis-drawn
    zig = [(x, 0 if x % 2 == 0 else 50) for x in range(1300)] + [(1300, 60), (0, 60)]
    with pytest.raises(RuntimeError, match="overflowed"):
        rasterize_polygons([zig], 1400, 64, "cuda")


def test_parse_xml_mask_device_equals_reference_style_host_mask(tmp_path):
    """XML annotation file -> device mask == the PIL mask ``parse_xml_mask`` draws (same vertices, same fill)."""
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import parse_xml_mask
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing.lesion_mask import parse_xml_mask_device
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(20000, 16000, seed=4)
    rng = np.random.default_rng(8)
    xml = ['<?xml version="1.0"?><ASAP_Annotations><Annotations>']
    for a in range(5):
        cx, cy = rng.uniform(2000, 18000), rng.uniform(2000, 14000)
        nv = int(rng.integers(20, 300))
        ang = np.linspace(0, 2 * np.pi, nv, endpoint=False)
        r = rng.uniform(300, 2500) * (1 + 0.2 * np.sin(4 * ang))
        xml.append(f'<Annotation Name="_{a}" Type="Polygon" PartOfGroup="Tumor" Color="#F4FA58"><Coordinates>')
        xml += [f'<Coordinate Order="{i}" X="{cx + r[i] * np.cos(ang[i]):.4f}" Y="{cy + r[i] * np.sin(ang[i]):.4f}" />' for i in range(nv)]
        xml.append('</Coordinates></Annotation>')
    xml.append('</Annotations></ASAP_Annotations>')
    path = tmp_path / "tumor_900.xml"
    path.write_text("\n".join(xml))
    for level in (2, 3):
        dims = slide.level_dimensions[level]
        want = np.asarray(parse_xml_mask(str(path), dims, slide))
        got = parse_xml_mask_device(str(path), dims, slide, "cuda").cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want) and want.any()
