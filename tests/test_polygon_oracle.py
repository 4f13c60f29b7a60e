"""The polygon-fill restatement (oracle/pil_polygon.py) pinned against the installed Pillow: random polygons of every
flavour (self-intersecting, star-shaped, axis-aligned staircases, repeated vertices, annotation-like curves, vertices
outside the image) must rasterise to the same bytes as ``ImageDraw.polygon(xy, outline=255, fill=255)`` -- the call the
reference makes in ``parse_xml_mask`` (src/main.py:392-409)."""
import numpy as np
import pytest

from oracle import pil_polygon as pp


def random_polygon(rng, t, w, h):
    kind = t % 5
    nv = int(rng.integers(3, 12 if kind else 40))
    if kind == 0:
        return [(int(rng.integers(-10, w + 10)), int(rng.integers(-10, h + 10))) for _ in range(nv)]
    if kind == 1:
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        r = rng.uniform(2, max(w, h) / 2, nv)
        return [(int(cx + r[i] * np.cos(ang[i])), int(cy + r[i] * np.sin(ang[i]))) for i in range(nv)]
    if kind == 2:
        xy, x, y = [], int(rng.integers(0, w)), int(rng.integers(0, h))
        for _ in range(nv):
            if rng.random() < 0.5:
                x = int(rng.integers(0, w))
            else:
                y = int(rng.integers(0, h))
            xy.append((x, y))
        return xy
    if kind == 3:
        xy = [(int(rng.integers(0, 8)), int(rng.integers(0, 8))) for _ in range(nv)]
        if rng.random() < 0.3:
            xy.append(xy[0])
        return xy
    cx, cy = rng.uniform(w * 0.2, w * 0.8), rng.uniform(h * 0.2, h * 0.8)
    nv = int(rng.integers(20, 120))
    ang = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    r = (0.2 + 0.15 * np.sin(3 * ang + rng.uniform(0, 6)) + 0.05 * rng.standard_normal(nv)) * min(w, h)
    return [(int(cx + r[i] * np.cos(ang[i])), int(cy + r[i] * np.sin(ang[i]))) for i in range(nv)]


def pil_mask(polys, w, h):
    from PIL import Image, ImageDraw
    im = Image.new("L", (w, h), 0)
    d = ImageDraw.Draw(im)
    for xy in polys:
        d.polygon([tuple(map(int, p)) for p in xy], outline=255, fill=255)
    return np.array(im)


@pytest.mark.parametrize("seed", [0, 1])
def test_restatement_equals_installed_pillow(seed):
    rng = np.random.default_rng(seed)
    for t in range(700):
        w, h = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        xy = random_polygon(rng, t, w, h)
        assert np.array_equal(pp.polygon_mask([xy], w, h), pil_mask([xy], w, h)), (w, h, xy)


def test_known_answers_and_degenerate_inputs():
    # the discontiguous-corner rule: the apex row is pulled one pixel towards the previous row's span
    assert pp.polygon_mask([[(2, 3), (7, 6), (1, 3)]], 12, 10)[6].tolist() == [0] * 6 + [255, 255] + [0] * 4
    assert pp.polygon_mask([[(7, 2), (4, 7), (6, 2)]], 12, 10)[7].tolist() == [0] * 4 + [255] + [0] * 7
    for polys in ([[(3, 3), (3, 3)]], [[(1, 1), (5, 1)]], [[(1, 1), (1, 6)]], [[(2, 2), (6, 2), (6, 2), (2, 2)]], []):
        assert np.array_equal(pp.polygon_mask(polys, 9, 9), pil_mask(polys, 9, 9))
    # several overlapping annotations on one mask, some outside the image
    rng = np.random.default_rng(5)
    polys = [random_polygon(rng, t, 200, 150) for t in range(12)]
    assert np.array_equal(pp.polygon_mask(polys, 200, 150), pil_mask(polys, 200, 150))
