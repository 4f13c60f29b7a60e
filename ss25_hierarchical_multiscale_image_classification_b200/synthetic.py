"""Deterministic synthetic whole-slide pyramids.

Every pixel is a pure integer function of ``(seed, level, x, y, c)`` so the
same slide can be regenerated on any host without shipping gigabytes of
fixtures.  A slide is a white-ish background (value 244..247, rejected by the
reference's ``mean > 240`` tissue test, reference ``src/main.py:718-720``), a
union of seeded tissue ellipses filled with blocky texture + noise, an
optional "faint" strip whose mean sits just under the threshold, and 1-3
lesion ellipses inside the tissue (the rasterised lesion mask that
``parse_xml_mask`` would return, reference ``src/main.py:372-410``).

``SyntheticSlide`` duck-types the three OpenSlide members the hot path uses
(reference ``src/main.py:654-655,693-697``): ``level_dimensions``,
``level_downsamples`` and ``read_region``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["hash32", "SyntheticSlide", "make_level", "make_lesion_mask", "seeded_resnet18"]

_M32 = np.uint32(0xFFFFFFFF)


def hash32(x, y, c, seed):
    """Counter-based 32-bit integer hash (all arithmetic mod 2**32)."""
    x = np.asarray(x, dtype=np.uint32)
    y = np.asarray(y, dtype=np.uint32)
    c = np.asarray(c, dtype=np.uint32)
    with np.errstate(over="ignore"):
        h = (x * np.uint32(0x9E3779B1) + y * np.uint32(0x85EBCA77)
             + c * np.uint32(0xC2B2AE3D) + np.uint32(seed & 0xFFFFFFFF) * np.uint32(0x27D4EB2F))
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x2C1B3C6D)
        h ^= h >> np.uint32(12)
        h *= np.uint32(0x297A2D39)
        h ^= h >> np.uint32(15)
    return h


def _ellipses(seed: int, n: int, w: int, h: int, lo: float, hi: float, inside=None):
    """Seeded ellipse list ``(cx, cy, rx, ry)`` in integer pixels of a (w, h) level."""
    out = []
    for i in range(n):
        r = [int(hash32(i, k, 7, seed)) / 4294967296.0 for k in range(4)]
        if inside is None:
            cx, cy = int((0.2 + 0.6 * r[0]) * w), int((0.2 + 0.6 * r[1]) * h)
        else:  # place inside a parent ellipse
            pcx, pcy, prx, pry = inside[i % len(inside)]
            cx = int(pcx + (r[0] - 0.5) * prx)
            cy = int(pcy + (r[1] - 0.5) * pry)
        rx = max(1, int((lo + (hi - lo) * r[2]) * w))
        ry = max(1, int((lo + (hi - lo) * r[3]) * h))
        out.append((cx, cy, rx, ry))
    return out


def _inside(ells, xs, ys):
    """Boolean [len(ys), len(xs)] membership in the union of ellipses (int64 exact)."""
    xs = xs.astype(np.int64)[None, :]
    ys = ys.astype(np.int64)[:, None]
    m = np.zeros((ys.shape[0], xs.shape[1]), dtype=bool)
    for cx, cy, rx, ry in ells:
        m |= ((xs - cx) * ry) ** 2 + ((ys - cy) * rx) ** 2 <= (rx * ry) ** 2
    return m


def _geometry(seed: int, level: int, w: int, h: int):
    # Geometry is defined in level-independent fractions, so every level shows
    # the same tissue layout (levels are generated independently, not by
    # downsampling: SURVEY.md §8d).
    tissue = _ellipses(seed * 31 + 1, 3, w, h, 0.12, 0.30)
    lesion = _ellipses(seed * 31 + 2, 2, w, h, 0.03, 0.08, inside=tissue)
    faint_x0 = int(0.55 * w)
    faint_x1 = int(0.62 * w)
    return tissue, lesion, faint_x0, faint_x1


def make_level(seed: int, level: int, w: int, h: int, y0: int = 0, y1: int | None = None,
               x0: int = 0, x1: int | None = None) -> np.ndarray:
    """RGB uint8 ``[y1-y0, x1-x0, 3]`` window of the synthetic level image."""
    y1 = h if y1 is None else y1
    x1 = w if x1 is None else x1
    tissue, _, fx0, fx1 = _geometry(seed, level, w, h)
    xs = np.arange(x0, x1, dtype=np.uint32)
    ys = np.arange(y0, y1, dtype=np.uint32)
    X = xs[None, :, None]
    Y = ys[:, None, None]
    C = np.arange(3, dtype=np.uint32)[None, None, :]
    s = (seed * 4 + level) & 0xFFFFFFFF
    hn = hash32(X, Y, C, s)                                  # per-pixel noise
    hb = hash32(X >> np.uint32(5), Y >> np.uint32(5), np.uint32(9), s)  # 32x32 block texture
    tint = np.array([10, -20, 5], dtype=np.int32)[None, None, :]
    val = (150 + (hb % np.uint32(96)).astype(np.int32) - 48 + tint
           + (hn & np.uint32(31)).astype(np.int32) - 16)
    bg = 244 + ((hn >> np.uint32(8)) & np.uint32(3)).astype(np.int32)
    faint = 236 + ((hn >> np.uint32(12)) & np.uint32(7)).astype(np.int32)
    tis = _inside(tissue, xs, ys)[:, :, None]
    in_faint = ((xs >= fx0) & (xs < fx1))[None, :, None]
    img = np.where(tis, np.where(in_faint, faint, val), bg)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_lesion_mask(seed: int, level: int, w: int, h: int, y0: int = 0, y1: int | None = None) -> np.ndarray:
    """uint8 {0,255} ``[y1-y0, w]`` rasterised lesion mask at this level."""
    y1 = h if y1 is None else y1
    _, lesion, _, _ = _geometry(seed, level, w, h)
    xs = np.arange(0, w, dtype=np.uint32)
    ys = np.arange(y0, y1, dtype=np.uint32)
    return np.where(_inside(lesion, xs, ys), 255, 0).astype(np.uint8)


class SyntheticSlide:
    """In-memory pyramid with the OpenSlide members the hot path calls.

    ``level_downsamples[L] == 2**L`` exactly and ``level_dimensions[L] ==
    (W0 >> L, H0 >> L)``.  ``levels`` may be passed explicitly (list of
    ``[H, W, 3]`` uint8 arrays) to wrap arbitrary test images.
    """

    def __init__(self, width0: int | None = None, height0: int | None = None, seed: int = 1234,
                 n_levels: int = 4, levels: list[np.ndarray] | None = None, name: str = "tumor_900",
                 with_lesion: bool = True):
        self.name = name
        self.seed = seed
        self.with_lesion = with_lesion
        if levels is not None:
            self._levels = {i: np.ascontiguousarray(a) for i, a in enumerate(levels)}
            self.level_dimensions = tuple((a.shape[1], a.shape[0]) for a in levels)
        else:
            self._levels = {}
            self.level_dimensions = tuple((width0 >> l, height0 >> l) for l in range(n_levels))
        self.level_count = len(self.level_dimensions)
        self.level_downsamples = tuple(float(2 ** l) for l in range(self.level_count))
        self.dimensions = self.level_dimensions[0]

    # -- array access -----------------------------------------------------
    def level_array(self, level: int) -> np.ndarray:
        if level not in self._levels:
            w, h = self.level_dimensions[level]
            self._levels[level] = make_level(self.seed, level, w, h)
        return self._levels[level]

    def lesion_mask(self, level: int) -> np.ndarray | None:
        if not self.with_lesion:
            return None
        w, h = self.level_dimensions[level]
        return make_lesion_mask(self.seed, level, w, h)

    # -- OpenSlide duck type ------------------------------------------------
    def read_region(self, location, level, size):
        """RGBA PIL image; ``location`` is in level-0 pixels (reference src/main.py:693-697)."""
        from PIL import Image
        ds = int(self.level_downsamples[level])
        x, y = int(location[0]) // ds, int(location[1]) // ds
        w, h = int(size[0]), int(size[1])
        arr = self.level_array(level)
        H, W = arr.shape[:2]
        out = np.zeros((h, w, 4), dtype=np.uint8)  # OpenSlide: transparent black outside the slide
        xs, ys = max(x, 0), max(y, 0)
        xe, ye = min(x + w, W), min(y + h, H)
        if xe > xs and ye > ys:
            out[ys - y:ye - y, xs - x:xe - x, :3] = arr[ys:ye, xs:xe]
            out[ys - y:ye - y, xs - x:xe - x, 3] = 255
        return Image.fromarray(out, "RGBA")

    def close(self):
        pass


def seeded_resnet18(seed: int = 0, classifier: bool = True):
    """Seeded random-init torchvision resnet18 (+ ``Linear(512,2)`` head), eval mode -- the "random-init ResNet18"
    of BASELINE.json's configs (the reference's trained checkpoint is not in its checkout, and its
    ``--extract_features`` runs on random-init weights anyway).

    BatchNorm statistics and affine parameters are perturbed away from (0, 1, 1, 0) so that BN folding is
    actually exercised."""
    import torch
    import torchvision

    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    net = torchvision.models.resnet18(weights=None)
    if classifier:
        net.fc = torch.nn.Linear(512, 2)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.75 + 0.5 * torch.rand(m.num_features, generator=g))
                m.weight.copy_(0.8 + 0.4 * torch.rand(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return net.eval()
