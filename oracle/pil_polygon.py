"""CPU restatement of Pillow's filled-polygon rasteriser -- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference rasterises CAMELYON16 lesion annotations with ``ImageDraw.polygon(coords, outline=255, fill=255)`` on an
"L" image (``src/main.py:388-409``, ``int(x * scale)`` vertices).  Pillow is a third-party dependency that is not vendored
in the reference (``src/requirements.txt``: ``pillow @ file:///...`` conda build, version not recoverable; the container
has Pillow 12.2.0), so its published algorithm is restated here and PINNED against the installed Pillow on thousands of
random polygons (``tests/test_polygon_oracle.py``) and against the reference's own ``parse_xml_mask`` run under the
harness (``tests/golden/polygon_masks.npz``).

Pillow ``src/libImaging/Draw.c``:
  * ``ImageDraw.polygon``: with ``outline == fill`` only the fill is drawn (``ink != fill_ink`` gate in ImageDraw.py).
  * ``ImagingDrawPolygon``: edge list from consecutive vertices (consecutive collinear horizontal edges are merged into the
    previous edge's x range), closing edge unless the last vertex equals the first.
  * ``add_edge``: ``dx = (float)(x1 - x0) / (y1 - y0)`` in float32; horizontal edges ``d = 0, dx = 0``.
  * ``polygon_generic``: horizontal edges are drawn directly with ``hline``; for every scan line ``ymin..ymax`` the float32
    intersections ``(y - y0) * dx + x0`` of all edges with ``ymin <= y <= ymax`` are collected (the intersection of an edge's
    LAST row is duplicated when it is not the polygon's last row; the "connect discontiguous corners" rule rewrites one
    earlier entry), sorted, and consecutive pairs are filled from ``ROUND_UP(xx[i-1])`` to ``ROUND_DOWN(xx[i])`` with the
    running ``x_pos`` clamp.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def _round_up(f: float) -> int:
    return int(math.floor(F32(f + F32(0.5)))) if f >= 0.0 else -int(math.floor(F32(F32(abs(f)) + F32(0.5))))


def _round_down(f: float) -> int:
    return int(math.ceil(F32(f - F32(0.5)))) if f >= 0.0 else -int(math.ceil(F32(F32(abs(f)) - F32(0.5))))


class Edge:
    __slots__ = ("d", "x0", "y0", "xmin", "ymin", "xmax", "ymax", "dx")

    def __init__(self, x0, y0, x1, y1):
        self.xmin, self.xmax = (x0, x1) if x0 <= x1 else (x1, x0)
        self.ymin, self.ymax = (y0, y1) if y0 <= y1 else (y1, y0)
        if y0 == y1:
            self.d, self.dx = 0, F32(0.0)
        else:
            self.dx = F32(F32(x1 - x0) / F32(y1 - y0))
            self.d = 1 if y0 == self.ymin else -1
        self.x0, self.y0 = x0, y0


def build_edges(xy):
    """``ImagingDrawPolygon`` (fill branch): list of integer ``(x, y)`` vertices -> edge list."""
    count = len(xy)
    edges = []
    i = 0
    for i in range(count - 1):
        x0, y0 = xy[i]
        x1, y1 = xy[i + 1]
        if y0 == y1 and i != 0 and y0 == xy[i - 1][1]:
            last = edges[-1]
            if x1 > x0 and x0 > xy[i - 1][0]:
                last.xmax = x1
                continue
            elif x1 < x0 and x0 < xy[i - 1][0]:
                last.xmin = x1
                continue
        edges.append(Edge(x0, y0, x1, y1))
    i = count - 1
    if xy[i][0] != xy[0][0] or xy[i][1] != xy[0][1]:
        edges.append(Edge(xy[i][0], xy[i][1], xy[0][0], xy[0][1]))
    return edges


def _x_at(e: Edge, y: int):
    return F32(F32(F32(y - e.y0) * e.dx) + F32(e.x0))


def _hline(img, x0, y, x1, ink):
    h, w = img.shape
    if 0 <= y < h:
        if x0 < 0:
            x0 = 0
        elif x0 >= w:
            return
        if x1 < 0:
            return
        elif x1 >= w:
            x1 = w - 1
        if x0 <= x1:
            img[y, x0:x1 + 1] = ink


def _roundf(v) -> np.float32:
    """C ``roundf``: half away from zero."""
    v = float(v)
    return F32(math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5))


def fill_polygon(img: np.ndarray, xy, ink: int = 255):
    """``polygon_generic`` (Pillow 12.2.0, 8-bit path) on a uint8 ``[H, W]`` array, in place.

    Read back from the installed binary (``PIL/_imaging*.so``, ``polygon_generic``), since the C source is not in the
    container: per scan line, every non-horizontal edge active on the row contributes ``(y - y0) * dx + x0`` (float32
    multiply then add).  An edge ENDING on the row (and the row is not the polygon's last) contributes it twice.  An edge
    with a VERTEX on the row (its first row, or its last row on the polygon's last row) looks for an EARLIER edge that has
    a vertex at the same rounded x on this row and is active on the adjacent row (next row; previous row for an ending
    edge): if this corner lies more than one pixel beyond BOTH edges' positions on the adjacent row, the intersection is
    pulled to ``roundf(max) + 1`` / ``roundf(min) - 1`` so that the corner stays 8-connected to the adjacent row's span.
    Pairs of the sorted list are filled from ``ROUND_UP`` to ``ROUND_DOWN`` (no running clamp on the 8-bit path)."""
    if len(xy) <= 0:
        return
    h, w = img.shape
    e = build_edges(xy)
    if len(e) <= 0:
        return
    table = []
    ymin, ymax = h - 1, 0
    for ed in e:
        if ymin > ed.ymin:
            ymin = ed.ymin
        if ymax < ed.ymax:
            ymax = ed.ymax
        if ed.ymin == ed.ymax:
            _hline(img, ed.xmin, ed.ymin, ed.xmax, ink)
            continue
        table.append(ed)
    if ymin < 0:
        ymin = 0
    if ymax > h:
        ymax = h
    for y in range(ymin, ymax + 1):
        xx = []
        for i, cur in enumerate(table):
            if not (cur.ymin <= y <= cur.ymax):
                continue
            x = _x_at(cur, y)
            if y == cur.ymax and y < ymax:
                xx.append(x)
                xx.append(x)
                continue
            if (y == cur.ymin or y == cur.ymax) and cur.dx != 0:
                adj = y - 1 if y == cur.ymax else y + 1
                for k in range(i):
                    other = table[k]
                    if y != other.ymin and y != other.ymax:
                        continue
                    if other.dx == 0:
                        continue
                    if _roundf(x) != _roundf(_x_at(other, y)):
                        continue
                    if adj < other.ymin or adj > other.ymax:
                        continue
                    a, b = _x_at(cur, adj), _x_at(other, adj)
                    one = F32(1.0)
                    if x > F32(a + one) and x > F32(b + one):
                        x = F32(_roundf(max(a, b)) + one)
                    elif F32(a - one) > x and F32(b - one) > x:
                        x = F32(_roundf(min(a, b)) - one)
                    break
            xx.append(x)
        xx.sort()
        for i in range(1, len(xx), 2):
            _hline(img, _round_up(float(xx[i - 1])), y, _round_down(float(xx[i])), ink)


def polygon_mask(polys, width: int, height: int) -> np.ndarray:
    """uint8 ``[H, W]`` mask: every polygon (list of integer ``(x, y)``) filled with 255, in order."""
    img = np.zeros((height, width), np.uint8)
    for xy in polys:
        fill_polygon(img, [(int(x), int(y)) for x, y in xy], 255)
    return img
