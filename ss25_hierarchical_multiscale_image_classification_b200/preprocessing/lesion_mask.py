"""Lesion annotations -> lesion mask IN DEVICE MEMORY.

The reference rasterises every CAMELYON16 annotation on the host with Pillow -- ``parse_xml_mask`` (``src/main.py:372-410``):
vertices ``(int(X * level_w / w0), int(Y * level_h / h0))``, ``ImageDraw.polygon(coords, outline=255, fill=255)`` on an "L"
image of the level size -- and the patch loop then crops that image (``src/main.py:705-716``).  Here the XML walk stays on the
host (a few kilobytes of vertices) and the rasterisation runs on the GPU, bit-exact against Pillow's polygon fill
(``hipac_polygon_fill``, ``csrc/polygon.cu``), directly into the 16-byte-pitched mask buffer the tile scan reads: no
``H x W`` host image, no host scan, no mask upload.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .tensor_api import alloc_level_image


def annotation_polygons(xml_path, level_dims, slide):
    """The vertex lists ``parse_xml_mask`` would draw: one ``int32 [n, 2]`` array per ``Annotation/Coordinates`` node, in
    document order, scaled by ``level_dims / level_dimensions[0]`` and truncated with ``int()`` (``src/main.py:388-405``).
    Returns ``None`` on a parse error (the reference prints and carries on without a mask)."""
    import xml.etree.ElementTree as ET
    try:
        tree = ET.parse(xml_path)
    except ET.ParseError as e:
        print(f"\033[91m[ERROR]\033[0m Error parsing XML file {xml_path}: {e}")
        return None
    base_dims = slide.level_dimensions[0]
    scale_x = level_dims[0] / base_dims[0]
    scale_y = level_dims[1] / base_dims[1]
    polys = []
    for annotation in tree.getroot().iter("Annotation"):
        for coordinates_node in annotation.findall("Coordinates"):
            coords = []
            for coord_node in coordinates_node.findall("Coordinate"):
                try:
                    x = float(coord_node.get("X"))
                    y = float(coord_node.get("Y"))
                    coords.append((int(x * scale_x), int(y * scale_y)))
                except (ValueError, TypeError) as e:
                    print(f"\033[93mWarning: Could not parse coordinate (X,Y) from XML for {xml_path}: {e}\033[0m")
                    continue
            if coords:
                polys.append(np.asarray(coords, dtype=np.int32).reshape(-1, 2))
    return polys


class PolygonSet:
    """Integer polygons packed for ``hipac_polygon_fill``: host copies (the library derives scan ranges from them) plus the
    vertex array on one device."""

    def __init__(self, polys, device):
        polys = [np.ascontiguousarray(np.asarray(p, dtype=np.int32).reshape(-1, 2)) for p in polys]
        if any(len(p) < 2 for p in polys):
            # Pillow refuses such a polygon (ImageDraw.polygon -> TypeError); the reference's caller then reports the
            # annotation file as unparsable and carries on without a mask (src/main.py:668-675)
            raise TypeError("coordinate list must contain at least 2 coordinates")
        self.n = len(polys)
        self.offsets = np.zeros(self.n + 1, dtype=np.int32)
        for i, p in enumerate(polys):
            self.offsets[i + 1] = self.offsets[i] + len(p)
        self.xy = np.concatenate(polys).astype(np.int32) if self.n and self.offsets[-1] else np.zeros((0, 2), np.int32)
        self.device = torch.device(device)
        self.d_xy = torch.from_numpy(self.xy if len(self.xy) else np.zeros((1, 2), np.int32)).to(self.device)
        self.ws = torch.zeros((int(_lib.lib().hipac_polygon_workspace_bytes(self.n)),), dtype=torch.uint8, device=self.device)


def rasterize_polygons(polys, width: int, height: int, device="cuda", y_begin: int = 0, n_rows: int | None = None,
                       out: torch.Tensor | None = None, stream=None, check: bool = True) -> torch.Tensor:
    """uint8 ``[n_rows, W]`` device mask (16-byte row pitch) of rows ``[y_begin, y_begin + n_rows)`` of the ``height x width``
    level image: every polygon (a ``PolygonSet`` or a list of integer ``(x, y)`` vertex arrays) filled with 255 exactly as
    ``ImageDraw.polygon(xy, outline=255, fill=255)`` would.  ``check`` reads the overflow flag back (synchronises)."""
    ps = polys if isinstance(polys, PolygonSet) else PolygonSet(polys, device)
    n_rows = height - y_begin if n_rows is None else int(n_rows)
    mask = out if out is not None else alloc_level_image(n_rows, width, ps.device, channels=1)
    if int(mask.shape[0]) < n_rows or int(mask.shape[1]) != width or mask.dtype != torch.uint8 or mask.stride(1) != 1:
        raise ValueError("out must be a uint8 [>= n_rows, W] tensor with unit column stride")
    l = _lib.lib()
    st = stream or torch.cuda.current_stream(ps.device)
    with torch.cuda.device(ps.device):
        _lib.check(l.hipac_polygon_fill(ps.xy.ctypes.data, ps.offsets.ctypes.data, ps.n, ps.d_xy.data_ptr(), mask.data_ptr(), int(height),
                                        int(width), int(mask.stride(0)), int(y_begin), n_rows, 1, ps.ws.data_ptr(), int(ps.ws.numel()),
                                        st.cuda_stream), "hipac_polygon_fill")
        if check and l.hipac_polygon_overflowed(ps.ws.data_ptr(), st.cuda_stream) != 0:
            raise RuntimeError("a scan line crosses more than 1024 polygon edges; the device rasteriser's row buffer overflowed")
    return mask[:n_rows]


def parse_xml_mask_device(xml_path, level_dims, slide, device="cuda"):
    """``parse_xml_mask`` with the mask rasterised on the GPU: uint8 ``[H, W]`` device tensor, or ``None`` on a parse error."""
    polys = annotation_polygons(xml_path, level_dims, slide)
    if polys is None:
        return None
    return rasterize_polygons(polys, int(level_dims[0]), int(level_dims[1]), device)
