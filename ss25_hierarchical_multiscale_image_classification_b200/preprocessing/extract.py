"""Drop-in ``extract_patches`` / ``extract_patches_per_slide`` / ``parse_xml_mask``.

Same names, keyword arguments, defaults, directory layout, file names, log lines and
print-and-continue error behaviour as the reference (``src/main.py:609-732``, ``252-370``,
``372-410``).  What changed is *who decides*: the candidate grid, white padding, tissue test and
lesion vote run on the GPU (``hipac_tile_scan``); the host only reads the level image from the slide
reader, uploads it in row slabs and -- in this PNG compatibility mode -- encodes the surviving
full-size patches exactly where the reference put them, so ``PatchDataset`` / ``extract_features`` /
the reference's own downstream code consume the output unchanged.

The slide reader is duck-typed OpenSlide (``level_dimensions``, ``level_downsamples``,
``read_region``); pass ``slide_opener=`` to inject one (tests use ``SyntheticSlide``), otherwise
``openslide.OpenSlide`` is imported lazily.

Differences on purpose:
  * pixels are read as full-width row slabs at ``(0, int(y0 * downsample))`` and patches are cut out of the slab on the
    GPU, where the reference issues one ``read_region((int(x * ds), int(y * ds)), level, (w, h))`` per candidate
    (``src/main.py:693-697``).  For pyramids whose ``level_downsamples`` are exact integers (``2 ** level``: every
    synthetic slide here, and CAMELYON16 levels whose level-0 size is divisible by ``2 ** level``) the two address the
    same level pixels.  With a non-integer downsample (odd level-0 dimensions) ``int(x * ds)`` is not a multiple of the
    downsample and the reader resolves the level-0 location to a level pixel by its own rounding, so a per-patch read
    may start up to one LEVEL pixel off the slab's column ``x``; parity with the reference is only claimed (and tested)
    for exact downsamples.
  * a full-width RGBA ``read_region`` of a slab is a large host allocation; ``max_slab_bytes`` bounds it.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from PIL import Image, ImageDraw

from .tensor_api import alloc_level_image, extract_patches_tensor, grid_shape, patch_and_stride, upload_level_rows


class bcolors:  # same tags as the reference (src/main.py:35-44)
    INFO = '\033[95m'
    WARNING = '\033[93m'
    ERROR = '\033[91m'
    ENDC = '\033[0m'


def parse_xml_mask(xml_path, level_dims, slide):
    """CAMELYON16 XML annotation -> PIL "L" mask at ``level_dims`` (reference ``src/main.py:372-410``).

    Coordinates are scaled by ``level_dims / level_dimensions[0]`` and truncated with ``int()``; every
    ``Annotation/Coordinates`` polygon is drawn with fill = outline = 255.  Uses the standard-library
    XML parser (the reference's lxml is not a dependency here)."""
    import xml.etree.ElementTree as ET
    try:
        tree = ET.parse(xml_path)
    except ET.ParseError as e:
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Error parsing XML file {xml_path}: {e}")
        return None
    base_dims = slide.level_dimensions[0]
    scale_x = level_dims[0] / base_dims[0]
    scale_y = level_dims[1] / base_dims[1]
    mask = Image.new("L", tuple(level_dims), 0)
    draw = ImageDraw.Draw(mask)
    for annotation in tree.getroot().iter("Annotation"):
        for coordinates_node in annotation.findall("Coordinates"):
            coords = []
            for coord_node in coordinates_node.findall("Coordinate"):
                try:
                    x = float(coord_node.get("X"))
                    y = float(coord_node.get("Y"))
                    coords.append((int(x * scale_x), int(y * scale_y)))
                except (ValueError, TypeError) as e:
                    print(f"{bcolors.WARNING}Warning: Could not parse coordinate (X,Y) from XML for {xml_path}: {e}{bcolors.ENDC}")
                    continue
            if coords:
                draw.polygon(coords, outline=255, fill=255)
    return mask


def _default_opener(path):
    import openslide  # lazy: only needed for real slides
    return openslide.OpenSlide(path)


def read_level_rows(slide, level: int, y0: int, y1: int) -> np.ndarray:
    """RGB uint8 ``[y1-y0, W, 3]`` rows of a level via ``read_region`` (location in level-0 pixels,
    reference ``src/main.py:693-697``)."""
    ds = slide.level_downsamples[level]
    width, _ = slide.level_dimensions[level]
    region = slide.read_region((0, int(y0 * ds)), level, (width, y1 - y0)).convert("RGB")
    return np.array(region)   # own, writable copy (torch.from_numpy needs one)


def scan_slide(slide, level: int, mask, stride=None, patch_size: int = 224, device="cuda",
               max_slab_bytes: int = 2 << 30, want_u8: bool = False, layout=None, on_slab=None):
    """Run the GPU tile scan over a whole slide level in row slabs that fit ``max_slab_bytes``.

    ``mask``: the rasterised lesion mask (uint8 ``[H, W]`` numpy array, uploaded slab by slab), a
    ``lesion_mask.PolygonSet`` (the annotation polygons: every slab's mask rows are rasterised on the GPU, Pillow-exact,
    nothing is uploaded), or ``None`` (every patch "normal").

    Returns ``(coords int32 [N,2], labels uint8 [N])`` as numpy arrays in the reference's emission order,
    plus whatever ``on_slab(pb, slab_rgb, y0)`` collected (it is called once per slab with the
    ``PatchBatch`` in slab-relative coordinates)."""
    width, height = slide.level_dimensions[level]
    P, S = patch_and_stride(level, stride, patch_size)
    nx, ny = grid_shape(width, height, S)
    rows_budget = max(P + S, int(max_slab_bytes // (width * 3)))
    rows_per_slab = max(1, (rows_budget - (P - S)) // S)          # grid rows per slab
    all_coords, all_labels = [], []
    for i0 in range(0, ny, rows_per_slab):
        i1 = min(ny, i0 + rows_per_slab)
        y0, y1 = i0 * S, min(height, (i1 - 1) * S + P)
        rgb = read_level_rows(slide, level, y0, y1)
        # device slab with 16-byte row pitch (streaming-pass precondition), filled by one 2-D copy per buffer
        img = alloc_level_image(y1 - y0, width, device)
        upload_level_rows(img, np.ascontiguousarray(rgb))
        m = None
        if mask is not None and not isinstance(mask, np.ndarray):
            from .lesion_mask import rasterize_polygons
            m = rasterize_polygons(mask, width, height, device, y_begin=y0, n_rows=y1 - y0)
        elif mask is not None:
            m = alloc_level_image(y1 - y0, width, device, channels=1)
            upload_level_rows(m, np.ascontiguousarray(mask[y0:y1]))
        # a slab that ends above the image bottom has complete data for its grid rows, so tiling it as its own
        # image (rows [0, i1-i0) of the slab) gives exactly the patches of grid rows [i0, i1)
        pb = extract_patches_tensor(img, m, level, stride=stride, patch_size=patch_size, row_range=(0, i1 - i0),
                                    layout=layout, want_u8=want_u8)
        c = pb.coords.cpu().numpy().copy()
        if on_slab is not None:
            on_slab(pb, rgb, y0)
        c[:, 1] += y0
        all_coords.append(c)
        all_labels.append(pb.labels.cpu().numpy())
    coords = np.concatenate(all_coords) if all_coords else np.zeros((0, 2), np.int32)
    labels = np.concatenate(all_labels) if all_labels else np.zeros((0,), np.uint8)
    order = np.lexsort((coords[:, 1], coords[:, 0]))                # x outer, y inner (src/main.py:682-683)
    return coords[order], labels[order]


def _save_patches(rgb_level_rows, y0, coords, labels, P, width, height, prefix, patch_save_dir):
    """PNG compatibility mode: full-size P x P patches, white padded, ``{prefix}_x{x}_y{y}_{label}.png``
    (reference ``src/main.py:699-703, 722-726``)."""
    n = 0
    for (x, y), lab in zip(coords, labels):
        ys = y - y0
        pw, ph = min(P, width - x), min(P, height - y)
        region = Image.fromarray(np.ascontiguousarray(rgb_level_rows[ys:ys + ph, x:x + pw]), "RGB")
        if pw < P or ph < P:
            padded = Image.new("RGB", (P, P), (255, 255, 255))
            padded.paste(region, (0, 0))
            region = padded
        name = f"{prefix}_x{x}_y{y}_{'tumor' if lab else 'normal'}.png"
        path = os.path.join(patch_save_dir, name)
        if not os.path.exists(path):
            region.save(path)
        n += 1
    return n


def _extract_one(file, wsi_dir, level_dir, annot_dir_train, annot_dir_test, level, stride, patch_size_arg, pad,
                 slide_opener, device, max_slab_bytes, skip_needs_both_labels=False):
    prefix = file.replace(".tif", "")
    patch_save_dir = os.path.join(level_dir, prefix)
    existing = os.listdir(patch_save_dir) if os.path.exists(patch_save_dir) else []
    done = len(existing) > 0
    if skip_needs_both_labels:
        # extract_patches_per_slide only skips a slide that already has BOTH a *_normal.png and a *_tumor.png
        # (reference src/main.py:286-292): a partially extracted slide is re-run (existing files are kept, 725-726)
        done = done and any(f.endswith("_normal.png") for f in existing) and any(f.endswith("_tumor.png") for f in existing)
    if done:
        print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Patches for {file} already extracted, skipping.")
        return None
    os.makedirs(patch_save_dir, exist_ok=True)
    wsi_path = os.path.join(wsi_dir, file)
    xml_name = file.replace(".tif", ".xml")
    xml_path = os.path.join(annot_dir_test if file.startswith("test_") else annot_dir_train, xml_name)
    try:
        slide = slide_opener(wsi_path)
    except Exception as e:
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Could not open {wsi_path}: {e}")
        return None
    P, S = patch_and_stride(level, stride, patch_size_arg)
    width, height = slide.level_dimensions[level]
    if pad:
        pad_w = (P - width % P) % P
        pad_h = (P - height % P) % P
    else:
        pad_w = pad_h = 0
    mask = None
    if os.path.exists(xml_path):
        try:
            # same XML walk and vertex arithmetic as parse_xml_mask; the polygons are rasterised on the GPU per slab
            # (bit-exact against the Pillow call the reference makes), so no level-sized host mask is ever built
            from .lesion_mask import PolygonSet, annotation_polygons
            polys = annotation_polygons(xml_path, (width, height), slide)
            mask = PolygonSet(polys, device) if polys is not None else None
        except Exception as e:
            print(f"{bcolors.WARNING}[WARNING]{bcolors.ENDC} Failed to parse XML for {file}: {e}")
    else:
        print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} No annotation found for {file} in {xml_path}, treating as normal.")
    print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Processing {file} at level {level} (size: {width}x{height}, "
          f"padded: {width + pad_w}x{height + pad_h})")
    count = [0]

    def on_slab(pb, rgb, y0):
        c = pb.coords.cpu().numpy().copy()
        c[:, 1] += y0
        count[0] += _save_patches(rgb, y0, c, pb.labels.cpu().numpy(), P, width, height, prefix, patch_save_dir)

    scan_slide(slide, level, mask, stride=stride, patch_size=patch_size_arg, device=device,
               max_slab_bytes=max_slab_bytes, on_slab=on_slab)
    print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Patch extraction complete for {file} at level {level}. "
          f"Total patches: {count[0]}")
    return count[0]


def extract_patches(patch_size=224, level=3, stride=None, pad=True, only_tumor=False, test=False, *,
                    slide_opener=None, device="cuda", max_slab_bytes: int = 2 << 30):
    """Reference ``extract_patches`` (``src/main.py:609-732``): every ``*.tif`` under
    ``./data/camelyon16/train/img`` -> PNG patches under ``./data/camelyon16/patches/level_{level}/{slide}/``.

    ``only_tumor`` and ``test`` are accepted and ignored, as in the reference.  Note the reference's stride
    rule: ``stride = stride or patch_size`` with ``patch_size`` still the ARGUMENT (default 224), so the
    default grid is 224 px at every level."""
    print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Extracting patches at level {level}...")
    cwd = os.getcwd()
    wsi_dir = os.path.join(cwd, "data", "camelyon16", "train", "img")
    annot_dir_train = os.path.join(cwd, "data", "camelyon16", "train", "mask", "annotations")
    annot_dir_test = os.path.join(cwd, "data", "camelyon16", "test", "mask", "annotations")
    level_dir = os.path.join(cwd, "data", "camelyon16", "patches", f"level_{level}")
    os.makedirs(level_dir, exist_ok=True)
    opener = slide_opener or _default_opener
    for file in os.listdir(wsi_dir):
        if not file.endswith(".tif"):
            continue
        _extract_one(file, wsi_dir, level_dir, annot_dir_train, annot_dir_test, level, stride, patch_size, pad, opener,
                     device, max_slab_bytes)


def extract_patches_per_slide(slide_path="tumor_109", patch_size=224, level=3, stride=None, pad=True, only_tumor=False, *,
                              slide_opener=None, device="cuda", max_slab_bytes: int = 2 << 30):
    """Reference ``extract_patches_per_slide`` (``src/main.py:252-370``): same as ``extract_patches`` for one slide file."""
    cwd = os.getcwd()
    wsi_dir, file = os.path.split(slide_path)
    if not wsi_dir:
        wsi_dir = os.path.join(cwd, "data", "camelyon16", "train", "img")
    if not file.endswith(".tif"):
        file += ".tif"
    annot_dir_train = os.path.join(cwd, "data", "camelyon16", "train", "mask", "annotations")
    annot_dir_test = os.path.join(cwd, "data", "camelyon16", "test", "mask", "annotations")
    level_dir = os.path.join(cwd, "data", "camelyon16", "patches", f"level_{level}")
    os.makedirs(level_dir, exist_ok=True)
    return _extract_one(file, wsi_dir, level_dir, annot_dir_train, annot_dir_test, level, stride, patch_size, pad,
                        slide_opener or _default_opener, device, max_slab_bytes, skip_needs_both_labels=True)
