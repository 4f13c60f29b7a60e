"""CPU oracle for the HiPAC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of what the reference computes between "level image in
memory" and "normalised 224x224 tensor + label per surviving patch", plus a
torch-fp32 ResNet18 for the floating-point stage.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package; the product package never does.

Parity pin: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c), so this oracle is pinned differentially: ``oracle/ref_harness.py`` runs
the reference's own, unmodified ``extract_patches`` / ``PatchDataset`` /
``transforms.Resize`` / ``ResNet18FeatureExtractor`` in the dev container on
synthetic slides and ``tests/golden/make_golden.py`` freezes their outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those fixtures (and against the installed Pillow directly).

Reference lines restated (all under ``/root/reference``):
  * grid / stride / patch size per level ....... src/main.py:609-615, 682-691
  * white padding of border patches ............ src/main.py:699-703
  * lesion label from the rasterised mask ...... src/main.py:668-677, 705-716
  * tissue test ``np.mean(patch) > 240`` ....... src/main.py:718-720
  * output naming ``{prefix}_x{x}_y{y}_{label}`` src/main.py:722
  * Resize((224,224)) -> ToTensor -> Normalize .. src/main.py:812-818
  * Pillow ``Image.resize(BILINEAR)`` (third-party, not vendored; Pillow
    ``src/libImaging/Resample.c``: ``precompute_coeffs``,
    ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` then
    ``ImagingResampleVertical_8bpc``; container has Pillow 12.2.0)
  * label map of file names .................... src/datasets/patch_dataset.py:27-33
  * ResNet18 trunk / classifier ................ src/models/resnet.py:22-77
"""
from __future__ import annotations

import math

import numpy as np

PATCH_SIZES = {0: 1792, 1: 896, 2: 448, 3: 224}   # src/main.py:614
OUT = 224                                          # src/main.py:814
TISSUE_THRESHOLD = 240                             # src/main.py:719
PRECISION_BITS = 32 - 8 - 2                        # Pillow Resample.c
IMAGENET_MEAN = (0.485, 0.456, 0.406)              # src/main.py:816
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------
# stage 1a: candidate grid (src/main.py:609-615, 682-691)
# --------------------------------------------------------------------------
def patch_and_stride(level: int, stride=None, patch_size_arg: int = 224):
    """``stride = stride or patch_size`` is evaluated BEFORE the level map
    overwrites ``patch_size`` (src/main.py:611 vs 614-615)."""
    s = stride or patch_size_arg
    p = PATCH_SIZES.get(level, 224)
    return p, s


def candidate_grid(width: int, height: int, level: int, stride=None, patch_size_arg: int = 224):
    """All (x, y) the reference visits, in its order: x outer, y inner
    (src/main.py:682-686).  The padded loop bounds are neutralised by the
    ``x >= width or y >= height`` guard."""
    p, s = patch_and_stride(level, stride, patch_size_arg)
    xs = np.arange(0, width, s, dtype=np.int64)
    ys = np.arange(0, height, s, dtype=np.int64)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    return p, s, np.stack([gx.ravel(), gy.ravel()], axis=1)


# --------------------------------------------------------------------------
# stage 1b: padded patch, tissue test, lesion label
# --------------------------------------------------------------------------
def padded_patch(level_img: np.ndarray, x: int, y: int, p: int) -> np.ndarray:
    """P x P x 3 patch, partial regions pasted at (0,0) on white (src/main.py:688-703)."""
    h, w = level_img.shape[:2]
    pw, ph = min(p, w - x), min(p, h - y)
    out = np.full((p, p, 3), 255, dtype=np.uint8)
    out[:ph, :pw] = level_img[y:y + ph, x:x + pw]
    return out


def tissue_sum_limit(p: int) -> int:
    """keep <=> sum(u8) <= 240*3*P^2 (integer form of ``np.mean(patch) > 240 -> reject``)."""
    return TISSUE_THRESHOLD * 3 * p * p


def tissue_keep(patch: np.ndarray) -> bool:
    return int(patch.sum(dtype=np.uint64)) <= tissue_sum_limit(patch.shape[0])


def lesion_label(mask: np.ndarray | None, x: int, y: int, p: int) -> int:
    """1 if any mask pixel > 0 inside [x,x+P) x [y,y+P), out of bounds = 0
    (PIL ``crop`` zero-fills; src/main.py:705-716).  No mask -> 0 ("normal")."""
    if mask is None:
        return 0
    return int(np.any(mask[y:y + p, x:x + p] > 0))


def patch_name(prefix: str, x: int, y: int, label: int) -> str:
    return f"{prefix}_x{x}_y{y}_{'tumor' if label else 'normal'}.png"  # src/main.py:722


# --------------------------------------------------------------------------
# stage 1c: Pillow antialiased bilinear resize, 8 bits per channel
# --------------------------------------------------------------------------
def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the triangle filter.

    Returns ``(xmin[out], count[out], kk[out, ksize] int64)``."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int64)
    cnt = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)          # C (int) cast truncates toward zero
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = []
        ww = 0.0
        for x in range(n):
            a = (x + lo - center + 0.5) * ss
            a = -a if a < 0.0 else a
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx] = lo
        cnt[xx] = n
    return xmin, cnt, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    in_size = img.shape[axis]
    xmin, cnt, kk = pil_bilinear_coeffs(in_size, out_size)
    ksize = kk.shape[1]
    idx = np.minimum(xmin[:, None] + np.arange(ksize)[None, :], in_size - 1)   # taps past count have kk == 0
    src = np.moveaxis(img, axis, 0)                                            # [in, ...] uint8
    # 2^21 + 255 * sum(kk) < 2^31, so int32 accumulation is exact
    acc = np.full((out_size,) + src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int32)
    kk32 = kk.astype(np.int32)
    for t in range(ksize):
        if not kk32[:, t].any():
            continue
        acc += src[idx[:, t]].astype(np.int32) * kk32[:, t].reshape((-1,) + (1,) * (src.ndim - 1))
    out = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(patch: np.ndarray, out_size: int = OUT) -> np.ndarray:
    """``PIL.Image.resize((out,out), BILINEAR)`` on an HxWx3 uint8 array: horizontal
    pass to a uint8 intermediate, then vertical pass; identity when sizes match."""
    h, w = patch.shape[:2]
    if (h, w) == (out_size, out_size):
        return patch.copy()
    tmp = _resample_axis(patch, out_size, axis=1) if w != out_size else patch
    return _resample_axis(tmp, out_size, axis=0) if h != out_size else tmp


# --------------------------------------------------------------------------
# stage 1d: ToTensor + Normalize (src/main.py:815-816)
# --------------------------------------------------------------------------
def normalize_lut() -> np.ndarray:
    """float32 [256, 3]: ``(u8/255 - mean_c) / std_c`` exactly as torch computes it in fp32."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)
    std = np.asarray(IMAGENET_STD, dtype=np.float32)
    return ((v[:, None] - mean[None, :]) / std[None, :]).astype(np.float32)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 bit pattern (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    return (r >> np.uint32(16)).astype(np.uint16)


def normalize_u8(img224: np.ndarray) -> np.ndarray:
    """float32 HWC normalised image from a uint8 HWC image."""
    lut = normalize_lut()
    c = np.arange(3)
    return lut[img224, c[None, None, :]] if img224.ndim == 3 else lut[img224, c]


# --------------------------------------------------------------------------
# full stage 1
# --------------------------------------------------------------------------
def extract_patches_oracle(level_img: np.ndarray, mask: np.ndarray | None, level: int, stride=None,
                           patch_size_arg: int = 224, want_images: bool = True, row_range=None):
    """Survivors of one level image, in the reference's emission order.

    Returns dict(coords int32[N,2] (x,y), labels uint8[N], images uint8[N,224,224,3] | None,
    candidates int, patch, stride).  ``row_range=(i0,i1)`` restricts to grid rows
    ``y // stride in [i0,i1)`` (the multi-GPU shard unit)."""
    h, w = level_img.shape[:2]
    p, s, grid = candidate_grid(w, h, level, stride, patch_size_arg)
    coords, labels, images = [], [], []
    n_cand = 0
    for x, y in grid:
        x, y = int(x), int(y)
        if row_range is not None and not (row_range[0] <= y // s < row_range[1]):
            continue
        n_cand += 1
        patch = padded_patch(level_img, x, y, p)
        label = lesion_label(mask, x, y, p)
        if not tissue_keep(patch):
            continue
        coords.append((x, y))
        labels.append(label)
        if want_images:
            images.append(pil_resize_bilinear(patch))
    return dict(
        coords=np.asarray(coords, dtype=np.int32).reshape(-1, 2),
        labels=np.asarray(labels, dtype=np.uint8),
        images=(np.stack(images) if images else np.zeros((0, OUT, OUT, 3), np.uint8)) if want_images else None,
        candidates=n_cand, patch=p, stride=s)


# --------------------------------------------------------------------------
# stage 2: torch fp32 ResNet18 (src/models/resnet.py:22-77)
# --------------------------------------------------------------------------
def make_resnet18(seed: int = 0, classifier: bool = True):
    """Seeded random-init torchvision resnet18 (+ Linear(512,2) head), eval mode: the shared synthetic-weights
    recipe (``synthetic.seeded_resnet18``), so both sides of every parity test load the same state dict."""
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import seeded_resnet18
    return seeded_resnet18(seed, classifier)


def resnet18_features_fp32(net, images_u8: np.ndarray, batch: int = 64):
    """fp32 CPU forward of the trunk on uint8 NHWC patches: (features f32[N,512], logits f32[N,k])."""
    import torch

    lut = torch.from_numpy(normalize_lut())
    feats, logits = [], []
    with torch.no_grad():
        for i in range(0, len(images_u8), batch):
            u8 = torch.from_numpy(np.ascontiguousarray(images_u8[i:i + batch])).long()
            x = torch.stack([lut[:, c][u8[..., c]] for c in range(3)], dim=1)   # NCHW fp32
            x = net.maxpool(net.relu(net.bn1(net.conv1(x))))
            x = net.layer4(net.layer3(net.layer2(net.layer1(x))))
            f = torch.flatten(net.avgpool(x), 1)
            feats.append(f)
            logits.append(net.fc(f))
    if not feats:
        return np.zeros((0, 512), np.float32), np.zeros((0, 2), np.float32)
    return torch.cat(feats).numpy(), torch.cat(logits).numpy()
