"""Drop-in ``extract_features`` (reference ``src/main.py:805-894``) and its artefact files.

Reads the PNG patch tree ``./data/camelyon16/patches/level_{level}`` that ``extract_patches`` wrote,
and writes the three artefacts every downstream consumer of the reference expects --
``patch_features_{level}.npy`` (N, 512) float32, ``patch_labels_{level}.npy`` (N,) int64 and
``patch_paths_{level}.txt`` -- in the current directory.

What differs from the reference, on purpose:
  * Resize((224,224)) / ToTensor / Normalize and the ResNet18 forward run on the GPU: decoded patches
    of one size are stacked into a tall "level image" and pushed through ``hipac_tile_scan`` with
    stride = patch size and the tissue rejection disabled, which yields the Pillow-exact 224x224
    normalised batch, then ``hipac_resnet18_forward``.
  * weights: the reference builds an ImageNet-pretrained ``ResNet18Classifier`` (needs the network) and
    then copies nothing from it because the key namespaces are disjoint (``src/main.py:852-859``), so
    its features come from ``ResNet18FeatureExtractor()``'s own weights.  This function uses exactly
    those: ``ResNet18FeatureExtractor()`` with its DEFAULT weight path (``src/models/resnet18_patch_classifier.pth``
    if present, else seeded-random init; ``seed=`` argument, the reference's init is unseeded).  ``model_path`` is
    accepted for signature compatibility; in the reference it only selects the classifier checkpoint whose weights
    never reach the feature model.
  * patches must be squares of side 224, 448, 896 or 1792 (what ``extract_patches`` writes); the reference's
    ``Resize`` would accept any PNG size.
  * row order follows the sorted path list instead of the reference's unseeded shuffle; the paths file
    identifies rows either way.

``extract_features_from_slides`` is the PNG-free equivalent: slide -> tile scan -> ResNet18 in HBM.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from PIL import Image

from . import features as _features
from .models.resnet import ResNet18FeatureExtractor
from .preprocessing.extract import _default_opener, bcolors, parse_xml_mask, read_level_rows
from .preprocessing.tensor_api import extract_patches_tensor, grid_shape, patch_and_stride

BATCH_SIZE = 512          # reference src/main.py:46
LEVEL_OF_SIZE = {1792: 0, 896: 1, 448: 2, 224: 3}


def _label_of(name: str):
    if "_tumor" in name:
        return 1
    if "_normal" in name:
        return 0
    return None


def _list_patches(patch_dir):
    paths, labels = [], []
    for root, _, files in sorted(os.walk(patch_dir)):
        for f in sorted(files):
            if not f.endswith(".png"):
                continue
            lab = _label_of(f)
            if lab is None:
                print(f"[WARNING] Could not determine label from filename: {f}")
                continue
            paths.append(os.path.join(root, f))
            labels.append(lab)
    return paths, labels


def features_of_patch_arrays(patches: list, packed: _features.PackedResNet18, device, chunk: int = 4096) -> torch.Tensor:
    """float32 ``[N,512]`` for a list of square RGB uint8 arrays whose side is 224/448/896/1792."""
    out = torch.empty((len(patches), 512), dtype=torch.float32, device=device)
    by_size = {}
    for i, a in enumerate(patches):
        if a.shape[0] != a.shape[1] or a.shape[0] not in LEVEL_OF_SIZE:
            raise ValueError(f"patch {i} has shape {a.shape}; expected a square of side 224, 448, 896 or 1792")
        by_size.setdefault(a.shape[0], []).append(i)
    for side, idx in by_size.items():
        stack = torch.from_numpy(np.concatenate([patches[i] for i in idx], axis=0)).to(device)   # [n*side, side, 3]
        pb = extract_patches_tensor(stack, None, LEVEL_OF_SIZE[side], stride=side, layout="s2d16", keep_all=True)
        assert len(pb) == len(idx)
        out[torch.as_tensor(idx, device=device)] = _features.extract_features_tensor(pb.batch, packed, chunk)
    return out


def extract_features(level=3, model_path="resnet18_patch_classifier.pth", *, device="cuda", seed=None,
                     batch_size=BATCH_SIZE):
    """Reference ``extract_features(level, model_path)``: PNG patches -> the three artefact files."""
    patch_dir = os.path.join(os.getcwd(), "data", "camelyon16", "patches", f"level_{level}")
    if not os.path.exists(patch_dir) or not os.listdir(patch_dir):
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Patch directory '{patch_dir}' does not exist or is empty. "
              "Please run patch extraction first.")
        return
    paths, labels = _list_patches(patch_dir)
    print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Extracting features from patches at level {level} with patch directory: "
          f"{patch_dir}, which exists: {os.path.exists(patch_dir)}")
    if seed is not None:
        torch.manual_seed(seed)
    model = ResNet18FeatureExtractor()   # default weight path, as the reference does (src/main.py:839); model_path only names the classifier checkpoint there
    packed = _features.pack_resnet18(model._tv_state(), device)
    feats = []
    for i in range(0, len(paths), batch_size):
        arrays = [np.asarray(Image.open(p).convert("RGB")) for p in paths[i:i + batch_size]]
        feats.append(features_of_patch_arrays(arrays, packed, device).cpu())
    if not feats:
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} No features were extracted. Check your patch directory and dataset. "
              "It might be that PatchDataset found no images, or data loader was empty.")
        return
    write_artefacts(level, torch.cat(feats).numpy(), np.asarray(labels), paths)


def write_artefacts(level, features: np.ndarray, labels: np.ndarray, paths: list, out_dir: str = "."):
    """``patch_features_{L}.npy`` / ``patch_labels_{L}.npy`` / ``patch_paths_{L}.txt`` (reference ``src/main.py:882-893``)."""
    f_path = os.path.join(out_dir, f"patch_features_{level}.npy")
    l_path = os.path.join(out_dir, f"patch_labels_{level}.npy")
    p_path = os.path.join(out_dir, f"patch_paths_{level}.txt")
    np.save(f_path, np.ascontiguousarray(features, dtype=np.float32))
    np.save(l_path, np.asarray(labels, dtype=np.int64))
    with open(p_path, "w") as f:
        for p in paths:
            f.write(f"{p}\n")
    print(f"{bcolors.INFO}[INFO]{bcolors.ENDC} Features saved to {f_path}, labels to {l_path}, paths to {p_path}")


def extract_features_from_slides(level=3, model_path="resnet18_patch_classifier.pth", stride=None, *, slide_opener=None,
                                 device="cuda", seed=None, max_slab_bytes: int = 2 << 30, out_dir: str = "."):
    """PNG-free path: every ``*.tif`` under ``./data/camelyon16/train/img`` -> tile scan -> ResNet18 -> artefacts.

    The paths file lists the names the PNG mode would have written
    (``.../level_{L}/{slide}/{slide}_x{x}_y{y}_{label}.png``), so the artefacts are interchangeable."""
    from .pipeline import process_level
    cwd = os.getcwd()
    wsi_dir = os.path.join(cwd, "data", "camelyon16", "train", "img")
    level_dir = os.path.join(cwd, "data", "camelyon16", "patches", f"level_{level}")
    if seed is not None:
        torch.manual_seed(seed)
    model = ResNet18FeatureExtractor()   # default weight path, as the reference does (src/main.py:839); model_path only names the classifier checkpoint there
    packed = _features.pack_resnet18(model._tv_state(), device)
    opener = slide_opener or _default_opener
    all_f, all_l, all_p = [], [], []
    for file in sorted(os.listdir(wsi_dir)):
        if not file.endswith(".tif"):
            continue
        prefix = file[:-4]
        try:
            slide = opener(os.path.join(wsi_dir, file))
        except Exception as e:
            print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Could not open {file}: {e}")
            continue
        width, height = slide.level_dimensions[level]
        sub = "test" if file.startswith("test_") else "train"
        xml_path = os.path.join(cwd, "data", "camelyon16", sub, "mask", "annotations", prefix + ".xml")
        mask = None
        if os.path.exists(xml_path):
            try:
                m = parse_xml_mask(xml_path, (width, height), slide)
                mask = np.asarray(m) if m is not None else None
            except Exception as e:
                print(f"{bcolors.WARNING}[WARNING]{bcolors.ENDC} Failed to parse XML for {file}: {e}")
        P, S = patch_and_stride(level, stride)
        nx, ny = grid_shape(width, height, S)
        rows_budget = max(P + S, int(max_slab_bytes // (width * 3)))
        rows_per_slab = max(1, (rows_budget - (P - S)) // S)
        coords, labels, feats = [], [], []
        for i0 in range(0, ny, rows_per_slab):
            i1 = min(ny, i0 + rows_per_slab)
            y0, y1 = i0 * S, min(height, (i1 - 1) * S + P)
            img = torch.from_numpy(np.ascontiguousarray(read_level_rows(slide, level, y0, y1))).to(device)
            m = torch.from_numpy(np.array(mask[y0:y1])).to(device) if mask is not None else None
            r = process_level(img, m, level, packed, stride=stride, row_range=(0, i1 - i0))
            c = r.coords.cpu().numpy().copy()
            c[:, 1] += y0
            coords.append(c), labels.append(r.labels.cpu().numpy()), feats.append(r.features.cpu().numpy())
        coords, labels, feats = np.concatenate(coords), np.concatenate(labels), np.concatenate(feats)
        order = np.lexsort((coords[:, 1], coords[:, 0]))
        for k in order:
            x, y = coords[k]
            all_p.append(os.path.join(level_dir, prefix, f"{prefix}_x{x}_y{y}_{'tumor' if labels[k] else 'normal'}.png"))
        all_f.append(feats[order]), all_l.append(labels[order].astype(np.int64))
    if not all_f:
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} No features were extracted.")
        return
    write_artefacts(level, np.concatenate(all_f), np.concatenate(all_l), all_p, out_dir)
