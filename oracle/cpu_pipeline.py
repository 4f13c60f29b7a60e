"""CPU port of the reference's hot path, structured like the reference -- TEST/BENCH INFRASTRUCTURE.

``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` time this on the GPU box's host
cores (kind = "port": the reference itself is Python that needs ``/root/reference``, which is not
shipped to the GPU box, so its algorithm is restated here with the SAME libraries it uses: numpy for
the tissue mean, Pillow for paste/crop/resize, torchvision transforms, torch fp32 for ResNet18).

Stage 1 follows ``extract_patches`` (reference ``src/main.py:682-727``): a single-threaded Python loop
over the x-outer/y-inner grid, ``read_region`` -> white-padded PIL patch -> mask crop -> ``np.mean``.
The PNG write/read between the stages (``src/main.py:722-726`` / ``patch_dataset.py:73``) is skipped
unless ``png_roundtrip=True`` (it only adds time to the reference's side).
Stage 2 follows ``extract_features`` (``src/main.py:812-818, 861-871``): ``Resize((224,224))`` ->
``ToTensor`` -> ``Normalize`` per patch, batched fp32 forward with all host threads.
"""
from __future__ import annotations

import io
import time

import numpy as np


def stage1_reference_loop(slide, level: int, stride=None, mask: np.ndarray | None = None, png_roundtrip: bool = False,
                          max_candidates: int | None = None):
    """Returns (patches: list[PIL.Image], coords, labels, n_candidates)."""
    from PIL import Image

    patch_size = 224
    stride = stride or patch_size
    patch_size = {0: 1792, 1: 896, 2: 448, 3: 224}.get(level, 224)
    downsample = slide.level_downsamples[level]
    width, height = slide.level_dimensions[level]
    mask_img = Image.fromarray(mask, "L") if mask is not None else None
    patches, coords, labels = [], [], []
    n_cand = 0
    for x in range(0, width, stride):
        for y in range(0, height, stride):
            if max_candidates is not None and n_cand >= max_candidates:
                break
            n_cand += 1
            patch_w = min(patch_size, width - x)
            patch_h = min(patch_size, height - y)
            region = slide.read_region((int(x * downsample), int(y * downsample)), level, (patch_w, patch_h)).convert("RGB")
            if patch_w < patch_size or patch_h < patch_size:
                padded = Image.new("RGB", (patch_size, patch_size), (255, 255, 255))
                padded.paste(region, (0, 0))
                region = padded
            if mask_img is not None:
                mask_patch = mask_img.crop((x, y, x + patch_size, y + patch_size))
                label = 1 if np.any(np.array(mask_patch) > 0) else 0
            else:
                label = 0
            if np.mean(np.array(region)) > 240:
                continue
            if png_roundtrip:
                buf = io.BytesIO()
                region.save(buf, format="PNG")
                buf.seek(0)
                region = Image.open(buf).convert("RGB")
            patches.append(region)
            coords.append((x, y))
            labels.append(label)
    return patches, np.asarray(coords, np.int32).reshape(-1, 2), np.asarray(labels, np.uint8), n_cand


def stage2_reference_loop(patches, net, batch: int = 64):
    """``Resize -> ToTensor -> Normalize`` per patch + batched fp32 trunk forward; returns f32 [N,512]."""
    import torch
    from torchvision import transforms

    tf = transforms.Compose([
        transforms.Resize((224, 224)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    outs = []
    with torch.no_grad():
        for i in range(0, len(patches), batch):
            x = torch.stack([tf(p) for p in patches[i:i + batch]])
            x = net.maxpool(net.relu(net.bn1(net.conv1(x))))
            x = net.layer4(net.layer3(net.layer2(net.layer1(x))))
            outs.append(torch.flatten(net.avgpool(x), 1))
    return torch.cat(outs).numpy() if outs else np.zeros((0, 512), np.float32)


def time_reference_path(slide, level: int, stride, mask, net, max_candidates: int, batch: int = 64,
                        png_roundtrip: bool = False):
    """Run both stages on a bounded sample; returns dict with seconds and counts."""
    import torch

    t0 = time.perf_counter()
    patches, coords, labels, n_cand = stage1_reference_loop(slide, level, stride, mask, png_roundtrip, max_candidates)
    t1 = time.perf_counter()
    feats = stage2_reference_loop(patches, net, batch)
    t2 = time.perf_counter()
    return dict(stage1_s=t1 - t0, stage2_s=t2 - t1, total_s=t2 - t0, candidates=n_cand, survivors=len(patches),
                threads=torch.get_num_threads(), features=feats, coords=coords, labels=labels)
