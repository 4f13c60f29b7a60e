"""Per-patch tumour-probability heatmap and the CAMELYON16 ``prob,x,y`` CSV.

The reference expects such CSVs (``src/main.py:1173-1181`` feeds ``utils/evaluation_FROC.py:67-88``:
one ``probability,x,y`` row per detection, x/y in LEVEL-0 pixels) but contains no code that produces
them, and BASELINE.json's configs[3] asks for a per-patch heatmap.  Both are thin views of the
classifier logits the hot path already returns; nothing here touches pixels.
"""
from __future__ import annotations

import numpy as np
import torch


def tumor_probability(logits: torch.Tensor) -> torch.Tensor:
    """softmax(logits)[:, 1] (class 1 = tumour, reference ``label_map`` ``src/datasets/patch_dataset.py:14``)."""
    return torch.softmax(logits.float(), dim=1)[:, 1]


def heatmap(coords: torch.Tensor, logits: torch.Tensor, width: int, height: int, stride: int, fill: float = 0.0) -> torch.Tensor:
    """float32 ``[ny, nx]`` grid over the candidate lattice: p(tumour) at surviving patches, ``fill`` elsewhere."""
    nx, ny = (width + stride - 1) // stride, (height + stride - 1) // stride
    out = torch.full((ny, nx), fill, dtype=torch.float32, device=logits.device)
    if coords.shape[0]:
        ix = (coords[:, 0] // stride).long()
        iy = (coords[:, 1] // stride).long()
        out[iy, ix] = tumor_probability(logits)
    return out


def write_froc_csv(path: str, coords, logits, level: int, patch: int, threshold: float = 0.0) -> int:
    """``prob,x,y`` rows (patch CENTRE in level-0 pixels) for patches with p(tumour) >= threshold; returns the row count."""
    prob = tumor_probability(torch.as_tensor(logits)).cpu().numpy()
    c = torch.as_tensor(coords).cpu().numpy().astype(np.int64)
    ds = 2 ** level
    n = 0
    with open(path, "w") as f:
        for p, (x, y) in zip(prob, c):
            if p >= threshold:
                f.write(f"{p:.6f},{(x + patch // 2) * ds},{(y + patch // 2) * ds}\n")
                n += 1
    return n
