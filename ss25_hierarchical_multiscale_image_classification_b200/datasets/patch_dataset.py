"""``PatchDataset`` with the reference's constructor and item contract (``src/datasets/patch_dataset.py:8-85``).

Kept for API compatibility at the boundary of the hot path: the fast path never materialises PNG files,
but ``extract_patches``' PNG mode and any reference-produced patch tree can be read back through this
class exactly as before -- recursive ``*.png`` discovery, label from the file name (``_tumor`` -> 1,
``_normal`` -> 0), optional class balancing / sub-sampling, shuffle, per-class transforms, items
``(image, label, path)``.  One addition: ``seed=`` makes the reference's unseeded ``random`` calls
reproducible; ``seed=None`` keeps its behaviour.
"""
from __future__ import annotations

import glob
import os
import random
from collections import Counter, defaultdict

from PIL import Image
from torch.utils.data import Dataset


class PatchDataset(Dataset):
    def __init__(self, root_dir, transform=None, tumor_transform=None, normal_transform=None, balanced=False,
                 max_samples=None, slide_names=None, seed=None):
        self.tumor_transform = tumor_transform if tumor_transform is not None else transform
        self.normal_transform = normal_transform if normal_transform is not None else transform
        self.transform = transform
        self.label_map = {"_normal": 0, "_tumor": 1}
        rng = random.Random(seed) if seed is not None else random
        by_class = defaultdict(list)
        for path in glob.glob(os.path.join(root_dir, "**", "*.png"), recursive=True):
            if slide_names is not None:
                slide_dir = os.path.relpath(path, root_dir).split(os.sep)[0]
                if slide_dir not in slide_names:
                    continue
            name = os.path.basename(path)
            if "_tumor" in name:
                by_class[1].append(path)
            elif "_normal" in name:
                by_class[0].append(path)
            else:
                print(f"[WARNING] Could not determine label from filename: {name}")
        paths, labels = [], []
        if balanced:
            floor = min(len(v) for v in by_class.values())
            for label, items in by_class.items():
                count = min(floor, max_samples) if max_samples else floor
                picked = rng.sample(items, min(count, len(items)))
                paths.extend(picked)
                labels.extend([label] * len(picked))
        else:
            for label, items in by_class.items():
                if max_samples:
                    items = rng.sample(items, min(len(items), max_samples))
                paths.extend(items)
                labels.extend([label] * len(items))
        order = list(range(len(paths)))
        rng.shuffle(order)
        self.image_paths = [paths[i] for i in order]
        self.labels = [labels[i] for i in order]
        counts = Counter(self.labels)
        print(f"[INFO] PatchDataset initialized: {len(self.labels)} total patches.")
        print(f"[INFO] Tumor patches: {counts.get(1, 0)} | Normal patches: {counts.get(0, 0)}")
        print(f"[INFO] Label distribution: {dict(counts)}")

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, idx):
        path, label = self.image_paths[idx], self.labels[idx]
        image = Image.open(path).convert("RGB")
        if label == 1 and self.tumor_transform:
            image = self.tumor_transform(image)
        elif label == 0 and self.normal_transform:
            image = self.normal_transform(image)
        elif self.transform:
            image = self.transform(image)
        return image, label, path

    def get_class_counts(self):
        return dict(Counter(self.labels))
