"""Small stage-1 + stage-2 workload for compute-sanitizer (memcheck / racecheck): every level through the TMA streaming
pass on pitched images of awkward sizes, the cp.async fallback, the direct path, level 3, and one device-count forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ss25_hierarchical_multiscale_image_classification_b200 import features
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_enqueue, extract_patches_tensor
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import seeded_resnet18

rng = np.random.default_rng(0)


def pitched(a):
    h, w = a.shape[:2]
    wp = (w + 15) // 16 * 16
    buf = torch.zeros((h, wp) + tuple(a.shape[2:]), dtype=torch.uint8, device="cuda")
    buf[:, :w] = torch.from_numpy(a).cuda()
    return buf[:, :w]


n = 0
for level, w, h in [(0, 2101, 1900), (1, 1203, 1000), (2, 905, 700), (3, 500, 300), (2, 449, 225)]:
    img = rng.integers(100, 256, size=(h, w, 3), dtype=np.uint8)
    msk = np.zeros((h, w), np.uint8)
    msk[h // 3: h // 3 + 9, w // 4: w // 2] = 3
    msk[h - 1, w - 1] = 1
    for mode, pit in [("fused", True), ("fused_legacy", True), ("auto", False), ("direct", False)]:
        a = extract_patches_tensor(pitched(img) if pit else torch.from_numpy(img).cuda(), pitched(msk) if pit else torch.from_numpy(msk).cuda(),
                                   level, layout="s2d16", want_u8=True, mode=mode)
        n += len(a)
packed = features.pack_resnet18(seeded_resnet18(0, True).state_dict(), "cuda")
img = rng.integers(0, 230, size=(700, 905, 3), dtype=np.uint8)
pend = extract_patches_enqueue(pitched(img), None, 2, layout="s2d16")
f, l = features.classify_tensor(pend.batch, packed, 8192, count=pend.count)
pb = pend.resolve()
torch.cuda.synchronize()
print("survivors", n, len(pb), float(f[:len(pb)].abs().sum()))
