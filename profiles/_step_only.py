"""Resident bench steps only (no e2e / CPU legs): the command ncu wraps for the per-step captures under profiles/."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import seeded_resnet18
from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
import __graft_entry__ as ge
ge.build()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
img_h, msk_h, (i0, i1, y0, H) = bench.build_slab(1, 0)
img, msk = img_h.cuda(), msk_h.cuda()
packed = features.pack_resnet18(seeded_resnet18(seed=0, classifier=True).state_dict(), "cuda")
for _ in range(steps):
    r = pipeline.process_level(img, msk, bench.LEVEL, packed)
torch.cuda.synchronize()
print(len(r), r.candidates)
