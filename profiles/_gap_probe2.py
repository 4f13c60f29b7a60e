"""Steady-state time of stage 1 alone, ResNet18 alone and the full step (CUDA events over 10 back-to-back iterations),
next to the per-kernel profiler sum: separates launch gaps from clock/power effects."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ss25_hierarchical_multiscale_image_classification_b200 import _lib, features, pipeline
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import seeded_resnet18
import __graft_entry__ as ge
ge.build()
img_h, msk_h, _ = bench.build_slab(1, 0)
img, msk = img_h.cuda(), msk_h.cuda()
packed = features.pack_resnet18(seeded_resnet18(0, True).state_dict(), "cuda")
pb = extract_patches_tensor(img, msk, 0, layout="s2d16")
batch = pb.batch.clone()

def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t1 = timed(lambda: extract_patches_tensor(img, msk, 0, layout="s2d16"))
t2 = timed(lambda: features.classify_tensor(batch, packed, 8192))
t3 = timed(lambda: pipeline.process_level(img, msk, 0, packed))
_lib.profile(True)
for _ in range(2): pipeline.process_level(img, msk, 0, packed)
torch.cuda.synchronize()
rep = _lib.profile_report(); _lib.profile(False)
s1 = sum(v["ms"] for k, v in rep.items() if not k.startswith(("conv", "avgpool"))) / 2
s2 = sum(v["ms"] for k, v in rep.items() if k.startswith(("conv", "avgpool"))) / 2
print(f"stage1 alone {t1:.3f} ms (kernels {s1:.3f})  resnet alone {t2:.3f} ms (kernels {s2:.3f})  full step {t3:.3f} ms")
