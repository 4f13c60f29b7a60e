"""Where does the step time outside the kernels go?  CUDA-event time of each stage call vs the sum of its kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ss25_hierarchical_multiscale_image_classification_b200 import _lib, features
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import seeded_resnet18
import __graft_entry__ as ge
ge.build()
img_h, msk_h, _ = bench.build_slab(1, 0)
img, msk = img_h.cuda(), msk_h.cuda()
packed = features.pack_resnet18(seeded_resnet18(0, True).state_dict(), "cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(8):
    prof = it >= 5
    _lib.profile(prof)
    ev[0].record()
    pb = extract_patches_tensor(img, msk, 0, layout="s2d16")
    ev[1].record()
    f, l = features.classify_tensor(pb.batch, packed)
    ev[2].record()
    torch.cuda.synchronize()
    rep = _lib.profile_report() if prof else {}
    s1 = sum(v["ms"] for k, v in rep.items() if not k.startswith(("conv", "avgpool")))
    s2 = sum(v["ms"] for k, v in rep.items() if k.startswith(("conv", "avgpool")))
    print(f"iter {it}: stage1 call {ev[0].elapsed_time(ev[1]):.3f} ms (kernels {s1:.3f})  stage2 call {ev[1].elapsed_time(ev[2]):.3f} ms (kernels {s2:.3f})  total {ev[0].elapsed_time(ev[2]):.3f}")
