import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import make_level, make_lesion_mask
H = W = 8192
img = torch.from_numpy(make_level(1234, 0, 16384, 16384, 4096, 4096 + H, 4096, 4096 + W)).cuda()
msk = torch.from_numpy(make_lesion_mask(1234, 0, 16384, 16384, 4096, 4096 + H)[:, 4096:4096 + W].copy()).cuda()
for _ in range(3):
    pb = extract_patches_tensor(img, msk, 0, layout="s2d16")
torch.cuda.synchronize()
print(len(pb), pb.candidates)
