"""CPU study: how often does a bf16-activation ResNet18 (fp32 accumulate, the arithmetic of the CUDA path) flip the
classifier argmax against the fp32 oracle, and with which margins?  Dev-container tool (no GPU needed).

    python profiles/_argmax_margin_study.py [size] [seed]

Patches are the level-3 survivors (P = S = 224) of a synthetic slide, i.e. the population tests/test_argmax_gpu.py uses.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hipac_oracle as orc  # noqa: E402
from ss25_hierarchical_multiscale_image_classification_b200.synthetic import make_level  # noqa: E402


def bf(x):
    return x.bfloat16().float()


def folded(net, conv, bn):
    sd = net.state_dict()
    scale = sd[f"{bn}.weight"] / torch.sqrt(sd[f"{bn}.running_var"] + 1e-5)
    w = bf(sd[f"{conv}.weight"] * scale[:, None, None, None])
    b = sd[f"{bn}.bias"] - sd[f"{bn}.running_mean"] * scale
    return w, b


def emulated_forward(net, x, last_fp32=False):
    """x: fp32 NCHW normalised, already bf16-rounded.  Mirrors the kernels' rounding points."""
    F = torch.nn.functional
    w, b = folded(net, "conv1", "bn1")
    y = bf(torch.relu(F.conv2d(x, w, b, stride=2, padding=3)))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, layer in enumerate(["layer1", "layer2", "layer3", "layer4"]):
        for blk in range(2):
            p = f"{layer}.{blk}"
            stride = 2 if (li > 0 and blk == 0) else 1
            w1, b1 = folded(net, f"{p}.conv1", f"{p}.bn1")
            w2, b2 = folded(net, f"{p}.conv2", f"{p}.bn2")
            t = bf(torch.relu(F.conv2d(y, w1, b1, stride=stride, padding=1)))
            o = F.conv2d(t, w2, b2, stride=1, padding=1)
            if stride == 2:
                wd, bd = folded(net, f"{p}.downsample.0", f"{p}.downsample.1")
                o = o + F.conv2d(y, wd, bd, stride=2)
            else:
                o = o + y
            o = torch.relu(o)
            last = layer == "layer4" and blk == 1
            y = o if (last and last_fp32) else bf(o)
    f = y.mean((2, 3))
    return f, net.fc(f)


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 8960
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4321
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    img = make_level(seed, 3, size, size)
    want = orc.extract_patches_oracle(img, None, 3)
    imgs = want["images"]
    print(f"{len(imgs)} survivors of {want['candidates']} candidates ({time.time() - t0:.0f} s)", flush=True)
    net = orc.make_resnet18(seed=0, classifier=True)
    ref_f, ref_l = orc.resnet18_features_fp32(net, imgs)
    lut = torch.from_numpy(orc.normalize_lut())
    for last_fp32 in (False, True):
        fs, ls = [], []
        with torch.no_grad():
            for i in range(0, len(imgs), 64):
                u8 = torch.from_numpy(np.ascontiguousarray(imgs[i:i + 64])).long()
                x = bf(torch.stack([lut[:, c][u8[..., c]] for c in range(3)], dim=1))
                f, l = emulated_forward(net, x, last_fp32)
                fs.append(f), ls.append(l)
        f, l = torch.cat(fs).numpy(), torch.cat(ls).numpy()
        cos = (f * ref_f).sum(1) / (np.linalg.norm(f, axis=1) * np.linalg.norm(ref_f, axis=1))
        maxrel = np.abs(f - ref_f).max(1) / np.abs(ref_f).max(1)
        margin = np.abs(ref_l[:, 0] - ref_l[:, 1])
        flips = np.nonzero(l.argmax(1) != ref_l.argmax(1))[0]
        dl = np.abs((l[:, 1] - l[:, 0]) - (ref_l[:, 1] - ref_l[:, 0]))
        # stress head: same weight direction, bias moved to the population median so that both classes occur
        w = (net.fc.weight[1] - net.fc.weight[0]).detach().numpy()
        d_ref, d_got = ref_f @ w, f @ w
        for scale_name, med in (("median", np.median(d_ref)), ("q10", np.quantile(d_ref, 0.1))):
            fl = np.nonzero((d_got > med) != (d_ref > med))[0]
            print(f"  centred head ({scale_name}): spread std {d_ref.std():.4f}, proj err max {np.abs(d_got - d_ref).max():.2e}, "
                  f"flips {len(fl)}/{len(imgs)} = {100 * len(fl) / len(imgs):.3f} % at margins {np.abs(d_ref[fl] - med).round(5).tolist()[:12]}")
        print(f"last_fp32={last_fp32}: min cos {cos.min():.6f} max rel {maxrel.max():.2e}; logit-diff error max {dl.max():.2e} "
              f"median {np.median(dl):.2e}; margin quantiles {np.quantile(margin, [0, .001, .01, .1, .5]).round(5).tolist()}; "
              f"flips {len(flips)}/{len(imgs)} margins {margin[flips].round(6).tolist()} class balance {ref_l.argmax(1).mean():.3f}", flush=True)


if __name__ == "__main__":
    main()
