"""Classifier argmax agreement at scale (north_star bar: >= 99.9 % vs the reference's fp32 torch path).

More than 2000 tissue patches of a synthetic slide go through the CUDA path (tile scan -> bf16 batch -> tcgen05 ResNet18
-> fp32 avgpool + fc) and through the fp32 CPU oracle (reference src/models/resnet.py:57-77, src/main.py:1004-1009:
``outputs.argmax(1)``).  No margin escape hatch: every disagreement counts, and is printed with its fp32 logit margin.

Two heads:
  * the BASELINE head (seeded random-init ``Linear(512, 2)``, the configuration BASELINE.json names): the >= 99.9 % bar;
  * a STRESS head with the same weight direction but its bias moved to the population median, so that the decision
    boundary cuts through the middle of the patch population (both classes ~50 %).  Patches whose fp32 margin is below
    the bf16 path's logit error (~1e-2 of a logit spread of ~0.08) must flip under ANY reduced-precision arithmetic; the
    test reports that rate and bounds the logit error instead of hiding the flips.
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from oracle import hipac_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIZE, SEED, LEVEL = 19040, 4321, 3          # 85 x 85 = 7225 candidates of 224 x 224 at level 3


def _level_image():
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import make_level
    img = np.empty((SIZE, SIZE, 3), np.uint8)

    def fill(r):
        img[r:r + 256] = make_level(SEED, LEVEL, SIZE, SIZE, r, min(r + 256, SIZE))

    with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as ex:
        list(ex.map(fill, range(0, SIZE, 256)))
    return img


def test_argmax_agreement_on_2000_plus_patches():
    from ss25_hierarchical_multiscale_image_classification_b200 import features
    from ss25_hierarchical_multiscale_image_classification_b200.preprocessing import extract_patches_tensor
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    img = _level_image()
    pb = extract_patches_tensor(torch.from_numpy(img).cuda(), None, LEVEL, layout="s2d16", want_u8=True)
    n = len(pb)
    assert n >= 2000, n
    net = orc.make_resnet18(seed=0, classifier=True)
    packed = features.pack_resnet18(net.state_dict(), "cuda")
    feats, logits = features.classify_tensor(pb.batch, packed)
    torch.cuda.synchronize()
    feats, logits = feats.cpu().numpy(), logits.cpu().numpy()
    u8 = pb.images_u8.cpu().numpy()
    # the uint8 patches the GPU produced are the reference's (identity resize at level 3): spot-check against the oracle
    want = orc.extract_patches_oracle(img[:2240], None, LEVEL, want_images=True)
    k = len(want["coords"])
    sel = np.nonzero(pb.coords.cpu().numpy()[:, 1] < 2240 - 223)[0]
    assert k == len(sel) and np.array_equal(u8[sel], want["images"])
    ref_f, ref_l = orc.resnet18_features_fp32(net, u8)

    cos = (feats * ref_f).sum(1) / (np.linalg.norm(feats, axis=1) * np.linalg.norm(ref_f, axis=1))
    maxrel = np.abs(feats - ref_f).max(1) / np.abs(ref_f).max(1)
    assert cos.min() >= 0.9995 and maxrel.max() <= 1e-2, (cos.min(), maxrel.max())

    # ---- BASELINE head: the bar, no forgiveness ----
    margin = np.abs(ref_l[:, 1] - ref_l[:, 0])
    flips = np.nonzero(logits.argmax(1) != ref_l.argmax(1))[0]
    agree = 1.0 - len(flips) / n
    d_err = np.abs((logits[:, 1] - logits[:, 0]) - (ref_l[:, 1] - ref_l[:, 0]))
    print(f"\nBASELINE head: n = {n}, agreement {100 * agree:.3f} %, flips at fp32 margins {margin[flips].round(6).tolist()}, "
          f"min margin {margin.min():.4f}, logit-difference error max {d_err.max():.2e} median {np.median(d_err):.2e}")
    assert agree >= 0.999, f"{len(flips)} argmax flips of {n}: margins {margin[flips].tolist()}"

    # ---- STRESS head: bias at the population median (GPU side: same features, fp32 dot product as k_avgpool_fc) ----
    w = (net.fc.weight[1] - net.fc.weight[0]).detach().numpy().astype(np.float32)
    d_ref, d_gpu = ref_f @ w, feats @ w
    med = np.float32(np.median(d_ref))
    sflips = np.nonzero((d_gpu > med) != (d_ref > med))[0]
    perr = np.abs(d_gpu - d_ref)
    spread = float(d_ref.std())
    rate = len(sflips) / n
    print(f"STRESS head (median-centred): class balance {float((d_ref > med).mean()):.3f}, logit spread (std) {spread:.4f}, "
          f"projection error max {perr.max():.2e} median {np.median(perr):.2e}, flips {len(sflips)} / {n} = {100 * rate:.2f} % "
          f"at fp32 margins <= {np.abs(d_ref[sflips] - med).max() if len(sflips) else 0:.2e}")
    # the bf16 path's logit error stays a small fraction of the logit spread, and only patches inside that band flip
    assert perr.max() <= 0.25 * spread, (perr.max(), spread)
    assert rate <= 0.03, rate
    rec = {"n": n, "baseline_head": {"agreement": agree, "flips": int(len(flips)), "min_margin_fp32": float(margin.min()),
                                     "logit_diff_error_max": float(d_err.max())},
           "stress_head": {"class_balance": float((d_ref > med).mean()), "flip_rate": rate, "flips": int(len(sflips)),
                           "logit_spread_std": spread, "projection_error_max": float(perr.max()),
                           "max_margin_of_a_flip": float(np.abs(d_ref[sflips] - med).max()) if len(sflips) else 0.0},
           "min_cos": float(cos.min()), "max_rel": float(maxrel.max())}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(rec, open(os.path.join(out_dir, "argmax_report.json"), "w"), indent=1)
