"""Tests that need TWO GPUs in one box (skipped otherwise; run with ``gpurun --gpus 2``)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs2
def test_two_devices_in_one_process():
    """Scan + forward on cuda:0, then on cuda:1, from ONE process: the library's per-device state (dynamic shared-memory
    opt-in, constant tables, SM count, count event) must follow the current device."""
    from oracle import hipac_oracle as orc
    from ss25_hierarchical_multiscale_image_classification_b200 import features, pipeline
    from ss25_hierarchical_multiscale_image_classification_b200.synthetic import SyntheticSlide
    slide = SyntheticSlide(6000, 5200, seed=21)
    img, mask = slide.level_array(2), slide.lesion_mask(2)
    net = orc.make_resnet18(seed=0, classifier=True)
    outs = []
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(d):
            packed = features.pack_resnet18(net.state_dict(), d)
            r = pipeline.process_level(torch.from_numpy(img).to(d), torch.from_numpy(mask).to(d), 2, packed)
            torch.cuda.synchronize(d)
            outs.append((r.coords.cpu(), r.labels.cpu(), r.features.cpu(), r.logits.cpu()))
    want = orc.extract_patches_oracle(img, mask, 2, want_images=False)
    for o in outs:
        assert np.array_equal(o[0].numpy(), want["coords"]) and np.array_equal(o[1].numpy(), want["labels"])
        assert torch.equal(o[2], outs[0][2]) and torch.equal(o[3], outs[0][3])


@needs2
@pytest.mark.parametrize("level", [1, 3])
def test_two_ranks_equal_one_rank_bitwise(level):
    """torchrun, 2 ranks over NCCL: gathered (coords, labels, feature bits, logit bits) == the single-rank run."""
    port = 29500 + (os.getpid() % 400) + level
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py"), "--level", str(level),
           "--width", "5000" if level == 1 else "3000", "--rows-per-rank", "2300" if level == 1 else "1300"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        open(os.path.join(out_dir, f"check_sharded_l{level}.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    err = "\n".join(l for l in r.stderr.splitlines() if "Error" in l or "error" in l or "Traceback" in l or "  File" in l)
    assert r.returncode == 0, (r.stdout[-1500:], err[-3000:])
    rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["equals_single_rank"] and rec["all_ranks_hold_identical_bytes"] and rec["world"] == 2
    assert rec["survivors_gathered"] == rec["survivors_single_rank"] > 0
