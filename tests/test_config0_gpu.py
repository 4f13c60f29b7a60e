"""BASELINE configs[0] (4096 x 4096 level-3 image, 361 candidates) end to end on BOTH arms of `bench.py --config 0`: the
unmodified reference runs the whole config on the host (PNG files, DataLoader), our kernels run it on the GPU with the
weights the reference drew, and every survivor is compared (file names = coords + labels, 512-d features)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config0_whole_config_equals_reference_as_written():
    sys.path.insert(0, ROOT)
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("baseline/_ref (the reference's own source files) is not present on this box")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "0", "--steps", "2", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    par = line["parity"]
    assert line["same_config"] and line["cpu_baseline"]["kind"] == "reference"
    assert par["n_candidates"] == 361 and par["n_survivors_ours"] == par["n_survivors_reference"] > 50
    assert par["file_names_equal"] and par["labels_equal_reference_npy"]
    assert par["min_cos"] >= 0.9995 and par["max_rel"] <= 1e-2 and par["pass"]      # north_star's feature tolerance
