"""tools/run_config.py (the other BASELINE configurations through the public API) at toy sizes: runs, prints one JSON
line, and its rank shares add up to the single-rank result."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_config.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_heatmap_rank_shares_add_up():
    whole = _run("heatmap", "--size", "3000")
    parts = [_run("heatmap", "--size", "3000", "--ranks", "3", "--rank", str(r)) for r in range(3)]
    assert whole["grid"] == [14, 14] and whole["candidates_this_rank"] == 196
    assert sum(p["candidates_this_rank"] for p in parts) == 196
    assert sum(p["survivors_this_rank"] for p in parts) == whole["survivors_this_rank"] > 0
    assert abs(sum(p["heatmap_sum"] for p in parts) - whole["heatmap_sum"]) < 1e-2 * max(1.0, whole["heatmap_sum"])


def test_pyramid_and_batch_run():
    p = _run("pyramid", "--size", "4096")
    assert p["candidates"] == {"0": 361, "1": 100, "2": 25, "3": 9} and sum(p["survivors"].values()) > 0
    b = _run("batch", "--slides", "2", "--size", "2048", "--level", "2")
    assert b["slides"] == 2 and b["survivors_gathered"] == b["survivors_this_rank"] > 0
