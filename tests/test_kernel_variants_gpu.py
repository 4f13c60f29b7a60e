"""The conv-stack kernels that are NOT on the default path stay correct: every A/B switch of csrc/resnet18.cu selects a
different kernel for some layers (single-CTA row / im2col kernels instead of the CTA-pair ones, register epilogues
instead of the TMA-staged ones, streamed instead of resident weights, no programmatic dependent launch).  The switches
are read once per process, so each combination runs the per-layer / fused-stem / fused-projection / end-to-end parity
tests of test_stage2_gpu.py in a subprocess of its own."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "single_cta_everywhere": {"HIPAC_CTA_PAIRS": "0"},
    "single_cta_stem_layer1_layer34": {"HIPAC_CTA_PAIRS_STEM": "0", "HIPAC_CTA_PAIRS_C64": "0", "HIPAC_CTA_PAIRS_C256": "0"},
    "register_epilogues_no_pdl_streamed_weights": {"HIPAC_TMA_EPILOGUE": "0", "HIPAC_TMA_EPILOGUE_C128": "0", "HIPAC_PDL": "0",
                                                   "HIPAC_RESIDENT_B": "0", "HIPAC_CTA_PAIRS_C64": "0"},
    "pair_im2col_c128": {"HIPAC_CTA_PAIRS_C128_IM2COL": "1"},
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_non_default_kernels_keep_parity(name):
    env = dict(os.environ, **VARIANTS[name])
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_stage2_gpu.py"), "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
           "-k", "conv_layer or fused_stem or fused_projection or fp32_oracle or golden"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, f"{name}: {tail}"
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
