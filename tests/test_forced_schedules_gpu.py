"""The op-level parity tests again under every forced kernel schedule.  The library picks tile / persistent / CTA-pair
(128- or 160-column) kernels and TMA / pair weight-gradient kernels per shape; with the default heuristics the small
shapes of tests/test_ops_gpu.py only ever reach some of them, so each schedule is forced here (the overrides are read
once per process, hence subprocesses) to keep ragged tiles, odd tile counts and K tails covered for all of them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FORCED = [
    {"CAVP_IGEMM_WS": "0"},                                   # one tile per CTA
    {"CAVP_IGEMM_WS": "1"},                                   # persistent
    {"CAVP_IGEMM_WS": "2", "CAVP_IGEMM_BN160": "0"},          # CTA pair, 128-column tiles
    {"CAVP_IGEMM_WS": "2", "CAVP_IGEMM_BN160": "1"},          # CTA pair, 160-column tiles
    {"CAVP_IGEMM_WS": "2", "CAVP_IGEMM_BN256": "1"},          # CTA pair, 256-column tiles wherever N % 256 == 0
    {"CAVP_IGEMM_BN256": "0"},                                # never the 256-column kernel
    {"CAVP_IGEMM_WS": "2", "CAVP_IGEMM_LIN": "0"},            # linear layers on the generic pair kernel's fast path
    {"CAVP_WGRAD_TMA": "1", "CAVP_WGRAD_PAIR": "1"},          # every weight gradient through the TMA + pair kernel
    {"CAVP_WGRAD_TMA": "1", "CAVP_WGRAD_PAIR": "0"},          # TMA-fed single-CTA weight gradient
    {"CAVP_WGRAD_TMA": "0"},                                  # thread-gathered weight gradient
]


@pytest.mark.parametrize("env", FORCED, ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_ops_under_forced_schedule(env):
    full = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_ops_gpu.py"), "-m", "gpu", "-x",
                        "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=full, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
