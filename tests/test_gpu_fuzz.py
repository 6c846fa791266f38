"""GPU (-m gpu): randomized differential test of the fused scoring kernel against the oracle
(scripts/gpu_fuzz_score.py): random shapes x operand dtypes x CTA modes x column-split on/off x DAC on/off."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 11])
def test_fused_scoring_fuzz(cuda_lib, seed):
    env = {k: v for k, v in os.environ.items() if k not in ("CCAL_SCORE_CTAS", "CCAL_SCORE_NOSPLIT")}
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz_score.py"), str(seed), "20"],
                         capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0 and "fuzz ok" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
