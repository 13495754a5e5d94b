"""A short run of the random-configuration parity sweep (tools/fuzz_parity.py): random batch sizes, keypoint counts
(1-point images, per-pair counts), input dimensions, precisions and launch modes against the CPU oracle."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_random_configurations_against_oracle():
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "fuzz_parity.py"), "24", "3"], capture_output=True, text=True,
                         timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "24 of 24 cases ok" in res.stdout
