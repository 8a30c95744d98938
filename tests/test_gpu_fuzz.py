"""Random-shape sweep through every forward kernel variant (tools/fuzz_shapes.py): ragged N, odd group counts,
K = 64*odd (generic kernel), every M regime boundary (1..4 packed FP4, <= 8, <= 16, token-blocked, tensor-core),
bias on/off, activation magnitudes 1e-3..50 — all within the 1e-2 parity gate of the dequantise-then-FP32-GEMM
reference."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("seed", [7, 8])
def test_random_shapes(seed):
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "fuzz_shapes.py"), "70", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "bad 0" in r.stdout
