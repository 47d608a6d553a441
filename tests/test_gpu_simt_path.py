"""The FFMA kernels stay covered: re-run the layer parity tests in a subprocess with the tensor-core path disabled."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layers_with_tensor_cores_disabled():
    env = dict(os.environ, PTGNN_B200_DISABLE_TC="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_layers.py", "tests/test_gpu_gnn.py", "-q", "-m", "gpu",
                        "-x", "-k", "not full_size"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
